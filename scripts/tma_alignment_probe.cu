#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include "../momentum_b200/csrc/ik_ptx.cuh"  // build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe scripts/tma_alignment_probe.cu -lcuda; run: ./tma_probe <swizzle 0|32|64|128> <c0> <c1>
using namespace mb2;
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap m, int c0, int c1, int c2, float* out) {
  extern __shared__ __align__(1024) float sm[];
  float* tiles = sm + (((1024u - (smemAddr(sm) & 1023u)) & 1023u) >> 2);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(tiles + 256);
  const uint32_t b = smemAddr(bar);
  if (threadIdx.x == 0) { mbarInit(b, 1); fenceBarrierInit(); mbarExpectTx(b, 1024); }
  __syncthreads();
  if (threadIdx.x == 0) tmaLoad3d(smemAddr(tiles), &m, c0, c1, c2, b);
  mbarWait(b, 0);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = tiles[i];
}
int main(int argc, char** argv) {
  int sw = argc > 1 ? atoi(argv[1]) : 64;
  int B = 2, rows = 221, ld = 224;
  std::vector<float> h(size_t(B) * rows * ld);
  for (size_t i = 0; i < h.size(); ++i) h[i] = float(i % 100000);
  float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 1024);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t st[2] = {(cuuint64_t)ld * 4, (cuuint64_t)rows * ld * 4};
  cuuint32_t box[3] = {16, 16, 1}, es[3] = {1, 1, 1};
  CUtensorMapSwizzle s = sw == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : sw == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = ((EncodeFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, s, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d (swizzle %d)\n", (int)r, sw);
  int c0 = argc > 2 ? atoi(argv[2]) : 15, c1 = argc > 3 ? atoi(argv[3]) : 3;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
  k<<<1, 64, 8192>>>(m, c0, c1, 1, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  float out[256]; cudaMemcpy(out, o, 1024, cudaMemcpyDeviceToHost);
  for (int r2 = 0; r2 < 3; ++r2) { for (int c = 0; c < 16; ++c) printf("%7.0f", out[r2 * 16 + c]); printf("\n"); }
  printf("expect row0: %f ...\n", h[size_t(1) * rows * ld + size_t(c1) * ld + c0]);
  return 0;
}
