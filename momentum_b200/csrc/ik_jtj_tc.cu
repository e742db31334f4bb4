// K2 on the 5th-generation tensor cores: per-instance [J r]^T [J r] (lower triangle) with
// tcgen05.mma kind::tf32, fp32 accumulators in TMEM, operands staged global -> shared by TMA.
//
//   reference computation: jtj.selfadjointView<Lower>().rankUpdate(J^T); jtr += J^T r
//   (momentum/solver/solver_function.cpp:113-116, gauss_newton_solver.cpp:215-216)
//
// Data layout. The device Jacobian of an instance is R = numCols + 1 "column vectors" of ldJ floats
// (column c of J at J + c*ldJ, the residual r as vector numCols): a K-major [R x K] operand, which is
// exactly what both UMMA operands want for D = X X^T with X = [J r]^T. D's last row is J^T r, so Jtr
// costs nothing extra. One TMA box (boxRows x 32 floats, SWIZZLE_128B) per 32-row K block lands the
// whole K slab of an instance; out-of-range rows are zero-filled by TMA.
//
// Precision. kind::tf32 keeps 10 mantissa bits, so a single pass gives ~1e-3 relative error, which
// would change the Gauss-Newton path of the under-determined IK problems (m < n). The default is the
// 3xTF32 split: x = hi + lo with hi = tf32(x), lo = tf32(x - hi); D += hi*hi + hi*lo + lo*hi. A group of
// converter warps derives hi (in place) and lo (second buffer, same swizzled positions) from the raw
// fp32 slab that TMA delivered.
//
// Work items. The result is cut into row tiles of 128 (UMMA M = 128); row tile t needs the columns 128 t .. only (consumers read the
// upper triangle). An item is (row tile t, a chunk of at most 256 of those columns) = one accumulator of at most 256 TMEM columns:
//   * diagonal item: columns 128 t .. 128 t + 255 - ONE box of up to 256 operand rows, whose first 128 rows are also the A operand;
//   * far item (only when numCols + 1 > 128 t + 256): the remaining columns - a B box of those rows plus a separate 128-row A box.
// numCols + 1 <= 256 gives the two diagonal items of the first version of this kernel; up to 512 (bodyhands300: 425) there are at most
// four row tiles and at most one far item per tile. Two TMEM slots of 256 columns alternate between consecutive items, so the epilogue
// of one item drains while the MMAs of the next run.
//
// Roles in one persistent CTA (1 CTA / SM, 448 threads):
//   warp 0        TMA producer (one elected lane)
//   warp 1        TMEM allocator + MMA issuer (one elected lane)
//   warps 2..9    hi/lo converters
//   warps 10..13  epilogue: TMEM -> registers -> shared (32x32 transpose stage) -> global rows (upper triangle) of the symmetric matrix,
//                 every store instruction covering four full 128-byte lines
// Pipelines: smem ring full -> converted -> (MMA) -> empty; TMEM full/empty between MMA and epilogue.
#include "ik_jtj_tc.cuh"
#include "ik_ptx.cuh"

#include <cuda.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

namespace mb2 {

namespace {

constexpr int kTcThreads = 448;      // warp 0 TMA, warp 1 MMA, warps 2-9 converters, warps 10-13 epilogue
constexpr int kConvThreads = 256;
constexpr int kStageRowFloats = 36;  // epilogue staging row: 32 floats + 4 pad (16-byte aligned, conflict-free float4 rows)
constexpr int kKBlock = 32;         // floats per K block = one 128-byte swizzle row
constexpr int kRowBytes = 128;
constexpr int kUmmaK = 8;           // tf32: 32 bytes of K per instruction

constexpr int kMaxItems = 8;
struct TcItem {
  int aRow0;    // first operand row of the A box = first row of the row tile (128 t)
  int bRow0;    // first operand row of the B box = first column of the chunk
  int n;        // UMMA N: columns of the chunk, a multiple of 16, <= 256
  int bBoxRows; // 256 or 128: which tensor map loads the B box
  int sepA;     // far item: the A rows are not in the B box, a second (128-row) box is loaded behind the B region
};
struct TcParams {
  int batch;
  int ns, numCols, ldJ, kBlocks;
  float* H;
  int ldH;
  const int32_t* active;
  int passes;     // 3 or 1
  int numItems;
  TcItem items[kMaxItems];
  int bRegionRows; // operand rows reserved for the B box in a stage (256, or 128 when every item fits a 128-row box)
  int planeRows;   // bRegionRows (+ 128 when some item needs a separate A box): rows of one hi (or lo) plane of a stage
  int tmemCols;    // 2 slots x 256 (or the power of two that holds 2 x the widest item)
  int slotCols;
  int stages;
  size_t hStride;
  float* G;       // optional [batch][ldG]: J^T r as a contiguous vector (the scheduled Cholesky bulk-copies it)
  int ldG;
  int profile;    // MB2_TC_PROFILE=1: block 0 prints per-role wait/busy cycles (debug aid, off by default)
};

__device__ __forceinline__ void ummaTf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmemD),
      "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void ummaCommit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t toTf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void tmemLoad16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmemLoad32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) (8 rows * 128 B)
// | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t makeSmemDesc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6) | a_format TF32 (2) [7,10) | b_format TF32 (2) [10,13)
// | a_major K (0) [15] | b_major K (0) [16] | N>>3 [17,23) | M>>4 [24,29)
__device__ __forceinline__ uint32_t makeInstrDesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(kTcThreads, 1) jtjTensorKernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmapHalf, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smemRaw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smemAddr(smemRaw) + 1023u) & ~1023u;
  const uint32_t slabBytes = (uint32_t)p.planeRows * kRowBytes;        // one K block of every operand row of an item: [B box | A box of far items]
  const uint32_t stageBytes = slabBytes * (p.passes == 3 ? 2u : 1u);   // hi [+ lo]
  const uint32_t aRegionOff = (uint32_t)p.bRegionRows * kRowBytes;     // where a far item's A box sits inside a plane
  const uint32_t barBase = base + stageBytes * p.stages;
  auto fullBar = [&](int s) { return barBase + 8u * s; };
  auto convBar = [&](int s) { return barBase + 8u * (p.stages + s); };
  auto emptyBar = [&](int s) { return barBase + 8u * (2 * p.stages + s); };
  // Two TMEM slots with their own full / empty barriers alternate between consecutive work items (file header)
  auto tmemFullBar = [&](int t) { return barBase + 8u * (3 * p.stages + t); };
  auto tmemEmptyBar = [&](int t) { return barBase + 8u * (3 * p.stages + 2 + t); };
  const uint32_t tmemSlot = barBase + 8u * (3 * p.stages + 4);
  const uint32_t stageOff = ((tmemSlot + 16u - base) + 15u) & ~15u; // epilogue transpose stages: 4 warps x 32 rows x 36 floats
  uint8_t* gen = smemRaw + (base - smemAddr(smemRaw)); // generic pointer to the aligned base

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbarInit(fullBar(s), 1); mbarInit(convBar(s), kConvThreads); mbarInit(emptyBar(s), 1); }
    for (int t = 0; t < 2; ++t) { mbarInit(tmemFullBar(t), 1); mbarInit(tmemEmptyBar(t), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmemSlot), "r"((uint32_t)p.tmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcFenceBefore();
  __syncthreads();
  tcFenceAfter();
  const uint32_t tmemBase = *reinterpret_cast<volatile uint32_t*>(gen + (tmemSlot - base));

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      long long wEmpty = 0, tStart = clock64();
      for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        if (p.active != nullptr && p.active[b] == 0) continue;
        for (int it = 0; it < p.numItems; ++it) {
          const TcItem I = p.items[it];
          const uint32_t bBytes = (uint32_t)I.bBoxRows * kRowBytes;
          for (int kb = 0; kb < p.kBlocks; ++kb) {
            long long t0 = clock64();
            mbarWaitRelaxed(emptyBar(s), ph ^ 1u);
            wEmpty += clock64() - t0;
            mbarExpectTx(fullBar(s), bBytes + (I.sepA ? 128u * kRowBytes : 0u));
            tmaLoad3d(base + stageBytes * s, I.bBoxRows == 128 ? &tmapHalf : &tmap, kb * kKBlock, I.bRow0, b, fullBar(s));
            if (I.sepA) tmaLoad3d(base + stageBytes * s + aRegionOff, &tmapHalf, kb * kKBlock, I.aRow0, b, fullBar(s));
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
      if ((p.profile & 1) && blockIdx.x == 0) printf("tc-profile producer: total %lld waitEmpty %lld\n", clock64() - tStart, wEmpty);
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0, tph[2] = {0, 0}, count = 0;
      long long wTmem = 0, wConv = 0, tStart = clock64();
      for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        if (p.active != nullptr && p.active[b] == 0) continue;
        for (int it = 0; it < p.numItems; ++it, ++count) {
          const TcItem I = p.items[it];
          const int slot = int(count & 1u);
          long long t0 = clock64();
          mbarWait(tmemEmptyBar(slot), tph[slot] ^ 1u); // the epilogue has drained the item that used this slot before
          wTmem += clock64() - t0;
          tcFenceAfter();
          const uint32_t dcol = tmemBase + uint32_t(slot * p.slotCols);
          const uint32_t idesc = makeInstrDesc(128, I.n);
          const uint32_t aOff = I.sepA ? aRegionOff : 0u; // diagonal item: the A rows are the first 128 rows of the B box
          for (int kb = 0; kb < p.kBlocks; ++kb) {
            t0 = clock64();
            mbarWait(convBar(s), ph);
            wConv += clock64() - t0;
            tcFenceAfter();
            const uint32_t hi = base + stageBytes * s, lo = hi + slabBytes;
            for (int pass = 0; pass < p.passes; ++pass) {
              const uint32_t aBase = (pass == 2 ? lo : hi) + aOff; // hi*hi, hi*lo, lo*hi
              const uint32_t bBase = pass == 1 ? lo : hi;
#pragma unroll
              for (int k4 = 0; k4 < kKBlock / kUmmaK; ++k4) {
                const uint32_t acc = (kb == 0 && pass == 0 && k4 == 0) ? 0u : 1u;
                ummaTf32(dcol, makeSmemDesc(aBase + k4 * 32u), makeSmemDesc(bBase + k4 * 32u), idesc, acc);
              }
            }
            ummaCommit(emptyBar(s)); // smem slab reusable once these MMAs have read it
            if (kb == p.kBlocks - 1) ummaCommit(tmemFullBar(slot));
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
          tph[slot] ^= 1u;
        }
      }
      if ((p.profile & 1) && blockIdx.x == 0) printf("tc-profile mma: total %lld waitTmemEmpty %lld waitConverted %lld\n", clock64() - tStart, wTmem, wConv);
    }
  } else if (warp < 10) {
    // ---------------- converters: raw fp32 -> tf32 hi (in place) and lo ----------------
    const int ct = threadIdx.x - 64; // 0..255
    int s = 0;
    uint32_t ph = 0;
    long long wFull = 0, tStart = clock64();
    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
      if (p.active != nullptr && p.active[b] == 0) continue;
      for (int it = 0; it < p.numItems; ++it) {
        const TcItem I = p.items[it];
        // two segments of the plane: the B box, and (far items) the A box behind the B region
        const int vecsB = I.bBoxRows * kRowBytes / 16, vecsA = I.sepA ? 128 * kRowBytes / 16 : 0, offA = int(aRegionOff / 16u);
        for (int kb = 0; kb < p.kBlocks; ++kb) {
          long long t0 = clock64();
          mbarWaitRelaxed(fullBar(s), ph);
          wFull += clock64() - t0;
          float4* hi = reinterpret_cast<float4*>(gen + stageBytes * s);
          float4* lo = reinterpret_cast<float4*>(gen + stageBytes * s + slabBytes);
#pragma unroll 4
          for (int j = ct; j < vecsB + vecsA; j += kConvThreads) {
            const int i = j < vecsB ? j : offA + (j - vecsB);
            const float4 x = hi[i];
            uint4 h;
            h.x = toTf32(x.x); h.y = toTf32(x.y); h.z = toTf32(x.z); h.w = toTf32(x.w);
            reinterpret_cast<uint4*>(hi)[i] = h;
            if (p.passes == 3) {
              uint4 l;
              l.x = toTf32(x.x - __uint_as_float(h.x)); l.y = toTf32(x.y - __uint_as_float(h.y));
              l.z = toTf32(x.z - __uint_as_float(h.z)); l.w = toTf32(x.w - __uint_as_float(h.w));
              reinterpret_cast<uint4*>(lo)[i] = l;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy stores -> visible to the tensor core (async proxy)
          mbarArrive(convBar(s));
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
    if ((p.profile & 1) && blockIdx.x == 0 && ct == 0) printf("tc-profile converter: total %lld waitFull %lld\n", clock64() - tStart, wFull);
  } else {
    // ---------------- epilogue: TMEM -> global, column-major lower triangle of [JtJ; Jtr] ----------------
    const int q = warp & 3; // TMEM lane quarter this warp may access
    float* stage = reinterpret_cast<float*>(gen + stageOff) + (warp - 10) * 32 * kStageRowFloats;
    uint32_t eph[2] = {0, 0}, count = 0;
    long long wFullT = 0, tStart = clock64(), tLd = 0, nChunks = 0, tSts = 0, tStg = 0;
    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
      if (p.active != nullptr && p.active[b] == 0) continue;
      float* H = p.H + (size_t)b * p.hStride;
      for (int it = 0; it < p.numItems; ++it, ++count) {
        const TcItem I = p.items[it];
        const int slot = int(count & 1u);
        long long t0 = clock64();
        mbarWaitRelaxed(tmemFullBar(slot), eph[slot]);
        wFullT += clock64() - t0;
        tcFenceAfter();
        const int rowBase = I.aRow0 + q * 32;                     // first of this warp's 32 rows of [J r]^T [J r]
        const int cb = I.bRow0;                                   // first matrix column of the item's accumulator
        const int row = rowBase + lane;
        const int i = row < p.ns ? row : (row == p.numCols ? p.ns : -1); // row of the (ns+1) system; -1: not wanted
        if (__reduce_max_sync(0xffffffffu, i) < 0) { tcFenceBefore(); mbarArrive(tmemEmptyBar(slot)); eph[slot] ^= 1u; continue; } // warp-uniform: nothing to write
        const int nT = I.n;
        const uint32_t colBase = tmemBase + ((uint32_t)(q * 32) << 16) + uint32_t(slot * p.slotCols);
        if (p.ns == p.numCols) {
          // solver path: full rows. 32x32 blocks go TMEM -> registers -> shared (row = TMEM lane) -> global, re-mapped so
          // that one store instruction writes four rows x 128 contiguous bytes.
          // the four rows x 128 bytes this lane stores per instruction are the same for every chunk of the tile
          float* rowPtr[8];
          int rowIo[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rowAbs = rowBase + 4 * k + (lane >> 3);
            rowIo[k] = rowAbs < p.ns ? rowAbs : (rowAbs == p.numCols ? p.ns : (1 << 30)); // 1 << 30: never stored (col + 3 >= row fails)
            rowPtr[k] = H + (size_t)(rowIo[k] < (1 << 30) ? rowIo[k] : 0) * p.ldH + 4 * (lane & 7);
          }
          for (int c0 = 0; c0 < nT; c0 += 32) {
            if (cb + c0 + 32 <= rowBase && !(cb + c0 <= p.numCols && p.numCols < cb + c0 + 32)) continue; // entirely left of the diagonal
            float v[32];
            const long long tl0 = p.profile ? clock64() : 0;
            if (c0 + 32 <= nT) tmemLoad32(colBase + (uint32_t)c0, v);
            else { tmemLoad16(colBase + (uint32_t)c0, v); for (int k = 16; k < 32; ++k) v[k] = 0.f; }
            if (p.profile) { tLd += clock64() - tl0; ++nChunks; }
#pragma unroll
            for (int g4 = 0; g4 < 8; ++g4)
              *reinterpret_cast<float4*>(stage + lane * kStageRowFloats + 4 * g4) = make_float4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
            __syncwarp();
            if (p.profile) { tSts += clock64() - tl0; }
            if (p.G != nullptr && cb + c0 <= p.numCols && p.numCols < cb + c0 + 32 && i >= 0 && i < p.ns) // column numCols = J^T r
              p.G[(size_t)b * p.ldG + i] = stage[lane * kStageRowFloats + (p.numCols - cb - c0)];
            const int c = cb + c0 + 4 * (lane & 7); // matrix column of this lane's float4
            float4 vals[8]; // all shared-memory reads first: the compiler cannot move them across the global stores itself (possible aliasing)
#pragma unroll
            for (int k = 0; k < 8; ++k) vals[k] = *reinterpret_cast<const float4*>(stage + (4 * k + (lane >> 3)) * kStageRowFloats + 4 * (lane & 7));
            if (!(p.profile & 2) && c < p.ldH) {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (c + 3 >= rowIo[k]) *reinterpret_cast<float4*>(rowPtr[k] + (cb + c0)) = vals[k]; // upper triangle (col >= row) only: everything downstream reads H(min, max)
            }
            __syncwarp();
            if (p.profile) { tStg += clock64() - tl0; }
          }
        } else {
          // leading-block request (ns < numCols, getJtJR parity entry): columns ns.. are skipped except the residual column
          float* Hrow = H + (size_t)(i < 0 ? 0 : i) * p.ldH;
          for (int c0 = 0; c0 < nT; c0 += 16) {
            float v[16];
            tmemLoad16(colBase + (uint32_t)c0, v);
            if (i < 0) continue;
#pragma unroll
            for (int cc = 0; cc < 16; ++cc) {
              const int c = cb + c0 + cc;
              if (c < p.ns) Hrow[c] = v[cc];
              else if (c == p.numCols) Hrow[p.ns] = v[cc];
            }
          }
        }
        tcFenceBefore();
        mbarArrive(tmemEmptyBar(slot));
        eph[slot] ^= 1u;
      }
    }
    if ((p.profile & 1) && blockIdx.x == 0 && lane == 0)
      printf("tc-profile epilogue q%d: total %lld waitTmemFull %lld chunks %lld | cumulative per chunk: tmemLoad %lld +sts %lld +stg %lld\n", q, clock64() - tStart, wFullT, nChunks,
             tLd / (nChunks ? nChunks : 1), tSts / (nChunks ? nChunks : 1), tStg / (nChunks ? nChunks : 1));
  }
  tcFenceBefore();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"((uint32_t)p.tmemCols) : "memory");
  }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encodeTiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int roundUpI(int v, int m) { return (v + m - 1) / m * m; }

struct Shape {
  int rows, r16, numItems, bRegionRows, planeRows, maxN;
  bool anySepA;
  TcItem items[kMaxItems];
};
Shape shapeFor(int numCols) {
  Shape s{};
  s.rows = numCols + 1;
  s.r16 = roundUpI(s.rows, 16);
  const int tiles = (s.rows + 127) / 128;
  for (int t = tiles - 1; t >= 0; --t) { // last row tile first, as the kernel's roles walk them
    const int c0 = 128 * t, left = s.r16 - c0;
    TcItem d{};
    d.aRow0 = c0; d.bRow0 = c0; d.n = left < 256 ? left : 256; d.bBoxRows = d.n <= 128 ? 128 : 256; d.sepA = 0;
    s.items[s.numItems++] = d;
    if (left > 256) { // the columns beyond the diagonal box: a far item with its own A box
      TcItem f{};
      f.aRow0 = c0; f.bRow0 = c0 + 256; f.n = left - 256; f.bBoxRows = f.n <= 128 ? 128 : 256; f.sepA = 1;
      s.items[s.numItems++] = f;
      s.anySepA = true;
    }
  }
  for (int i = 0; i < s.numItems; ++i) { s.bRegionRows = s.bRegionRows > s.items[i].bBoxRows ? s.bRegionRows : s.items[i].bBoxRows; s.maxN = s.maxN > s.items[i].n ? s.maxN : s.items[i].n; }
  s.planeRows = s.bRegionRows + (s.anySepA ? 128 : 0);
  return s;
}

} // namespace

cudaError_t makeTensorMap3d(CUtensorMap* map, const float* base, const uint64_t dims[3], const uint64_t strideBytes[2], const uint32_t box[3], int swizzleBytes) {
  if (encodeTiled() == nullptr) return cudaErrorNotSupported;
  const cuuint64_t d[3] = {dims[0], dims[1], dims[2]};
  const cuuint64_t st[2] = {strideBytes[0], strideBytes[1]};
  const cuuint32_t bx[3] = {box[0], box[1], box[2]};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUtensorMapSwizzle sw = swizzleBytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzleBytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                              : swizzleBytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUresult r = encodeTiled()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

bool jtjTensorSupported(int ns, int numCols, int ldJ) {
  if (encodeTiled() == nullptr) return false;
  return ns >= 1 && ns <= numCols && numCols + 1 <= 512 && (ldJ % kKBlock) == 0; // at most four row tiles, at most one far item per tile
}

cudaError_t launchJtJTensor(const JtJArgs& a, int passes, cudaStream_t stream) {
  if (!jtjTensorSupported(a.ns, a.numCols, a.ldJ)) return cudaErrorNotSupported;
  const Shape sh = shapeFor(a.numCols);
  // TMA descriptors over the device Jacobian [batch][numCols + 1][ldJ] (innermost first): 256-row and 128-row boxes of one 32-float K block
  CUtensorMap map;
  const uint64_t dims[3] = {(uint64_t)a.ldJ, (uint64_t)(a.numCols + 1), (uint64_t)a.batch};
  const uint64_t strides[2] = {(uint64_t)a.ldJ * sizeof(float), (uint64_t)(a.numCols + 1) * a.ldJ * sizeof(float)};
  const uint32_t box[3] = {(uint32_t)kKBlock, 256u, 1u};
  cudaError_t me = makeTensorMap3d(&map, a.jacobian, dims, strides, box, 128);
  if (me != cudaSuccess) return me;
  CUtensorMap mapHalf;
  const uint32_t boxHalf[3] = {(uint32_t)kKBlock, 128u, 1u};
  me = makeTensorMap3d(&mapHalf, a.jacobian, dims, strides, boxHalf, 128);
  if (me != cudaSuccess) return me;
  TcParams p{};
  p.batch = a.batch;
  p.ns = a.ns;
  p.numCols = a.numCols;
  p.ldJ = a.ldJ;
  p.kBlocks = (a.kRows + kKBlock - 1) / kKBlock;
  p.H = a.H;
  p.ldH = a.ldH;
  p.active = a.active;
  p.passes = passes == 3 ? 3 : 1;
  p.numItems = sh.numItems;
  for (int i = 0; i < sh.numItems; ++i) p.items[i] = sh.items[i];
  p.bRegionRows = sh.bRegionRows;
  p.planeRows = sh.planeRows;
  p.slotCols = 32;
  while (p.slotCols < sh.maxN) p.slotCols <<= 1;
  p.tmemCols = 2 * p.slotCols; // a power of two >= 64
  p.hStride = a.hStride;
  p.G = a.g;
  p.ldG = a.ldG;
  p.profile = getenv("MB2_TC_PROFILE") != nullptr ? atoi(getenv("MB2_TC_PROFILE")) : 0;
  const size_t stageBytes = size_t(sh.planeRows) * kRowBytes * (p.passes == 3 ? 2 : 1);
  int stages = int((196 * 1024) / stageBytes);
  if (stages > 6) stages = 6;
  if (stages < 2) return cudaErrorInvalidConfiguration;
  p.stages = stages;
  const size_t smem = stageBytes * stages + 1024 /*alignment slack*/ + 8 * (3 * stages + 4) + 48 + 4 * 32 * kStageRowFloats * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(jtjTensorKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.batch < sms ? a.batch : sms;
  jtjTensorKernel<<<grid, kTcThreads, smem, stream>>>(map, mapHalf, p);
  return cudaGetLastError();
}

} // namespace mb2
