// Device building blocks of the level-scheduled tile-sparse Cholesky (see ik_chol_sched.h).
//
// Storage: every structurally non-zero lower tile is a 16x16 fp32 block in shared memory. Rows are
// 64 bytes; the four float4 groups of a row are XOR-swizzled with ((row >> 1) & 3) so that the two
// access shapes used everywhere — "16 lanes read 16 consecutive floats of one row" and "lane r reads
// float4 group g of row r" — are both bank-conflict free.
// Panel tiles L(I,K) are stored TRANSPOSED once solved (T[c][r] = L[r][c]): update tasks then read
// both operands along rows (float2 / float4), and the substitution phases read along rows too.
//
// Diagonal tiles and the substitutions are mapped to half-warps (one 16-row tile each, width-16 shuffles); the 16x16x16
// products (panel tiles, update tasks, the Gram kernel) are one warp each on mma.sync with the three-term TF32 split.
// The host build (tests/emu) runs the same arithmetic lane by lane so that the schedule and the tile algebra are
// validated without a GPU.
#pragma once

#include <cmath>
#include <cstring>

#include "ik_types.h"

#include <vector_types.h> // float2 / float4 (CUDA toolkit header, usable from plain C++ too)

namespace mb2 {

struct CholSchedDev {
  int32_t n, nPad, numTileCols, numTiles, numLevels;
  // every table below lives in one contiguous int32 blob (copied to shared memory by the kernels)
  const int32_t* blob;
  int32_t blobInts;
  const int32_t* perm;        // [nPad] slot -> device column, -1 = padding
  const int32_t* pos;         // [n] device column -> slot (monotone: device columns are in elimination order)
  const int32_t* tileIdTable; // [numTileCols^2]
  const int32_t* tileRow;
  const int32_t* tileCol;
  const int32_t* diagTile;
  const int32_t* levelColStart;
  const int32_t* levelCols;
  const int32_t* levelPanelStart;
  const int32_t* panelTile;
  const int32_t* panelDiag;
  const int32_t* levelTaskStart;
  const int32_t* taskDst;
  const int32_t* taskPairStart;
  const int32_t* pairA;
  const int32_t* pairB;
  const int32_t* levelVTaskStart;
  const int32_t* vtaskRow;
  const int32_t* vtaskSrcStart;
  const int32_t* vsrcTile;
  const int32_t* vsrcCol;
  const int32_t* colPanelStart;
  const int32_t* colPanelTile;
  const int32_t* colPanelRow;
  const int32_t* levelOrderStart8;  // update-task assignment for 8 warps per instance: [numLevels + 1] into taskOrder8 ([rounds][8], -1 = idle)
  const int32_t* taskOrder8;
  const int32_t* levelOrderStart16; // ... for the 16-warp CTA of wide systems
  const int32_t* taskOrder16;
  const int32_t* tileInfo;    // [numTiles][3] {gi0, gj0, validI | validJ << 8 | diag << 16}, see ik_chol_sched.h
};
// view of the same schedule with every table pointer moved to a copy of the blob at `newBlob`
// Every pointer is re-derived FROM newBlob (newBlob + element offset): nvcc assumes kernel-parameter pointers address
// global memory, so "old pointer + delta" would still be loaded with ld.global even when the copy lives in shared memory.
MB2_HD CholSchedDev rebaseSchedule(const CholSchedDev& S, const int32_t* newBlob) {
  CholSchedDev R = S;
  const int32_t* o = S.blob;
  R.blob = newBlob;
#define MB2_RB(f) R.f = newBlob + (S.f - o);
  MB2_RB(perm) MB2_RB(pos) MB2_RB(tileIdTable) MB2_RB(tileRow) MB2_RB(tileCol) MB2_RB(diagTile) MB2_RB(levelColStart) MB2_RB(levelCols)
  MB2_RB(levelPanelStart) MB2_RB(panelTile) MB2_RB(panelDiag) MB2_RB(levelTaskStart) MB2_RB(taskDst) MB2_RB(taskPairStart) MB2_RB(pairA) MB2_RB(pairB)
  MB2_RB(levelVTaskStart) MB2_RB(vtaskRow) MB2_RB(vtaskSrcStart) MB2_RB(vsrcTile) MB2_RB(vsrcCol) MB2_RB(colPanelStart) MB2_RB(colPanelTile)
  MB2_RB(colPanelRow) MB2_RB(levelOrderStart8) MB2_RB(taskOrder8) MB2_RB(levelOrderStart16) MB2_RB(taskOrder16) MB2_RB(tileInfo)
#undef MB2_RB
  return R;
}

MB2_HD int tileIdx(int r, int c) { return r * 16 + ((((c >> 2) ^ ((r >> 1) & 3)) << 2) | (c & 3)); }
MB2_HD int tileGrp(int r, int g) { return r * 16 + ((g ^ ((r >> 1) & 3)) << 2); }

// Padding pass: after the 16x16 box of H has landed in tile storage (TMA on the device), rows >= validJ / columns >= validI
// belong to the NEXT parameters of the elimination order, not to this tile: overwrite them with the identity extension.
// One work item = storage row c, float4 group g (matrix rows 4g..4g+3).
MB2_HD void cholPadGroup(float* tile, int info, int c, int g) {
  const int vI = info & 0xFF, vJ = (info >> 8) & 0xFF, diag = (info >> 16) & 1;
  if (c < vJ && 4 * g + 3 < vI) return;
  float* dst = tile + tileGrp(c, g);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = 4 * g + q;
    if (c >= vJ || r >= vI) dst[q] = (diag && r == c) ? 1.f : 0.f;
  }
}

// ---- tile-sparse Gram (see GramPlan in ik_chol_sched.h): a warp owns one tile and accumulates it on the tensor cores ----
// strips: [strip][16 columns][4 rows] floats, so S[c][k] of a strip is float 4 c + k. Two pairs at a time form the
// 16 x 16 x 8 product out(r, c) += sum_k A[r][k] B[c][k] (k 0..3 from the first pair, 4..7 from the second; lists are padded
// to even length with an all-zero strip) = two mma.sync.m16n8k8 (column halves) per term of the three-term TF32 split
// hi*hi + hi*lo + lo*hi with fp32 accumulation (fp32-class accuracy). hi = nearest tf32 of x, lo = x - hi (exact; the tensor
// core reads its leading 10 mantissa bits: a 2^-22 relative perturbation). Every fragment is one conflict-free 128-byte warp read:
//   a0 = A[g][t], a1 = A[g + 8][t], b(h) = B[8 h + g][t]   with g = lane >> 2, t = lane & 3  ->  float index lane (+ 32).
// Accumulators d[h][0..3] follow the mma C layout: (row g, cols 8h + 2t, 8h + 2t + 1), (row g + 8, same cols).
MB2_HD float tf32High(float x) { // nearest tf32 (ties away from zero), 10 mantissa bits: what cvt.rna.tf32.f32 returns for finite x, but on
                                  // the integer ALU (the conversion instruction runs on the quarter-rate XU pipe)
#if defined(__CUDA_ARCH__)
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
#else
  uint32_t u;
  std::memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
#endif
}
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void mmaTf32K8(float d[4], float a0, float a1, float a2, float a3, float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)), "r"(__float_as_uint(a2)), "r"(__float_as_uint(a3)), "r"(__float_as_uint(b0)),
                 "r"(__float_as_uint(b1)));
}
#endif
// d: running sums of the leading term, added in fp32 registers (round to nearest) after every step: the tensor core's own
// accumulator truncates, which over a few hundred steps becomes a visible bias. small: the two correction terms, accumulated
// inside the tensor core (they are 2^-11 of the result: their truncation does not matter); added to d once at the end.
// One k = 8 step of out += A B^T on the tensor cores with the three-term TF32 split (see gramTilePairs for the accuracy notes):
// a0..a3 / b[h][0..1] are the raw fp32 fragment values of mma.m16n8k8 (A: rows g, g + 8 x k = t, t + 4; B: n = 8 h + g).
// d += hi*hi in fp32 registers (round to nearest), small += lo*hi + hi*lo inside the tensor core.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void mma3xTf32Step(float d[2][4], float small[2][4], float a0, float a1, float a2, float a3, float b00, float b01, float b10, float b11) {
  const float a0h = tf32High(a0), a1h = tf32High(a1), a2h = tf32High(a2), a3h = tf32High(a3);
  const float b00h = tf32High(b00), b01h = tf32High(b01), b10h = tf32High(b10), b11h = tf32High(b11);
  const float a0l = a0 - a0h, a1l = a1 - a1h, a2l = a2 - a2h, a3l = a3 - a3h; // (the tensor core reads the leading bits of lo)
  const float b00l = b00 - b00h, b01l = b01 - b01h, b10l = b10 - b10h, b11l = b11 - b11h;
  mmaTf32K8(small[0], a0l, a1l, a2l, a3l, b00h, b01h); mmaTf32K8(small[1], a0l, a1l, a2l, a3l, b10h, b11h);
  mmaTf32K8(small[0], a0h, a1h, a2h, a3h, b00l, b01l); mmaTf32K8(small[1], a0h, a1h, a2h, a3h, b10l, b11l);
  float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f};
  mmaTf32K8(t0, a0h, a1h, a2h, a3h, b00h, b01h); mmaTf32K8(t1, a0h, a1h, a2h, a3h, b10h, b11h);
#pragma unroll
  for (int e = 0; e < 4; ++e) { d[0][e] += t0[e]; d[1][e] += t1[e]; }
}
#endif
// host emulation of the same step for one output: sum over the step's eight k of the split products
MB2_HD void mma3xEmulate(float& d, float& small, const float av[8], const float bv[8]) {
  float lo = 0.f, mid = 0.f, hi = 0.f;
  for (int k = 0; k < 8; ++k) {
    const float ah = tf32High(av[k]), bh = tf32High(bv[k]), al = av[k] - ah, bl = bv[k] - bh;
    lo += al * bh; mid += ah * bl; hi += ah * bh;
  }
  small += lo + mid;
  d += hi;
}

MB2_HD void gramTilePairs(const float* strips, int oa0, int ob0, int oa1, int ob1, int lane, float d[2][4], float small[2][4]) {
  const float* A0 = strips + oa0; // float offsets of the four strips
  const float* B0 = strips + ob0;
  const float* A1 = strips + oa1;
  const float* B1 = strips + ob1;
#if defined(__CUDA_ARCH__)
  mma3xTf32Step(d, small, A0[lane], A0[32 + lane], A1[lane], A1[32 + lane], B0[lane], B1[lane], B0[32 + lane], B1[32 + lane]); // b[h][k half]
#else
  const int g = lane >> 2, t = lane & 3; // host emulation: the lane's eight outputs from the same three-term split
  for (int h = 0; h < 2; ++h)
    for (int e = 0; e < 4; ++e) {
      const int r = g + 8 * (e >> 1), c = 8 * h + 2 * t + (e & 1);
      float lo = 0.f, mid = 0.f, hi = 0.f;
      for (int k = 0; k < 8; ++k) {
        const float av = k < 4 ? A0[4 * r + k] : A1[4 * r + k - 4], bv = k < 4 ? B0[4 * c + k] : B1[4 * c + k - 4];
        const float ah = tf32High(av), bh = tf32High(bv), al = av - ah, bl = bv - bh;
        lo += al * bh; mid += ah * bl; hi += ah * bh;
      }
      small[h][e] += lo + mid;
      d[h][e] += hi;
    }
#endif
}
// quads: {A0, B0, A1, B1} float offsets per step (GramPlan::quad, pair lists padded to even length), staged in shared memory
MB2_HD void gramTileAccumulate(const float* strips, const int32_t* quads, int q0, int q1, int lane, float d[2][4]) {
  float small[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int q = q0; q < q1; ++q) {
#if defined(__CUDA_ARCH__)
    const int4 o = *reinterpret_cast<const int4*>(quads + 4 * q); // one broadcast read
    gramTilePairs(strips, o.x, o.y, o.z, o.w, lane, d, small);
#else
    gramTilePairs(strips, quads[4 * q], quads[4 * q + 1], quads[4 * q + 2], quads[4 * q + 3], lane, d, small);
#endif
  }
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int e = 0; e < 4; ++e) d[h][e] += small[h][e];
}
// float offsets of the lane's eight outputs in tile storage T[c][r] = H(r, c) (same for every tile: computed once per kernel)
MB2_HD void gramLaneOffsets(int lane, int off[8]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int e = 0; e < 4; ++e) off[4 * h + e] = tileIdx(8 * h + 2 * t + (e & 1), g + 8 * (e >> 1));
}
// writes the lane's accumulators with the identity extension on padded rows/columns and the damping added to the real diagonal
// (info = validI | validJ << 8 | diag << 16 as in tileInfo); full off-diagonal tiles take the plain path
MB2_HD void gramTileStore(float* tile, const float d[2][4], int info, float lambda, int lane, const int off[8]) {
  const int vI = info & 0xFF, vJ = (info >> 8) & 0xFF, diag = (info >> 16) & 1;
  if (vI == 16 && vJ == 16 && !diag) {
#pragma unroll
    for (int i = 0; i < 8; ++i) tile[off[i]] = d[i >> 2][i & 3];
    return;
  }
  const int g = lane >> 2, t = lane & 3;
  const bool rowOk[2] = {g < vI, g + 8 < vI}; // the lane owns two rows and four columns: six comparisons instead of sixteen
  bool colOk[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) colOk[i] = 8 * (i >> 1) + 2 * t + (i & 1) < vJ;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int e = 0; e < 4; ++e) tile[off[4 * h + e]] = (rowOk[e >> 1] && colOk[2 * h + (e & 1)]) ? d[h][e] : 0.f;
  if (diag) { // the lane's outputs on the diagonal: r == c <=> g (+8) == 8 h + 2 t (+1)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = g + 8 * (e >> 1), c = 8 * h + 2 * t + (e & 1);
        if (r == c) tile[off[4 * h + e]] = (r < vI) ? d[h][e] + lambda : 1.f;
      }
  }
}
// entry hl of block K of J^T r: sum over the strips of tile column K of strip[hl][0..3] . r[4q..4q+3]
MB2_HD float gramVectorEntry(const float* strips, const float* resid, const int32_t* colStrip, const int32_t* stripRow, int s0, int s1, int hl) {
  float g0 = 0.f, g1 = 0.f;
  for (int k = s0; k < s1; ++k) {
    const int sidx = colStrip[k];
    const float4 sv = *reinterpret_cast<const float4*>(strips + size_t(sidx) * 64 + 4 * hl);
    const float4 rv = *reinterpret_cast<const float4*>(resid + stripRow[sidx]);
    g0 += sv.x * rv.x + sv.y * rv.y;
    g1 += sv.z * rv.z + sv.w * rv.w;
  }
  return g0 + g1;
}

MB2_HD void tileLoadRow(const float* tile, int r, float* a) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 v = *reinterpret_cast<const float4*>(tile + tileGrp(r, g));
    a[4 * g] = v.x; a[4 * g + 1] = v.y; a[4 * g + 2] = v.z; a[4 * g + 3] = v.w;
  }
}
MB2_HD void tileStoreRow(float* tile, int r, const float* a) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float4 v;
    v.x = a[4 * g]; v.y = a[4 * g + 1]; v.z = a[4 * g + 2]; v.w = a[4 * g + 3];
    *reinterpret_cast<float4*>(tile + tileGrp(r, g)) = v;
  }
}

// ---- phase A: Cholesky of a diagonal tile, forward solve of its 16 right-hand-side entries, and W = L^-1 ----
// hl = lane within the half-warp (0..15), hmask = shuffle mask of the half-warp.
// A non-positive pivot is replaced by `fallback` (the damping) and reported through *fail.
// On return the tile holds W = L(K,K)^-1 (lower triangular, zeros above the diagonal): every later use of the diagonal
// block (panel solve, backward substitution) is then a 16x16 product with W — no dependent 16-step chain on the
// critical path of a level. L(K,K) itself is not needed again.
MB2_HD void cholDiagTile(float* tile, float* y16, int hl, unsigned hmask, float fallback, int* fail) {
  // Square-root-free elimination keeps the per-step dependency chain short (pivot broadcast -> reciprocal -> one multiply ->
  // one fused multiply-add); the 1/sqrt(pivot) scalings and the right-hand side ride along off that chain:
  //   A = U D U^T (U unit lower, U[r][k] = a_r[k] / d_k),  L = U D^1/2,  y = L^-1 g = D^-1/2 U^-1 g.
  // W = L^-1 = D^-1/2 U^-1 comes out of the SAME sixteen steps (Gauss-Jordan): the row operations that eliminate A are applied to an
  // identity matrix E as well, row_r(E) -= U[r][k] row_k(E) for r > k, which leaves E = U^-1. One dependent chain of sixteen steps per
  // diagonal tile instead of two (factorise, then invert by substitution), and no shared-memory traffic inside it.
  // The gather fills the tile as S[c][r] = H(r, c), a symmetric tile: matrix row hl is storage column hl.
#if defined(__CUDA_ARCH__)
  float a[16], e[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { a[k] = tile[tileIdx(k, hl)]; e[k] = (k == hl) ? 1.f : 0.f; }
  float z = y16[hl], rdSelf = 0.f;
  __syncwarp(hmask); // every lane has read its column before any lane overwrites the tile
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    float piv = __shfl_sync(hmask, a[k], k, 16);
    if (!(piv > 0.f)) { piv = fallback; if (hl == k) *fail = 1; }
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(piv)); // one MUFU on the dependency chain (1 ulp; the step is damped Gauss-Newton)
    const float rdk = rsqrtf(piv);
    if (hl == k) rdSelf = rdk;
    const float zk = __shfl_sync(hmask, z, k, 16);
    const float t = a[k] * inv;           // U[hl][k] for hl > k
    const float tm = hl > k ? t : 0.f;    // rows <= k of E (and of the right-hand side) are finished
#pragma unroll
    for (int j = k + 1; j < 16; ++j) a[j] -= t * __shfl_sync(hmask, a[k], j, 16);
#pragma unroll
    for (int j = 0; j < k; ++j) e[j] -= tm * __shfl_sync(hmask, e[j], k, 16);
    e[k] -= tm;                           // row k of E has a one on the diagonal
    z -= tm * zk;
  }
  y16[hl] = z * rdSelf;
  // row hl of W = D^-1/2 U^-1 (zero above the diagonal, 1/sqrt(d) on it), written as one swizzled row: conflict-free float4 stores
#pragma unroll
  for (int j = 0; j < 16; ++j) e[j] *= rdSelf;
  tileStoreRow(tile, hl, e);
#else
  if (hl != 0) return; // host emulation: one caller plays the sixteen lanes in lock step with the same operation order
  (void)hmask;
  float A[16][16], E[16][16], z[16], rd[16]; // A[lane][k], E[lane][j]
  for (int l = 0; l < 16; ++l) { for (int k = 0; k < 16; ++k) { A[l][k] = tile[tileIdx(k, l)]; E[l][k] = (k == l) ? 1.f : 0.f; } z[l] = y16[l]; }
  for (int k = 0; k < 16; ++k) {
    float piv = A[k][k];
    if (!(piv > 0.f)) { piv = fallback; *fail = 1; }
    const float inv = 1.f / piv;
    rd[k] = 1.f / sqrtf(piv);
    const float zk = z[k];
    float colk[16], ek[16];
    for (int l = 0; l < 16; ++l) { colk[l] = A[l][k]; ek[l] = E[k][l]; }
    for (int l = 0; l < 16; ++l) {
      const float t = colk[l] * inv;
      const float tm = l > k ? t : 0.f;
      for (int j = k + 1; j < 16; ++j) A[l][j] -= t * colk[j];
      for (int j = 0; j < k; ++j) E[l][j] -= tm * ek[j];
      E[l][k] -= tm;
      z[l] -= tm * zk;
    }
  }
  for (int l = 0; l < 16; ++l) {
    y16[l] = z[l] * rd[l];
    float w[16];
    for (int j = 0; j < 16; ++j) w[j] = E[l][j] * rd[l];
    tileStoreRow(tile, l, w);
  }
#endif
}

// ---- 16x16x16 tile products on the tensor cores (mma.sync m16n8k8, three-term TF32 split, fp32 accumulate) ----
// out(r, c) += sum_k A(r, k) Bop(c, k); A is a panel tile in transposed storage TA[k][r]. B is either another panel tile
// (kMajorB: TB[k][c]) or a row-major tile (W[c][k]). The k index of the fragments is permuted ({0,1,4,5} / {2,3,6,7} per step of 8)
// so that the XOR-swizzled rows give conflict-free (k-major) or 2-way (row-major) shared-memory reads; any permutation is valid
// as long as A and B use the same one. Lane's outputs follow the mma C layout: d[h][e] = out(g + 8 (e >> 1), 8 h + 2 t + (e & 1)).
template <bool kMajorB>
MB2_HD void tileProduct(const float* TA, const float* TB, int lane, float d[2][4], float small[2][4]) {
  const int g = lane >> 2, t = lane & 3;
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int k0 = 8 * ks + ((t & 1) | ((t & 2) << 1)), k1 = k0 + 2;
    const float a0 = TA[tileIdx(k0, g)], a1 = TA[tileIdx(k0, g + 8)], a2 = TA[tileIdx(k1, g)], a3 = TA[tileIdx(k1, g + 8)];
    float b00, b01, b10, b11;
    if (kMajorB) { b00 = TB[tileIdx(k0, g)]; b01 = TB[tileIdx(k1, g)]; b10 = TB[tileIdx(k0, g + 8)]; b11 = TB[tileIdx(k1, g + 8)]; }
    else { b00 = TB[tileIdx(g, k0)]; b01 = TB[tileIdx(g, k1)]; b10 = TB[tileIdx(g + 8, k0)]; b11 = TB[tileIdx(g + 8, k1)]; }
    mma3xTf32Step(d, small, a0, a1, a2, a3, b00, b01, b10, b11);
  }
#else
  for (int h = 0; h < 2; ++h)
    for (int e = 0; e < 4; ++e) {
      const int r = g + 8 * (e >> 1), c = 8 * h + 2 * t + (e & 1);
      for (int ks = 0; ks < 2; ++ks) {
        float av[8], bv[8];
        for (int k = 0; k < 8; ++k) { av[k] = TA[tileIdx(8 * ks + k, r)]; bv[k] = kMajorB ? TB[tileIdx(8 * ks + k, c)] : TB[tileIdx(c, 8 * ks + k)]; }
        mma3xEmulate(d[h][e], small[h][e], av, bv);
      }
    }
#endif
}

// ---- phase B: X = A(I,K) L(K,K)^-T = A W^T for one panel tile, a warp per tile; the result replaces A (transposed storage) ----
// split in two so that every lane has read the tile before any lane overwrites it (device: __syncwarp in between)
MB2_HD void cholPanelProduct(const float* tile, const float* diagW, int lane, float out[2][4]) {
  float small[2][4];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int e = 0; e < 4; ++e) { out[h][e] = 0.f; small[h][e] = 0.f; }
  tileProduct<false>(tile, diagW, lane, out, small);
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int e = 0; e < 4; ++e) out[h][e] += small[h][e];
}
MB2_HD void cholPanelStore(float* tile, int lane, const float out[2][4]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int e = 0; e < 4; ++e) tile[tileIdx(8 * h + 2 * t + (e & 1), g + 8 * (e >> 1))] = out[h][e]; // T[c][r] = X(r, c)
}

// ---- phase C: one update task, D(I,J) -= sum_pairs L(I,K) L(J,K)^T; a warp per destination tile ----
MB2_HD void cholUpdateTask(float* tiles, const CholSchedDev& S, int task, int lane) {
  float d[2][4], small[2][4];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int e = 0; e < 4; ++e) { d[h][e] = 0.f; small[h][e] = 0.f; }
  for (int p = S.taskPairStart[task]; p < S.taskPairStart[task + 1]; ++p)
    tileProduct<true>(tiles + size_t(S.pairA[p]) * 256, tiles + size_t(S.pairB[p]) * 256, lane, d, small);
  float* D = tiles + size_t(S.taskDst[task]) * 256;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int e = 0; e < 4; ++e) D[tileIdx(g + 8 * (e >> 1), 8 * h + 2 * t + (e & 1))] -= d[h][e] + small[h][e]; // storage row = index of the A operand (the schedule orders each pair accordingly)
}

// ---- phase C (vector part): y_I -= sum L(I,K) y_K over this level's columns; lane hl = row ----
MB2_HD void cholVectorTask(const float* tiles, float* y, const CholSchedDev& S, int vtask, int hl) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f; // four partial sums: the dependent chain is 4 deep instead of 16 per source tile
  for (int p = S.vtaskSrcStart[vtask]; p < S.vtaskSrcStart[vtask + 1]; ++p) {
    const float* T = tiles + size_t(S.vsrcTile[p]) * 256;
    const float* yk = y + S.vsrcCol[p] * 16;
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      s0 += T[tileIdx(c, hl)] * yk[c];
      s1 += T[tileIdx(c + 1, hl)] * yk[c + 1];
      s2 += T[tileIdx(c + 2, hl)] * yk[c + 2];
      s3 += T[tileIdx(c + 3, hl)] * yk[c + 3];
    }
  }
  y[S.vtaskRow[vtask] * 16 + hl] -= (s0 + s1) + (s2 + s3);
}

// ---- backward substitution for one tile column K: y_K <- L(K,K)^-T (y_K - sum_I L(I,K)^T y_I) ----
MB2_HD void cholBackwardColumn(const float* tiles, float* y, const CholSchedDev& S, int K, int hl, unsigned hmask) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int p = S.colPanelStart[K]; p < S.colPanelStart[K + 1]; ++p) {
    float t[16];
    tileLoadRow(tiles + size_t(S.colPanelTile[p]) * 256, hl, t); // row c = hl of the transposed tile: L[r][c], r = 0..15
    const float* yi = y + S.colPanelRow[p] * 16;
#pragma unroll
    for (int r = 0; r < 16; r += 4) { s0 += t[r] * yi[r]; s1 += t[r + 1] * yi[r + 1]; s2 += t[r + 2] * yi[r + 2]; s3 += t[r + 3] * yi[r + 3]; }
  }
  const float s = y[K * 16 + hl] - ((s0 + s1) + (s2 + s3));
  // x_K = W^T s with W = L(K,K)^-1 stored by phase A: stage s, then lane c accumulates column c of W
  const float* W = tiles + size_t(S.diagTile[K]) * 256;
#if defined(__CUDA_ARCH__)
  float x0 = 0.f, x1 = 0.f;
#pragma unroll
  for (int r = 0; r < 16; r += 2) {
    x0 += W[tileIdx(r, hl)] * __shfl_sync(hmask, s, r, 16);
    x1 += W[tileIdx(r + 1, hl)] * __shfl_sync(hmask, s, r + 1, 16);
  }
  y[K * 16 + hl] = x0 + x1;
#else
  // host emulation: lanes run one after the other, so stage the sums, then lane 15 (last) finishes the block
  y[K * 16 + hl] = s;
  if (hl != 15) return;
  float sv[16];
  for (int r = 0; r < 16; ++r) sv[r] = y[K * 16 + r];
  for (int c = 0; c < 16; ++c) {
    float x0 = 0.f, x1 = 0.f;
    for (int r = 0; r < 16; r += 2) { x0 += W[tileIdx(r, c)] * sv[r]; x1 += W[tileIdx(r + 1, c)] * sv[r + 1]; }
    y[K * 16 + c] = x0 + x1;
  }
#endif
}

} // namespace mb2
