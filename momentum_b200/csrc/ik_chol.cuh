// Blocked LLT building blocks for K3, structured like Eigen's llt_inplace<Lower> (which is what
// GaussNewtonSolverT::llt_ runs, gauss_newton_solver.cpp:251): left-looking unblocked factorisation
// of each diagonal block, right-sided triangular solve of the panel below it, rank-bs update of the
// trailing matrix. Block size follows Eigen's rule (n < 32: one unblocked block; else
// clamp((n/8)/16*16, 8, 128)), capped at 32 here. A non-positive pivot stops the factorisation
// exactly where Eigen's would (info() = NumericalIssue is ignored by the reference solver, so the
// partially factored matrix is still used by the two triangular solves).
//
// The matrix is (n+1) x lda row-major: rows 0..n-1 hold the lower triangle of JtJ + lambda I, row n
// holds Jtr; the panel solve/trailing update treat row n like any other row, which turns it into
// y = L^-1 Jtr without a separate forward substitution.
//
// Functions are __host__ __device__ and take (tid, nthreads) so tests/emu can run them lane by lane.
#pragma once

#include <cmath>

#include "ik_types.h"

#include <vector_types.h> // float2 / float4 (CUDA toolkit header, usable from plain C++ too)

namespace mb2 {

constexpr int kCholThreads = 256;

struct CholCtx {
  float* A;
  int lda;
  int n;
  float* P;  // [NB][ldp] transposed panel (P[c][i] = L[k+bs+i][k+c])
  int ldp;
  int* fail; // [0]: 0 ok, else failing column + 1
};

MB2_HD int cholBlockSize(int n, int maxNb) {
  if (n < 32) return n < 1 ? 1 : n; // Eigen: size < 32 -> unblocked
  int bs = n / 8;
  bs = (bs / 16) * 16;
  bs = bs < 8 ? 8 : (bs > 128 ? 128 : bs);
  return bs < maxNb ? bs : maxNb;
}
MB2_HD int cholCompletedColumns(int failColumn, int blockSize) { return (failColumn / blockSize) * blockSize; }

// x = A(kk,kk) - ||A(kk, k..kk-1)||^2, kk = k + jj   (llt_inplace::unblocked)
MB2_HD float cholDiagPivot(const CholCtx& c, int k, int jj) {
  const float* row = c.A + size_t(k + jj) * c.lda;
  float x = row[k + jj];
  for (int j = 0; j < jj; ++j) x -= row[k + j] * row[k + j];
  return x;
}
// column kk of the diagonal block: A21 = (A21 - A20 * A10^T) / sqrt(x); lanes own rows
MB2_HD void cholDiagColumn(const CholCtx& c, int k, int bs, int jj, float x, int lane) {
  const float d = sqrtf(x);
  const float* rk = c.A + size_t(k + jj) * c.lda;
  for (int i = jj + 1 + lane; i < bs; i += 32) {
    float* ri = c.A + size_t(k + i) * c.lda;
    float s = ri[k + jj];
    for (int j = 0; j < jj; ++j) s -= ri[k + j] * rk[k + j];
    ri[k + jj] = s / d;
  }
  if (lane == 0) c.A[size_t(k + jj) * c.lda + k + jj] = d;
}

// A21 <- A21 * L11^-T for rows k+bs .. n (inclusive: the appended Jtr row); also scatter into P
template <int NB>
MB2_HD void cholPanelSolve(const CholCtx& c, int k, int bs, int tid, int nthreads) {
  for (int i = k + bs + tid; i <= c.n; i += nthreads) {
    float* ri = c.A + size_t(i) * c.lda + k;
    float x[NB];
#pragma unroll
    for (int cc = 0; cc < NB; ++cc) {
      if (cc < bs) {
        const float* rl = c.A + size_t(k + cc) * c.lda + k;
        float s = ri[cc];
#pragma unroll
        for (int j = 0; j < NB; ++j)
          if (j < cc) s -= x[j] * rl[j];
        x[cc] = s / rl[cc];
        ri[cc] = x[cc];
        c.P[cc * c.ldp + (i - k - bs)] = x[cc];
      }
    }
  }
}

// A22 <- A22 - A21 A21^T on the lower triangle (rows up to n inclusive, columns up to n-1)
template <int NB>
MB2_HD void cholTrailingUpdate(const CholCtx& c, int k, int bs, int tid, int nthreads) {
  const int base = k + bs;
  const int R = c.n + 1 - base; // rows incl. appended row
  const int C = c.n - base;     // columns
  if (C <= 0) return;
  const int TI = (R + 3) >> 2;
  const int total = TI * (TI + 1) / 2;
  for (int t = tid; t < total; t += nthreads) {
    int ti = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while (ti * (ti + 1) / 2 > t) --ti;
    const int tj = t - ti * (ti + 1) / 2;
    if (4 * tj >= C) continue;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
    for (int cc = 0; cc < bs; ++cc) {
      const float4 av = *reinterpret_cast<const float4*>(c.P + cc * c.ldp + 4 * ti);
      const float4 bv = *reinterpret_cast<const float4*>(c.P + cc * c.ldp + 4 * tj);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] += a4[r] * b4[q];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = base + 4 * ti + r;
      if (i > c.n) continue;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = base + 4 * tj + q;
        if (j <= i && j < c.n) c.A[size_t(i) * c.lda + j] -= acc[r][q];
      }
    }
  }
}

// serial forward substitution from column k0 on (failure path only)
MB2_HD void cholForwardFrom(const CholCtx& c, int k0, float* y) {
  for (int col = k0; col < c.n; ++col) {
    const float* row = c.A + size_t(col) * c.lda;
    float s = y[col];
    for (int j = k0; j < col; ++j) s -= row[j] * y[j];
    y[col] = s / row[col];
  }
}

} // namespace mb2
