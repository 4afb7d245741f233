// Device helpers shared by the persistent Stiefel tCG kernels (fp64 tensor-core
// building blocks, exact Gram accumulation, named barriers).
#pragma once
#include "tcg.cuh"

namespace ob200 {

constexpr int ST_P = 32;
constexpr int ST_NB = 128;
constexpr int PS = 33;   // row stride of the staged p block (grouped k-map, conflict free)
constexpr int WS = 36;   // row stride of staged W / Y / S / G (natural k-map, conflict free)

constexpr size_t SM_P = 0;
constexpr size_t SM_W = SM_P + sizeof(double) * ST_NB * PS;
constexpr size_t SM_Y = SM_W + sizeof(double) * ST_NB * WS;
constexpr size_t SM_S = SM_Y + sizeof(double) * ST_NB * WS;
constexpr size_t SM_G = SM_S + sizeof(double) * ST_P * WS;
constexpr size_t SM_ACC = SM_G + sizeof(double) * ST_P * WS;
constexpr size_t SM_TOTAL = SM_ACC + sizeof(u64) * ACC_NSCAL * KUL_STRIDE;

struct StiefelArgs {
  unsigned long long n_rows;   // local rows
  const unsigned short *A;     // bf16 blocks
  const double *Y;
  const double *S;             // p x p (device), sym(Y^T A Y)
  double op_norm_bound;
};

// --- step 2: W strip (8 rows x 32 cols) = A_strip * Pblk + Pstrip * Sneg -----
// acc[t][c] = W[8w + lane/4][8t + 2*(lane%4) + c]
__device__ __forceinline__ void strip_apply(const unsigned short *Ablock, const double *Psm,
                                            const double *Sneg, int warp, int lane,
                                            double (&acc)[4][2]) {
  const int m = lane >> 2, j = lane & 3;
  const unsigned short *Arow = Ablock + (size_t)(8 * warp + m) * ST_NB;
  uint2 av[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) av[u] = __ldg(reinterpret_cast<const uint2 *>(Arow + 16 * u + 4 * j));
#pragma unroll
  for (int t = 0; t < 4; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const unsigned short h[4] = {(unsigned short)(av[u].x & 0xffffu), (unsigned short)(av[u].x >> 16),
                                 (unsigned short)(av[u].y & 0xffffu), (unsigned short)(av[u].y >> 16)};
#pragma unroll
    for (int w4 = 0; w4 < 4; ++w4) {
      const double a = bf16_bits_to_double(h[w4]);
      const double *Prow = Psm + (16 * u + 4 * j + w4) * PS + m;
#pragma unroll
      for (int t = 0; t < 4; ++t) dmma884(acc[t][0], acc[t][1], a, Prow[8 * t]);
    }
  }
  if (Sneg) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double a = Psm[(8 * warp + m) * PS + 4 * q + j];
      const double *Srow = Sneg + (4 * q + j) * WS + m;
#pragma unroll
      for (int t = 0; t < 4; ++t) dmma884(acc[t][0], acc[t][1], a, Srow[8 * t]);
    }
  }
}

// --- step 3: 8x8 tile (mt, nt) of X^T Z over one staged 128-row block ----------
__device__ __forceinline__ void gram_tile(const double *Xsm, const double *Zsm, int tile, int lane,
                                          double &g0, double &g1) {
  const int mt = tile >> 2, nt = tile & 3, m = lane >> 2, j = lane & 3;
  g0 = 0.0; g1 = 0.0;
#pragma unroll 8
  for (int q = 0; q < 32; ++q) {
    const int krow = 4 * q + j;
    dmma884(g0, g1, Xsm[krow * WS + 8 * mt + m], Zsm[krow * WS + 8 * nt + m]);
  }
}

// same over 64 rows (one half block)
__device__ __forceinline__ void gram_tile_half(const double *Xsm, const double *Zsm, int tile, int lane,
                                               double &g0, double &g1) {
  const int mt = tile >> 2, nt = tile & 3, m = lane >> 2, j = lane & 3;
  g0 = 0.0; g1 = 0.0;
#pragma unroll 8
  for (int q = 0; q < 16; ++q) {
    const int krow = 4 * q + j;
    dmma884(g0, g1, Xsm[krow * WS + 8 * mt + m], Zsm[krow * WS + 8 * nt + m]);
  }
}

// Both tiles (mt, nt0), (mt, nt0 + 1) of one 64-row half with four independent partial accumulators
// per tile over k (the fp64 MMA has a long dependent-issue latency: 8 chains of 4 steps instead of
// 2 chains of 16); the partials are added in a fixed order.
__device__ __forceinline__ void gram_pair_half_split(const double *Xsm, const double *Zsm, int mt, int nt0, int lane,
                                                     double &g00, double &g01, double &g10, double &g11) {
  const int m = lane >> 2, j = lane & 3;
  double a0[4][2], a1[4][2];
#pragma unroll
  for (int s = 0; s < 4; ++s) { a0[s][0] = a0[s][1] = a1[s][0] = a1[s][1] = 0.0; }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int krow = 4 * (4 * s + q) + j;
      const double x = Xsm[krow * WS + 8 * mt + m];
      dmma884(a0[s][0], a0[s][1], x, Zsm[krow * WS + 8 * nt0 + m]);
      dmma884(a1[s][0], a1[s][1], x, Zsm[krow * WS + 8 * (nt0 + 1) + m]);
    }
  }
  g00 = (a0[0][0] + a0[1][0]) + (a0[2][0] + a0[3][0]);
  g01 = (a0[0][1] + a0[1][1]) + (a0[2][1] + a0[3][1]);
  g10 = (a1[0][0] + a1[1][0]) + (a1[2][0] + a1[3][0]);
  g11 = (a1[0][1] + a1[1][1]) + (a1[2][1] + a1[3][1]);
}

// out strip += X_strip * M with vectorised operand loads: lane (m, j) reads its 8 CONTIGUOUS elements
// X[row m][8j .. 8j+7] (four 16-byte loads) and k-step q contracts over the column set {8j + q}; the
// B operand is indexed with the same permutation, so no shuffles are needed.  M is staged with row
// stride GS (odd: the four j groups fall on disjoint bank halves).  Two k-halves run as independent
// accumulator chains (the fp64 MMA has a long dependent-issue latency) and are added at the end.
constexpr int GS = 33;
__device__ __forceinline__ void strip_rightmul_load(const double *Xrow /* row of this lane or null */, int lane,
                                                    double2 (&x)[4]) {
  const int j = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = Xrow ? ldcg2(Xrow + 8 * j + 2 * i) : make_double2(0.0, 0.0);
}
__device__ __forceinline__ void strip_rightmul_v(const double2 (&x)[4], const double *Msm, int lane,
                                                 double (&acc)[4][2]) {
  const int m = lane >> 2, j = lane & 3;
  double acc2[4][2];
#pragma unroll
  for (int t = 0; t < 4; ++t) acc2[t][0] = acc2[t][1] = 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double a_lo = (q & 1) ? x[q >> 1].y : x[q >> 1].x;              // column 8j + q
    const double a_hi = (q & 1) ? x[2 + (q >> 1)].y : x[2 + (q >> 1)].x;  // column 8j + 4 + q
    const double *Mlo = Msm + (8 * j + q) * GS + m;
    const double *Mhi = Msm + (8 * j + 4 + q) * GS + m;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      dmma884(acc[t][0], acc[t][1], a_lo, Mlo[8 * t]);
      dmma884(acc2[t][0], acc2[t][1], a_hi, Mhi[8 * t]);
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) { acc[t][0] += acc2[t][0]; acc[t][1] += acc2[t][1]; }
}

// out strip = C + X_strip * M  (M staged with stride WS); X read from global
__device__ __forceinline__ void strip_rightmul(const double *Xrow /* row of this lane or null */,
                                               const double *Msm, int lane, double (&acc)[4][2]) {
  const int m = lane >> 2, j = lane & 3;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const double a = Xrow ? __ldcg(Xrow + 4 * q + j) : 0.0;
    const double *Mrow = Msm + (4 * q + j) * WS + m;
#pragma unroll
    for (int t = 0; t < 4; ++t) dmma884(acc[t][0], acc[t][1], a, Mrow[8 * t]);
  }
}

__device__ __forceinline__ void gram_accumulate(double g0, double g1, double inv_q, i64 (&gfix)[4],
                                                unsigned *ovf) {
  const Fix2 f0 = fix2_from_double(g0, inv_q, ovf);
  const Fix2 f1 = fix2_from_double(g1, inv_q, ovf);
  gfix[0] += f0.hi; gfix[1] += f0.lo; gfix[2] += f1.hi; gfix[3] += f1.lo;
}

__device__ __forceinline__ void gram_flush(u64 *set, int tile, int lane, i64 (&gfix)[4], unsigned ovf) {
  const int mt = tile >> 2, nt = tile & 3, m = lane >> 2, j = lane & 3;
  const int e = (8 * mt + m) * ST_P + 8 * nt + 2 * j;
  u64 *g = set + ACC_GRAM_OFF + 2 * e;
  if (gfix[0]) atomicAdd(g + 0, (u64)gfix[0]);
  if (gfix[1]) atomicAdd(g + 1, (u64)gfix[1]);
  if (gfix[2]) atomicAdd(g + 2, (u64)gfix[2]);
  if (gfix[3]) atomicAdd(g + 3, (u64)gfix[3]);
  gfix[0] = gfix[1] = gfix[2] = gfix[3] = 0;
  if (ovf) atomicOr((unsigned long long *)(set + ACC_FLAG_OFF), 1ull);
}

// quantum for the fixed-point Gram: |entry| <= bound  =>  e = ilogb(bound) + 2
__host__ __device__ __forceinline__ int gram_exponent(double bound) {
  if (!(bound > 0.0) || !(bound < 1.0e300)) return 0;
  return ilogb(bound) + 2;
}

enum { NB_FULL = 1, NB_EMPTY = 3, NB_MSYNC = 5, NB_MSYNC2 = 6, NB_MSYNC3 = 7, NB_LSYNC = 8 };
__device__ __forceinline__ void nbar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void nbar_arrive(int id, int n) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}


}  // namespace ob200
