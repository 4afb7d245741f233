// Block kernels of the LOBPCG path (reference include/Optimization/LinearAlgebra/LOBPCG.h:131-337, BASELINE config C4):
// block operator apply (diagonal, 7-point 3-D Laplacian), tall-skinny Gram S^T Z, block update S C, residual
// R = AX - BX diag(theta) with column norms, and the small glue kernels around the Rayleigh-Ritz pencil.
// Block vectors are row-major m x k with a leading dimension (k contiguous doubles per row): every kernel streams
// whole rows, coalesced.  All cross-CTA sums go through per-CTA partial buffers reduced in a fixed order
// (deterministic: the grid is fixed by the SM count, the row partition is static).
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace ob200 {

constexpr int LB_THREADS = 256;
constexpr int LB_KMAX = 192;      // 3 * nx, nx <= 64

// ---- operator apply --------------------------------------------------------------------------------
// out[r, c] = d[r] * in[r, c]   (d == nullptr: scalar `alpha`)
__global__ void __launch_bounds__(LB_THREADS) blk_diag_kernel(unsigned long long m, int k, const double *d, double alpha,
                                                              const double *in, int ldi, double *out, int ldo) {
  const unsigned long long total = m * (unsigned long long)k;
  for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < total;
       e += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long r = e / k;
    const int c = (int)(e - r * k);
    out[r * ldo + c] = (d ? d[r] : alpha) * in[r * ldi + c];
  }
}
// 7-point Laplacian, Dirichlet boundary, grid gx x gy x gz (x fastest): out = 6 in - sum of existing neighbours
// has_lo / has_hi: the block vector is one z-slab of a row-sharded grid and carries a ghost plane below / above its
// first / last plane (filled from the neighbouring rank before the call); otherwise the slab ends at the Dirichlet boundary.
__global__ void __launch_bounds__(LB_THREADS) blk_stencil7_kernel(unsigned gx, unsigned gy, unsigned gz, int k, const double *in,
                                                                  int ldi, double *out, int ldo, int has_lo, int has_hi) {
  const unsigned long long m = (unsigned long long)gx * gy * gz, total = m * (unsigned long long)k;
  const unsigned long long sx = 1, sy = gx, sz = (unsigned long long)gx * gy;
  for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < total;
       e += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long r = e / k;
    const int c = (int)(e - r * k);
    const unsigned x = (unsigned)(r % gx), y = (unsigned)((r / gx) % gy), z = (unsigned)(r / sz);
    const double *p = in + r * ldi + c;
    double v = 6.0 * p[0];
    if (x > 0) v -= p[-(long long)(sx * ldi)];
    if (x + 1 < gx) v -= p[sx * ldi];
    if (y > 0) v -= p[-(long long)(sy * ldi)];
    if (y + 1 < gy) v -= p[sy * ldi];
    if (z > 0 || has_lo) v -= p[-(long long)(sz * ldi)];
    if (z + 1 < gz || has_hi) v -= p[sz * ldi];
    out[r * ldo + c] = v;
  }
}

// ---- Gram on the fp64 tensor cores: partial[bx][i][j] = sum over the CTA's rows of A[r, i] * B[r, 64 by + j] -------
// (i < k1 <= 192, j < min(64, k2 - 64 by)).  The Grams of LOBPCG are symmetric (S^T A S with A symmetric, S^T B S), and the
// Rayleigh-Ritz step only uses one triangle (like the reference's self-adjoint eigensolver): only the block-upper
// triangle is formed, i.e. column block `by` gets the rows i < 64 (by + 1) -- 6 of 9 blocks at ns = 192.
// mma.sync.m8n8k4.f64: the "A" operand is the transposed S tile (A[i][r] = S[r][i]), the "B" operand the Z tile; the
// row tiles of the block are dealt round-robin to the 8 warps (up to 3 each), every warp takes all eight column tiles
// (up to 24 accumulator tiles = 24 independent MMA chains, enough to cover the long dependent-issue latency of the
// fp64 MMA).  Shared-memory row strides are 4 (mod 16) doubles: conflict-free fragments.
__device__ __forceinline__ void lb_dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
constexpr int GR_TR = 32;          // rows per shared-memory tile (8 k-steps of 4 rows)
constexpr int GR_LDA = 196;        // 192 + 4
constexpr int GR_LDB = 68;         // 64 + 4
constexpr int GR_THREADS = 512;    // 16 warps: the (up to 24) row tiles of the block are dealt round-robin, two per warp
constexpr int GR_STAGE = GR_TR * (GR_LDA + GR_LDB);   // doubles per pipeline stage (66 KB)
// 16-byte asynchronous copy global -> shared, zero-filled when `ok` is false (src-size 0)
__device__ __forceinline__ void gr_cp16(double *dst, const double *src, bool ok) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
// Two-stage cp.async pipeline over 32-row tiles: the tile after the one being multiplied streams in meanwhile (the first
// version loaded a tile, synchronised, multiplied, synchronised: ncu showed the fp64 MMA pipe 33 % active and the warps
// waiting on the tile loads -- long scoreboard -- most of the time).  One CTA of 16 warps per SM.
__global__ void __launch_bounds__(GR_THREADS, 1) blk_gram_kernel(unsigned long long m, const double *A, int lda, int k1,
                                                                 const double *B, int ldb, int k2, double *partial) {
  extern __shared__ double gsm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;   // fragment coordinates: (row / col index, k index)
  const int j0 = 64 * blockIdx.y;
  const int kb = min(64, k2 - j0);
  const int it_n = min((k1 + 7) >> 3, 8 * ((int)blockIdx.y + 1));   // row tiles of the block-upper triangle
  const int ka = min(8 * it_n, LB_KMAX);                              // columns of A this block needs (a multiple of 8)
  const unsigned long long r_lo = m * blockIdx.x / gridDim.x, r_hi = m * (blockIdx.x + 1ull) / gridDim.x;
  // rows 16-byte aligned and of even length: asynchronous 16-byte copies (the bases of LOBPCG: ld = 3 nx or nx even);
  // otherwise scalar loads, same pipeline structure without the overlap
  const bool vec = !(lda & 1) && !(ldb & 1) && !((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) &&
                   !(k1 & 1) && !(kb & 1);
  double acc[2][8][2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int v = 0; v < 8; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
  auto fetch = [&](unsigned long long r0, int stage) {
    double *As = gsm + (size_t)stage * GR_STAGE, *Bs = As + GR_TR * GR_LDA;
    const int rows = (int)min((unsigned long long)GR_TR, r_hi - r0);
    if (vec) {
      const int ka2 = ka >> 1;                              // 16-byte chunks per row of the A tile
      for (int e = tid; e < GR_TR * ka2; e += GR_THREADS) {
        const int rr = e / ka2, c2 = e - rr * ka2;
        const bool ok = rr < rows && 2 * c2 < k1;
        gr_cp16(As + rr * GR_LDA + 2 * c2, ok ? A + (r0 + rr) * lda + 2 * c2 : A, ok);
      }
      for (int e = tid; e < GR_TR * 32; e += GR_THREADS) {
        const int rr = e >> 5, c2 = e & 31;
        const bool ok = rr < rows && 2 * c2 < kb;
        gr_cp16(Bs + rr * GR_LDB + 2 * c2, ok ? B + (r0 + rr) * ldb + j0 + 2 * c2 : B, ok);
      }
    } else {
      for (int e = tid; e < GR_TR * ka; e += GR_THREADS) {
        const int rr = e / ka, c = e - rr * ka;
        As[rr * GR_LDA + c] = (rr < rows && c < k1) ? A[(r0 + rr) * lda + c] : 0.0;
      }
      for (int e = tid; e < GR_TR * 64; e += GR_THREADS) {
        const int rr = e >> 6, c = e & 63;
        Bs[rr * GR_LDB + c] = (rr < rows && c < kb) ? B[(r0 + rr) * ldb + j0 + c] : 0.0;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int stage = 0;
  if (r_lo < r_hi) fetch(r_lo, 0);
  for (unsigned long long r0 = r_lo; r0 < r_hi; r0 += GR_TR) {
    const bool more = r0 + GR_TR < r_hi;
    if (more) {
      fetch(r0 + GR_TR, stage ^ 1);                         // (its buffer was released by the barrier that ended the last pass)
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const double *As = gsm + (size_t)stage * GR_STAGE, *Bs = As + GR_TR * GR_LDA;
#pragma unroll 2
    for (int ks = 0; ks < GR_TR / 4; ++ks) {
      const int kr = 4 * ks + fk;
      double bf[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) bf[v] = Bs[kr * GR_LDB + 8 * v + fr];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int itile = warp + 16 * u;
        if (itile < it_n) {                   // warp-uniform
          const double af = As[kr * GR_LDA + 8 * itile + fr];
#pragma unroll
          for (int v = 0; v < 8; ++v) lb_dmma(acc[u][v][0], acc[u][v][1], af, bf[v]);
        }
      }
    }
    __syncthreads();
    stage ^= 1;
  }
  double *out = partial + (size_t)blockIdx.x * k1 * k2;
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int v = 0; v < 8; ++v)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int i = 8 * (warp + 16 * u) + fr, j = 8 * v + 2 * fk + c;
        if (warp + 16 * u < it_n && i < k1 && j < kb) out[(size_t)i * k2 + j0 + j] = acc[u][v][c];
      }
}
// G[e] = sum_b partial[b][e], fixed order
__global__ void blk_reduce_kernel(const double *partial, int nb, int count, double *G) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[(size_t)b * count + e];
    G[e] = s;
  }
}

// ---- block update on the fp64 tensor cores: out[m x n2] = S[m x k] * C[k x n2]   (C row-major, ldc; n2 <= 64) ----------
// 64-row tiles; warp w owns row tile w (8 rows) and all eight 8-column tiles: 8 independent MMA chains of up to 48 k-steps.
constexpr int GM_TR = 64;
constexpr int GM_LDS = 196, GM_LDC = 68;
__global__ void __launch_bounds__(LB_THREADS) blk_gemm_kernel(unsigned long long m, const double *S, int lds, int k,
                                                              const double *C, int ldc, int n2, double *out, int ldo) {
  extern __shared__ double sm[];
  double *Cs = sm;                              // [LB_KMAX][GM_LDC]
  double *Ss = sm + (size_t)LB_KMAX * GM_LDC;   // [GM_TR][GM_LDS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int k4 = (k + 3) & ~3;
  for (int e = tid; e < k4 * 64; e += LB_THREADS) {
    const int kk = e >> 6, c = e & 63;
    Cs[kk * GM_LDC + c] = (kk < k && c < n2) ? C[(size_t)kk * ldc + c] : 0.0;
  }
  const unsigned long long r_lo = m * blockIdx.x / gridDim.x, r_hi = m * (blockIdx.x + 1ull) / gridDim.x;
  for (unsigned long long r0 = r_lo; r0 < r_hi; r0 += GM_TR) {
    const int rows = (int)min((unsigned long long)GM_TR, r_hi - r0);
    __syncthreads();
    for (int e = tid; e < GM_TR * k4; e += LB_THREADS) {
      const int rr = e / k4, c = e - rr * k4;
      Ss[rr * GM_LDS + c] = (rr < rows && c < k) ? S[(r0 + rr) * lds + c] : 0.0;
    }
    __syncthreads();
    double acc[8][2];
#pragma unroll
    for (int v = 0; v < 8; ++v) acc[v][0] = acc[v][1] = 0.0;
    for (int k0 = 0; k0 < k4; k0 += 4) {
      const double af = Ss[(8 * warp + fr) * GM_LDS + k0 + fk];
#pragma unroll
      for (int v = 0; v < 8; ++v) lb_dmma(acc[v][0], acc[v][1], af, Cs[(k0 + fk) * GM_LDC + 8 * v + fr]);
    }
    const int rr = 8 * warp + fr;
    if (rr < rows)
#pragma unroll
      for (int v = 0; v < 8; ++v)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = 8 * v + 2 * fk + c;
          if (col < n2) out[(r0 + rr) * ldo + col] = acc[v][c];
        }
  }
}

// ---- fused LOBPCG update (LOBPCG.h:239 + 249):  P = S[:, nx:ns] C[nx:ns, :],  X = S[:, :nx] C[:nx, :] + P ----------------
// One pass over S instead of two and ns nx instead of (2 ns - nx) nx multiply-adds per row: the P part of the
// contraction is shared.  Same tiling as blk_gemm_kernel, two accumulator sets per warp.
__global__ void __launch_bounds__(LB_THREADS) blk_update_kernel(unsigned long long m, const double *S, int lds, int ns, int nx,
                                                                const double *C, int ldc, double *Xo, int ldx, double *Po,
                                                                int ldp) {
  extern __shared__ double sm[];
  double *Cs = sm;                              // [LB_KMAX][GM_LDC]
  double *Ss = sm + (size_t)LB_KMAX * GM_LDC;   // [GM_TR][GM_LDS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int k4 = (ns + 3) & ~3;
  for (int e = tid; e < k4 * 64; e += LB_THREADS) {
    const int kk = e >> 6, c = e & 63;
    Cs[kk * GM_LDC + c] = (kk < ns && c < nx) ? C[(size_t)kk * ldc + c] : 0.0;
  }
  const unsigned long long r_lo = m * blockIdx.x / gridDim.x, r_hi = m * (blockIdx.x + 1ull) / gridDim.x;
  for (unsigned long long r0 = r_lo; r0 < r_hi; r0 += GM_TR) {
    const int rows = (int)min((unsigned long long)GM_TR, r_hi - r0);
    __syncthreads();
    for (int e = tid; e < GM_TR * k4; e += LB_THREADS) {
      const int rr = e / k4, c = e - rr * k4;
      Ss[rr * GM_LDS + c] = (rr < rows && c < ns) ? S[(r0 + rr) * lds + c] : 0.0;
    }
    __syncthreads();
    double ax[8][2], ap[8][2];
#pragma unroll
    for (int v = 0; v < 8; ++v) ax[v][0] = ax[v][1] = ap[v][0] = ap[v][1] = 0.0;
    for (int k0 = 0; k0 < k4; k0 += 4) {
      const double af = Ss[(8 * warp + fr) * GM_LDS + k0 + fk];
      const bool lowp = k0 < nx, highp = k0 + 3 >= nx;      // k-step touches the X part / the (W, P) part (warp-uniform)
      const bool mine_low = k0 + fk < nx;
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const double c = Cs[(k0 + fk) * GM_LDC + 8 * v + fr];
        if (lowp) lb_dmma(ax[v][0], ax[v][1], af, (highp && !mine_low) ? 0.0 : c);
        if (highp) lb_dmma(ap[v][0], ap[v][1], af, (lowp && mine_low) ? 0.0 : c);
      }
    }
    const int rr = 8 * warp + fr;
    if (rr < rows)
#pragma unroll
      for (int v = 0; v < 8; ++v)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = 8 * v + 2 * fk + c;
          if (col < nx) {
            Po[(r0 + rr) * ldp + col] = ap[v][c];
            Xo[(r0 + rr) * ldx + col] = ax[v][c] + ap[v][c];
          }
        }
  }
}

// ---- R = AX - BX diag(theta); per-CTA partial column sums of R^2 and X^2 -----------------------------------------
__global__ void __launch_bounds__(LB_THREADS) blk_residual_kernel(unsigned long long m, int nx, const double *AX, const double *BX,
                                                                  int ldb, const double *X, int ldx, const double *theta,
                                                                  double *R, double *partial /* [grid][2 nx] */) {
  __shared__ double s_r[4][64], s_x[4][64];
  const int tid = threadIdx.x, c = tid & 63, g = tid >> 6;     // 4 row groups x 64 columns
  const unsigned long long r_lo = m * blockIdx.x / gridDim.x, r_hi = m * (blockIdx.x + 1ull) / gridDim.x;
  double rr = 0.0, xx = 0.0;
  if (c < nx) {
    const double th = theta[c];
    for (unsigned long long r = r_lo + g; r < r_hi; r += 4) {
      const double x = X[r * ldx + c];
      const double v = AX[r * nx + c] - BX[r * ldb + c] * th;
      R[r * nx + c] = v;
      rr = fma(v, v, rr);
      xx = fma(x, x, xx);
    }
  }
  s_r[g][c] = rr;
  s_x[g][c] = xx;
  __syncthreads();
  if (tid < 64 && tid < nx) {
    partial[(size_t)blockIdx.x * 2 * nx + tid] = (s_r[0][tid] + s_r[1][tid]) + (s_r[2][tid] + s_r[3][tid]);
    partial[(size_t)blockIdx.x * 2 * nx + nx + tid] = (s_x[0][tid] + s_x[1][tid]) + (s_x[2][tid] + s_x[3][tid]);
  }
}
// sum of squares of a block (Frobenius norm^2), per-CTA partials
__global__ void __launch_bounds__(LB_THREADS) blk_sumsq_kernel(unsigned long long total, const double *V, double *partial) {
  __shared__ double s[LB_THREADS];
  double a = 0.0;
  const unsigned long long lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1ull) / gridDim.x;
  for (unsigned long long e = lo + threadIdx.x; e < hi; e += LB_THREADS) a = fma(V[e], V[e], a);
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = LB_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

// ---- Rayleigh-Ritz glue (ns x ns, one CTA) ---------------------------------------------------------------------
// equilibration (LOBPCG.h:56-59): D = 1/sqrt(diag(GB)); EA = sym(D GA D), EB = sym(D GB D)   (symmetric: row-major ==
// column-major); out of place
__global__ void rr_equilibrate_kernel(int ns, const double *GA, const double *GB, double *EA, double *EB, double *D) {
  __shared__ double sD[LB_KMAX];
  for (int i = threadIdx.x; i < ns; i += blockDim.x) {
    sD[i] = 1.0 / sqrt(GB[i * ns + i]);
    D[i] = sD[i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ns * ns; e += blockDim.x) {
    const int i = e / ns, j = e - i * ns;
    const double w = sD[i] * sD[j];
    const int ib = i >> 6, jb = j >> 6;     // only the block-upper triangle of the Grams was formed
    const double a = ib == jb ? 0.5 * (GA[i * ns + j] + GA[j * ns + i]) : (ib < jb ? GA[i * ns + j] : GA[j * ns + i]);
    const double b = ib == jb ? 0.5 * (GB[i * ns + j] + GB[j * ns + i]) : (ib < jb ? GB[i * ns + j] : GB[j * ns + i]);
    EA[e] = a * w;
    EB[e] = b * w;
  }
}
// C[k][j] (row-major, ld ns) = D[k] * Z(k, j), Z column-major eigenvectors from the solver (LOBPCG.h:61)
__global__ void rr_scale_transpose_kernel(int ns, const double *Z, const double *D, double *C) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ns * ns; e += gridDim.x * blockDim.x) {
    const int k = e / ns, j = e - k * ns;
    C[e] = D[k] * Z[k + (size_t)j * ns];
  }
}

// ---- launchers -------------------------------------------------------------------------------------------
static int lb_grid(unsigned long long total, int sm_count) {
  unsigned long long g = (total + LB_THREADS - 1) / LB_THREADS;
  const unsigned long long cap = (unsigned long long)sm_count * 8ull;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}
cudaError_t launch_blk_diag(unsigned long long m, int k, const double *d, double alpha, const double *in, int ldi, double *out,
                            int ldo, int sm_count, cudaStream_t st) {
  blk_diag_kernel<<<lb_grid(m * k, sm_count), LB_THREADS, 0, st>>>(m, k, d, alpha, in, ldi, out, ldo);
  return cudaGetLastError();
}
// The same stencil with 16-byte accesses and the grid coordinates computed once per row (the scalar version above does
// four 64-bit divisions per element: it ran at 1.8 TB/s): a warp takes a row, lane l the column pairs l, l + 32, l + 64.
// Requires even k / ldi / ldo and 16-byte aligned bases (the bases of LOBPCG with an even block size), m < 2^32.
__global__ void __launch_bounds__(LB_THREADS) blk_stencil7_vec_kernel(unsigned gx, unsigned gy, unsigned gz, int k, const double *in,
                                                                      int ldi, double *out, int ldo, int has_lo, int has_hi) {
  const unsigned m = gx * gy * gz, sz = gx * gy;
  const int lane = threadIdx.x & 31, k2 = k >> 1;
  const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
  for (unsigned r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < m; r += nwarps) {
    const unsigned z = r / sz, rem = r - z * sz, y = rem / gx, x = rem - y * gx;
    const bool xm = x > 0, xp = x + 1 < gx, ym = y > 0, yp = y + 1 < gy, zm = z > 0 || has_lo, zp = z + 1 < gz || has_hi;
    const double *row = in + (size_t)r * ldi;
    double *orow = out + (size_t)r * ldo;
    for (int c2 = lane; c2 < k2; c2 += 32) {
      const double2 *p = reinterpret_cast<const double2 *>(row) + c2;
      const double2 c = *p;
      const double2 z0 = make_double2(0.0, 0.0);
      const double2 a = xm ? *(p - (ldi >> 1)) : z0, b = xp ? *(p + (ldi >> 1)) : z0;
      const double2 d = ym ? *(p - (size_t)gx * (ldi >> 1)) : z0, e = yp ? *(p + (size_t)gx * (ldi >> 1)) : z0;
      const double2 f = zm ? *(p - (size_t)sz * (ldi >> 1)) : z0, g = zp ? *(p + (size_t)sz * (ldi >> 1)) : z0;
      // same operation order as the scalar kernel: ((((((6 c) - a) - b) - d) - e) - f) - g, absent neighbours skipped
      double2 v = make_double2(6.0 * c.x, 6.0 * c.y);
      if (xm) { v.x -= a.x; v.y -= a.y; }
      if (xp) { v.x -= b.x; v.y -= b.y; }
      if (ym) { v.x -= d.x; v.y -= d.y; }
      if (yp) { v.x -= e.x; v.y -= e.y; }
      if (zm) { v.x -= f.x; v.y -= f.y; }
      if (zp) { v.x -= g.x; v.y -= g.y; }
      *(reinterpret_cast<double2 *>(orow) + c2) = v;
    }
  }
}
static bool stencil_vec_ok(unsigned gx, unsigned gy, unsigned gz, int k, const double *in, int ldi, const double *out, int ldo) {
  return !(k & 1) && !(ldi & 1) && !(ldo & 1) && !((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) &&
         (unsigned long long)gx * gy * gz < (1ull << 32);
}

cudaError_t launch_blk_stencil7(unsigned gx, unsigned gy, unsigned gz, int k, const double *in, int ldi, double *out, int ldo,
                                int sm_count, cudaStream_t st) {
  if (stencil_vec_ok(gx, gy, gz, k, in, ldi, out, ldo))
    blk_stencil7_vec_kernel<<<sm_count * 8, LB_THREADS, 0, st>>>(gx, gy, gz, k, in, ldi, out, ldo, 0, 0);
  else
    blk_stencil7_kernel<<<lb_grid((unsigned long long)gx * gy * gz * k, sm_count), LB_THREADS, 0, st>>>(gx, gy, gz, k, in, ldi, out, ldo, 0, 0);
  return cudaGetLastError();
}
cudaError_t launch_blk_stencil7_slab(unsigned gx, unsigned gy, unsigned gz, int k, const double *in, int ldi, double *out, int ldo,
                                     int has_lo, int has_hi, int sm_count, cudaStream_t st) {
  if (stencil_vec_ok(gx, gy, gz, k, in, ldi, out, ldo))
    blk_stencil7_vec_kernel<<<sm_count * 8, LB_THREADS, 0, st>>>(gx, gy, gz, k, in, ldi, out, ldo, has_lo, has_hi);
  else
    blk_stencil7_kernel<<<lb_grid((unsigned long long)gx * gy * gz * k, sm_count), LB_THREADS, 0, st>>>(gx, gy, gz, k, in, ldi, out, ldo,
                                                                                                   has_lo, has_hi);
  return cudaGetLastError();
}

// ---- row-sharded LOBPCG (one process per GPU): exchanges through the ranks' halo buffers (NVLink peer stores) ------------
// All-reduce of `count` doubles: every rank stores its values into slot [rank] of every rank's region, raises a flag per
// destination, waits for all sources and sums the slots IN RANK ORDER (every rank forms the same sum, bit for bit, so the
// replicated Rayleigh-Ritz steps stay identical).
struct LobPeers {
  double *ar[MAX_RANKS];     // rank q's all-reduce region for the current parity
  size_t slot_stride;        // doubles per source slot
};
__global__ void __launch_bounds__(512) lob_allreduce_kernel(CommDev cm, unsigned long long gphase, LobPeers pe, double *buf,
                                                            int count, int *abort_flag) {
  for (int q = 0; q < cm.world; ++q) {
    double *dst = pe.ar[q] + (size_t)cm.rank * pe.slot_stride;
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  const int slot = (int)(gphase % ACC_SLOTS);
  if ((int)threadIdx.x < cm.world) {
    st_release_sys_u64(cm.flags[threadIdx.x] + slot * MAX_RANKS + cm.rank, gphase + 1ull);
    const unsigned long long *f = cm.flags[cm.rank] + slot * MAX_RANKS + threadIdx.x;
    unsigned long long spins = 0;
    while (ld_acquire_sys_u64(f) < gphase + 1ull) {
      if (++spins > (1ull << 26)) { atomicExch(abort_flag, 1); break; }
      if (spins > 256) __nanosleep(64);
    }
    __threadfence();
  }
  __syncthreads();
  const double *mine = pe.ar[cm.rank];
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    double v = __ldcg(mine + i);
    for (int q = 1; q < cm.world; ++q) v += __ldcg(mine + (size_t)q * pe.slot_stride + i);
    buf[i] = v;
  }
}
cudaError_t launch_lob_allreduce(const CommDev &cm, unsigned long long gphase, double *const *peer_region, size_t slot_stride,
                                 double *buf, int count, int *abort_flag, cudaStream_t st) {
  LobPeers pe;
  for (int q = 0; q < MAX_RANKS; ++q) pe.ar[q] = q < cm.world ? peer_region[q] : nullptr;
  pe.slot_stride = slot_stride;
  lob_allreduce_kernel<<<1, 512, 0, st>>>(cm, gphase, pe, buf, count, abort_flag);
  return cudaGetLastError();
}

// Ghost planes of a z-slab: push my first / last plane (k columns of a row-major block vector with leading dimension ld)
// into the neighbours' receive regions, the last CTA to finish raises their flags; the pull kernel waits for my own flags
// and copies the received planes into my ghost planes.
__global__ void __launch_bounds__(LB_THREADS) lob_plane_push_kernel(PlaneXchg px, const double *interior, int ld, int k,
                                                                    unsigned long long plane_rows, unsigned long long planes) {
  const unsigned long long per = plane_rows * (unsigned long long)k;
  for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < 2 * per;
       e += (unsigned long long)gridDim.x * blockDim.x) {
    const bool top = e >= per;
    const unsigned long long i = top ? e - per : e, r = i / k, c = i - r * k;
    if (!top && px.lo_dst) px.lo_dst[i] = interior[r * ld + c];                                     // my first plane -> rank r-1
    if (top && px.hi_dst) px.hi_dst[i] = interior[((planes - 1) * plane_rows + r) * ld + c];      // my last plane -> rank r+1
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(px.counter, 1u);
    if (done == gridDim.x - 1) {
      __threadfence_system();
      if (px.lo_flag) st_release_sys_u64(px.lo_flag, px.seq);
      if (px.hi_flag) st_release_sys_u64(px.hi_flag, px.seq);
    }
  }
}
__global__ void __launch_bounds__(LB_THREADS) lob_plane_pull_kernel(PlaneXchg px, double *interior, int ld, int k,
                                                                    unsigned long long plane_rows, unsigned long long planes,
                                                                    int has_lo, int has_hi, int *abort_flag) {
  if (threadIdx.x == 0) {
    for (int d = 0; d < 2; ++d) {
      if (!(d ? has_hi : has_lo)) continue;
      unsigned long long spins = 0;
      while (ld_acquire_sys_u64(px.my_flags + d) < px.seq) {
        if (++spins > (1ull << 26)) { atomicExch(abort_flag, 1); break; }
        if (spins > 256) __nanosleep(64);
      }
    }
    __threadfence();
  }
  __syncthreads();
  const unsigned long long per = plane_rows * (unsigned long long)k;
  for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < 2 * per;
       e += (unsigned long long)gridDim.x * blockDim.x) {
    const bool top = e >= per;
    const unsigned long long i = top ? e - per : e, r = i / k, c = i - r * k;
    if (!top && has_lo) (interior - plane_rows * ld)[r * ld + c] = __ldcg(px.from_below + i);          // ghost plane below
    if (top && has_hi) (interior + planes * plane_rows * ld)[r * ld + c] = __ldcg(px.from_above + i);  // ghost plane above
  }
}
cudaError_t launch_lob_plane_exchange(const PlaneXchg &px, double *interior, int ld, int k, unsigned long long plane_rows,
                                      unsigned long long planes, int has_lo, int has_hi, int *abort_flag, int sm_count,
                                      cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(px.counter, 0, sizeof(unsigned), st);
  if (e) return e;
  const int grid = lb_grid(2 * plane_rows * (unsigned long long)k, sm_count);
  lob_plane_push_kernel<<<grid, LB_THREADS, 0, st>>>(px, interior, ld, k, plane_rows, planes);
  lob_plane_pull_kernel<<<grid, LB_THREADS, 0, st>>>(px, interior, ld, k, plane_rows, planes, has_lo, has_hi, abort_flag);
  return cudaGetLastError();
}
// G (k1 x k2, row-major) = A^T B ; partial: scratch of nb * k1 * k2 doubles
cudaError_t launch_blk_gram(unsigned long long m, const double *A, int lda, int k1, const double *B, int ldb, int k2,
                            double *partial, int nb, double *G, cudaStream_t st) {
  static bool attr = false;
  const size_t smem = sizeof(double) * 2 * (size_t)GR_STAGE;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(blk_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    attr = true;
  }
  dim3 grid(nb, (k2 + 63) / 64);
  blk_gram_kernel<<<grid, GR_THREADS, smem, st>>>(m, A, lda, k1, B, ldb, k2, partial);
  blk_reduce_kernel<<<(k1 * k2 + 255) / 256, 256, 0, st>>>(partial, nb, k1 * k2, G);
  return cudaGetLastError();
}
cudaError_t launch_blk_gemm(unsigned long long m, const double *S, int lds, int k, const double *C, int ldc, int n2, double *out,
                            int ldo, int nb, cudaStream_t st) {
  static bool attr = false;
  const size_t smem = sizeof(double) * ((size_t)LB_KMAX * GM_LDC + (size_t)GM_TR * GM_LDS);
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(blk_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    attr = true;
  }
  blk_gemm_kernel<<<nb, LB_THREADS, smem, st>>>(m, S, lds, k, C, ldc, n2, out, ldo);
  return cudaGetLastError();
}
// The same update with a two-stage cp.async pipeline over 32-row tiles of S (rows of S 16-byte aligned: the bases of
// LOBPCG with an even block size).  The synchronous version above spends more time waiting for a tile than multiplying
// it.  Warp w: row tile w & 3 (8 rows), column half w >> 2 (four 8-column tiles): eight accumulator chains as before.
constexpr int UP_TR = 32;
constexpr int UP_STAGE = UP_TR * GM_LDS;
__global__ void __launch_bounds__(LB_THREADS, 1) blk_update_pipe_kernel(unsigned long long m, const double *S, int lds, int ns,
                                                                      int nx, const double *C, int ldc, double *Xo, int ldx,
                                                                      double *Po, int ldp) {
  extern __shared__ double sm[];
  double *Cs = sm;                              // [LB_KMAX][GM_LDC]
  double *Sb = sm + (size_t)LB_KMAX * GM_LDC;   // two stages of [UP_TR][GM_LDS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int rt = warp & 3, ch = warp >> 2;
  const int k4 = (ns + 3) & ~3, k2c = k4 >> 1;  // 16-byte chunks per row
  const bool vec_out = !(ldp & 1) && !(ldx & 1) && !((reinterpret_cast<uintptr_t>(Po) | reinterpret_cast<uintptr_t>(Xo)) & 15);
  for (int e = tid; e < k4 * 64; e += LB_THREADS) {
    const int kk = e >> 6, c = e & 63;
    Cs[kk * GM_LDC + c] = (kk < ns && c < nx) ? C[(size_t)kk * ldc + c] : 0.0;
  }
  const unsigned long long r_lo = m * blockIdx.x / gridDim.x, r_hi = m * (blockIdx.x + 1ull) / gridDim.x;
  auto fetch = [&](unsigned long long r0, int stage) {
    double *Ss = Sb + (size_t)stage * UP_STAGE;
    const int rows = (int)min((unsigned long long)UP_TR, r_hi - r0);
    for (int e = tid; e < UP_TR * k2c; e += LB_THREADS) {
      const int rr = e / k2c, c2 = e - rr * k2c;
      // bytes of this chunk that hold columns < ns (0, 8 or 16): the rest is zero-filled by the copy
      const int nbytes = rr < rows ? max(0, min(16, 8 * (ns - 2 * c2))) : 0;
      const unsigned d = (unsigned)__cvta_generic_to_shared(Ss + rr * GM_LDS + 2 * c2);
      const double *src = nbytes ? S + (r0 + rr) * lds + 2 * c2 : S;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int stage = 0;
  if (r_lo < r_hi) fetch(r_lo, 0);
  for (unsigned long long r0 = r_lo; r0 < r_hi; r0 += UP_TR) {
    const int rows = (int)min((unsigned long long)UP_TR, r_hi - r0);
    if (r0 + UP_TR < r_hi) {
      fetch(r0 + UP_TR, stage ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();                              // (also orders the first pass after the staging of C)
    const double *Ss = Sb + (size_t)stage * UP_STAGE;
    double ax[4][2], ap[4][2];
#pragma unroll
    for (int v = 0; v < 4; ++v) ax[v][0] = ax[v][1] = ap[v][0] = ap[v][1] = 0.0;
    for (int k0 = 0; k0 < k4; k0 += 4) {
      const double af = Ss[(8 * rt + fr) * GM_LDS + k0 + fk];
      const bool lowp = k0 < nx, highp = k0 + 3 >= nx;      // k-step touches the X part / the (W, P) part (warp-uniform)
      const bool mine_low = k0 + fk < nx;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const double c = Cs[(k0 + fk) * GM_LDC + 32 * ch + 8 * v + fr];
        if (lowp) lb_dmma(ax[v][0], ax[v][1], af, (highp && !mine_low) ? 0.0 : c);
        if (highp) lb_dmma(ap[v][0], ap[v][1], af, (lowp && mine_low) ? 0.0 : c);
      }
    }
    const int rr = 8 * rt + fr;
    if (rr < rows)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int col = 32 * ch + 8 * v + 2 * fk;
        if (col + 1 < nx && vec_out) {
          *reinterpret_cast<double2 *>(Po + (r0 + rr) * ldp + col) = make_double2(ap[v][0], ap[v][1]);
          *reinterpret_cast<double2 *>(Xo + (r0 + rr) * ldx + col) = make_double2(ax[v][0] + ap[v][0], ax[v][1] + ap[v][1]);
        } else {
#pragma unroll
          for (int c = 0; c < 2; ++c)
            if (col + c < nx) {
              Po[(r0 + rr) * ldp + col + c] = ap[v][c];
              Xo[(r0 + rr) * ldx + col + c] = ax[v][c] + ap[v][c];
            }
        }
      }
    __syncthreads();
    stage ^= 1;
  }
}
cudaError_t launch_blk_update(unsigned long long m, const double *S, int lds, int ns, int nx, const double *C, int ldc, double *Xo,
                              int ldx, double *Po, int ldp, int nb, cudaStream_t st) {
  static bool attr = false;
  const size_t smem = sizeof(double) * ((size_t)LB_KMAX * GM_LDC + (size_t)GM_TR * GM_LDS);
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(blk_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    attr = true;
  }
  const bool pipe = !(lds & 1) && !(reinterpret_cast<uintptr_t>(S) & 15);
  if (pipe) {
    static bool attr2 = false;
    const size_t smem2 = sizeof(double) * ((size_t)LB_KMAX * GM_LDC + 2 * (size_t)UP_STAGE);
    if (!attr2) {
      cudaError_t e = cudaFuncSetAttribute(blk_update_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      if (e) return e;
      attr2 = true;
    }
    blk_update_pipe_kernel<<<nb, LB_THREADS, smem2, st>>>(m, S, lds, ns, nx, C, ldc, Xo, ldx, Po, ldp);
    return cudaGetLastError();
  }
  blk_update_kernel<<<nb, LB_THREADS, smem, st>>>(m, S, lds, ns, nx, C, ldc, Xo, ldx, Po, ldp);
  return cudaGetLastError();
}
// R, and norms2[0:nx] = column sums of R^2, norms2[nx:2nx] = column sums of X^2
cudaError_t launch_blk_residual(unsigned long long m, int nx, const double *AX, const double *BX, int ldb, const double *X,
                                int ldx, const double *theta, double *R, double *partial, int nb, double *norms2,
                                cudaStream_t st) {
  const int g = nb * 8;   // several CTAs per SM: the kernel is a plain stream with one load in flight per thread and array
  blk_residual_kernel<<<g, LB_THREADS, 0, st>>>(m, nx, AX, BX, ldb, X, ldx, theta, R, partial);
  blk_reduce_kernel<<<1, 256, 0, st>>>(partial, g, 2 * nx, norms2);
  return cudaGetLastError();
}
cudaError_t launch_blk_sumsq(unsigned long long total, const double *V, double *partial, int nb, double *out, cudaStream_t st) {
  blk_sumsq_kernel<<<nb, LB_THREADS, 0, st>>>(total, V, partial);
  blk_reduce_kernel<<<1, 32, 0, st>>>(partial, nb, 1, out);
  return cudaGetLastError();
}
cudaError_t launch_rr_equilibrate(int ns, const double *GA, const double *GB, double *EA, double *EB, double *D,
                                  cudaStream_t st) {
  rr_equilibrate_kernel<<<1, 256, 0, st>>>(ns, GA, GB, EA, EB, D);
  return cudaGetLastError();
}
cudaError_t launch_rr_scale_transpose(int ns, const double *Z, const double *D, double *C, cudaStream_t st) {
  rr_scale_transpose_kernel<<<(ns * ns + 255) / 256, 256, 0, st>>>(ns, Z, D, C);
  return cudaGetLastError();
}

}  // namespace ob200
