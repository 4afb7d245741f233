// Level-1 primitives with the exact order-independent reduction.  These are
// what the generic (arbitrary user Hessian functor) tCG path and the TNT outer
// loop use for `metric(x, a, b)` (reference TNT.h:382,387,493,511-512,575,579)
// and the vector statements of IterativeSolvers.h:211-256,336,374-420.
#include "tcg.cuh"

namespace ob200 {

__device__ __forceinline__ void l1_load_run(const double *base, unsigned long long N,
                                            unsigned long long e0, int lane, double2 (&v)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned long long e = e0 + 2ull * (unsigned)(lane + 32 * i);
    if (e + 1 < N) v[i] = ldcg2(base + e);
    else {
      v[i].x = (e < N) ? __ldcg(base + e) : 0.0;
      v[i].y = 0.0;
    }
  }
}
__device__ __forceinline__ void l1_store_run(double *base, unsigned long long N, unsigned long long e0,
                                             int lane, const double2 (&v)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned long long e = e0 + 2ull * (unsigned)(lane + 32 * i);
    if (e + 1 < N) stcg2(base + e, v[i]);
    else if (e < N) __stcg(base + e, v[i].x);
  }
}

struct DotsArgs {
  const double *a[4];
  const double *b[4];
  int count;
};

// up to 4 inner products in one pass; unit of determinism = 256-element run
__global__ void __launch_bounds__(TCG_THREADS) dots_kernel(unsigned long long N, DotsArgs d, u64 *set) {
  __shared__ u64 sacc[4 * KUL_STRIDE];
  for (int i = threadIdx.x; i < 4 * KUL_STRIDE; i += blockDim.x) sacc[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long units = (N + 255ull) / 256ull;
  for (unsigned long long u = (unsigned long long)blockIdx.x * TCG_WARPS + warp; u < units;
       u += (unsigned long long)gridDim.x * TCG_WARPS) {
    const unsigned long long e0 = u * 256ull;
    for (int c = 0; c < d.count; ++c) {
      double2 x[4], y[4];
      l1_load_run(d.a[c], N, e0, lane, x);
      if (d.b[c] != d.a[c]) l1_load_run(d.b[c], N, e0, lane, y);
      double part = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double2 yy = (d.b[c] != d.a[c]) ? y[i] : x[i];
        part = fma(x[i].x, yy.x, part);
        part = fma(x[i].y, yy.y, part);
      }
      part = warp_sum(part);
      if (lane == 0) kul_add_atomic(sacc + c * KUL_STRIDE, part);
    }
  }
  __syncthreads();
  flush_scalars(sacc, set, d.count);
}

// Prologue of the stand-alone Stiefel HVP in ONE launch (it used to be seven stream operations): the exact <V,V> (bound of
// the fixed-point projection Gram), the content checksum of A that validates the cached digit planes, and the clearing
// of what the persistent kernel expects to be zero.  `vv_acc` / `sum` alternate between two copies from call to call:
// this launch accumulates into one copy (zero on entry) and clears the other for the next call.  The last CTA to finish
// rounds <V,V> once (same exact accumulator and finalisation as dots_kernel + finalize_many_kernel).
__global__ void __launch_bounds__(TCG_THREADS)
hvp_prologue_kernel(unsigned long long N, const double *V, const uint4 *A16, unsigned long long nvec, u64 *vv_acc,
                    unsigned long long *sum, u64 *vv_acc_next, unsigned long long *sum_next, u64 *zero_words,
                    unsigned long long n_zero, unsigned *barrier_words, unsigned *done_counter, double *vv_out) {
  __shared__ u64 sacc[KUL_STRIDE];
  __shared__ int s_last;
  for (int i = threadIdx.x; i < KUL_STRIDE; i += blockDim.x) sacc[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long gthreads = (unsigned long long)gridDim.x * blockDim.x;
  // (a) content checksum of A (stiefel_checksum_kernel's hash; the loads go first: they are the long pole)
  unsigned long long acc = 0;
  for (unsigned long long i = gtid; i < nvec; i += gthreads) acc += a_checksum_term(__ldg(A16 + i), i);
  // (b) <V,V>, unit of determinism = 256-element run
  const unsigned long long units = (N + 255ull) / 256ull;
  for (unsigned long long u = (unsigned long long)blockIdx.x * TCG_WARPS + warp; u < units;
       u += (unsigned long long)gridDim.x * TCG_WARPS) {
    double2 x[4];
    l1_load_run(V, N, u * 256ull, lane, x);
    double part = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      part = fma(x[i].x, x[i].x, part);
      part = fma(x[i].y, x[i].y, part);
    }
    part = warp_sum(part);
    if (lane == 0) kul_add_atomic(sacc, part);
  }
  // (c) what the persistent kernel and the next call expect to be zero
  for (unsigned long long i = gtid; i < n_zero; i += gthreads) zero_words[i] = 0;
  if (blockIdx.x == 0) {
    if (threadIdx.x < 16) barrier_words[threadIdx.x] = 0u;
    for (int i = threadIdx.x; i < KUL_STRIDE; i += blockDim.x) vv_acc_next[i] = 0;
    if (threadIdx.x == 0) *sum_next = 0ull;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0 && acc) atomicAdd(sum, acc);
  __syncthreads();
  flush_scalars(sacc, vv_acc, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(done_counter, 1u) + 1u == gridDim.x);
  }
  __syncthreads();
  if (s_last && warp == 0) {
    __threadfence();
    const double v = kul_finalize_warp([vv_acc](int j) { return __ldcg(vv_acc + j); });
    if (lane == 0) {
      *vv_out = v;
      *done_counter = 0u;
    }
  }
}

__global__ void finalize_many_kernel(const u64 *set, int count, double *out) {
  const int warp = threadIdx.x >> 5;
  if (blockIdx.x == 0 && warp < count) {
    const u64 *p = set + warp * KUL_STRIDE;
    const double v = kul_finalize_warp([p](int j) { return p[j]; });
    if ((threadIdx.x & 31) == 0) out[warp] = v;
  }
}

__global__ void __launch_bounds__(TCG_THREADS)
axpby_kernel(unsigned long long N, double alpha, const double *x, double beta, const double *y, double *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long units = (N + 255ull) / 256ull;
  for (unsigned long long u = (unsigned long long)blockIdx.x * TCG_WARPS + warp; u < units;
       u += (unsigned long long)gridDim.x * TCG_WARPS) {
    const unsigned long long e0 = u * 256ull;
    double2 xv[4], yv[4], o[4];
    l1_load_run(x, N, e0, lane, xv);
    if (y) l1_load_run(y, N, e0, lane, yv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (y) {
        o[i].x = fma(alpha, xv[i].x, beta * yv[i].x);
        o[i].y = fma(alpha, xv[i].y, beta * yv[i].y);
      } else {
        o[i].x = alpha * xv[i].x;
        o[i].y = alpha * xv[i].y;
      }
    }
    l1_store_run(out, N, e0, lane, o);
  }
}

__global__ void __launch_bounds__(TCG_THREADS)
hadamard_kernel(unsigned long long N, const double *d, const double *x, double *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long units = (N + 255ull) / 256ull;
  for (unsigned long long u = (unsigned long long)blockIdx.x * TCG_WARPS + warp; u < units;
       u += (unsigned long long)gridDim.x * TCG_WARPS) {
    const unsigned long long e0 = u * 256ull;
    double2 dv[4], xv[4], o[4];
    l1_load_run(d, N, e0, lane, dv);
    l1_load_run(x, N, e0, lane, xv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[i].x = dv[i].x * xv[i].x;
      o[i].y = dv[i].y * xv[i].y;
    }
    l1_store_run(out, N, e0, lane, o);
  }
}

// out = x / a with a true IEEE division per element (LSQR / TNLS normalise with `v /= Scalar`,
// reference IterativeSolvers.h:707-799; a multiply by the reciprocal would round differently)
__global__ void __launch_bounds__(TCG_THREADS)
div_kernel(unsigned long long N, const double *x, double a, double *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long units = (N + 255ull) / 256ull;
  for (unsigned long long u = (unsigned long long)blockIdx.x * TCG_WARPS + warp; u < units;
       u += (unsigned long long)gridDim.x * TCG_WARPS) {
    const unsigned long long e0 = u * 256ull;
    double2 xv[4], o[4];
    l1_load_run(x, N, e0, lane, xv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[i].x = __ddiv_rn(xv[i].x, a);
      o[i].y = __ddiv_rn(xv[i].y, a);
    }
    l1_store_run(out, N, e0, lane, o);
  }
}

static int l1_grid(unsigned long long N, int sm_count) {
  const unsigned long long units = (N + 255ull) / 256ull;
  unsigned long long g = (units + TCG_WARPS - 1) / TCG_WARPS;
  const unsigned long long cap = (unsigned long long)sm_count * 2ull;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

cudaError_t launch_dots(unsigned long long N, int count, const double *const *a, const double *const *b,
                        u64 *set, int sm_count, cudaStream_t st) {
  DotsArgs d;
  d.count = count;
  for (int i = 0; i < 4; ++i) {
    d.a[i] = i < count ? a[i] : nullptr;
    d.b[i] = i < count ? b[i] : nullptr;
  }
  dots_kernel<<<l1_grid(N, sm_count), TCG_THREADS, 0, st>>>(N, d, set);
  return cudaGetLastError();
}
cudaError_t launch_hvp_prologue(unsigned long long N, const double *V, const unsigned short *A, unsigned long long nvec,
                                u64 *vv_acc, unsigned long long *sum, u64 *vv_acc_next, unsigned long long *sum_next,
                                u64 *zero_words, unsigned long long n_zero, unsigned *barrier_words,
                                unsigned *done_counter, double *vv_out, int sm_count, cudaStream_t st) {
  hvp_prologue_kernel<<<2 * sm_count, TCG_THREADS, 0, st>>>(N, V, reinterpret_cast<const uint4 *>(A), nvec, vv_acc, sum,
                                                            vv_acc_next, sum_next, zero_words, n_zero, barrier_words,
                                                            done_counter, vv_out);
  return cudaGetLastError();
}
cudaError_t launch_finalize_many(const u64 *set, int count, double *out, cudaStream_t st) {
  finalize_many_kernel<<<1, 32 * (count < 1 ? 1 : count), 0, st>>>(set, count, out);
  return cudaGetLastError();
}
cudaError_t launch_axpby(unsigned long long N, double alpha, const double *x, double beta, const double *y,
                         double *out, int sm_count, cudaStream_t st) {
  axpby_kernel<<<l1_grid(N, sm_count), TCG_THREADS, 0, st>>>(N, alpha, x, beta, y, out);
  return cudaGetLastError();
}
cudaError_t launch_div(unsigned long long N, const double *x, double a, double *out, int sm_count, cudaStream_t st) {
  div_kernel<<<l1_grid(N, sm_count), TCG_THREADS, 0, st>>>(N, x, a, out);
  return cudaGetLastError();
}
cudaError_t launch_hadamard(unsigned long long N, const double *d, const double *x, double *out,
                            int sm_count, cudaStream_t st) {
  hadamard_kernel<<<l1_grid(N, sm_count), TCG_THREADS, 0, st>>>(N, d, x, out);
  return cudaGetLastError();
}

}  // namespace ob200
