// Shared device/host helpers for the fused tCG kernels (sm_100a).
//
// Determinism model: every inner product that feeds a
// decision of the Steihaug-Toint loop (reference IterativeSolvers.h:290,305-307,
// 320,341,347,408) is reduced as
//   fixed unit (8-row strip / 256-element run) -> double partial in a fixed
//   lane order -> EXACT integer accumulation (Kulisch long accumulator).
// Integer addition is associative, so the result does not depend on which
// warp / CTA / GPU processed which unit, on the grid size, or on atomics
// ordering: runs are bit-reproducible and independent of the GPU count.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <math.h>

namespace ob200 {

// ---------------------------------------------------------------------------
// Kulisch-style long accumulator: 32-bit digits held in int64 limbs.
// value = sum_j limb[j] * 2^(32 j - KUL_BIAS)
// ---------------------------------------------------------------------------
constexpr int KUL_BIAS = 1088;      // > 1074 (smallest subnormal exponent), multiple of 32
constexpr int KUL_LIMBS = 68;       // covers up to 2^(68*32-1088) = 2^1088
constexpr int KUL_STRIDE = 72;      // limbs + [68] non-finite counter, padded

typedef unsigned long long u64;
typedef long long i64;

// Split x into (at most) three signed 32-bit digits at limb j, j+1, j+2.
// Returns false for inf / nan.
__host__ __device__ __forceinline__ bool kul_decompose(double x, int &j, i64 &d0, i64 &d1, i64 &d2) {
  u64 bits;
#ifdef __CUDA_ARCH__
  bits = (u64)__double_as_longlong(x);
#else
  memcpy(&bits, &x, 8);
#endif
  const int e = (int)((bits >> 52) & 0x7ff);
  u64 mant = bits & 0xFFFFFFFFFFFFFull;
  if (e == 0x7ff) return false;
  int shift;
  if (e == 0) {
    shift = -1074 + KUL_BIAS;
  } else {
    mant |= (1ull << 52);
    shift = e - 1075 + KUL_BIAS;
  }
  j = shift >> 5;
  const int off = shift & 31;
  const u64 lo = mant << off;                          // low 64 bits of mant * 2^off
  const u64 hi = off ? (mant >> (64 - off)) : 0ull;    // bits above 64 (mant < 2^53, off < 32)
  d0 = (i64)(lo & 0xFFFFFFFFull);
  d1 = (i64)(lo >> 32);
  d2 = (i64)hi;
  if (bits >> 63) { d0 = -d0; d1 = -d1; d2 = -d2; }
  return true;
}

#ifdef __CUDACC__
// Add x into a (shared- or global-memory) accumulator with 64-bit atomics.
__device__ __forceinline__ void kul_add_atomic(u64 *acc, double x) {
  if (x == 0.0) return;
  int j;
  i64 d0, d1, d2;
  if (!kul_decompose(x, j, d0, d1, d2)) {  // inf / nan poisons the sum
    atomicAdd(acc + KUL_LIMBS, 1ull);
    return;
  }
  if (d0) atomicAdd(acc + j, (u64)d0);
  if (d1) atomicAdd(acc + j + 1, (u64)d1);
  if (d2) atomicAdd(acc + j + 2, (u64)d2);
}
#endif

// host twin (tests): plain adds
inline void kul_add_host(u64 *acc, double x) {
  if (x == 0.0) return;
  int j;
  i64 d0, d1, d2;
  if (!kul_decompose(x, j, d0, d1, d2)) { acc[KUL_LIMBS] += 1; return; }
  acc[j] += (u64)d0;
  acc[j + 1] += (u64)d1;
  acc[j + 2] += (u64)d2;
}

// Correctly rounded (round-to-nearest-even) double value of an accumulator.
// `ld(j)` returns limb j.  Executed redundantly and identically per CTA.
template <class Load>
__host__ __device__ inline double kul_finalize(Load ld) {
  if (ld(KUL_LIMBS) != 0) {
    const u64 nanbits = 0x7ff8000000000000ull;
    double nanv;
    memcpy(&nanv, &nanbits, 8);
    return nanv;
  }
  i64 raw[KUL_LIMBS];
  i64 carry = 0;
#pragma unroll 1
  for (int j = 0; j < KUL_LIMBS; ++j) {
    raw[j] = (i64)ld(j);
    carry = (raw[j] + carry) >> 32;   // arithmetic shift: floor division
  }
  const bool neg = carry < 0;
  uint32_t dig[KUL_LIMBS];
  carry = 0;
  int top = -1;
#pragma unroll 1
  for (int j = 0; j < KUL_LIMBS; ++j) {
    const i64 t = (neg ? -raw[j] : raw[j]) + carry;
    dig[j] = (uint32_t)(t & 0xFFFFFFFFll);
    carry = t >> 32;
    if (dig[j]) top = j;
  }
  if (top < 0) return 0.0;
  const uint32_t d_top = dig[top];
  const uint32_t d_mid = top >= 1 ? dig[top - 1] : 0u;
  const uint32_t d_low = top >= 2 ? dig[top - 2] : 0u;
  bool sticky = false;
#pragma unroll 1
  for (int j = 0; j < top - 2; ++j) sticky = sticky || (dig[j] != 0);
  // 96-bit integer I = d_top*2^64 + d_mid*2^32 + d_low, value = I * 2^(32*(top-2) - BIAS)
  const u64 hi64 = ((u64)d_top << 32) | (u64)d_mid;
#ifdef __CUDA_ARCH__
  const int lz = __clzll((i64)hi64);
#else
  const int lz = __builtin_clzll(hi64);
#endif
  u64 m64 = hi64;
  uint32_t rem = d_low;
  if (lz) {  // d_top != 0 -> lz <= 31
    m64 = (hi64 << lz) | ((u64)d_low >> (32 - lz));
    rem = d_low << lz;
  }
  if (rem != 0 || sticky) m64 |= 1ull;  // sticky bit, far below the rounding position
#ifdef __CUDA_ARCH__
  const double m = __ull2double_rn(m64);
#else
  const double m = (double)m64;
#endif
  const int ex = 32 * (top - 1) - KUL_BIAS - lz;
  const double v = scalbn(m, ex);
  return neg ? -v : v;
}

#ifdef __CUDACC__
// Warp-cooperative finalize: all 32 lanes call it with the same accumulator and
// all return the same correctly rounded value as kul_finalize.  The limbs are
// fetched with three coalesced loads; the serial carry chain only runs over the
// window of non-zero limbs (a handful in practice) and is executed redundantly
// by every lane from warp shuffles, so there is no local-memory array and no
// dependent global-memory latency chain.
template <class Load>
__device__ __forceinline__ double kul_finalize_warp(Load ld) {
  const int lane = threadIdx.x & 31;
  const i64 v0 = (i64)ld(lane);
  const i64 v1 = (i64)ld(lane + 32);
  const i64 v2 = (lane + 64 <= KUL_LIMBS) ? (i64)ld(lane + 64) : 0;   // limbs 64..67 and the non-finite counter [68]
  const unsigned bad = __ballot_sync(0xffffffffu, lane + 64 == KUL_LIMBS && v2 != 0);
  if (bad) return __longlong_as_double(0x7ff8000000000000ll);
  const unsigned b0 = __ballot_sync(0xffffffffu, v0 != 0);
  const unsigned b1 = __ballot_sync(0xffffffffu, v1 != 0);
  const unsigned b2 = __ballot_sync(0xffffffffu, v2 != 0) & 0xfu;
  if ((b0 | b1 | b2) == 0) return 0.0;
  const int lo = b0 ? (__ffs(b0) - 1) : (b1 ? 32 + __ffs(b1) - 1 : 64 + __ffs(b2) - 1);
  int hi = b2 ? 64 + (31 - __clz(b2)) : (b1 ? 32 + (31 - __clz(b1)) : (31 - __clz(b0)));
  hi = min(hi + 3, KUL_LIMBS - 1);   // room for the carry out of the top non-zero limb (< 2^33) and the sign
  auto limb = [&](int j) -> i64 {
    const i64 src = j < 32 ? v0 : (j < 64 ? v1 : v2);
    return __shfl_sync(0xffffffffu, src, j & 31);
  };
  i64 carry = 0;
  // (not unrolled: this runs once per reduction on one warp; unrolled copies of the shuffle-select chains cost ~1000
  // instructions per call site, and the persistent kernels are instruction-cache bound between their phases)
#pragma unroll 1
  for (int j = lo; j <= hi; ++j) carry = (limb(j) + carry) >> 32;
  const bool neg = carry < 0;
  carry = 0;
  // sliding window over the magnitude digits: (d_top, d_mid, d_low, sticky) at the last non-zero digit
  uint32_t h0 = 0, h1 = 0, h2 = 0;
  bool stk = false;
  uint32_t t0 = 0, t1 = 0, t2 = 0;
  bool tstk = false;
  int top = -1;
#pragma unroll 1
  for (int j = lo; j <= hi; ++j) {
    const i64 r = limb(j);
    const i64 t = (neg ? -r : r) + carry;
    const uint32_t d = (uint32_t)(t & 0xFFFFFFFFll);
    carry = t >> 32;
    stk = stk || (h2 != 0);
    h2 = h1; h1 = h0; h0 = d;
    if (d) { top = j; t0 = h0; t1 = h1; t2 = h2; tstk = stk; }
  }
  if (top < 0) return 0.0;
  const u64 hi64 = ((u64)t0 << 32) | (u64)t1;
  const int lz = __clzll((i64)hi64);
  u64 m64 = hi64;
  uint32_t rem = t2;
  if (lz) {
    m64 = (hi64 << lz) | ((u64)t2 >> (32 - lz));
    rem = t2 << lz;
  }
  if (rem != 0 || tstk) m64 |= 1ull;
  const double m = __ull2double_rn(m64);
  const int ex = 32 * (top - 1) - KUL_BIAS - lz;
  const double v = scalbn(m, ex);
  return neg ? -v : v;
}
#endif

// ---------------------------------------------------------------------------
// Bounded two-limb fixed point for the p x p Gram of the tangent projection.
// x is quantised to q = 2^(e-90) with |x| < 2^e guaranteed by the caller's
// bound; integer = hi * 2^45 + lo, both accumulated exactly in int64.
// ---------------------------------------------------------------------------
struct Fix2 { i64 hi, lo; };

__host__ __device__ __forceinline__ Fix2 fix2_from_double(double x, double inv_q /* 2^(90-e) */,
                                                 unsigned *overflow) {
  const double t = x * inv_q;                       // exact power-of-two scaling
  if (!(fabs(t) < 0x1p90)) { *overflow = 1u; Fix2 z = {0, 0}; return z; }
  const double h = floor(t * 0x1p-45);              // exact
  const double r = t - h * 0x1p45;                  // exact, in [0, 2^45)
  Fix2 f;
  f.hi = (i64)h;
  f.lo = (i64)rint(r);
  return f;
}
__host__ __device__ __forceinline__ double fix2_to_double(i64 hi, i64 lo, double q /* 2^(e-90) */) {
  // hi*2^45 exact (|hi| < 2^53); lo < 2^63 rounds once; the sum rounds once.
  return ((double)hi * 0x1p45 + (double)lo) * q;
}

// ---------------------------------------------------------------------------
// Register-resident exact accumulation of bounded partial sums (hot loops).
// A lane quantises its fp64 partial x (|x| <= bound < 2^e) to the integer
// t = x / 2^(e-90) split as hi * 2^45 + lo (lo signed, |lo| <= 2^44) and adds both
// halves to private int64 accumulators; integer adds are exact, so the total is
// independent of which warp / CTA / GPU handled which unit.  At the end of a phase
// the warp folds the 32 lanes with shuffles and lane 0 adds hi * 2^(e-45) and
// lo * 2^(e-90) into the CTA's Kulisch accumulator.
// ---------------------------------------------------------------------------
struct FixAcc { i64 hi, lo; };

__device__ __forceinline__ void fixacc_add(FixAcc &a, double x, double inv_q /* 2^(90-e) */, unsigned &ovf) {
  const double t = x * inv_q;                         // exact power-of-two scaling
  if (!(fabs(t) < 0x1p89)) { ovf = 1u; return; }      // also catches nan / inf
  const double h = rint(t * 0x1p-45);
  const double r = fma(-h, 0x1p45, t);                // exact, |r| <= 2^44
  a.hi += (i64)h;
  a.lo += (i64)rint(r);
}

#ifdef __CUDACC__
// acc += v * 2^ex  (v any int64), 64-bit shared-memory atomics on the 32-bit-digit limbs
__device__ __forceinline__ void kul_add_scaled_i64(u64 *acc, i64 v, int ex) {
  if (v == 0) return;
  const bool neg = v < 0;
  const u64 mag = neg ? (u64)(-v) : (u64)v;           // < 2^63
  const int P = ex + KUL_BIAS;                        // bit position of the LSB (callers keep it >= 0)
  const int j = P >> 5, off = P & 31;
  const u64 lo = mag << off;
  const u64 hi = off ? (mag >> (64 - off)) : 0ull;
  i64 d0 = (i64)(lo & 0xFFFFFFFFull), d1 = (i64)(lo >> 32), d2 = (i64)hi;
  if (neg) { d0 = -d0; d1 = -d1; d2 = -d2; }
  if (d0) atomicAdd(acc + j, (u64)d0);
  if (d1) atomicAdd(acc + j + 1, (u64)d1);
  if (d2) atomicAdd(acc + j + 2, (u64)d2);
}
// warp fold + flush into the CTA accumulator; e = exponent the lanes quantised with
__device__ __forceinline__ void fixacc_flush(FixAcc &a, u64 *kul, int e) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a.hi += __shfl_xor_sync(0xffffffffu, a.hi, o);
    a.lo += __shfl_xor_sync(0xffffffffu, a.lo, o);
  }
  if ((threadIdx.x & 31) == 0) {
    kul_add_scaled_i64(kul, a.hi, e - 45);
    kul_add_scaled_i64(kul, a.lo, e - 90);
  }
  a.hi = 0;
  a.lo = 0;
}
// e with sqrt(x2) < 2^e, from integer exponent arithmetic (no square root on the scalar critical path)
__device__ __forceinline__ int half_exponent(double x2) {
  if (!(x2 > 0.0) || !(x2 < 1.0e300)) return 0;
  return (ilogb(x2) + 2) >> 1;
}
// exponent for a bound: |x| <= bound  =>  e = ilogb(bound) + 2, clamped so that e - 90 + KUL_BIAS >= 0
__device__ __forceinline__ int fixacc_exponent(double bound) {
  if (!(bound > 0.0) || !(bound < 1.0e300)) return 0;
  const int e = ilogb(bound) + 2;
  return e < -990 ? -990 : e;
}
#endif

// ---------------------------------------------------------------------------
// Grid-wide barrier for the persistent kernels (cooperative launch guarantees
// co-residency).  Monotonic counter; `gen` is the caller's private generation.
// Returns false if the watchdog expired (never on a healthy run): callers then
// abandon the solve instead of hanging the device.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned atom_add_acqrel_u32(unsigned *p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---------------------------------------------------------------------------
// Multi-GPU exchange over NVLink peer memory (one process per GPU, buffers shared
// with CUDA IPC).  Every rank owns an inbox of ACC_SLOTS x MAX_RANKS accumulator
// sets plus one arrival flag per (slot, source rank).  After the local grid
// barrier the LAST CTA of a rank stores the rank's exact integer partial sums
// into every rank's inbox (plain stores on peer pointers), fences at system
// scope and raises the flags; every CTA of every rank then waits for all flags in
// its OWN memory and reads the reduced value as the integer sum over the source
// slots.  Integer addition is associative, so the result is bit-identical for any
// number of GPUs.  Flags carry a monotonically increasing global phase number.
// ---------------------------------------------------------------------------
constexpr int MAX_RANKS = 8;
constexpr int ACC_SLOTS = 3;
struct CommDev {
  int rank, world;
  unsigned long long epoch;                 // global phase counter at kernel start (same on all ranks)
  u64 *inbox[MAX_RANKS];                    // [ACC_SLOTS][MAX_RANKS][words_per_set] on rank r
  unsigned long long *flags[MAX_RANKS];     // [ACC_SLOTS][MAX_RANKS] on rank r
  int words_per_set;
};

// ghost-plane exchange of a z-slab between neighbouring ranks (row-sharded LOBPCG, lobpcg.cu)
struct PlaneXchg {
  double *lo_dst, *hi_dst;                    // (r-1)'s "from above" region, (r+1)'s "from below" region (or null)
  unsigned long long *lo_flag, *hi_flag;      // the flags to raise there
  const double *from_below, *from_above;      // my receive regions
  const unsigned long long *my_flags;         // [0] from below (rank r-1), [1] from above (rank r+1)
  unsigned long long seq;
  unsigned *counter;                          // CTA completion counter (zeroed before the push)
};

struct RedView {   // reduced word j = sum_r base0[r * stride + j]
  const u64 *base0;
  size_t stride;
  int world;
  __device__ __forceinline__ u64 load(int j) const {
    u64 v = __ldcg(base0 + j);
    for (int r = 1; r < world; ++r) v += __ldcg(base0 + (size_t)r * stride + j);
    return v;
  }
};

// Grid-wide (and, when cm.world > 1, machine-wide) barrier for the persistent
// kernels.  Cooperative launch guarantees co-residency.  `counter` is a
// monotonically increasing arrival counter, `gen` the caller's private
// generation.  On return `view` addresses the reduced copy of set[off, off+count).
// Returns false if the watchdog expired (a peer is gone): callers abandon the
// solve instead of hanging the device.
// `stamps` (optional, thread 0 only): [0] = time this CTA arrived, [1] = released.
__device__ __forceinline__ bool grid_reduce_barrier(unsigned *counter, unsigned &gen, int *abort_flag,
                                                    const CommDev &cm, unsigned long long gphase, u64 *set,
                                                    int off, int count, RedView &view,
                                                    unsigned long long *stamps = nullptr /* shared memory */,
                                                    unsigned leader = 0 /* thread that arrives / spins for the CTA */) {
  __shared__ int s_ok, s_last;
  __syncthreads();
  if (threadIdx.x == leader) {
    gen += 1;
    const unsigned target = gen * gridDim.x;
    if (cm.world > 1) __threadfence_system();   // this CTA's peer stores (halo rows) are ordered before the machine-wide flag
    else __threadfence();
    if (stamps) stamps[0] = globaltimer_ns();
    const unsigned old = atom_add_acqrel_u32(counter, 1u);
    int ok = 1;
    s_last = (old + 1u == target);
    if (cm.world == 1) {
      unsigned spins = 0;
      while (ld_acquire_u32(counter) < target) {
        if (++spins > (1u << 24)) {  // ~ seconds: a peer CTA is gone; bail out
          if (*((volatile int *)abort_flag) || spins > (1u << 25)) { ok = 0; break; }
        }
        if (spins > 64) __nanosleep(64);
      }
      if (stamps) stamps[1] = globaltimer_ns();
      if (!ok) atomicExch(abort_flag, 1);
      __threadfence();
    }
    s_ok = ok;
  }
  __syncthreads();
  view.world = cm.world;
  view.stride = (size_t)cm.words_per_set;
  if (cm.world == 1) {
    view.base0 = set;
    return s_ok != 0;
  }
  const int slot = (int)(gphase % ACC_SLOTS);
  const size_t slot_off = (size_t)(slot * MAX_RANKS) * cm.words_per_set;
  if (s_last) {   // CTA-uniform: this CTA completed the local reduction -> publish it to every rank
    __threadfence();
    // (all loads first: one L2 round trip, then the peer stores stream out back to back)
    constexpr int PUB_MAX = 8;                       // count <= PUB_MAX * blockDim.x (2632 words, >= 384 threads)
    u64 wv[PUB_MAX];
#pragma unroll
    for (int k = 0; k < PUB_MAX; ++k) {
      const int i = threadIdx.x + k * blockDim.x;
      wv[k] = (i < count) ? __ldcg(set + off + i) : 0ull;
    }
    for (int r = 0; r < cm.world; ++r) {
      u64 *dst = cm.inbox[r] + slot_off + (size_t)cm.rank * cm.words_per_set + off;
#pragma unroll
      for (int k = 0; k < PUB_MAX; ++k) {
        const int i = threadIdx.x + k * blockDim.x;
        if (i < count) dst[i] = wv[k];
      }
    }
    // the CTA barrier orders every thread's (weak) peer stores before the flag writers; their
    // st.release.sys then publishes them at system scope (one fence per destination, not per thread)
    __syncthreads();
    if ((int)threadIdx.x < cm.world)
      st_release_sys_u64(cm.flags[threadIdx.x] + slot * MAX_RANKS + cm.rank, gphase + 1ull);
  }
  if ((int)threadIdx.x < cm.world) {
    const unsigned long long *f = cm.flags[cm.rank] + slot * MAX_RANKS + threadIdx.x;
    unsigned spins = 0;
    while (ld_acquire_sys_u64(f) < gphase + 1ull) {
      if (++spins > (1u << 24)) {
        if (*((volatile int *)abort_flag) || spins > (1u << 25)) { atomicExch(abort_flag, 1); break; }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == leader) {
    if (stamps) stamps[1] = globaltimer_ns();
    s_ok = (*((volatile int *)abort_flag) == 0);
  }
  __syncthreads();
  view.base0 = cm.inbox[cm.rank] + slot_off;
  return s_ok != 0;
}

// One 16-byte vector of the content checksum of A (position-dependent 64-bit mix; the sum over the vectors is the
// checksum: order independent, so any grid computes the same value).
__device__ __forceinline__ unsigned long long a_checksum_term(const uint4 v, unsigned long long i) {
  const unsigned long long lo = ((unsigned long long)v.y << 32) | v.x, hi = ((unsigned long long)v.w << 32) | v.z;
  unsigned long long z = (lo ^ (i * 0x9E3779B97F4A7C15ull)) * 0xBF58476D1CE4E5B9ull;
  z ^= z >> 29;
  z += (hi ^ ((i + 0x632BE59BD9B4E019ull) * 0x94D049BB133111EBull)) * 0xD6E8FEB86659FD93ull;
  return z ^ (z >> 31);
}

// ---------------------------------------------------------------------------
// fp64 tensor-core MMA (DMMA): D(8x8) += A(8x4) * B(4x8)
//   a : A[row = lane/4][k = lane%4]
//   b : B[k = lane%4][col = lane/4]
//   c0,c1 : C[row = lane/4][col = 2*(lane%4) + {0,1}]
// ---------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// fp64 tensor-core MMA, large shape: D(16x8) += A(16x16) * B(16x8).  With g = lane / 4, t = lane % 4:
//   a[i] : A[row = g + 8 (i & 1)][k = t + 4 (i >> 1)],  i < 8
//   b[i] : B[k = t + 4 i][col = g],                      i < 4
//   c[i] : C[row = g + 8 (i >> 1)][col = 2 t + (i & 1)], i < 4
// (layout verified by tools/probe/dmma16.cu).  Same throughput per flop as m8n8k4 on this part, but ~400 cycles of
// dependent-issue latency for eight times the work (m8n8k4: ~210 cycles).
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, "
      "{%0,%1,%2,%3};"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
        "d"(b[2]), "d"(b[3]));
}

__device__ __forceinline__ double bf16_bits_to_double(unsigned short b) {
  return (double)__uint_as_float(((unsigned)b) << 16);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming (L2-only) vector accesses: no L1 allocation, no stale-L1 hazard for
// data produced by other CTAs earlier in the same persistent kernel
__device__ __forceinline__ double2 ldcg2(const double *p) {
  return __ldcg(reinterpret_cast<const double2 *>(p));
}
__device__ __forceinline__ void stcg2(double *p, double2 v) {
  __stcg(reinterpret_cast<double2 *>(p), v);
}

}  // namespace ob200
