// tcgen05 path for the block contraction W = A * P of the Stiefel Hessian
// (reference call site: the user's Hessian functor, IterativeSolvers.h:294).
//   stiefel_planes_kernel : one-time, per operator: bf16 A blocks -> sign-magnitude
//                           byte planes in the UMMA shared-memory image + per-block scale
//   stiefel_ap_tc_kernel  : stand-alone W = A * P (validation of the digit-plane scheme and
//                           building block of ob200_hvp); the persistent tCG kernel uses the
//                           same device functions.
#include "tcg.cuh"
#include "tc_common.cuh"

namespace ob200 {
using namespace tc;

// ---------------------------------------------------------------------------------
// planes[b] : 48 KB = [digit plane a2 | a1 | a0][128 rows x 128 B, SW128]   (int8)
// plane_exp[b] = e_lsb (A = 2^e_lsb * A'), or INT_MIN if the block is not representable
// with |A'| < 2^22 (then the operator stays on the fp64 tensor-core path).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stiefel_planes_kernel(const unsigned short *A, unsigned long long nblk,
                                                              unsigned char *planes, int *plane_exp,
                                                              int *unsupported) {
  __shared__ int s_lsb, s_msb;
  for (unsigned long long b = blockIdx.x; b < nblk; b += gridDim.x) {
    const unsigned short *Ab = A + b * (TC_NB * TC_NB);
    if (threadIdx.x == 0) { s_lsb = 1 << 30; s_msb = -(1 << 30); }
    __syncthreads();
    int lsb = 1 << 30, msb = -(1 << 30);
    bool bad = false;
    for (int i = threadIdx.x; i < TC_NB * TC_NB; i += blockDim.x) {
      const unsigned v = Ab[i];
      const int e = (v >> 7) & 0xff;
      const unsigned frac = v & 0x7f;
      if (e == 0xff) bad = true;                 // inf / nan
      if (e == 0 && frac == 0) continue;         // zero
      const unsigned m = e ? (0x80u | frac) : frac;            // 8-bit significand (subnormal: no hidden bit)
      const int ex = (e ? e : 1) - 127 - 7;                    // value = m * 2^ex
      lsb = min(lsb, ex + (__ffs(m) - 1));
      msb = max(msb, ex + (31 - __clz(m)));
    }
    atomicMin(&s_lsb, lsb);
    atomicMax(&s_msb, msb);
    if (bad) atomicExch(unsupported, 1);
    __syncthreads();
    const int e_lsb = (s_msb < s_lsb) ? 0 : s_lsb;             // all-zero block -> 0
    const bool ok = (s_msb < s_lsb) || (s_msb - s_lsb + 1 <= 22);
    if (threadIdx.x == 0) {
      plane_exp[b] = ok ? e_lsb : (int)0x80000000;
      if (!ok) atomicExch(unsupported, 1);
    }
    unsigned char *Pb = planes + b * (size_t)TC_ABLOCK;
    // one thread per (row, 16-byte chunk): 128 rows x 8 chunks of 16 k
    for (int idx = threadIdx.x; idx < TC_NB * 8; idx += blockDim.x) {
      const int r = idx >> 3, c = idx & 7;                     // image row (TMEM lane) r, k = 16 c .. 16 c + 15
      const int rs = (int)tc_row_of_lane((uint32_t)r);         // block row held by that lane
      uint32_t w2[4], w1[4], w0[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t x2 = 0, x1 = 0, x0 = 0;
#pragma unroll
        for (int z = 0; z < 4; ++z) {
          const unsigned v = Ab[rs * TC_NB + 16 * c + 4 * q + z];
          const int e = (v >> 7) & 0xff;
          const unsigned frac = v & 0x7f;
          int val = 0;
          if (ok && !(e == 0 && frac == 0) && e != 0xff) {
            const unsigned m = e ? (0x80u | frac) : frac;
            const int sh = (e ? e : 1) - 127 - 7 - e_lsb;      // value = m * 2^(sh + e_lsb)
            const unsigned mag = sh >= 0 ? (m << sh) : (m >> (-sh));   // exact, < 2^22
            val = (v & 0x8000u) ? -(int)mag : (int)mag;
          }
          const unsigned u = ((unsigned)val + 0x808080u) ^ 0x808080u;  // balanced base-256 digits in bytes 0..2
          x0 |= (u & 255u) << (8 * z);
          x1 |= ((u >> 8) & 255u) << (8 * z);
          x2 |= ((u >> 16) & 255u) << (8 * z);
        }
        w2[q] = x2; w1[q] = x1; w0[q] = x0;
      }
      const uint32_t off = sw128_chunk_off(r, c);
      *reinterpret_cast<uint4 *>(Pb + off) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
      *reinterpret_cast<uint4 *>(Pb + TC_APLANE + off) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
      *reinterpret_cast<uint4 *>(Pb + 2 * TC_APLANE + off) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
    }
    __syncthreads();
  }
}

// Position-dependent 64-bit checksum of the bf16 blocks of A (wrapping integer sum: order independent).  The
// digit planes are cached per operator; every solve re-validates the cache against this checksum, so an A that
// was updated in place -- or a different A that the allocator placed at the same address -- is never multiplied
// with stale planes.  Reads A once (25.6 MB at n = 1e5: a few microseconds).
__global__ void __launch_bounds__(256) stiefel_checksum_kernel(const uint4 *A16, unsigned long long nvec,
                                                                unsigned long long *out) {
  unsigned long long acc = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    acc += a_checksum_term(__ldg(A16 + i), i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

constexpr size_t APTC_SMEM = 1024 /*align slack*/ + TC_ABLOCK + TC_QBYTES + 256;

extern __shared__ __align__(16) unsigned char tc_smem_raw[];

__global__ void __launch_bounds__(256, 1)
stiefel_ap_tc_kernel(unsigned long long n_rows, const unsigned char *planes, const int *plane_exp, const double *P,
                     double *Wout, int frag_readback) {
  unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *Asm = base;
  unsigned char *Qsm = base + TC_ABLOCK;
  uint64_t *bars = reinterpret_cast<uint64_t *>(base + TC_ABLOCK + TC_QBYTES);   // [0] A landed, [1] MMAs done
  uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(bars + 2);
  double *s_max = reinterpret_cast<double *>(bars + 4);                          // 8 warp maxima
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_holder, TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const unsigned long long nblk = (n_rows + TC_NB - 1) / TC_NB;
  uint32_t parity = 0;
  for (unsigned long long b = blockIdx.x; b < nblk; b += gridDim.x) {
    const unsigned long long r0 = b * TC_NB;
    if (tid == 0) {
      mbar_expect_tx(&bars[0], TC_ABLOCK);
      bulk_g2s(Asm, planes + b * (size_t)TC_ABLOCK, TC_ABLOCK, &bars[0]);
    }
    // load the tile, block maximum
    const int cp = tid & 15, g = tid >> 4;
    double p[8][2];
    double mx = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const unsigned long long grow = r0 + 8 * g + i;
      double2 v = make_double2(0.0, 0.0);
      if (grow < n_rows) v = ldcg2(P + (size_t)grow * TC_N + 2 * cp);
      p[i][0] = v.x; p[i][1] = v.y;
      mx = fmax(mx, fmax(fabs(v.x), fabs(v.y)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_max[warp] = mx;
    __syncthreads();
    mx = s_max[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmax(mx, s_max[w]);
    // |P| < 2^E
    const int E = (mx > 0.0) ? (int)((__double_as_longlong(mx) >> 52) & 0x7ff) - 1023 + 1 : 0;
    slice_tile_to_smem(p, scalbn(1.0, 54 - E), Qsm, tid);
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(&bars[0], parity);
      tc_fence_after();
      issue_block_mmas(smem_u32(Asm), smem_u32(Qsm), tmem_base);
      umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], parity);
    tc_fence_after();
    if (frag_readback) {
      // the read-back of the persistent kernel's M warps: 16-lane groups in the mma.sync accumulator arrangement
      const int q4 = warp & 3, chalf = warp >> 2, m = lane >> 2, j = lane & 3;
      const double sc = scalbn(1.0, plane_exp[b] + E + 10);
#pragma unroll 1
      for (int g16 = 0; g16 < 2; ++g16) {
        double out[8];
        recombine_frag16(tmem_base + ((uint32_t)(32 * q4 + 16 * g16) << 16) + 16 * chalf, out);
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          const int row = 64 * g16 + 16 * q4 + m + ((k & 2) ? 8 : 0);
          const int col = 16 * chalf + ((k & 4) ? 8 : 0) + 2 * j;
          const unsigned long long grow = r0 + row;
          if (grow < n_rows) stcg2(Wout + (size_t)grow * TC_N + col, make_double2(out[k] * sc, out[k + 1] * sc));
        }
      }
    } else {
      const int q4 = warp & 3, chalf = warp >> 2;                 // TMEM lane quarter, column half
      const int row = (int)tc_row_of_lane((uint32_t)(32 * q4 + lane));
      double out[16];
      recombine_row16(tmem_base + ((uint32_t)(32 * q4) << 16) + 16 * chalf, out);
      const double sc = scalbn(1.0, plane_exp[b] + E + 10);
      const unsigned long long grow = r0 + row;
      if (grow < n_rows) {
#pragma unroll
        for (int c = 0; c < 16; c += 2)
          stcg2(Wout + (size_t)grow * TC_N + 16 * chalf + c, make_double2(out[c] * sc, out[c + 1] * sc));
      }
    }
    tc_fence_before();
    __syncthreads();
    parity ^= 1;
  }
  if (warp == 0) tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// ---- host launchers ---------------------------------------------------------------------
cudaError_t launch_stiefel_planes(const unsigned short *A, unsigned long long nblk, unsigned char *planes,
                                  int *plane_exp, int *unsupported, int sm_count, cudaStream_t st) {
  unsigned long long grid = nblk < (unsigned long long)(4 * sm_count) ? nblk : (unsigned long long)(4 * sm_count);
  stiefel_planes_kernel<<<(unsigned)grid, 256, 0, st>>>(A, nblk, planes, plane_exp, unsupported);
  return cudaGetLastError();
}
cudaError_t launch_stiefel_ap_tc(unsigned long long n_rows, const unsigned char *planes, const int *plane_exp,
                                 const double *P, double *Wout, int frag_readback, int grid, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(stiefel_ap_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)APTC_SMEM);
    if (e) return e;
    attr = true;
  }
  stiefel_ap_tc_kernel<<<grid, 256, APTC_SMEM, st>>>(n_rows, planes, plane_exp, P, Wout, frag_readback);
  return cudaGetLastError();
}
cudaError_t launch_stiefel_checksum(const unsigned short *A, unsigned long long nblk, unsigned long long *out,
                                    int sm_count, cudaStream_t st) {
  const unsigned long long nvec = nblk * (TC_NB * TC_NB * sizeof(unsigned short) / sizeof(uint4));
  stiefel_checksum_kernel<<<4 * sm_count, 256, 0, st>>>(reinterpret_cast<const uint4 *>(A), nvec, out);
  return cudaGetLastError();
}
size_t stiefel_planes_bytes(unsigned long long nblk) { return (size_t)nblk * TC_ABLOCK; }

}  // namespace ob200
