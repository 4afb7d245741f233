// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers (sm_100a) and the exact
// bf16 digit-plane scheme used to run the block contraction A * P on the 5th
// generation tensor cores without giving up fp64 accuracy.
//
// Scheme.  A block of A (bf16) is a block-fixed-point matrix:
// A = sa * A', A' integer, |A'| < 2^16, sa = 2^e_lsb.  Sign-magnitude byte planes
//   A' = sgn * (hi * 256 + lo),   hi, lo in [0, 255]      (both exact in bf16)
// are precomputed once per operator in the UMMA K-major SWIZZLE_128B shared-memory
// image.  The fp64 tangent tile P (128 x 32) is scaled by the block maximum,
// |P| < 2^E, rounded to a 56-bit fixed-point magnitude F and cut into seven
// sign-magnitude 8-bit digits  F = sum_t d_t 2^(48-8t)  (each exact in bf16).
// Every product plane_h x digit_t is < 2^16 and a K = 128 sum of them < 2^23, so
// the fp32 accumulation in TMEM is EXACT; the eight accumulators
//   D_u = hi * digit_u + lo * digit_(u-1)        (two pairs each: < 2^24, exact)
// are recombined in fp64:  (A P) = 2^(e_lsb + E) * sum_u D_u 2^(-8u).
// The only rounding is the 2^(E-56) quantisation of P (below fp64 resolution of
// the block maximum) and the final fp64 Horner sum.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ob200 {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

// ---- bulk async copy global -> shared (TMA 1-D; SASS UBLKCP) ---------------------
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMEM ------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *holder_smem, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {         // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 16 consecutive 32-bit columns: thread = TMEM lane (accumulator row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors ---------------------------------------------------------------
// K-major, SWIZZLE_128B, 16-bit elements: rows of 64 elements (128 B), 8-row groups 1024 B
// apart (SBO), LBO = 1 (unused for swizzled K-major), descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16, A = B = bf16 (K-major), D = f32, M = 128, N = n (multiple of 16, <= 256)
__host__ __device__ constexpr uint32_t idesc_bf16_m128(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; single thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- layout of the operand images -------------------------------------------------------
// One K-major SW128 tile: `rows` rows x 64 bf16 (128 B per row).  Byte offset of the
// 16-byte chunk c (8 elements: k = 8c .. 8c+7) of row r:
__host__ __device__ __forceinline__ uint32_t sw128_chunk_off(uint32_t r, uint32_t c) {
  return r * 128u + ((c ^ (r & 7u)) << 4);
}
constexpr uint32_t TC_NB = 128;                       // block rows / cols (M and K)
constexpr uint32_t TC_N = 32;                         // p
constexpr uint32_t TC_ATILE = TC_NB * 128;            // 16 KB: 128 rows x 64 k
constexpr uint32_t TC_APLANE = 2 * TC_ATILE;          // 32 KB: K = 128 -> 2 k-blocks
constexpr uint32_t TC_ABLOCK = 2 * TC_APLANE;         // 64 KB: hi + lo planes   [plane][kb]
constexpr uint32_t TC_QTILE = TC_N * 128;             // 4 KB: one digit slice, 32 rows(n) x 64 k
constexpr int TC_SLICES = 7;
constexpr uint32_t TC_QKB = TC_SLICES * TC_QTILE;     // 28 KB: the 7 slices stacked along N (224 rows) for one k-block
constexpr uint32_t TC_QBYTES = 2 * TC_QKB;            // 56 KB                     [kb][slice][n]
constexpr int TC_NACC = TC_SLICES + 1;                // 8 accumulators x 32 columns = 256 TMEM columns
constexpr uint32_t TC_TMEM_COLS = 256;

// bf16 bit pattern of the integer d in [0, 255] with sign bit sgn (0 / 0x8000): exact
__device__ __forceinline__ uint32_t bf16_of_u8(uint32_t d, uint32_t sgn) {
  const float f = __uint_as_float(0x4B000000u | d) - 8388608.0f;   // (float)d, exact
  return (__float_as_uint(f) >> 16) | sgn;
}

// Issue the MMAs of one block.  The seven digit slices are stacked along N, so one instruction
// multiplies an A plane with all of them (N = 224: the 4 KB A tile is read from shared memory once per
// 224 columns instead of once per 32).  The lo plane accumulates one slice to the right of the hi
// plane (columns [32, 256) vs [0, 224)), which realises  D_u = hi * d_u + lo * d_(u-1)  in place:
//   first k-step : lo -> [32,256) (overwrite) ; hi -> [0,32) (overwrite) and [32,224) (accumulate)
//   other k-steps: lo -> [32,256), hi -> [0,224), accumulate.                       17 instructions.
__device__ __forceinline__ void issue_block_mmas(uint32_t a_base, uint32_t q_base, uint32_t tmem_base) {
  const uint64_t ad0 = umma_desc_k_sw128(a_base);
  const uint64_t bd0 = umma_desc_k_sw128(q_base);
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
    for (int k16 = 0; k16 < 4; ++k16) {
      const uint64_t a_hi = ad0 + (uint64_t)((kb * TC_ATILE + 32 * k16) >> 4);
      const uint64_t a_lo = a_hi + (uint64_t)(TC_APLANE >> 4);
      const uint64_t bd = bd0 + (uint64_t)((kb * TC_QKB + 32 * k16) >> 4);
      if (kb == 0 && k16 == 0) {
        umma_bf16(tmem_base + TC_N, a_lo, bd, idesc_bf16_m128(224), 0u);
        umma_bf16(tmem_base, a_hi, bd, idesc_bf16_m128(32), 0u);
        umma_bf16(tmem_base + TC_N, a_hi, bd + (uint64_t)(TC_QTILE >> 4), idesc_bf16_m128(192), 1u);
      } else {
        umma_bf16(tmem_base + TC_N, a_lo, bd, idesc_bf16_m128(224), 1u);
        umma_bf16(tmem_base, a_hi, bd, idesc_bf16_m128(224), 1u);
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// Digit slicing of a 128 x 32 fp64 tile into the seven bf16 digit images.
// Thread mapping (256 threads): cp = tid & 15 -> columns n = 2cp, 2cp+1 ; g = tid >> 4 ->
// rows k = 8g .. 8g+7, i.e. exactly one 16-byte chunk per (slice, n).
// `p[i][z]` = P[8g + i][2cp + z]; scale = 2^(56 - E) with |P| < 2^E over the whole tile.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void slice_tile_to_smem(const double (&p)[8][2], double scale, unsigned char *Qsm,
                                                   int tid) {
  const int cp = tid & 15, g = tid >> 4;
  const int kb = g >> 3, c = g & 7;
#pragma unroll
  for (int z = 0; z < 2; ++z) {
    const int n = 2 * cp + z;
    unsigned long long F[8];
    uint32_t sg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      F[i] = (unsigned long long)__double2ll_rn(fabs(p[i][z]) * scale);     // < 2^56
      sg[i] = (uint32_t)(((unsigned long long)__double_as_longlong(p[i][z])) >> 63) << 15;
    }
    const uint32_t off = kb * TC_QKB + sw128_chunk_off(n, c);
#pragma unroll
    for (int t = 0; t < TC_SLICES; ++t) {
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t d0 = (uint32_t)(F[2 * q] >> (48 - 8 * t)) & 255u;
        const uint32_t d1 = (uint32_t)(F[2 * q + 1] >> (48 - 8 * t)) & 255u;
        w[q] = bf16_of_u8(d0, sg[2 * q]) | (bf16_of_u8(d1, sg[2 * q + 1]) << 16);
      }
      *reinterpret_cast<uint4 *>(Qsm + t * TC_QTILE + off) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// fp64 recombination of the eight fp32 accumulators of this thread's row:
// out[c] = sum_u D_u[row][col0 + c] 2^(-8u)   (Horner, smallest first), 16 columns.
// The eight TMEM loads of an 8-column group are issued back to back and waited for once.
__device__ __forceinline__ void recombine_row16(uint32_t tmem_row_addr /* lane | col0 */, double (&out)[16]) {
#pragma unroll
  for (int hcol = 0; hcol < 2; ++hcol) {
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.0;
#pragma unroll
    for (int ub = TC_NACC - 4; ub >= 0; ub -= 4) {          // accumulators 7..4, then 3..0
      uint32_t v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) tmem_ld8(tmem_row_addr + (ub + u) * TC_N + 8 * hcol, v[u]);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int u = 3; u >= 0; --u) acc[c] = fma(acc[c], 0x1p-8, (double)__uint_as_float(v[u][c]));
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) out[8 * hcol + c] = acc[c];
  }
}

}  // namespace tc
}  // namespace ob200
