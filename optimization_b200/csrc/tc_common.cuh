// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers (sm_100a) and the exact integer
// digit-plane scheme used to run the block contraction A * P on the 5th generation
// tensor cores without giving up fp64 accuracy.
//
// Scheme.  A block of A (bf16 storage) is a block-fixed-point matrix:
// A = 2^e_lsb * A', A' integer, |A'| < 2^22.  Its three balanced base-256 digits
//   A' = a2 * 2^16 + a1 * 2^8 + a0,   a_h in [-128, 127]            (int8 planes)
// are precomputed once per operator in the UMMA K-major SWIZZLE_128B shared-memory
// image.  The fp64 tangent tile P (128 x 32) is scaled by the block maximum,
// |P| < 2^E, rounded to the 55-bit signed fixed point F = rint(P * 2^(54-E)) and cut
// into seven balanced int8 digits  F = sum_i d_i 2^(8i).  The tensor cores
// (tcgen05.mma kind::i8, int32 accumulators in TMEM: integer arithmetic, exact)
// form   D_u = sum_{h + i = 8 - u} a_h * d_i ,  u = 0 .. 7   (|D_u| < 2^23),
// the pair (a0, d0) -- weight 2^-64 of the leading one -- being dropped, and
//   A P = 2^(e_lsb + E + 10) * sum_u D_u 2^(-8u)
// is recombined with 64-bit integer adds and one fp64 FMA per element.  The only
// roundings are the quantisation of P (2^-54 of the block maximum, below the fp64
// resolution of that maximum) and the final fp64 rounding.
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
// Same with a watchdog: a hand-off that does not complete within ~2 s is a protocol bug; trap (the launch fails with an
// error) instead of hanging the device.  The fast path costs nothing extra: the timer is read only while spinning.
__device__ __forceinline__ void mbar_wait_guarded(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      if (t1 - t0 > 2000000000ull) __trap();
    }
  }
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
// 16 lanes x 16 consecutive 32-bit columns in the mma.sync accumulator arrangement: thread (m = lane >> 2,
// j = lane & 3) receives   v[0], v[1] = row m,     columns 2j, 2j + 1        v[4], v[5] = row m,     columns 8 + 2j, 9 + 2j
//                          v[2], v[3] = row m + 8, columns 2j, 2j + 1        v[6], v[7] = row m + 8, columns 8 + 2j, 9 + 2j
// (rows relative to the lane field of taddr, which must be a multiple of 16 inside the warp's lane quarter).
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
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
// kind::i8, A = B = signed int8 (K-major), D = s32, M = 128, N = n (multiple of 16, <= 256)
__host__ __device__ constexpr uint32_t idesc_s8_m128(uint32_t n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; single thread
__device__ __forceinline__ void umma_s8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- layout of the operand images -------------------------------------------------------
// One K-major SW128 tile of int8: `rows` rows x 128 elements (128 B per row = the whole K = 128).
// Byte offset of the 16-byte chunk c (k = 16c .. 16c+15) of row r:
__host__ __device__ __forceinline__ uint32_t sw128_chunk_off(uint32_t r, uint32_t c) {
  return r * 128u + ((c ^ (r & 7u)) << 4);
}
// Row order of the A images (the M index of the MMA = TMEM lane of the accumulators).  TMEM lane l = 32 q + 16 g + i
// (q: the lane quarter a warp can read, warp % 4; g: 16-lane group; i < 16) holds block row
//   tc_row_of_lane(l) = 64 g + 16 q + i ,
// so the LOWER half block (rows 0..63) sits in the first 16 lanes of every quarter and the UPPER half in the last 16:
// whichever half (or both) a CTA owns, all its read-back warps take part.
__host__ __device__ __forceinline__ uint32_t tc_row_of_lane(uint32_t l) {
  return 64u * ((l >> 4) & 1u) + 16u * (l >> 5) + (l & 15u);
}
constexpr uint32_t TC_NB = 128;                       // block rows / cols (M and K)
constexpr uint32_t TC_N = 32;                         // p
constexpr int TC_PLANES = 3;
constexpr uint32_t TC_APLANE = TC_NB * 128;           // 16 KB: 128 rows x 128 k (int8)
constexpr uint32_t TC_ABLOCK = TC_PLANES * TC_APLANE; // 48 KB                      [plane h = 2, 1, 0]
constexpr int TC_SLICES = 7;
constexpr uint32_t TC_QTILE = TC_N * 128;             // 4 KB: one digit slice, 32 rows(n) x 128 k
constexpr uint32_t TC_QBYTES = TC_SLICES * TC_QTILE;  // 28 KB: slices stacked along N   [slice s <-> digit i = 6 - s][n]
constexpr int TC_NACC = 8;                            // 8 accumulators x 32 columns = 256 TMEM columns
constexpr uint32_t TC_TMEM_COLS = 256;
constexpr unsigned long long TC_DIGIT_BIAS = 0x0080808080808080ull;

// Issue the MMAs of one block (13 instructions).  The digit slices are stacked along N, so one
// instruction multiplies an A plane with all of them; plane h lands (2 - h) accumulators to the right,
// which realises  D_u = sum_{h+i = 8-u} a_h d_i  in place:
//   plane 2 (most significant) x slices 0..6 -> columns [0, 224)
//   plane 1                    x slices 0..6 -> columns [32, 256)
//   plane 0                    x slices 0..5 -> columns [64, 256)     (a0 x d0 dropped)
// K = 32 per instruction -> 4 k-steps.  First k-step: plane 1 overwrites [32,256), plane 2 overwrites
// [0,32) and accumulates into [32,224), plane 0 accumulates.
__device__ __forceinline__ void issue_block_mmas(uint32_t a_base, uint32_t q_base, uint32_t tmem_base) {
  const uint64_t ad0 = umma_desc_k_sw128(a_base);
  const uint64_t bd0 = umma_desc_k_sw128(q_base);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint64_t a2 = ad0 + (uint64_t)((32 * ks) >> 4);
    const uint64_t a1 = a2 + (uint64_t)(TC_APLANE >> 4);
    const uint64_t a0 = a1 + (uint64_t)(TC_APLANE >> 4);
    const uint64_t bd = bd0 + (uint64_t)((32 * ks) >> 4);
    if (ks == 0) {
      umma_s8(tmem_base + TC_N, a1, bd, idesc_s8_m128(224), 0u);
      umma_s8(tmem_base, a2, bd, idesc_s8_m128(32), 0u);
      umma_s8(tmem_base + TC_N, a2, bd + (uint64_t)(TC_QTILE >> 4), idesc_s8_m128(192), 1u);
    } else {
      umma_s8(tmem_base + TC_N, a1, bd, idesc_s8_m128(224), 1u);
      umma_s8(tmem_base, a2, bd, idesc_s8_m128(224), 1u);
    }
    umma_s8(tmem_base + 2 * TC_N, a0, bd, idesc_s8_m128(192), 1u);
  }
}

// ---------------------------------------------------------------------------------
// Digit slicing of a 128 x 32 fp64 tile into the seven int8 digit images.
// Thread mapping (256 threads): cp = tid & 15 -> columns n = 2cp, 2cp+1 ; g = tid >> 4 ->
// rows k = 8g .. 8g+7, i.e. one 8-byte half chunk per (slice, n).
// `p[i][z]` = P[8g + i][2cp + z]; scale = 2^(54 - E) with |P| < 2^E over the whole tile.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void slice_tile_to_smem(const double (&p)[8][2], double scale, unsigned char *Qsm,
                                                   int tid) {
  const int cp = tid & 15, g = tid >> 4;
  const uint32_t c = g >> 1, hbyte = (g & 1) * 8;
#pragma unroll
  for (int z = 0; z < 2; ++z) {
    const uint32_t n = 2 * cp + z;
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // balanced digits: byte j of ((F + bias) ^ bias) is the int8 digit d_j
      const unsigned long long u =
          ((unsigned long long)__double2ll_rn(p[i][z] * scale) + TC_DIGIT_BIAS) ^ TC_DIGIT_BIAS;
      lo[i] = (uint32_t)u;
      hi[i] = (uint32_t)(u >> 32);
    }
    const uint32_t off = sw128_chunk_off(n, c) + hbyte;
#pragma unroll
    for (int s = 0; s < TC_SLICES; ++s) {
      const int d = 6 - s;                                    // digit index held by slice s
      const uint32_t sel = (d & 3) | (((d & 3) + 4) << 4);    // byte d of a -> pos 0, byte d of b -> pos 1
      uint32_t w[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t x0 = d < 4 ? lo[4 * q] : hi[4 * q], x1 = d < 4 ? lo[4 * q + 1] : hi[4 * q + 1];
        const uint32_t x2 = d < 4 ? lo[4 * q + 2] : hi[4 * q + 2], x3 = d < 4 ? lo[4 * q + 3] : hi[4 * q + 3];
        const uint32_t t01 = __byte_perm(x0, x1, sel), t23 = __byte_perm(x2, x3, sel);
        w[q] = __byte_perm(t01, t23, 0x5410);
      }
      *reinterpret_cast<uint2 *>(Qsm + s * TC_QTILE + off) = make_uint2(w[0], w[1]);
    }
  }
}

// Recombination in the fragment arrangement of tmem_ld_16x256b_x2: out[k] <-> v[k] of that load, i.e. this thread's
// 2 rows x 4 columns of   sum_u D_u 2^(-8u)   over the 16 lanes / 16 columns addressed by taddr (accumulator 0).
__device__ __forceinline__ void recombine_frag16(uint32_t taddr, double (&out)[8]) {
  long long part[2][8];
#pragma unroll
  for (int gq = 0; gq < 2; ++gq) {
    uint32_t v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) tmem_ld_16x256b_x2(taddr + (4 * gq + u) * TC_N, v[u]);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 8; ++c)
      part[gq][c] = (long long)(int)v[3][c] + ((long long)(int)v[2][c] << 8) + ((long long)(int)v[1][c] << 16) +
                    ((long long)(int)v[0][c] << 24);
  }
  // |part| < 2^48: exact int64 -> double through the 2^52 + 2^51 bias (one integer add and one DADD instead of the
  // slow 64-bit I2F conversion)
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double p1 = __longlong_as_double(0x4338000000000000ll + part[1][c]) - 6755399441055744.0;
    const double p0 = __longlong_as_double(0x4338000000000000ll + part[0][c]) - 6755399441055744.0;
    out[c] = fma(p1, 0x1p-32, p0) * 0x1p-24;
  }
}

// Recombination of the eight int32 accumulators of this thread's row:
// out[c] = sum_u D_u[row][col0 + c] 2^(-8u), 16 columns; 64-bit integer adds, one fp64 FMA.
__device__ __forceinline__ void recombine_row16(uint32_t tmem_row_addr /* lane | col0 */, double (&out)[16]) {
#pragma unroll
  for (int hcol = 0; hcol < 2; ++hcol) {
    long long part[2][8];
#pragma unroll
    for (int gq = 0; gq < 2; ++gq) {                          // accumulators 0..3 -> part[0], 4..7 -> part[1]
      uint32_t v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) tmem_ld8(tmem_row_addr + (4 * gq + u) * TC_N + 8 * hcol, v[u]);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c)
        part[gq][c] = (long long)(int)v[3][c] + ((long long)(int)v[2][c] << 8) + ((long long)(int)v[1][c] << 16) +
                      ((long long)(int)v[0][c] << 24);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {   // |part| < 2^48: exact int64 -> double through the 2^52 + 2^51 bias
      const double p1 = __longlong_as_double(0x4338000000000000ll + part[1][c]) - 6755399441055744.0;
      const double p0 = __longlong_as_double(0x4338000000000000ll + part[0][c]) - 6755399441055744.0;
      out[8 * hcol + c] = fma(p1, 0x1p-32, p0) * 0x1p-24;
    }
  }
}

}  // namespace tc
}  // namespace ob200
