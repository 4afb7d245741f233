// C ABI of liboptimization_b200.so (see include/optimization_b200.h).
// Host-side dispatch only: every arithmetic statement of the hot path runs in
// the CUDA kernels of tcg_elementwise.cu / tcg_stiefel.cu / level1.cu.  There is
// no CPU fallback: without a usable GPU ob200_create fails.
#include "../../include/optimization_b200.h"
#include "tcg.cuh"

#include <cusolverDn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace ob200 {
cudaError_t launch_tcg_exchange(const CommDev &cm, unsigned long long gphase, u64 *set, int off, int count,
                                int mode, int *abort_flag, cudaStream_t st);
cudaError_t launch_tcg_init(const TcgCommon &a, int grid, cudaStream_t st);
cudaError_t launch_tcg_finalize(const u64 *acc, int slot, double *out, cudaStream_t st);
cudaError_t launch_tcg_diag(const TcgCommon &a, const double *hdiag, int grid, cudaStream_t st);
cudaError_t launch_tcg_stiefel(const TcgCommon &a, unsigned long long n_rows, const unsigned short *A,
                               const double *Y, const double *S_dev, double op_norm_bound, int grid,
                               cudaStream_t stm);
cudaError_t launch_stiefel_apply(unsigned long long n_rows, const unsigned short *A, const double *V,
                                 const double *S_dev, const double *Y, double *Wout, u64 *set, double inv_q,
                                 int grid, cudaStream_t stm);
cudaError_t launch_stiefel_gram(unsigned long long n_rows, const double *X, const double *Z, u64 *set,
                                double inv_q, int grid, cudaStream_t stm);
cudaError_t launch_stiefel_rowgemm(unsigned long long n_rows, const double *W, double cW, const double *X,
                                   const double *M_dev, double *out, int grid, cudaStream_t stm);
cudaError_t launch_stiefel_absrowsum(const unsigned short *A, unsigned long long nrows_padded,
                                     unsigned long long *out_bits, cudaStream_t stm);
int gram_exponent_host(double bound);
cudaError_t launch_stiefel_planes(const unsigned short *A, unsigned long long nblk, unsigned char *planes,
                                  int *plane_exp, int *unsupported, int sm_count, cudaStream_t st);
cudaError_t launch_stiefel_ap_tc(unsigned long long n_rows, const unsigned char *planes, const int *plane_exp,
                                 const double *P, double *Wout, int frag_readback, int grid, cudaStream_t st);
size_t stiefel_planes_bytes(unsigned long long nblk);
cudaError_t launch_stiefel_checksum(const unsigned short *A, unsigned long long nblk, unsigned long long *out,
                                    int sm_count, cudaStream_t st);
cudaError_t launch_tcg_stiefel_tc(const TcgCommon &a, unsigned long long n_rows, const unsigned short *A,
                                  const double *Y, const double *S_dev, double op_norm_bound,
                                  const unsigned char *planes, const int *plane_exp, int grid, cudaStream_t stm,
                                  int hvp_mode = 0, const unsigned long long *planes_sum_dev = nullptr,
                                  unsigned long long planes_sum_expected = 0);
cudaError_t launch_tcg_stiefel_v6(const TcgCommon &a, unsigned long long n_rows, const unsigned short *A,
                                  const double *Y, const double *S_dev, double op_norm_bound,
                                  const unsigned char *planes, const int *plane_exp, int sm_count, cudaStream_t stm);
cudaError_t launch_tcg_sphere(const TcgCommon &a, const double *d, const double *Ut, unsigned long long ldu, const double *x,
                              const double *w, const double *sigma_host, int k, double lambda, int grid, cudaStream_t st);
cudaError_t launch_sphere_tdot(unsigned long long N, const double *Ut, unsigned long long ldu, const double *sigma_host, int k,
                               const double *v, u64 *set, int sm_count, cudaStream_t st);
cudaError_t launch_sphere_scale_t(const u64 *set, const double *sigma_host, int k, double *st_dev, cudaStream_t st);
cudaError_t launch_sphere_apply(unsigned long long N, const double *d, const double *Ut, unsigned long long ldu, int k,
                                const double *st_dev, const double *v, double *out, int sm_count, cudaStream_t st);
cudaError_t launch_sphere_combine(unsigned long long N, const double *Av, const double *x, const double *v, double c,
                                  double lambda, double *out, int sm_count, cudaStream_t st);
struct SparseArgs {
  int kind;
  int r;
  unsigned long long units;
  const unsigned long long *rowptr;
  const unsigned *colidx;
  const double *blocks, *lambda, *X;
  unsigned gx, gy, gz;
  // row-sharded CSR3: halo exchange of p
  unsigned long long n_halo;
  const double *halo;                        // this rank's halo buffer (n_halo poses x 3 r)
  const unsigned *send_idx;
  unsigned long long send_ptr[MAX_RANKS + 1];
  double *peer_halo[MAX_RANKS];              // rank q's halo buffer, already offset to where this rank's rows go
};
cudaError_t launch_tcg_sparse(const TcgCommon &a, const SparseArgs &sp, int sm_count, cudaStream_t st);
cudaError_t launch_sparse_apply(unsigned long long N, const SparseArgs &sp, const double *V, double *out, int sm_count,
                                cudaStream_t st);
cudaError_t launch_csr3_model(const SparseArgs &sp, const double *X, double *lambda_out, double *grad, u64 *set,
                              int sm_count, cudaStream_t st);
cudaError_t launch_csr3_retract(unsigned long long N, int r, const double *X, const double *V, double *out, int *bad,
                                cudaStream_t st);
cudaError_t launch_blk_diag(unsigned long long m, int k, const double *d, double alpha, const double *in, int ldi, double *out,
                            int ldo, int sm_count, cudaStream_t st);
cudaError_t launch_blk_stencil7(unsigned gx, unsigned gy, unsigned gz, int k, const double *in, int ldi, double *out, int ldo,
                                int sm_count, cudaStream_t st);
cudaError_t launch_blk_stencil7_slab(unsigned gx, unsigned gy, unsigned gz, int k, const double *in, int ldi, double *out, int ldo,
                                     int has_lo, int has_hi, int sm_count, cudaStream_t st);
cudaError_t launch_lob_allreduce(const CommDev &cm, unsigned long long gphase, double *const *peer_region, size_t slot_stride,
                                 double *buf, int count, int *abort_flag, cudaStream_t st);
cudaError_t launch_lob_plane_exchange(const PlaneXchg &px, double *interior, int ld, int k, unsigned long long plane_rows,
                                      unsigned long long planes, int has_lo, int has_hi, int *abort_flag, int sm_count,
                                      cudaStream_t st);
cudaError_t launch_blk_gram(unsigned long long m, const double *A, int lda, int k1, const double *B, int ldb, int k2,
                            double *partial, int nb, double *G, cudaStream_t st);
cudaError_t launch_blk_gemm(unsigned long long m, const double *S, int lds, int k, const double *C, int ldc, int n2, double *out,
                            int ldo, int nb, cudaStream_t st);
cudaError_t launch_blk_update(unsigned long long m, const double *S, int lds, int ns, int nx, const double *C, int ldc, double *Xo,
                              int ldx, double *Po, int ldp, int nb, cudaStream_t st);
cudaError_t launch_blk_residual(unsigned long long m, int nx, const double *AX, const double *BX, int ldb, const double *X,
                                int ldx, const double *theta, double *R, double *partial, int nb, double *norms2,
                                cudaStream_t st);
cudaError_t launch_blk_sumsq(unsigned long long total, const double *V, double *partial, int nb, double *out, cudaStream_t st);
cudaError_t launch_rr_equilibrate(int ns, const double *GA, const double *GB, double *EA, double *EB, double *D,
                                  cudaStream_t st);
cudaError_t launch_rr_scale_transpose(int ns, const double *Z, const double *D, double *C, cudaStream_t st);
cudaError_t launch_dots(unsigned long long N, int count, const double *const *a, const double *const *b,
                        u64 *set, int sm_count, cudaStream_t st);
cudaError_t launch_finalize_many(const u64 *set, int count, double *out, cudaStream_t st);
cudaError_t launch_hvp_prologue(unsigned long long N, const double *V, const unsigned short *A, unsigned long long nvec,
                                u64 *vv_acc, unsigned long long *sum, u64 *vv_acc_next, unsigned long long *sum_next,
                                u64 *zero_words, unsigned long long n_zero, unsigned *barrier_words,
                                unsigned *done_counter, double *vv_out, int sm_count, cudaStream_t st);
cudaError_t launch_axpby(unsigned long long N, double alpha, const double *x, double beta, const double *y,
                         double *out, int sm_count, cudaStream_t st);
cudaError_t launch_hadamard(unsigned long long N, const double *d, const double *x, double *out,
                            int sm_count, cudaStream_t st);
cudaError_t launch_div(unsigned long long N, const double *x, double a, double *out, int sm_count, cudaStream_t st);
}  // namespace ob200

using namespace ob200;

struct ob200_context {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  uint64_t launches = 0;
  cusolverDnHandle_t solver = nullptr;   // LOBPCG: dense generalised eigensolve of the Rayleigh-Ritz pencil (library)
  unsigned char *lob_ws = nullptr;       // LOBPCG workspace slab (grow-only, reused across calls)
  size_t lob_cap = 0;
  // workspace
  size_t vec_capacity = 0;        // doubles per work vector
  double *r = nullptr, *p0 = nullptr, *p1 = nullptr, *Hp = nullptr, *gs = nullptr; // gs: staging for g/s (host entry)
  double *gen_vec = nullptr;      // the unfused loop's own r | p | Hp | v | scratch (stpcg_generic)
  size_t gen_capacity = 0;
  size_t gs_capacity = 0;
  u64 *acc = nullptr;             // ACC_SETS * ACC_WORDS
  unsigned *barrier = nullptr;    // [0] counter, [1] abort flag (int)
  TcgDeviceResult *dres = nullptr;
  double *dscal = nullptr;        // 8 doubles
  double *dmat = nullptr;         // 2 * 32*32 doubles (S, M)
  double *drot = nullptr;         // v6 kernel: Q | Q^T | lambda (eigen-decomposition of S), 3 * 32*32 doubles
  double *yrot = nullptr;         // v6 kernel: Y Q (n x 32)
  size_t yrot_capacity = 0;
  bool rot_valid = false;         // (drot, yrot) belong to the point (rot_Y, rot_S)
  const double *rot_Y = nullptr;
  uint64_t rot_n = 0;
  double rot_S[1024];
  unsigned long long *dbits = nullptr;
  // pinned host mirrors
  TcgDeviceResult *hres = nullptr;
  double *hscal = nullptr;
  u64 *hacc = nullptr;            // ACC_WORDS
  double *hmat = nullptr;         // 32*32
  // multi-GPU
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  unsigned long long *dbg = nullptr;   // per-CTA phase timers (ob200_debug_phase_times)
  // tcgen05 operand cache: digit planes of the last block-diagonal A seen (keyed by pointer + size)
  unsigned char *planes = nullptr;
  int *plane_exp = nullptr;       // [nblk] + [nblk] = unsupported flag
  const void *planes_key = nullptr;
  uint64_t planes_n = 0;
  size_t planes_cap = 0;
  bool planes_ok = false;
  double S_cache[1024];                    // last S uploaded to dmat[0..1023] (skip the pageable H2D copy when unchanged)
  bool S_cache_valid = false;
  unsigned long long *blk_stats = nullptr; // per-block maxima of r / p for the v6 Stiefel kernel (10 words per block)
  size_t blk_stats_cap = 0;
  unsigned long long planes_sum = 0;       // checksum of the A the planes were built from
  unsigned long long *dsum = nullptr;      // device / pinned-host word for the per-solve checksum of A
  u64 *hvp_ws = nullptr;                   // stand-alone HVP prologue: 2 x (KUL_STRIDE accumulator words + checksum word) + counter
  int hvp_par = 0;                         // which copy the next call accumulates into
  unsigned long long *hsum = nullptr;
  int opt_tcgen05 = 1;            // 1: tcgen05 kernel v6 (warp-specialised register budgets, eigenbasis of S), 2: tcgen05 kernel v4, 0: fp64 MMA kernel
  int last_path = 0;              // 1 = tcgen05 kernel, 0 = fp64 tensor-core kernel
  // multi-GPU exchange (CUDA IPC peer memory)
  CommDev cm;                     // rank, world, epoch, peer pointers
  u64 *comm_buf = nullptr;        // own inbox + flags
  void *peer_base[MAX_RANKS] = {nullptr};
  unsigned long long plane_seq = 0;   // ghost-plane exchanges issued so far (row-sharded LOBPCG; same on every rank)
  unsigned long long ar_seq = 0;      // all-reduces through the halo buffer issued so far
  double *halo_buf = nullptr;     // halo buffer of a row-sharded sparse operator (peers store into it)
  size_t halo_bytes = 0;
  void *halo_peer[MAX_RANKS] = {nullptr};
};

static const size_t COMM_INBOX_WORDS = (size_t)ACC_SLOTS * MAX_RANKS * ACC_WORDS;
static const size_t COMM_FLAG_WORDS = (size_t)ACC_SLOTS * MAX_RANKS;

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                   \
      return OB200_CUDA_ERROR;                                                          \
    }                                                                                   \
  } while (0)

static int fail(ob200_context *ctx, int code, const char *msg) {
  if (ctx) ctx->err = msg;
  return code;
}

extern "C" {

int ob200_version(void) { return 100; }

int ob200_create(int device, void *stream, ob200_context **out) {
  if (!out) return OB200_INVALID_ARGUMENT;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    fprintf(stderr, "optimization_b200: no usable CUDA device %d (%s); there is no CPU fallback\n", device,
            e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range");
    return OB200_CUDA_ERROR;
  }
  ob200_context *ctx = new ob200_context;
  ctx->device = device;
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  if (!prop.cooperativeLaunch) { delete ctx; return OB200_UNSUPPORTED; }
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  CK(cudaMalloc(&ctx->acc, sizeof(u64) * ACC_SETS * ACC_WORDS));
  CK(cudaMalloc(&ctx->barrier, 64));
  CK(cudaMalloc(&ctx->dres, sizeof(TcgDeviceResult)));
  CK(cudaMalloc(&ctx->dscal, sizeof(double) * 8));
  CK(cudaMalloc(&ctx->dmat, sizeof(double) * 2 * 32 * 32));
  CK(cudaMalloc(&ctx->drot, sizeof(double) * 3 * 32 * 32));
  CK(cudaMalloc(&ctx->dbits, 64));
  CK(cudaMalloc(&ctx->dsum, 8));
  CK(cudaMalloc(&ctx->hvp_ws, sizeof(u64) * (2 * (KUL_STRIDE + 1) + 1)));
  CK(cudaMemset(ctx->hvp_ws, 0, sizeof(u64) * (2 * (KUL_STRIDE + 1) + 1)));
  CK(cudaMallocHost(&ctx->hsum, 8));
  CK(cudaMallocHost(&ctx->hres, sizeof(TcgDeviceResult)));
  CK(cudaMallocHost(&ctx->hscal, sizeof(double) * 8));
  CK(cudaMallocHost(&ctx->hacc, sizeof(u64) * ACC_WORDS));
  CK(cudaMallocHost(&ctx->hmat, sizeof(double) * 32 * 32));
  CK(cudaEventCreate(&ctx->ev0));
  CK(cudaEventCreate(&ctx->ev1));
  memset(&ctx->cm, 0, sizeof(ctx->cm));
  ctx->cm.world = 1;
  ctx->cm.words_per_set = ACC_WORDS;
  *out = ctx;
  return OB200_OK;
}

int ob200_destroy(ob200_context *ctx) {
  if (!ctx) return OB200_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->r); cudaFree(ctx->p0); cudaFree(ctx->p1); cudaFree(ctx->Hp); cudaFree(ctx->gs);
  cudaFree(ctx->gen_vec);
  cudaFree(ctx->acc); cudaFree(ctx->barrier); cudaFree(ctx->dres); cudaFree(ctx->dscal);
  cudaFree(ctx->drot); cudaFree(ctx->yrot);
  cudaFree(ctx->dmat); cudaFree(ctx->dbits); cudaFree(ctx->dsum); cudaFree(ctx->hvp_ws); cudaFreeHost(ctx->hsum);
  cudaFreeHost(ctx->hres); cudaFreeHost(ctx->hscal); cudaFreeHost(ctx->hacc); cudaFreeHost(ctx->hmat);
  cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
  for (int r = 0; r < MAX_RANKS; ++r)
    if (ctx->peer_base[r]) cudaIpcCloseMemHandle(ctx->peer_base[r]);
  cudaFree(ctx->comm_buf);
  for (int r = 0; r < MAX_RANKS; ++r)
    if (ctx->halo_peer[r] && r != ctx->cm.rank) cudaIpcCloseMemHandle(ctx->halo_peer[r]);
  cudaFree(ctx->halo_buf);
  cudaFree(ctx->dbg);
  cudaFree(ctx->planes);
  cudaFree(ctx->plane_exp);
  cudaFree(ctx->blk_stats);
  if (ctx->solver) cusolverDnDestroy(ctx->solver);
  cudaFree(ctx->lob_ws);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return OB200_OK;
}

const char *ob200_last_error(const ob200_context *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int ob200_sm_count(const ob200_context *ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t ob200_kernel_launches(const ob200_context *ctx) { return ctx ? ctx->launches : 0; }

int ob200_synchronize(ob200_context *ctx) {
  if (!ctx) return OB200_INVALID_ARGUMENT;
  CK(cudaStreamSynchronize(ctx->stream));
  return OB200_OK;
}

int ob200_debug_phase_times(ob200_context *ctx, int enable, uint64_t *out4_max, uint64_t *out4_min) {
  if (!ctx) return OB200_INVALID_ARGUMENT;
  const size_t words = 4 * 1024 + 256;
  if (enable && !ctx->dbg) {
    CK(cudaMalloc(&ctx->dbg, words * 8));
    CK(cudaMemset(ctx->dbg, 0, words * 8));
  }
  if (ctx->dbg && out4_max && out4_min) {
    std::vector<unsigned long long> h(words);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(h.data(), ctx->dbg, words * 8, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 4; ++k) { out4_max[k] = 0; out4_min[k] = ~0ull; }
    for (size_t b = 0; b < 1024; ++b) {
      if (!(h[4 * b] | h[4 * b + 1] | h[4 * b + 2] | h[4 * b + 3])) continue;
      for (int k = 0; k < 4; ++k) {
        if (h[4 * b + k] > out4_max[k]) out4_max[k] = h[4 * b + k];
        if (h[4 * b + k] < out4_min[k]) out4_min[k] = h[4 * b + k];
      }
    }
    if (getenv("OB200_TIMELINE")) {
      if (getenv("OB200_TIMELINE")[0] == '5') {   // v6 kernel: all 48 stamps relative to the end of the previous iteration's phase A loop start
        unsigned long long t0 = ~0ull;
        for (int k = 0; k < 192; ++k) if (h[4096 + k] && h[4096 + k] < t0) t0 = h[4096 + k];
        fprintf(stderr, "v6 timeline (ns rel. to the earliest stamp):");
        for (int k = 0; k < 64; ++k) if (h[4096 + k]) fprintf(stderr, " [%d]%lld", k, (long long)(h[4096 + k] - t0));
        fprintf(stderr, "\nper block [RP_FULL, L done, M frags, MMA done, M read back, M done, G start, G done]:\n");
        for (int b = 0; b < 8; ++b) {
          fprintf(stderr, "  blk %d:", b);
          for (int e = 0; e < 8; ++e) fprintf(stderr, " %6lld", h[4096 + 64 + 8 * b + e] ? (long long)(h[4096 + 64 + 8 * b + e] - t0) : -1ll);
          fprintf(stderr, "\n");
        }
      }
      fprintf(stderr, "timeline (ns rel. to L start):");
      for (int k = 0; k < 18; ++k) fprintf(stderr, " [%d]%lld", k, (long long)(h[4096 + k] - h[4096]));
      fprintf(stderr, "\nphase B timeline (ns rel. to barrier-A release):");
      for (int k = 0; k < 14; ++k) fprintf(stderr, " [%d]%lld", k, (long long)(h[4096 + 20 + k] - h[4096 + 20]));
      fprintf(stderr, "\n");
    }
    CK(cudaMemset(ctx->dbg, 0, words * 8));
  }
  if (!enable && ctx->dbg) { cudaFree(ctx->dbg); ctx->dbg = nullptr; }
  return OB200_OK;
}

int ob200_set_option(ob200_context *ctx, const char *name, int value) {
  if (!ctx || !name) return OB200_INVALID_ARGUMENT;
  if (!strcmp(name, "tcgen05")) { ctx->opt_tcgen05 = value; return OB200_OK; }
  return fail(ctx, OB200_INVALID_ARGUMENT, "unknown option");
}
int ob200_last_path(const ob200_context *ctx) { return ctx ? ctx->last_path : -1; }

int ob200_comm_export(ob200_context *ctx, void *handle_out) {
  if (!ctx || !handle_out) return OB200_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->comm_buf) {
    CK(cudaMalloc(&ctx->comm_buf, sizeof(u64) * (COMM_INBOX_WORDS + COMM_FLAG_WORDS)));
    CK(cudaMemset(ctx->comm_buf, 0, sizeof(u64) * (COMM_INBOX_WORDS + COMM_FLAG_WORDS)));
    CK(cudaDeviceSynchronize());
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == OB200_COMM_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ctx->comm_buf));
  memcpy(handle_out, &h, sizeof(h));
  return OB200_OK;
}

int ob200_comm_connect(ob200_context *ctx, int rank, int world, const void *handles) {
  if (!ctx || !handles || world < 1 || world > MAX_RANKS || rank < 0 || rank >= world)
    return fail(ctx, OB200_INVALID_ARGUMENT, "comm_connect: bad rank / world (max 8 ranks)");
  if (!ctx->comm_buf) return fail(ctx, OB200_INVALID_ARGUMENT, "comm_connect before comm_export");
  CK(cudaSetDevice(ctx->device));
  for (int r = 0; r < world; ++r) {
    u64 *base;
    if (r == rank) {
      base = ctx->comm_buf;
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char *)handles + (size_t)r * OB200_COMM_HANDLE_BYTES, sizeof(h));
      void *p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      ctx->peer_base[r] = p;
      base = (u64 *)p;
    }
    ctx->cm.inbox[r] = base;
    ctx->cm.flags[r] = (unsigned long long *)(base + COMM_INBOX_WORDS);
  }
  ctx->cm.rank = rank;
  ctx->cm.world = world;
  ctx->cm.epoch = 0;
  return OB200_OK;
}

int ob200_halo_create(ob200_context *ctx, uint64_t bytes, void *handle_out) {
  if (!ctx || !handle_out) return OB200_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  for (int r = 0; r < MAX_RANKS; ++r) {
    if (ctx->halo_peer[r] && r != ctx->cm.rank) cudaIpcCloseMemHandle(ctx->halo_peer[r]);
    ctx->halo_peer[r] = nullptr;
  }
  cudaFree(ctx->halo_buf);
  ctx->halo_buf = nullptr;
  ctx->halo_bytes = 0;
  ctx->plane_seq = 0;
  ctx->ar_seq = 0;
  CK(cudaMalloc(&ctx->halo_buf, bytes ? bytes : 256));
  CK(cudaMemset(ctx->halo_buf, 0, bytes ? bytes : 256));
  CK(cudaDeviceSynchronize());
  ctx->halo_bytes = bytes;
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ctx->halo_buf));
  memcpy(handle_out, &h, sizeof(h));
  return OB200_OK;
}
int ob200_halo_connect(ob200_context *ctx, const void *handles) {
  if (!ctx || !handles) return OB200_INVALID_ARGUMENT;
  if (!ctx->halo_buf) return fail(ctx, OB200_INVALID_ARGUMENT, "halo_connect before halo_create");
  CK(cudaSetDevice(ctx->device));
  for (int r = 0; r < ctx->cm.world; ++r) {
    if (r == ctx->cm.rank) { ctx->halo_peer[r] = ctx->halo_buf; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + (size_t)r * OB200_COMM_HANDLE_BYTES, sizeof(h));
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->halo_peer[r] = p;
  }
  return OB200_OK;
}

int ob200_comm_rank(const ob200_context *ctx) { return ctx ? ctx->cm.rank : 0; }
int ob200_comm_world(const ob200_context *ctx) { return ctx ? ctx->cm.world : 1; }

// ---- memory helpers ---------------------------------------------------------
int ob200_malloc(ob200_context *ctx, size_t bytes, void **p) {
  if (!ctx || !p) return OB200_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMalloc(p, bytes ? bytes : 16));
  return OB200_OK;
}
int ob200_free(ob200_context *ctx, void *p) {
  if (!ctx) return OB200_INVALID_ARGUMENT;
  if (p && p == ctx->planes_key) { ctx->planes_key = nullptr; ctx->planes_ok = false; }   // cached digit planes die with their A
  if (p && p == ctx->rot_Y) { ctx->rot_valid = false; ctx->rot_Y = nullptr; }              // so does the rotated copy of a freed Y
  CK(cudaFree(p));
  return OB200_OK;
}
int ob200_memcpy_h2d(ob200_context *ctx, void *dst, const void *src, size_t bytes) {
  if (!ctx) return OB200_INVALID_ARGUMENT;
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return OB200_OK;
}
int ob200_memcpy_d2h(ob200_context *ctx, void *dst, const void *src, size_t bytes) {
  if (!ctx) return OB200_INVALID_ARGUMENT;
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return OB200_OK;
}
int ob200_malloc_host(ob200_context *ctx, size_t bytes, void **p) {
  if (!ctx || !p) return OB200_INVALID_ARGUMENT;
  CK(cudaMallocHost(p, bytes ? bytes : 16));
  return OB200_OK;
}
int ob200_free_host(ob200_context *ctx, void *p) {
  if (!ctx) return OB200_INVALID_ARGUMENT;
  CK(cudaFreeHost(p));
  return OB200_OK;
}

}  // extern "C"

static void sym_eig32(const double *S, double *Q, double *lam);

// ---- internals ----------------------------------------------------------------
static int ensure_vectors(ob200_context *ctx, size_t N) {
  if (N <= ctx->vec_capacity) return OB200_OK;
  cudaFree(ctx->r); cudaFree(ctx->p0); cudaFree(ctx->p1); cudaFree(ctx->Hp);
  ctx->r = ctx->p0 = ctx->p1 = ctx->Hp = nullptr;
  ctx->vec_capacity = 0;
  const size_t bytes = sizeof(double) * (N + 64);
  CK(cudaMalloc(&ctx->r, bytes));
  CK(cudaMalloc(&ctx->p0, bytes));
  CK(cudaMalloc(&ctx->p1, bytes));
  CK(cudaMalloc(&ctx->Hp, bytes));
  ctx->vec_capacity = N;
  return OB200_OK;
}
static int ensure_staging(ob200_context *ctx, size_t N) {
  if (N <= ctx->gs_capacity) return OB200_OK;
  cudaFree(ctx->gs);
  ctx->gs = nullptr;
  ctx->gs_capacity = 0;
  CK(cudaMalloc(&ctx->gs, sizeof(double) * (2 * N + 64)));
  ctx->gs_capacity = N;
  return OB200_OK;
}

// Digit planes of A for the tcgen05 contraction (see tc_common.cuh), cached per operator.  The cache key is
// (device pointer, n, CONTENT CHECKSUM of A): planes_checksum_async() enqueues the checksum of the caller's A
// on every solve (one pass over A, a few microseconds), planes_validate() -- called after the stream has been
// synchronised -- rebuilds the planes whenever pointer, size or checksum differ.  An A updated in place, or a
// different A that the allocator placed at the address of a freed one, can therefore never meet stale planes.
// Sets ctx->planes_ok = false when some block is not 22-bit block-fixed-point: the caller then stays on the
// fp64 tensor-core path.
static int planes_checksum_async(ob200_context *ctx, const uint16_t *A, uint64_t n) {
  const unsigned long long nblk = (n + 127) / 128;
  CK(cudaMemsetAsync(ctx->dsum, 0, 8, ctx->stream));
  CK(launch_stiefel_checksum(A, nblk, ctx->dsum, ctx->sm_count, ctx->stream));
  ctx->launches += 1;
  CK(cudaMemcpyAsync(ctx->hsum, ctx->dsum, 8, cudaMemcpyDeviceToHost, ctx->stream));
  return OB200_OK;
}
static int planes_validate(ob200_context *ctx, const uint16_t *A, uint64_t n) {   // stream synchronised by the caller
  const unsigned long long sum = *ctx->hsum;
  if (ctx->planes_key == A && ctx->planes_n == n && ctx->planes_sum == sum) return OB200_OK;
  const unsigned long long nblk = (n + 127) / 128;
  const size_t bytes = stiefel_planes_bytes(nblk);
  ctx->planes_key = nullptr;
  if (bytes > ctx->planes_cap) {
    cudaFree(ctx->planes); cudaFree(ctx->plane_exp);
    ctx->planes = nullptr; ctx->plane_exp = nullptr; ctx->planes_cap = 0;
    CK(cudaMalloc(&ctx->planes, bytes));
    CK(cudaMalloc(&ctx->plane_exp, sizeof(int) * (nblk + 1)));
    ctx->planes_cap = bytes;
  }
  CK(cudaMemsetAsync(ctx->plane_exp + nblk, 0, sizeof(int), ctx->stream));
  CK(launch_stiefel_planes(A, nblk, ctx->planes, ctx->plane_exp, ctx->plane_exp + nblk, ctx->sm_count, ctx->stream));
  ctx->launches += 1;
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, ctx->plane_exp + nblk, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->planes_ok = (bad == 0);
  ctx->planes_key = A;
  ctx->planes_n = n;
  ctx->planes_sum = sum;
  return OB200_OK;
}
static int ensure_planes(ob200_context *ctx, const uint16_t *A, uint64_t n) {
  int rc = planes_checksum_async(ctx, A, n);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return planes_validate(ctx, A, n);
}

// Cross-rank fold of set[off, off+count) (no-op on one GPU).  All ranks call it in lockstep.
static int exchange(ob200_context *ctx, u64 *set, int off, int count, int mode = 0) {
  if (ctx->cm.world <= 1) return OB200_OK;
  CK(launch_tcg_exchange(ctx->cm, ctx->cm.epoch, set, off, count, mode, reinterpret_cast<int *>(ctx->barrier + 1),
                         ctx->stream));
  ctx->cm.epoch += 1;
  ctx->launches += 1;
  return OB200_OK;
}

// exact dot products (<= 4) -> host doubles; synchronises the stream
static int dots_sync(ob200_context *ctx, uint64_t N, int count, const double *const *a, const double *const *b,
                     double *out) {
  u64 *set = ctx->acc;  // set 0
  CK(cudaMemsetAsync(set, 0, sizeof(u64) * ACC_SCAL_WORDS, ctx->stream));
  CK(launch_dots(N, count, a, b, set, ctx->sm_count, ctx->stream));
  {
    int rc = exchange(ctx, set, 0, count * KUL_STRIDE);
    if (rc) return rc;
  }
  CK(launch_finalize_many(set, count, ctx->dscal, ctx->stream));
  ctx->launches += 2;
  CK(cudaMemcpyAsync(ctx->hscal, ctx->dscal, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < count; ++i) out[i] = ctx->hscal[i];
  return OB200_OK;
}

// read the fixed-point Gram of set 0 back to the host as doubles (p x p = 32 x 32)
static int read_gram(ob200_context *ctx, int e, double *G /* 1024 */, double *scal2 /* nullable: 2 scalars */) {
  {
    int rc = exchange(ctx, ctx->acc, 0, ACC_WORDS);
    if (rc) return rc;
  }
  if (scal2) {
    CK(launch_finalize_many(ctx->acc, 2, ctx->dscal, ctx->stream));
    ctx->launches += 1;
    CK(cudaMemcpyAsync(ctx->hscal, ctx->dscal, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaMemcpyAsync(ctx->hacc, ctx->acc, sizeof(u64) * ACC_WORDS, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->hacc[ACC_FLAG_OFF] != 0) return fail(ctx, OB200_NUMERIC_RANGE, "fixed-point Gram bound exceeded");
  const double q = std::ldexp(1.0, e - 90);
  for (int i = 0; i < 1024; ++i)
    G[i] = fix2_to_double((i64)ctx->hacc[ACC_GRAM_OFF + 2 * i], (i64)ctx->hacc[ACC_GRAM_OFF + 2 * i + 1], q);
  if (scal2) { scal2[0] = ctx->hscal[0]; scal2[1] = ctx->hscal[1]; }
  return OB200_OK;
}

static int check_params(ob200_context *ctx, const ob200_stpcg_params *p) {
  // reference IterativeSolvers.h:183-205 (max_iterations < 0 is dead code for size_t, l.187)
  if (!(p->Delta > 0)) return fail(ctx, OB200_INVALID_ARGUMENT, "Trust-region radius (Delta) must be a positive real value");
  if ((p->kappa_fgr < 0) || (p->kappa_fgr >= 1) || std::isnan(p->kappa_fgr))
    return fail(ctx, OB200_INVALID_ARGUMENT, "Target fractional reduction of the gradient norm (kappa_fgr) must be a real value in the range [0,1)");
  if ((p->theta < 0) || (p->theta > 1) || std::isnan(p->theta))
    return fail(ctx, OB200_INVALID_ARGUMENT, "Target superlinear convergence rate (theta) must be a real value in the range [0,1]");
  if ((p->epsilon <= 0) || (p->epsilon >= 1) || std::isnan(p->epsilon))
    return fail(ctx, OB200_INVALID_ARGUMENT, "Relative norm tolerance for declaring a vector to lie in the kernel of H (epsilon) should be a small positive number in the range (0,1)");
  return OB200_OK;
}

static int check_ldu(ob200_context *ctx, uint64_t n, uint64_t k, uint64_t ldu, const double *Ut) {
  const uint64_t ld = ldu ? ldu : n;
  if (k > 1 && (ld < n || (ld & 1))) return fail(ctx, OB200_INVALID_ARGUMENT, "U^T row stride (ldu) must be even and >= n");
  if (k && (reinterpret_cast<uintptr_t>(Ut) & 15)) return fail(ctx, OB200_INVALID_ARGUMENT, "U^T must be 16-byte aligned");
  return OB200_OK;
}
static int check_sphere(ob200_context *ctx, const ob200_operator *H) {
  if (H->p != 1) return fail(ctx, OB200_INVALID_ARGUMENT, "sphere operator acts on vectors (p == 1)");
  if (H->k > 16) return fail(ctx, OB200_UNSUPPORTED, "sphere low-rank operator supports k <= 16");
  if (!H->diag_dev || !H->x_dev || !H->Ax_dev || (H->k && (!H->U_dev || !H->sigma_host)))
    return fail(ctx, OB200_INVALID_ARGUMENT, "incomplete sphere operator");
  return check_ldu(ctx, H->n, H->k, H->ldu, H->U_dev);
}

// out = A v for A = diag(d) + U diag(sigma) U^T: low-rank sums (exact reduction), then one streaming pass
static int sphere_Av(ob200_context *ctx, uint64_t n, uint64_t k, const double *d, const double *Ut, uint64_t ldu,
                     const double *sigma_host, const double *v, double *out) {
  if (!ldu) ldu = n;
  cudaStream_t st = ctx->stream;
  double *st_dev = ctx->dmat;   // k doubles
  ctx->S_cache_valid = false;   // (dmat[0..] is also where the Stiefel HVP keeps S)
  if (k) {
    CK(cudaMemsetAsync(ctx->acc + ACC_GRAM_OFF, 0, sizeof(u64) * 16 * KUL_STRIDE, st));
    CK(launch_sphere_tdot(n, Ut, ldu, sigma_host, (int)k, v, ctx->acc, ctx->sm_count, st));
    ctx->launches += 1;
    int rc = exchange(ctx, ctx->acc, ACC_GRAM_OFF, (int)k * KUL_STRIDE);
    if (rc) return rc;
    CK(launch_sphere_scale_t(ctx->acc, sigma_host, (int)k, st_dev, st));
    ctx->launches += 1;
  }
  CK(launch_sphere_apply(n, d, Ut, ldu, (int)k, st_dev, v, out, ctx->sm_count, st));
  ctx->launches += 1;
  return OB200_OK;
}

static int sparse_args(ob200_context *ctx, const ob200_operator *H, SparseArgs *sp) {
  memset(sp, 0, sizeof(*sp));
  sp->kind = H->kind;
  sp->r = (int)H->p;
  if (H->kind == OB200_OP_BLOCK_CSR3) {
    if (H->p < 3 || H->p > 8) return fail(ctx, OB200_UNSUPPORTED, "block-CSR operator requires 3 <= p = r <= 8");
    if (H->n % 3) return fail(ctx, OB200_INVALID_ARGUMENT, "block-CSR operator: n must be 3 x (number of poses)");
    if (!H->csr_rowptr_dev || !H->csr_colidx_dev || !H->csr_blocks_dev || !H->csr_lambda_dev || !H->Y_dev)
      return fail(ctx, OB200_INVALID_ARGUMENT, "incomplete block-CSR operator");
    sp->units = H->n / 3;
    sp->rowptr = reinterpret_cast<const unsigned long long *>(H->csr_rowptr_dev);
    sp->colidx = H->csr_colidx_dev;
    sp->blocks = H->csr_blocks_dev;
    sp->lambda = H->csr_lambda_dev;
    sp->X = H->Y_dev;
    if (ctx->cm.world > 1) {
      if (!ctx->halo_buf || !ctx->halo_peer[ctx->cm.world - 1])
        return fail(ctx, OB200_INVALID_ARGUMENT, "row-sharded block-CSR operator: call ob200_halo_create / ob200_halo_connect first");
      if (ctx->cm.rank + 1 < ctx->cm.world && (uint64_t)(H->n / 3) % 256)   // (the last rank takes the ragged end)
        return fail(ctx, OB200_INVALID_ARGUMENT, "row-sharded block-CSR operator: local pose count must be a multiple of 256");
      if (H->csr_n_halo * 3 * H->p * sizeof(double) > ctx->halo_bytes) return fail(ctx, OB200_INVALID_ARGUMENT, "halo buffer too small");
      if (H->halo_send_ptr[ctx->cm.world] && !H->halo_send_idx_dev) return fail(ctx, OB200_INVALID_ARGUMENT, "missing halo send list");
      sp->n_halo = H->csr_n_halo;
      sp->halo = ctx->halo_buf;
      sp->send_idx = H->halo_send_idx_dev;
      for (int q = 0; q <= ctx->cm.world; ++q) sp->send_ptr[q] = H->halo_send_ptr[q];
      for (int q = 0; q < ctx->cm.world; ++q)
        sp->peer_halo[q] = static_cast<double *>(ctx->halo_peer[q]) + H->halo_dst_off[q] * 3 * H->p;
    }
  } else {
    if (!H->gx || !H->gy || !H->gz || (uint64_t)H->gx * H->gy * H->gz != H->n)
      return fail(ctx, OB200_INVALID_ARGUMENT, "stencil operator: gx gy gz must equal n");
    if (H->p > 0xffffu) return fail(ctx, OB200_UNSUPPORTED, "stencil operator: too many columns");
    sp->units = H->n;
    sp->gx = H->gx; sp->gy = H->gy; sp->gz = H->gz;
  }
  return OB200_OK;
}

// The v6 Stiefel kernel solves in the eigenbasis of S, where  p S  is an elementwise shift.
// Symmetric eigen-decomposition S = Q diag(lam) Q^T of a 32 x 32 matrix: Householder tridiagonalisation followed by the
// implicit QL iteration (the classical tred2 / tql2 pair), host, deterministic, ~30 us.  Q row-major, eigenvector k in
// column k.
static void sym_eig32(const double *S, double *Q, double *lam) {
  const int n = 32;
  double V[32][32], d[32], e[32];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i][j] = 0.5 * (S[i * n + j] + S[j * n + i]);
  // ---- tred2 ----
  for (int j = 0; j < n; ++j) d[j] = V[n - 1][j];
  for (int i = n - 1; i > 0; --i) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; ++j) { d[j] = V[i - 1][j]; V[i][j] = 0.0; V[j][i] = 0.0; }
    } else {
      for (int k = 0; k < i; ++k) { d[k] /= scale; h += d[k] * d[k]; }
      double f = d[i - 1];
      double g = std::sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      for (int j = 0; j < i; ++j) {
        f = d[j];
        V[j][i] = f;
        g = e[j] + V[j][j] * f;
        for (int k = j + 1; k <= i - 1; ++k) { g += V[k][j] * d[k]; e[k] += V[k][j] * f; }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * d[j]; }
      const double hh = f / (h + h);
      for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
      for (int j = 0; j < i; ++j) {
        f = d[j];
        g = e[j];
        for (int k = j; k <= i - 1; ++k) V[k][j] -= (f * e[k] + g * d[k]);
        d[j] = V[i - 1][j];
        V[i][j] = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; ++i) {
    V[n - 1][i] = V[i][i];
    V[i][i] = 1.0;
    const double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; ++k) d[k] = V[k][i + 1] / h;
      for (int j = 0; j <= i; ++j) {
        double g = 0.0;
        for (int k = 0; k <= i; ++k) g += V[k][i + 1] * V[k][j];
        for (int k = 0; k <= i; ++k) V[k][j] -= g * d[k];
      }
    }
    for (int k = 0; k <= i; ++k) V[k][i + 1] = 0.0;
  }
  for (int j = 0; j < n; ++j) { d[j] = V[n - 1][j]; V[n - 1][j] = 0.0; }
  V[n - 1][n - 1] = 1.0;
  e[0] = 0.0;
  // ---- tql2 ----
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = 2.220446049250313e-16;
  for (int l = 0; l < n; ++l) {
    tst1 = std::fmax(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n) {
      if (std::fabs(e[m]) <= eps * tst1) break;
      ++m;
    }
    if (m > l) {
      int iter = 0;
      do {
        ++iter;
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = std::hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c;
        const double el1 = e[l + 1];
        double s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; --i) {
          c3 = c2;
          c2 = c;
          s2 = s;
          g = c * e[i];
          h = c * p;
          r = std::hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; ++k) {
            h = V[k][i + 1];
            V[k][i + 1] = s * V[k][i] + c * h;
            V[k][i] = c * V[k][i] - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1 && iter < 200);
    }
    d[l] = d[l] + f;
    e[l] = 0.0;
  }
  for (int i = 0; i < n; ++i) {
    lam[i] = d[i];
    for (int j = 0; j < n; ++j) Q[i * n + j] = V[i][j];
  }
}

// ---- unfused Steihaug-Toint loop ----------------------------------------------------------------------------------------
// Reference IterativeSolvers.h:211-424 statement for statement (no `At`, no user function: how TNT calls it), with the
// vectors on the device: every inner product is the exact device reduction (dots_sync), every update a level-1 kernel.
// H is a host callback or any built-in operator (its stand-alone HVP, ob200_hvp); P is the pointwise Jacobi scaling, the
// projected Jacobi scaling of the Stiefel model, or a host callback.  Serves the operator x preconditioner pairs the fused
// kernels do not implement.
extern "C" int ob200_hvp(ob200_context *ctx, const ob200_operator *H, const double *v, double *out);
extern "C" int ob200_stiefel_project(ob200_context *ctx, uint64_t n, uint64_t p, const double *Y, const double *Z, double *out);
static int stpcg_generic(ob200_context *ctx, const ob200_operator *H, const ob200_precon *P, const double *g_dev,
                         const ob200_stpcg_params *prm, double *s_dev, ob200_stpcg_result *res) {
  const uint64_t N = H->n * H->p;
  int rc = ensure_vectors(ctx, N);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const uint64_t launches0 = ctx->launches;
  // (the built-in HVPs use the context's tCG vectors as their workspace: the loop keeps its own set)
  const uint64_t Npad = (N + 31ull) & ~31ull;   // 256-byte aligned vectors (the level-1 kernels use 16-byte accesses)
  if (ctx->gen_capacity < Npad) {
    cudaFree(ctx->gen_vec);
    ctx->gen_vec = nullptr;
    ctx->gen_capacity = 0;
    CK(cudaMalloc(&ctx->gen_vec, sizeof(double) * 5 * Npad));
    ctx->gen_capacity = Npad;
  }
  double *r = ctx->gen_vec, *p = r + Npad, *Hp = p + Npad, *vbuf = Hp + Npad, *scratch = vbuf + Npad;
  const bool has_P = P && P->kind != OB200_PRECON_NONE;
  auto apply_cb = [&](ob200_apply_fn fn, void *user, const double *in, double *out) -> int {
    CK(cudaStreamSynchronize(st));
    if (fn(user, in, out)) return fail(ctx, OB200_ABORTED, "operator / preconditioner callback reported a failure");
    return OB200_OK;
  };
  auto apply_H = [&](const double *in, double *out) -> int {                      // Hp = H(p), l.294
    if (H->kind == OB200_OP_HOST_CALLBACK) return apply_cb(H->apply, H->apply_user, in, out);
    return ob200_hvp(ctx, H, in, out);
  };
  auto precondition = [&](const double *rin, double *vout) -> int {               // v = P(r), l.383-386
    if (P->kind == OB200_PRECON_JACOBI || P->kind == OB200_PRECON_STIEFEL_PROJECTED_JACOBI) {
      const bool proj = P->kind == OB200_PRECON_STIEFEL_PROJECTED_JACOBI;
      CK(launch_hadamard(N, P->minv_dev, rin, proj ? scratch : vout, ctx->sm_count, st));
      ctx->launches += 1;
      if (proj) return ob200_stiefel_project(ctx, H->n, H->p, H->Y_dev, scratch, vout);
      return OB200_OK;
    }
    return apply_cb(P->apply, P->apply_user, rin, vout);
  };
  auto axpby = [&](double a, const double *x, double b, const double *y, double *out) -> int {
    CK(launch_axpby(N, a, x, b, y, out, ctx->sm_count, st));
    ctx->launches += 1;
    return OB200_OK;
  };
  CK(cudaMemsetAsync(s_dev, 0, sizeof(double) * N, st));                          // s_0 = 0 * g           (l.211)
  CK(cudaMemcpyAsync(r, g_dev, sizeof(double) * N, cudaMemcpyDeviceToDevice, st)); // r_0 = g              (l.214)
  const double *v = r;
  if (has_P) {
    if ((rc = precondition(r, vbuf))) return rc;
    v = vbuf;
  }
  if ((rc = axpby(-1.0, v, 0.0, nullptr, p))) return rc;                          // p_0 = -v_0            (l.256)
  double rv = 0.0;
  {
    const double *aa[1] = {r}, *bb[1] = {v};
    if ((rc = dots_sync(ctx, N, 1, aa, bb, &rv))) return rc;                      // <r_0, v_0>            (l.259)
  }
  const double r0_norm = std::sqrt(rv);                                           // l.275
  const double target = r0_norm * std::min(prm->kappa_fgr, std::pow(r0_norm, prm->theta));   // l.278-279
  const double Delta_2 = prm->Delta * prm->Delta;                                 // l.271
  double sk_M_2 = 0.0, sk_M_pk = 0.0, pk_M_2 = rv;                                // l.262-266
  uint64_t it = 0;
  int exit_reason = OB200_EXIT_MAX_ITERATIONS;
  double mnorm = 0.0;
  bool on_boundary = false;
  for (; it < prm->max_iterations; ++it) {                                        // l.285
    if (std::sqrt(rv) <= target) { exit_reason = OB200_EXIT_RESIDUAL; break; }    // l.290
    if ((rc = apply_H(p, Hp))) return rc;                                         // l.294
    double d3[3];
    {
      const double *aa[3] = {p, Hp, p}, *bb[3] = {Hp, Hp, p};
      if ((rc = dots_sync(ctx, N, 3, aa, bb, d3))) return rc;                     // kappa, |Hp|^2, |p|^2 (l.300-306)
    }
    const double kappa = d3[0];
    if (std::sqrt(d3[1]) / std::sqrt(d3[2]) < prm->epsilon) {                     // l.305-307: p in ker(H)
      double pr = 0.0;
      const double *aa[1] = {p}, *bb[1] = {r};
      if ((rc = dots_sync(ctx, N, 1, aa, bb, &pr))) return rc;                    // l.320
      double sgn = 1.0;
      if (pr < 0) { sgn = -1.0; sk_M_pk = -sk_M_pk; }                             // l.324-325
      const double sigma = (-sk_M_pk + std::sqrt(sk_M_pk * sk_M_pk + pk_M_2 * (Delta_2 - sk_M_2))) / pk_M_2;   // l.330-332
      if ((rc = axpby(1.0, s_dev, sgn * sigma, p, s_dev))) return rc;             // l.336
      exit_reason = OB200_EXIT_KERNEL;
      on_boundary = true;
      break;
    }
    const double alpha = rv / kappa;                                              // l.341
    const double skp1 = sk_M_2 + 2 * alpha * sk_M_pk + alpha * alpha * pk_M_2;    // l.344-345
    if (kappa <= 0 || skp1 > Delta_2) {                                           // l.347
      const double sigma = (-sk_M_pk + std::sqrt(sk_M_pk * sk_M_pk + pk_M_2 * (Delta_2 - sk_M_2))) / pk_M_2;   // l.355-357
      if ((rc = axpby(1.0, s_dev, sigma, p, s_dev))) return rc;                   // l.360
      exit_reason = OB200_EXIT_BOUNDARY;
      on_boundary = true;
      break;
    }
    if ((rc = axpby(1.0, s_dev, alpha, p, s_dev))) return rc;                     // l.374
    if ((rc = axpby(1.0, r, alpha, Hp, r))) return rc;                            // l.377
    if (has_P && (rc = precondition(r, vbuf))) return rc;                         // l.383-386
    double rv_new = 0.0;
    {
      const double *aa[1] = {r}, *bb[1] = {v};
      if ((rc = dots_sync(ctx, N, 1, aa, bb, &rv_new))) return rc;                // l.408
    }
    const double beta = rv_new / (alpha * kappa);                                 // l.412
    sk_M_2 = skp1;                                                                // l.415
    sk_M_pk = beta * (sk_M_pk + alpha * pk_M_2);                                  // l.416
    pk_M_2 = rv_new + beta * beta * pk_M_2;                                       // l.417
    if ((rc = axpby(-1.0, v, beta, p, p))) return rc;                             // l.420
    rv = rv_new;
  }
  mnorm = on_boundary ? prm->Delta : std::sqrt(sk_M_2);                           // l.334 / 359 / 424
  CK(cudaStreamSynchronize(st));
  res->update_step_M_norm = mnorm;
  res->num_iterations = it;
  res->exit_reason = exit_reason;
  res->r0_norm = r0_norm;
  res->final_rv = rv;
  res->kernel_launches = ctx->launches - launches0;
  res->solve_kernel_ms = 0.f;
  ctx->last_path = 3;
  return OB200_OK;
}

static int stpcg_device(ob200_context *ctx, const ob200_operator *H, const ob200_precon *P, const double *g_dev,
                        const ob200_stpcg_params *prm, double *s_dev, ob200_stpcg_result *res) {
  int rc = check_params(ctx, prm);
  if (rc) return rc;
  if (!H || !g_dev || !s_dev || !res) return fail(ctx, OB200_INVALID_ARGUMENT, "null argument");
  if (H->n == 0 || H->p == 0) return fail(ctx, OB200_INVALID_ARGUMENT, "empty operator");
  const int pkind = P ? P->kind : OB200_PRECON_NONE;
  if (H->kind == OB200_OP_HOST_CALLBACK || pkind == OB200_PRECON_HOST_CALLBACK || pkind == OB200_PRECON_STIEFEL_PROJECTED_JACOBI) {
    if (H->kind == OB200_OP_HOST_CALLBACK && !H->apply) return fail(ctx, OB200_INVALID_ARGUMENT, "callback operator without a function");
    if (pkind == OB200_PRECON_HOST_CALLBACK && !P->apply) return fail(ctx, OB200_INVALID_ARGUMENT, "callback preconditioner without a function");
    if ((pkind == OB200_PRECON_JACOBI || pkind == OB200_PRECON_STIEFEL_PROJECTED_JACOBI) && !P->minv_dev)
      return fail(ctx, OB200_INVALID_ARGUMENT, "Jacobi preconditioner without minv");
    if (pkind == OB200_PRECON_STIEFEL_PROJECTED_JACOBI && (H->kind != OB200_OP_STIEFEL_BLOCKDIAG || !H->Y_dev))
      return fail(ctx, OB200_INVALID_ARGUMENT, "the projected Jacobi preconditioner belongs to the Stiefel operator");
    if (pkind == OB200_PRECON_JACOBI && H->kind == OB200_OP_STIEFEL_BLOCKDIAG)
      return fail(ctx, OB200_UNSUPPORTED, "elementwise Jacobi does not preserve the Stiefel tangent space");
    if (ctx->cm.world > 1) return fail(ctx, OB200_UNSUPPORTED, "the unfused loop is single-GPU");
    CK(cudaSetDevice(ctx->device));
    return stpcg_generic(ctx, H, P, g_dev, prm, s_dev, res);
  }
  const uint64_t N = H->n * H->p;
  const double *minv = nullptr;
  if (P && P->kind == OB200_PRECON_JACOBI) {
    if (!P->minv_dev) return fail(ctx, OB200_INVALID_ARGUMENT, "Jacobi preconditioner without minv");
    minv = P->minv_dev;
  } else if (P && P->kind != OB200_PRECON_NONE) {
    return fail(ctx, OB200_UNSUPPORTED, "unknown preconditioner kind");
  }
  if (H->kind == OB200_OP_STIEFEL_BLOCKDIAG) {
    if (H->p != 32) return fail(ctx, OB200_UNSUPPORTED, "Stiefel block-diagonal operator requires p == 32");
    if (minv) return fail(ctx, OB200_UNSUPPORTED, "elementwise Jacobi does not preserve the Stiefel tangent space");
    if (!H->A_bf16_dev || !H->Y_dev || !H->S_host) return fail(ctx, OB200_INVALID_ARGUMENT, "incomplete Stiefel operator");
    if (!(H->op_norm_bound > 0)) return fail(ctx, OB200_INVALID_ARGUMENT, "op_norm_bound must be positive");
  } else if (H->kind == OB200_OP_DIAG) {
    if (!H->diag_dev) return fail(ctx, OB200_INVALID_ARGUMENT, "diag operator without diagonal");
  } else if (H->kind == OB200_OP_SPHERE_LOWRANK) {
    if ((rc = check_sphere(ctx, H))) return rc;
  } else if (H->kind == OB200_OP_BLOCK_CSR3 || H->kind == OB200_OP_STENCIL7) {
    SparseArgs chk;
    if ((rc = sparse_args(ctx, H, &chk))) return rc;
    if (ctx->cm.world > 1 && H->kind == OB200_OP_STENCIL7)
      return fail(ctx, OB200_UNSUPPORTED, "the stencil tCG operator is single-GPU in this version");
  } else {
    return fail(ctx, OB200_UNSUPPORTED, "operator kind not supported by the fused tCG path");
  }
  CK(cudaSetDevice(ctx->device));
  if ((rc = ensure_vectors(ctx, N))) return rc;
  const bool want_tc = (H->kind == OB200_OP_STIEFEL_BLOCKDIAG && ctx->opt_tcgen05);
  bool use_tc = false;
  if (want_tc && (rc = planes_checksum_async(ctx, H->A_bf16_dev, H->n))) return rc;   // validated after the sync below
  const uint64_t launches0 = ctx->launches;
  cudaStream_t st = ctx->stream;
  // v6 kernel: the solve runs in the eigenbasis of S = Q Lambda Q^T -- g~ = g Q, Y~ = Y Q, s = s~ Q^T.  The iteration is
  // the same Steihaug-Toint loop on an orthogonally rotated copy of the problem (Frobenius products, the projection and
  // the trust-region norm are invariant), with  p S  reduced to an elementwise shift.
  const int tc_gen = ctx->opt_tcgen05;
  const bool rotate = want_tc && tc_gen == 1;
  const double *Y_solve = H->Y_dev;
  if (rotate) {
    CK(cudaStreamSynchronize(st));
    if ((rc = planes_validate(ctx, H->A_bf16_dev, H->n))) return rc;
  }
  const bool rotated = rotate && ctx->planes_ok;
  if (rotated) {
    if (N > ctx->yrot_capacity) {
      cudaFree(ctx->yrot);
      ctx->yrot = nullptr; ctx->yrot_capacity = 0;
      ctx->rot_valid = false;
      CK(cudaMalloc(&ctx->yrot, sizeof(double) * (N + 64)));
      ctx->yrot_capacity = N;
    }
    const unsigned long long nblk_r = (H->n + 127) / 128;
    int grid_r = ctx->sm_count;
    if ((unsigned long long)grid_r > nblk_r) grid_r = (int)nblk_r;
    // (Q, Lambda, Y~) are kept for repeated solves at the same point: S = sym(Y^T A Y) identifies Y
    if (!(ctx->rot_valid && ctx->rot_Y == H->Y_dev && ctx->rot_n == H->n &&
          !memcmp(ctx->rot_S, H->S_host, sizeof(ctx->rot_S)))) {
      std::vector<double> rot(3 * 1024, 0.0);
      sym_eig32(H->S_host, rot.data(), rot.data() + 2048);
      for (int i = 0; i < 32; ++i)
        for (int j = 0; j < 32; ++j) rot[1024 + i * 32 + j] = rot[j * 32 + i];
      CK(cudaMemcpyAsync(ctx->drot, rot.data(), sizeof(double) * 3 * 1024, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));     // `rot` is pageable host memory
      CK(launch_stiefel_rowgemm(H->n, nullptr, 0.0, H->Y_dev, ctx->drot, ctx->yrot, grid_r, st));    // Y~ = Y Q
      ctx->launches += 1;
      memcpy(ctx->rot_S, H->S_host, sizeof(ctx->rot_S));
      ctx->rot_Y = H->Y_dev;
      ctx->rot_n = H->n;
      ctx->rot_valid = true;
    }
    CK(launch_stiefel_rowgemm(H->n, nullptr, 0.0, g_dev, ctx->drot, ctx->p0, grid_r, st));           // g~ = g Q
    ctx->launches += 1;
    Y_solve = ctx->yrot;
  }

  TcgCommon a;
  a.N = N;
  a.g = rotated ? ctx->p0 : g_dev;      // (p0 is first written by the second CG iteration)
  a.s = s_dev;
  a.r = ctx->r;
  a.p0 = ctx->p0;
  a.p1 = ctx->p1;
  a.Hp = ctx->Hp;
  a.minv = minv;
  a.rv0 = 0; a.target = 0;
  a.Delta = prm->Delta;
  a.epsilon = prm->epsilon;
  a.max_iterations = prm->max_iterations;
  a.acc = ctx->acc;
  a.barrier = ctx->barrier;
  a.abort_flag = reinterpret_cast<int *>(ctx->barrier + 1);
  a.result = ctx->dres;
  a.dbg = ctx->dbg;
  a.cm = ctx->cm;
  a.blk_stats = nullptr;
  a.nblk_stats = 0;
  if (want_tc && tc_gen == 1) {   // v6 kernel: block maxima of r (seeded by the init kernel) and p
    const size_t nblk = (H->n + 127) / 128;
    if (nblk > ctx->blk_stats_cap) {
      cudaFree(ctx->blk_stats);
      ctx->blk_stats = nullptr; ctx->blk_stats_cap = 0;
      CK(cudaMalloc(&ctx->blk_stats, sizeof(unsigned long long) * 10 * nblk));
      ctx->blk_stats_cap = nblk;
    }
    CK(cudaMemsetAsync(ctx->blk_stats, 0, sizeof(unsigned long long) * 10 * nblk, ctx->stream));
    a.blk_stats = ctx->blk_stats;
    a.nblk_stats = nblk;
  }

  // s = 0, r = g, <r, v>  (IterativeSolvers.h:211-266)
  CK(cudaMemsetAsync(ctx->acc, 0, sizeof(u64) * ACC_SETS * ACC_WORDS, st));
  CK(cudaMemsetAsync(ctx->barrier, 0, 64, st));
  CK(launch_tcg_init(a, ctx->sm_count * 2, st));
  if ((rc = exchange(ctx, ctx->acc, SC_RV * KUL_STRIDE, KUL_STRIDE))) return rc;
  CK(launch_tcg_finalize(ctx->acc, SC_RV, ctx->dscal, st));
  ctx->launches += 2;
  CK(cudaMemcpyAsync(ctx->hscal, ctx->dscal, sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaMemsetAsync(ctx->acc, 0, sizeof(u64) * ACC_WORDS, st));  // set 0 is reused by the loop
  if (H->kind == OB200_OP_STIEFEL_BLOCKDIAG) {
    CK(cudaMemcpyAsync(ctx->dmat, H->S_host, sizeof(double) * 32 * 32, cudaMemcpyHostToDevice, st));
    ctx->S_cache_valid = false;
  }
  CK(cudaStreamSynchronize(st));
  if (want_tc) {
    if (!rotate && (rc = planes_validate(ctx, H->A_bf16_dev, H->n))) return rc;
    use_tc = ctx->planes_ok;
  }
  const double rv0 = ctx->hscal[0];
  const double r0_norm = std::sqrt(rv0);                                            // l.275
  const double target = r0_norm * std::min(prm->kappa_fgr, std::pow(r0_norm, prm->theta));  // l.278-279
  a.rv0 = rv0;
  a.target = target;
  a.cm = ctx->cm;

  CK(cudaEventRecord(ctx->ev0, st));
  if (H->kind == OB200_OP_DIAG) {
    CK(launch_tcg_diag(a, H->diag_dev, ctx->sm_count, st));
  } else if (H->kind == OB200_OP_SPHERE_LOWRANK) {
    CK(launch_tcg_sphere(a, H->diag_dev, H->U_dev, H->ldu ? H->ldu : H->n, H->x_dev, H->Ax_dev, H->sigma_host,
                         (int)H->k, H->xAx, ctx->sm_count, st));
  } else if (H->kind == OB200_OP_BLOCK_CSR3 || H->kind == OB200_OP_STENCIL7) {
    SparseArgs sp;
    if ((rc = sparse_args(ctx, H, &sp))) return rc;
    CK(launch_tcg_sparse(a, sp, ctx->sm_count, st));
  } else {
    const unsigned long long nblk = (H->n + 127) / 128;
    int grid = ctx->sm_count;
    if ((unsigned long long)grid > nblk) grid = (int)nblk;
    if (use_tc && tc_gen == 2)                // previous generation of the tcgen05 kernel (row-sharded runs, A/B measurements)
      CK(launch_tcg_stiefel_tc(a, H->n, H->A_bf16_dev, H->Y_dev, ctx->dmat, H->op_norm_bound, ctx->planes,
                               ctx->plane_exp, grid, st));
    else if (use_tc)
      CK(launch_tcg_stiefel_v6(a, H->n, H->A_bf16_dev, Y_solve, ctx->drot + 2048 /* eigenvalues of S */, H->op_norm_bound,
                               ctx->planes, ctx->plane_exp, ctx->sm_count, st));
    else
      CK(launch_tcg_stiefel(a, H->n, H->A_bf16_dev, H->Y_dev, ctx->dmat, H->op_norm_bound, grid, st));
    ctx->last_path = use_tc ? (tc_gen == 2 ? 2 : 1) : 0;
  }
  CK(cudaEventRecord(ctx->ev1, st));
  ctx->launches += 1;
  if (rotated && use_tc) {               // s = s~ Q^T (row-wise, in place)
    const unsigned long long nblk_r = (H->n + 127) / 128;
    int grid_r = ctx->sm_count;
    if ((unsigned long long)grid_r > nblk_r) grid_r = (int)nblk_r;
    CK(launch_stiefel_rowgemm(H->n, nullptr, 0.0, s_dev, ctx->drot + 1024, s_dev, grid_r, st));
    ctx->launches += 1;
  }
  CK(cudaMemcpyAsync(ctx->hres, ctx->dres, sizeof(TcgDeviceResult), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  res->solve_kernel_ms = 0.f;
  cudaEventElapsedTime(&res->solve_kernel_ms, ctx->ev0, ctx->ev1);
  res->update_step_M_norm = ctx->hres->update_step_M_norm;
  res->num_iterations = ctx->hres->num_iterations;
  res->exit_reason = ctx->hres->exit_reason;
  res->r0_norm = r0_norm;
  res->final_rv = ctx->hres->final_rv;
  res->kernel_launches = ctx->launches - launches0;
  ctx->cm.epoch += ctx->hres->phases;
  if (ctx->hres->status == OB200_NUMERIC_RANGE) return fail(ctx, OB200_NUMERIC_RANGE, "fixed-point Gram bound exceeded or non-finite data");
  if (ctx->hres->status == OB200_ABORTED) return fail(ctx, OB200_ABORTED, "device grid barrier watchdog fired");
  return OB200_OK;
}

extern "C" {

int ob200_stpcg(ob200_context *ctx, const ob200_operator *H, const ob200_precon *P, const double *g_dev,
                const ob200_stpcg_params *params, double *s_dev, ob200_stpcg_result *result) {
  if (!ctx || !params) return OB200_INVALID_ARGUMENT;
  return stpcg_device(ctx, H, P, g_dev, params, s_dev, result);
}

int ob200_stpcg_host(ob200_context *ctx, const ob200_operator *H, const ob200_precon *P, const double *g_host,
                     const ob200_stpcg_params *params, double *s_host, ob200_stpcg_result *result) {
  if (!ctx || !params || !H || !g_host || !s_host) return OB200_INVALID_ARGUMENT;
  const uint64_t N = H->n * H->p;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_staging(ctx, N);
  if (rc) return rc;
  double *g_dev = ctx->gs, *s_dev = ctx->gs + N + (N & 1);
  CK(cudaMemcpyAsync(g_dev, g_host, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
  rc = stpcg_device(ctx, H, P, g_dev, params, s_dev, result);
  if (rc) return rc;
  CK(cudaMemcpyAsync(s_host, s_dev, sizeof(double) * N, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return OB200_OK;
}

static uint64_t op_bytes(const ob200_operator *H) {
  const uint64_t N = H->n * H->p;
  switch (H->kind) {
    case OB200_OP_DIAG: return 8 * N;                                   // d read once
    case OB200_OP_STIEFEL_BLOCKDIAG:                                    // A (bf16) + Y read twice
      return ((H->n + 127) / 128) * 128 * 128 * 2 + 2 * 8 * N;
    case OB200_OP_SPHERE_LOWRANK: return 8 * H->n * (4 + 2 * H->k);     // U twice, d, x, w = A x, p re-read
    case OB200_OP_BLOCK_CSR3:                                           // p re-read + X; blocks + indices; row pointer + Lambda
      return 2 * 8 * N + H->csr_nnz * (72 + 4) + (H->n / 3) * (8 + 72);
    case OB200_OP_STENCIL7: return 8 * N;                               // p re-read by the operator pass
    default: return 0;
  }
}
uint64_t ob200_stpcg_step_bytes(const ob200_operator *H, const ob200_precon *P) {
  if (!H) return 0;
  const uint64_t N = H->n * H->p;
  return 10 * 8 * N + op_bytes(H) + ((P && P->kind == OB200_PRECON_JACOBI) ? 2 * 8 * N : 0);
}
uint64_t ob200_hvp_bytes(const ob200_operator *H) {
  if (!H) return 0;
  return 2 * 8 * H->n * H->p + op_bytes(H);
}

int ob200_dot(ob200_context *ctx, uint64_t n, const double *a, const double *b, double *result) {
  if (!ctx || !a || !b || !result) return OB200_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  const double *aa[1] = {a}, *bb[1] = {b};
  return dots_sync(ctx, n, 1, aa, bb, result);
}
int ob200_dots(ob200_context *ctx, uint64_t n, int count, const double *const *a, const double *const *b,
               double *results) {
  if (!ctx || count < 1 || count > 4 || !a || !b || !results) return OB200_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  return dots_sync(ctx, n, count, a, b, results);
}
int ob200_axpby(ob200_context *ctx, uint64_t n, double alpha, const double *x, double beta, const double *y,
                double *out) {
  if (!ctx || !x || !out) return OB200_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  CK(launch_axpby(n, alpha, x, beta, y, out, ctx->sm_count, ctx->stream));
  ctx->launches += 1;
  return OB200_OK;
}
int ob200_div(ob200_context *ctx, uint64_t n, const double *x, double a, double *out) {
  if (!ctx || !x || !out) return OB200_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  CK(launch_div(n, x, a, out, ctx->sm_count, ctx->stream));
  ctx->launches += 1;
  return OB200_OK;
}
int ob200_hadamard(ob200_context *ctx, uint64_t n, const double *d, const double *x, double *out) {
  if (!ctx || !d || !x || !out) return OB200_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  CK(launch_hadamard(n, d, x, out, ctx->sm_count, ctx->stream));
  ctx->launches += 1;
  return OB200_OK;
}

int ob200_stiefel_model(ob200_context *ctx, uint64_t n, uint64_t p, const uint16_t *A, const double *Y,
                        double *S_host, double *f, double *grad_dev, double *op_norm_bound) {
  if (!ctx || !A || !Y || !S_host) return OB200_INVALID_ARGUMENT;
  if (p != 32) return fail(ctx, OB200_UNSUPPORTED, "Stiefel model requires p == 32");
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_vectors(ctx, n * p);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const unsigned long long nblk = (n + 127) / 128;
  int grid = ctx->sm_count;
  if ((unsigned long long)grid > nblk) grid = (int)nblk;
  // ||A||_inf
  CK(cudaMemsetAsync(ctx->dbits, 0, 8, st));
  CK(launch_stiefel_absrowsum(A, nblk * 128, ctx->dbits, st));
  ctx->launches += 1;
  if ((rc = exchange(ctx, reinterpret_cast<u64 *>(ctx->dbits), 0, 1, /*max*/ 1))) return rc;
  unsigned long long bits = 0;
  CK(cudaMemcpyAsync(&bits, ctx->dbits, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  double Ainf;
  memcpy(&Ainf, &bits, 8);
  if (!(Ainf > 0)) Ainf = 1.0;
  // W = A Y (into Hp workspace), Gram Y^T W, <Y, W>
  const int e = gram_exponent_host(Ainf * 4.0);
  CK(cudaMemsetAsync(ctx->acc, 0, sizeof(u64) * ACC_WORDS, st));
  CK(launch_stiefel_apply(n, A, Y, nullptr, Y, ctx->Hp, ctx->acc, std::ldexp(1.0, 90 - e), grid, st));
  ctx->launches += 1;
  std::vector<double> G(1024);
  double sc[2];
  if ((rc = read_gram(ctx, e, G.data(), sc))) return rc;
  double fro = 0.0;
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) {
      const double sij = 0.5 * (G[i * 32 + j] + G[j * 32 + i]);
      S_host[i * 32 + j] = sij;
      fro += sij * sij;
    }
  if (f) *f = 0.5 * sc[0];
  if (op_norm_bound) *op_norm_bound = Ainf + std::sqrt(fro);
  if (grad_dev) {
    for (int i = 0; i < 1024; ++i) ctx->hmat[i] = -S_host[i];
    CK(cudaMemcpyAsync(ctx->dmat + 1024, ctx->hmat, sizeof(double) * 1024, cudaMemcpyHostToDevice, st));
    CK(launch_stiefel_rowgemm(n, ctx->Hp, 1.0, Y, ctx->dmat + 1024, grad_dev, grid, st));   // A Y - Y S
    ctx->launches += 1;
    CK(cudaStreamSynchronize(st));  // hmat is reused
  }
  return OB200_OK;
}

int ob200_debug_block_apply(ob200_context *ctx, uint64_t n, const uint16_t *A, const double *V, double *out,
                            int use_tcgen05) {
  if (!ctx || !A || !V || !out) return OB200_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  const unsigned long long nblk = (n + 127) / 128;
  int grid = ctx->sm_count;
  if ((unsigned long long)grid > nblk) grid = (int)nblk;
  if (use_tcgen05) {
    int rc = ensure_planes(ctx, A, n);
    if (rc) return rc;
    if (!ctx->planes_ok) return fail(ctx, OB200_UNSUPPORTED, "A is not 16-bit block-fixed-point: tcgen05 path unavailable");
    CK(launch_stiefel_ap_tc(n, ctx->planes, ctx->plane_exp, V, out, use_tcgen05 == 2 ? 1 : 0, grid, ctx->stream));
  } else {
    CK(cudaMemsetAsync(ctx->acc, 0, sizeof(u64) * ACC_WORDS, ctx->stream));
    CK(launch_stiefel_apply(n, A, V, nullptr, nullptr, out, ctx->acc, 1.0, grid, ctx->stream));
  }
  ctx->launches += 1;
  CK(cudaStreamSynchronize(ctx->stream));
  return OB200_OK;
}

int ob200_hvp(ob200_context *ctx, const ob200_operator *H, const double *v, double *out) {
  if (!ctx || !H || !v || !out) return OB200_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  const uint64_t N = H->n * H->p;
  cudaStream_t st = ctx->stream;
  if (H->kind == OB200_OP_DIAG) {
    CK(launch_hadamard(N, H->diag_dev, v, out, ctx->sm_count, st));
    ctx->launches += 1;
    return OB200_OK;
  }
  if (H->kind == OB200_OP_SPHERE_LOWRANK) {
    int rc = check_sphere(ctx, H);
    if (rc) return rc;
    if ((rc = ensure_vectors(ctx, N))) return rc;
    double *Av = ctx->Hp;
    if ((rc = sphere_Av(ctx, H->n, H->k, H->diag_dev, H->U_dev, H->ldu, H->sigma_host, v, Av))) return rc;
    double xAv = 0.0;                                       // x^T A v
    const double *aa[1] = {H->x_dev}, *bb[1] = {Av};
    if ((rc = dots_sync(ctx, N, 1, aa, bb, &xAv))) return rc;
    CK(launch_sphere_combine(N, Av, H->x_dev, v, xAv, H->xAx, out, ctx->sm_count, st));
    ctx->launches += 1;
    CK(cudaStreamSynchronize(st));
    return OB200_OK;
  }
  if (H->kind == OB200_OP_BLOCK_CSR3 || H->kind == OB200_OP_STENCIL7) {
    SparseArgs sp;
    int rc = sparse_args(ctx, H, &sp);
    if (rc) return rc;
    if (ctx->cm.world > 1) return fail(ctx, OB200_UNSUPPORTED, "stand-alone HVP of a sparse operator is single-GPU in this version");
    CK(launch_sparse_apply(N, sp, v, out, ctx->sm_count, st));
    ctx->launches += 1;
    return OB200_OK;
  }
  if (H->kind != OB200_OP_STIEFEL_BLOCKDIAG || H->p != 32) return fail(ctx, OB200_UNSUPPORTED, "operator kind");
  if (!H->A_bf16_dev || !H->Y_dev || !H->S_host) return fail(ctx, OB200_INVALID_ARGUMENT, "incomplete Stiefel operator");
  int rc = ensure_vectors(ctx, N);
  if (rc) return rc;
  const unsigned long long nblk = (H->n + 127) / 128;
  int grid = ctx->sm_count;
  if ((unsigned long long)grid > nblk) grid = (int)nblk;
  // <V,V> stays on the device: it bounds the exact fixed-point Gram of the projection.  One GPU + cached digit planes:
  // <V,V>, the content checksum of A and every clearing the persistent kernel needs come from ONE prologue launch
  // (hvp_prologue_kernel); otherwise the separate kernels (and the cross-rank fold of <V,V>).
  if (ctx->opt_tcgen05 && !(ctx->planes_key == H->A_bf16_dev && ctx->planes_n == H->n)) {
    if ((rc = ensure_planes(ctx, H->A_bf16_dev, H->n))) return rc;
  }
  const bool fused_prologue = ctx->opt_tcgen05 && ctx->planes_ok && ctx->cm.world == 1;
  if (!fused_prologue) {
    const double *aa[1] = {v}, *bb[1] = {v};
    CK(cudaMemsetAsync(ctx->acc, 0, sizeof(u64) * ACC_SCAL_WORDS, st));
    CK(launch_dots(N, 1, aa, bb, ctx->acc, ctx->sm_count, st));
    if ((rc = exchange(ctx, ctx->acc, 0, KUL_STRIDE))) return rc;
    CK(launch_finalize_many(ctx->acc, 1, ctx->dscal, st));
    ctx->launches += 2;
  }
  if (!(ctx->S_cache_valid && !memcmp(ctx->S_cache, H->S_host, sizeof(ctx->S_cache)))) {
    memcpy(ctx->S_cache, H->S_host, sizeof(ctx->S_cache));
    CK(cudaMemcpyAsync(ctx->dmat, ctx->S_cache, sizeof(double) * 1024, cudaMemcpyHostToDevice, st));
    ctx->S_cache_valid = true;
  }
  if (ctx->opt_tcgen05) {
    // One persistent launch (the two fused phases of a CG step: contraction + Gram | projection), no host round trip:
    // the digit planes are validated ON THE DEVICE against the content checksum of A (stale: rebuild and relaunch).
    for (int attempt = 0; attempt < 2 && ctx->planes_ok; ++attempt) {
      unsigned long long *sum_dev = ctx->dsum;
      if (fused_prologue) {
        const int par = ctx->hvp_par;
        ctx->hvp_par ^= 1;
        u64 *cur = ctx->hvp_ws + (size_t)par * (KUL_STRIDE + 1), *nxt = ctx->hvp_ws + (size_t)(par ^ 1) * (KUL_STRIDE + 1);
        sum_dev = reinterpret_cast<unsigned long long *>(cur + KUL_STRIDE);
        CK(launch_hvp_prologue(N, v, H->A_bf16_dev, nblk * (128ull * 128ull * 2ull / 16ull), cur, sum_dev, nxt,
                               reinterpret_cast<unsigned long long *>(nxt + KUL_STRIDE), ctx->acc,
                               (unsigned long long)ACC_SETS * ACC_WORDS, ctx->barrier,
                               reinterpret_cast<unsigned *>(ctx->hvp_ws + 2 * (KUL_STRIDE + 1)), ctx->dscal, ctx->sm_count, st));
        ctx->launches += 1;
      } else {
        CK(cudaMemsetAsync(ctx->dsum, 0, 8, st));
        CK(launch_stiefel_checksum(H->A_bf16_dev, nblk, ctx->dsum, ctx->sm_count, st));
        CK(cudaMemsetAsync(ctx->acc, 0, sizeof(u64) * ACC_SETS * ACC_WORDS, st));
        CK(cudaMemsetAsync(ctx->barrier, 0, 64, st));
        ctx->launches += 1;
      }
      TcgCommon a;
      memset(&a, 0, sizeof(a));
      a.N = N;
      a.g = ctx->dscal;             // device scalar <V,V>
      a.s = out;
      a.r = const_cast<double *>(v);
      a.p0 = ctx->p0; a.p1 = ctx->p1;
      a.Hp = ctx->Hp;
      a.Delta = 1.0; a.epsilon = 1e-8;
      a.acc = ctx->acc;
      a.barrier = ctx->barrier;
      a.abort_flag = reinterpret_cast<int *>(ctx->barrier + 1);
      a.result = ctx->dres;
      a.cm = ctx->cm;
      CK(launch_tcg_stiefel_tc(a, H->n, H->A_bf16_dev, H->Y_dev, ctx->dmat, H->op_norm_bound, ctx->planes,
                               ctx->plane_exp, grid, st, 1, sum_dev, ctx->planes_sum));
      ctx->launches += 1;
      CK(cudaMemcpyAsync(ctx->hres, ctx->dres, sizeof(TcgDeviceResult), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (ctx->hres->status == 6) {                 // A changed under the cached planes: rebuild, once
        ctx->planes_key = nullptr;
        if ((rc = ensure_planes(ctx, H->A_bf16_dev, H->n))) return rc;
        continue;
      }
      ctx->cm.epoch += ctx->hres->phases;
      if (ctx->hres->status == OB200_NUMERIC_RANGE) return fail(ctx, OB200_NUMERIC_RANGE, "fixed-point Gram bound exceeded or non-finite data");
      if (ctx->hres->status == OB200_ABORTED) return fail(ctx, OB200_ABORTED, "device grid barrier watchdog fired");
      return OB200_OK;
    }
  }
  // A not block-fixed-point (or tcgen05 switched off): fp64 tensor-core kernels, Gram symmetrised by the host
  double vv = 0.0;
  CK(cudaMemcpyAsync(&vv, ctx->dscal, sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int e = gram_exponent_host(H->op_norm_bound * std::sqrt(vv) * 4.0);
  CK(cudaMemsetAsync(ctx->acc, 0, sizeof(u64) * ACC_WORDS, st));
  CK(launch_stiefel_apply(H->n, H->A_bf16_dev, v, ctx->dmat, H->Y_dev, ctx->Hp, ctx->acc,
                          std::ldexp(1.0, 90 - e), grid, st));
  ctx->launches += 1;
  std::vector<double> G(1024);
  if ((rc = read_gram(ctx, e, G.data(), nullptr))) return rc;
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) ctx->hmat[i * 32 + j] = -0.5 * (G[i * 32 + j] + G[j * 32 + i]);
  CK(cudaMemcpyAsync(ctx->dmat + 1024, ctx->hmat, sizeof(double) * 1024, cudaMemcpyHostToDevice, st));
  CK(launch_stiefel_rowgemm(H->n, ctx->Hp, 1.0, H->Y_dev, ctx->dmat + 1024, out, grid, st));
  ctx->launches += 1;
  CK(cudaStreamSynchronize(st));
  return OB200_OK;
}

int ob200_csr3_model(ob200_context *ctx, uint64_t N, uint64_t r, const uint64_t *rowptr, const uint32_t *colidx,
                     const double *blocks, const double *X, double *lambda_dev, double *f, double *grad_dev) {
  if (!ctx || !rowptr || !colidx || !blocks || !X || !lambda_dev) return OB200_INVALID_ARGUMENT;
  if (r < 3 || r > 8) return fail(ctx, OB200_UNSUPPORTED, "rotation-synchronisation model requires 3 <= r <= 8");
  CK(cudaSetDevice(ctx->device));
  SparseArgs sp;
  memset(&sp, 0, sizeof(sp));
  sp.kind = OB200_OP_BLOCK_CSR3;
  sp.r = (int)r;
  sp.units = N;
  sp.rowptr = reinterpret_cast<const unsigned long long *>(rowptr);
  sp.colidx = colidx;
  sp.blocks = blocks;
  cudaStream_t st = ctx->stream;
  CK(cudaMemsetAsync(ctx->acc, 0, sizeof(u64) * ACC_SCAL_WORDS, st));
  CK(launch_csr3_model(sp, X, lambda_dev, grad_dev, ctx->acc, ctx->sm_count, st));
  {   // row-sharded: X holds own poses followed by the halo poses, f is the sum over the ranks
    int rc = exchange(ctx, ctx->acc, 0, KUL_STRIDE);
    if (rc) return rc;
  }
  CK(launch_finalize_many(ctx->acc, 1, ctx->dscal, st));
  ctx->launches += 2;
  CK(cudaMemcpyAsync(ctx->hscal, ctx->dscal, sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (f) *f = ctx->hscal[0];
  return OB200_OK;
}

int ob200_csr3_retract(ob200_context *ctx, uint64_t N, uint64_t r, const double *X, const double *V, double *out) {
  if (!ctx || !X || !V || !out) return OB200_INVALID_ARGUMENT;
  if (r < 3 || r > 8) return fail(ctx, OB200_UNSUPPORTED, "retraction on St(3,r)^N requires 3 <= r <= 8");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  int *bad_dev = reinterpret_cast<int *>(ctx->dbits);
  CK(cudaMemsetAsync(bad_dev, 0, sizeof(int), st));
  CK(launch_csr3_retract(N, (int)r, X, V, out, bad_dev, st));
  ctx->launches += 1;
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, bad_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (bad) return fail(ctx, OB200_NUMERIC_RANGE, "retraction: a pose block X_i + V_i is rank deficient or not finite");
  return OB200_OK;
}

int ob200_sphere_model(ob200_context *ctx, uint64_t n, uint64_t k, const double *d, const double *Ut, uint64_t ldu,
                       const double *sigma_host, const double *x, double *Ax_dev, double *xAx, double *grad_dev) {
  if (!ctx || !d || !x || !Ax_dev || !xAx || (k && (!Ut || !sigma_host))) return OB200_INVALID_ARGUMENT;
  if (k > 16) return fail(ctx, OB200_UNSUPPORTED, "sphere low-rank operator supports k <= 16");
  CK(cudaSetDevice(ctx->device));
  int rc = check_ldu(ctx, n, k, ldu, Ut);
  if (rc) return rc;
  if ((rc = sphere_Av(ctx, n, k, d, Ut, ldu, sigma_host, x, Ax_dev))) return rc;
  const double *aa[1] = {x}, *bb[1] = {Ax_dev};
  if ((rc = dots_sync(ctx, n, 1, aa, bb, xAx))) return rc;     // f(x) = x^T A x
  if (grad_dev) {                                               // grad f(x) = 2 (A x - (x^T A x) x)
    CK(launch_sphere_combine(n, Ax_dev, x, nullptr, *xAx, 0.0, grad_dev, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return OB200_OK;
}

int ob200_sphere_retract(ob200_context *ctx, uint64_t n, const double *x, const double *v, double *out) {
  if (!ctx || !x || !v || !out) return OB200_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  CK(launch_axpby(n, 1.0, x, 1.0, v, out, ctx->sm_count, ctx->stream));       // z = x + v
  ctx->launches += 1;
  double zz = 0.0;
  const double *aa[1] = {out}, *bb[1] = {out};
  int rc = dots_sync(ctx, n, 1, aa, bb, &zz);
  if (rc) return rc;
  if (!(zz > 0)) return fail(ctx, OB200_NUMERIC_RANGE, "retraction: x + v is zero or not finite");
  const double inv = 1.0 / std::sqrt(zz);
  CK(launch_axpby(n, inv, out, 0.0, nullptr, out, ctx->sm_count, ctx->stream));  // z / ||z||
  ctx->launches += 1;
  return OB200_OK;
}

int ob200_stiefel_retract(ob200_context *ctx, uint64_t n, uint64_t p, const double *Y, const double *V,
                          double *out) {
  if (!ctx || !Y || !V || !out) return OB200_INVALID_ARGUMENT;
  if (p != 32) return fail(ctx, OB200_UNSUPPORTED, "Stiefel retraction requires p == 32");
  CK(cudaSetDevice(ctx->device));
  const uint64_t N = n * p;
  int rc = ensure_vectors(ctx, N);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const unsigned long long nblk = (n + 127) / 128;
  int grid = ctx->sm_count;
  if ((unsigned long long)grid > nblk) grid = (int)nblk;
  double *Z = ctx->Hp;
  CK(launch_axpby(N, 1.0, Y, 1.0, V, Z, ctx->sm_count, st));   // Z = Y + V
  ctx->launches += 1;
  double zz = 0.0;
  const double *aa[1] = {Z}, *bb[1] = {Z};
  if ((rc = dots_sync(ctx, N, 1, aa, bb, &zz))) return rc;
  const int e = gram_exponent_host(zz * 2.0);                  // |(Z^T Z)_ij| <= ||Z||_F^2
  CK(cudaMemsetAsync(ctx->acc, 0, sizeof(u64) * ACC_WORDS, st));
  CK(launch_stiefel_gram(n, Z, Z, ctx->acc, std::ldexp(1.0, 90 - e), grid, st));
  ctx->launches += 1;
  std::vector<double> M(1024), R(1024, 0.0), Ri(1024, 0.0);
  if ((rc = read_gram(ctx, e, M.data(), nullptr))) return rc;
  const int P = 32;
  for (int i = 0; i < P; ++i)
    for (int j = i; j < P; ++j) { const double s = 0.5 * (M[i * P + j] + M[j * P + i]); M[i * P + j] = s; M[j * P + i] = s; }
  // upper Cholesky M = R^T R
  for (int j = 0; j < P; ++j) {
    double s = M[j * P + j];
    for (int k = 0; k < j; ++k) s -= R[k * P + j] * R[k * P + j];
    if (!(s > 0)) return fail(ctx, OB200_NUMERIC_RANGE, "retraction: Y + V is rank deficient");
    const double rjj = std::sqrt(s);
    R[j * P + j] = rjj;
    for (int i = j + 1; i < P; ++i) {
      double t = M[j * P + i];
      for (int k = 0; k < j; ++k) t -= R[k * P + j] * R[k * P + i];
      R[j * P + i] = t / rjj;
    }
  }
  // Ri = R^{-1} (upper triangular), column by column: R * Ri = I
  for (int c = 0; c < P; ++c) {
    for (int i = c; i >= 0; --i) {
      double t = (i == c) ? 1.0 : 0.0;
      for (int k = i + 1; k <= c; ++k) t -= R[i * P + k] * Ri[k * P + c];
      Ri[i * P + c] = t / R[i * P + i];
    }
  }
  for (int i = 0; i < 1024; ++i) ctx->hmat[i] = Ri[i];
  CK(cudaMemcpyAsync(ctx->dmat + 1024, ctx->hmat, sizeof(double) * 1024, cudaMemcpyHostToDevice, st));
  CK(launch_stiefel_rowgemm(n, nullptr, 0.0, Z, ctx->dmat + 1024, out, grid, st));   // Q = Z R^{-1}
  ctx->launches += 1;
  CK(cudaStreamSynchronize(st));
  return OB200_OK;
}

int ob200_debug_sym_eig32(const double *S, double *Q, double *lam) {
  if (!S || !Q || !lam) return OB200_INVALID_ARGUMENT;
  sym_eig32(S, Q, lam);
  return OB200_OK;
}

int ob200_stiefel_project(ob200_context *ctx, uint64_t n, uint64_t p, const double *Y, const double *Z, double *out) {
  if (!ctx || !Y || !Z || !out) return OB200_INVALID_ARGUMENT;
  if (p != 32) return fail(ctx, OB200_UNSUPPORTED, "Stiefel projection requires p == 32");
  CK(cudaSetDevice(ctx->device));
  const uint64_t N = n * p;
  int rc = ensure_vectors(ctx, N);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const unsigned long long nblk = (n + 127) / 128;
  int grid = ctx->sm_count;
  if ((unsigned long long)grid > nblk) grid = (int)nblk;
  double zz = 0.0;
  const double *aa[1] = {Z}, *bb[1] = {Z};
  if ((rc = dots_sync(ctx, N, 1, aa, bb, &zz))) return rc;
  const int e = gram_exponent_host(std::sqrt(zz) * 2.0);       // |(Y^T Z)_ij| <= ||y_i|| ||z_j|| <= ||Z||_F
  CK(cudaMemsetAsync(ctx->acc, 0, sizeof(u64) * ACC_WORDS, st));
  CK(launch_stiefel_gram(n, Y, Z, ctx->acc, std::ldexp(1.0, 90 - e), grid, st));
  ctx->launches += 1;
  std::vector<double> G(1024);
  if ((rc = read_gram(ctx, e, G.data(), nullptr))) return rc;
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) ctx->hmat[i * 32 + j] = -0.5 * (G[i * 32 + j] + G[j * 32 + i]);
  CK(cudaMemcpyAsync(ctx->dmat + 1024, ctx->hmat, sizeof(double) * 1024, cudaMemcpyHostToDevice, st));
  CK(launch_stiefel_rowgemm(n, Z, 1.0, Y, ctx->dmat + 1024, out, grid, st));   // Z - Y sym(Y^T Z)
  ctx->launches += 1;
  CK(cudaStreamSynchronize(st));
  return OB200_OK;
}

}  // extern "C"

// ---- LOBPCG ------------------------------------------------------------------------------------------------
static int block_apply(ob200_context *ctx, const ob200_block_operator *Op, uint64_t m, int k, const double *in, int ldi,
                       double *out, int ldo) {
  cudaStream_t st = ctx->stream;
  switch (Op->kind) {
    case OB200_BLK_DIAG:
      if (!Op->diag_dev) return fail(ctx, OB200_INVALID_ARGUMENT, "diagonal block operator without diagonal");
      CK(launch_blk_diag(m, k, Op->diag_dev, 0.0, in, ldi, out, ldo, ctx->sm_count, st));
      break;
    case OB200_BLK_SCALAR:
      CK(launch_blk_diag(m, k, nullptr, Op->alpha, in, ldi, out, ldo, ctx->sm_count, st));
      break;
    case OB200_BLK_STENCIL7:
      if ((uint64_t)Op->gx * Op->gy * Op->gz != m) return fail(ctx, OB200_INVALID_ARGUMENT, "stencil grid does not match m");
      CK(launch_blk_stencil7(Op->gx, Op->gy, Op->gz, k, in, ldi, out, ldo, ctx->sm_count, st));
      break;
    default:
      return fail(ctx, OB200_UNSUPPORTED, "unknown block operator kind");
  }
  ctx->launches += 1;
  return OB200_OK;
}

namespace {
struct Slab {   // carves 256-byte aligned buffers out of the context's LOBPCG workspace; pass 1 sizes, pass 2 hands out
  unsigned char *base = nullptr;
  size_t off = 0;
  template <class T> void get(T **p, size_t count) {
    const size_t bytes = (sizeof(T) * (count ? count : 1) + 255) & ~(size_t)255;
    *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
    off += bytes;
  }
};
}  // namespace

// Rayleigh-Ritz (LOBPCG.h:53-62) on the device: GA, GB (ns x ns) -> theta (ascending), C (row-major, C^T GB C = I)
static int rayleigh_ritz(ob200_context *ctx, int ns, const double *GA, const double *GB, double *EA, double *EB, double *D,
                         double *theta, double *C, double *work, int lwork, int *info_dev) {
  cudaStream_t st = ctx->stream;
  CK(launch_rr_equilibrate(ns, GA, GB, EA, EB, D, st));
  if (cusolverDnDsygvd(ctx->solver, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, ns, EA, ns, EB, ns,
                       theta, work, lwork, info_dev) != CUSOLVER_STATUS_SUCCESS)
    return fail(ctx, OB200_CUDA_ERROR, "cusolverDnDsygvd failed");
  CK(launch_rr_scale_transpose(ns, EA, D, C, st));
  ctx->launches += 2;
  int info = 0;
  CK(cudaMemcpyAsync(&info, info_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (info != 0) return fail(ctx, OB200_NUMERIC_RANGE, "Rayleigh-Ritz pencil is not positive definite (S^T B S singular)");
  return OB200_OK;
}

extern "C" int ob200_block_apply(ob200_context *ctx, const ob200_block_operator *Op, uint64_t m, uint64_t k,
                                 const double *in, uint64_t ldi, double *out, uint64_t ldo) {
  if (!ctx || !Op || !in || !out || k == 0 || ldi < k || ldo < k) return OB200_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  return block_apply(ctx, Op, m, (int)k, in, (int)ldi, out, (int)ldo);
}

extern "C" int ob200_lobpcg(ob200_context *ctx, const ob200_block_operator *A, const ob200_block_operator *B,
                            const ob200_block_operator *T, uint64_t m, uint64_t nx64, double *X, uint64_t nev,
                            uint64_t max_iters, double tau, const double *Omega, double *theta_host, uint64_t *num_iters,
                            uint64_t *num_converged) {
  if (!ctx || !A || !X || !theta_host || !num_iters || !num_converged) return OB200_INVALID_ARGUMENT;
  // reference LOBPCG.h:148-155
  if (nev > nx64) return fail(ctx, OB200_INVALID_ARGUMENT, "Block size nx must be greater than or equal to the number nev of desired eigenpairs");
  if (nx64 > m) return fail(ctx, OB200_INVALID_ARGUMENT, "Block size nx must be less than or equal to the dimension m of the problem");
  if (nx64 == 0 || nx64 > 64) return fail(ctx, OB200_UNSUPPORTED, "block size nx must be in 1..64");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (!ctx->solver) {
    if (cusolverDnCreate(&ctx->solver) != CUSOLVER_STATUS_SUCCESS) return fail(ctx, OB200_CUDA_ERROR, "cusolverDnCreate failed");
  }
  cusolverDnSetStream(ctx->solver, st);
  const int nx = (int)nx64, nsmax = 3 * nx, nb = ctx->sm_count;
  const size_t mnx = (size_t)m * nx;
  int rc;
  // Row sharding (one process per GPU, ob200_comm_connect + ob200_halo_create / ob200_halo_connect): every rank holds a
  // contiguous block of rows (for the Laplacian: a slab of z-planes, A->gz = LOCAL planes); Grams, norms and residual
  // norms are all-reduced in rank order through the halo buffers, the Rayleigh-Ritz step runs replicated on identical
  // data, the stencil exchanges one ghost plane with each neighbouring rank before every apply.
  const int world = ctx->cm.world, rank = ctx->cm.rank;
  const bool sharded = world > 1;
  const bool slab = sharded && A->kind == OB200_BLK_STENCIL7;
  const size_t plane_rows = slab ? (size_t)A->gx * A->gy : 0;
  const size_t ghost = plane_rows * (size_t)nsmax;                         // doubles of one ghost plane of a basis buffer
  const size_t AR_MAX = (size_t)nsmax * nsmax + 64;
  const size_t HDR = 512;                                                  // doubles reserved for flags at the buffer head
  if (sharded) {
    if ((B && B->kind == OB200_BLK_STENCIL7) || (T && T->kind == OB200_BLK_STENCIL7))
      return fail(ctx, OB200_UNSUPPORTED, "row-sharded LOBPCG: only A may be the stencil operator");
    const size_t need = sizeof(double) * (HDR + 2 * (size_t)world * AR_MAX + 4 * plane_rows * (size_t)nsmax);
    if (!ctx->halo_buf || ctx->halo_bytes < need || !ctx->halo_peer[world - 1]) {
      ctx->err = "row-sharded LOBPCG: halo buffer of at least " + std::to_string(need) +
                 " bytes required (ob200_halo_create / ob200_halo_connect)";
      return OB200_INVALID_ARGUMENT;
    }
  }
  int *abort_dev = reinterpret_cast<int *>(ctx->barrier + 1);
  auto allreduce = [&](double *buf, int count) -> int {                    // sum over the ranks, in rank order, in place
    if (!sharded) return OB200_OK;
    double *regions[MAX_RANKS];
    const size_t par = (size_t)(ctx->ar_seq & 1ull);
    for (int q = 0; q < world; ++q) regions[q] = static_cast<double *>(ctx->halo_peer[q]) + HDR + par * world * AR_MAX;
    CK(launch_lob_allreduce(ctx->cm, ctx->cm.epoch, regions, AR_MAX, buf, count, abort_dev, st));
    ctx->cm.epoch += 1;
    ctx->ar_seq += 1;
    ctx->launches += 1;
    return OB200_OK;
  };
  auto apply_A = [&](int k, double *in, int ldi, double *out, int ldo) -> int {   // `in`: a basis buffer (ghost planes around it)
    if (!slab) return block_apply(ctx, A, m, k, in, ldi, out, ldo);
    const int has_lo = rank > 0, has_hi = rank + 1 < world;
    ctx->plane_seq += 1;
    const size_t par = (size_t)(ctx->plane_seq & 1ull), plane = plane_rows * (size_t)nsmax;
    auto region = [&](int q, int dir) {   // dir 0: "from below" (filled by q-1), 1: "from above" (filled by q+1)
      return static_cast<double *>(ctx->halo_peer[q]) + HDR + 2 * (size_t)world * AR_MAX + (2 * par + dir) * plane;
    };
    auto flag = [&](int q, int dir) { return reinterpret_cast<unsigned long long *>(ctx->halo_peer[q]) + 8 + dir; };
    PlaneXchg px;
    px.lo_dst = has_lo ? region(rank - 1, 1) : nullptr;
    px.hi_dst = has_hi ? region(rank + 1, 0) : nullptr;
    px.lo_flag = has_lo ? flag(rank - 1, 1) : nullptr;
    px.hi_flag = has_hi ? flag(rank + 1, 0) : nullptr;
    px.from_below = region(rank, 0);
    px.from_above = region(rank, 1);
    px.my_flags = flag(rank, 0);
    px.seq = ctx->plane_seq;
    px.counter = reinterpret_cast<unsigned *>(ctx->halo_buf) + 0;            // first word of my own header
    CK(launch_lob_plane_exchange(px, in, ldi, k, plane_rows, A->gz, has_lo, has_hi, abort_dev, ctx->sm_count, st));
    CK(launch_blk_stencil7_slab(A->gx, A->gy, A->gz, k, in, ldi, out, ldo, has_lo, has_hi, ctx->sm_count, st));
    ctx->launches += 3;
    return OB200_OK;
  };

  double *S = nullptr, *S2 = nullptr, *AS = nullptr, *BS = nullptr, *AX = nullptr, *BX = nullptr, *R = nullptr, *P = nullptr,
         *tmp = nullptr;
  double *GA = nullptr, *GB = nullptr, *EA = nullptr, *EB = nullptr, *D = nullptr, *theta = nullptr, *C = nullptr,
         *partial = nullptr, *norms2 = nullptr, *work = nullptr;
  int *info_dev = nullptr;
  const size_t nn = (size_t)nsmax * nsmax;
  int lwork = 0;
  if (cusolverDnDsygvd_bufferSize(ctx->solver, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, nsmax,
                                  ctx->dmat, nsmax, ctx->dmat, nsmax, ctx->dmat, &lwork) != CUSOLVER_STATUS_SUCCESS)   // size query only
    return fail(ctx, OB200_CUDA_ERROR, "cusolverDnDsygvd_bufferSize failed");
  for (int pass = 0; pass < 2; ++pass) {
    Slab mem;
    mem.base = pass ? ctx->lob_ws : nullptr;
    mem.get(&S, (size_t)m * nsmax + 2 * ghost);      // (+ one ghost plane below and above when the grid is sharded in z)
    mem.get(&S2, (size_t)m * nsmax + 2 * ghost);     // the update writes the new X straight into the next basis (no copies of X)
    mem.get(&AS, (size_t)m * nsmax);
    if (B) mem.get(&BS, (size_t)m * nsmax);
    mem.get(&AX, mnx);
    if (B) mem.get(&BX, mnx);
    mem.get(&R, mnx);
    mem.get(&P, mnx);
    mem.get(&tmp, mnx);
    mem.get(&GA, nn); mem.get(&GB, nn); mem.get(&EA, nn); mem.get(&EB, nn); mem.get(&C, nn);
    mem.get(&D, (size_t)nsmax); mem.get(&theta, (size_t)nsmax);
    mem.get(&partial, (size_t)nb * nn + (size_t)nb * 16 * nx);   // Gram partials (nb sets) / residual partials (8 nb sets)
    mem.get(&norms2, (size_t)2 * nx + 8);
    mem.get(&info_dev, (size_t)1);
    mem.get(&work, (size_t)lwork);
    if (!pass && mem.off > ctx->lob_cap) {
      cudaFree(ctx->lob_ws);
      ctx->lob_ws = nullptr;
      ctx->lob_cap = 0;
      CK(cudaMalloc(&ctx->lob_ws, mem.off));
      ctx->lob_cap = mem.off;
    }
  }
  S += ghost;      // interior of the basis buffers
  S2 += ghost;
  const size_t rowX = sizeof(double) * nx, rowS = sizeof(double) * nsmax;

  std::vector<double> h(2 * nx + 8), th(nsmax);
  auto frob = [&](const double *V, double *out) -> int {   // ||V||_F of an m x nx block
    CK(launch_blk_sumsq(mnx, V, partial, nb, norms2, st));
    ctx->launches += 2;
    { int rca = allreduce(norms2, 1); if (rca) return rca; }
    CK(cudaMemcpyAsync(h.data(), norms2, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *out = std::sqrt(h[0]);
    return OB200_OK;
  };

  // operator norm estimates (l.172-181) from the probe block
  const double *Om = Omega ? Omega : X;
  double nOm = 0, nA = 0, nB = 0;
  if ((rc = frob(Om, &nOm))) return rc;
  CK(cudaMemcpy2DAsync(S, rowS, Om, rowX, rowX, m, cudaMemcpyDeviceToDevice, st));        // (the stencil reads a basis buffer)
  if ((rc = apply_A(nx, S, nsmax, tmp, nx))) return rc;
  if ((rc = frob(tmp, &nA))) return rc;
  const double A2normest = nA / nOm;
  double B2normest = 1.0;
  if (B) {
    if ((rc = block_apply(ctx, B, m, nx, Om, nx, tmp, nx))) return rc;
    if ((rc = frob(tmp, &nB))) return rc;
    B2normest = nB / nOm;
  }

  // initial Rayleigh-Ritz on span(X0) (l.186-198)
  CK(cudaMemcpy2DAsync(S, rowS, X, rowX, rowX, m, cudaMemcpyDeviceToDevice, st));            // l.210 (first iteration)
  if ((rc = apply_A(nx, S, nsmax, AX, nx))) return rc;
  const double *BXp = X;
  if (B) { if ((rc = block_apply(ctx, B, m, nx, X, nx, BX, nx))) return rc; BXp = BX; }
  CK(launch_blk_gram(m, X, nx, nx, AX, nx, nx, partial, nb, GA, st));
  CK(launch_blk_gram(m, X, nx, nx, BXp, nx, nx, partial, nb, GB, st));
  ctx->launches += 4;
  if ((rc = allreduce(GA, nx * nx))) return rc;
  if ((rc = allreduce(GB, nx * nx))) return rc;
  if ((rc = rayleigh_ritz(ctx, nx, GA, GB, EA, EB, D, theta, C, work, lwork, info_dev))) return rc;
  CK(launch_blk_gemm(m, AX, nx, nx, C, nx, nx, tmp, nx, nb, st));           // AX <- AX C
  CK(cudaMemcpyAsync(AX, tmp, sizeof(double) * mnx, cudaMemcpyDeviceToDevice, st));
  CK(launch_blk_gemm(m, BXp, nx, nx, C, nx, nx, tmp, nx, nb, st));          // BX <- BX C  (B absent: X C, kept in tmp)
  ctx->launches += 2;
  const double *BXc = tmp;
  if (B) { CK(cudaMemcpyAsync(BX, tmp, sizeof(double) * mnx, cudaMemcpyDeviceToDevice, st)); BXc = BX; }
  CK(launch_blk_residual(m, nx, AX, BXc, nx, X, nx, theta, R, partial, nb, norms2, st));
  ctx->launches += 2;

  uint64_t nc = 0, it = 1;
  // the current eigenvector block lives in the first nx columns of the current basis buffer (leading dimension nsmax)
  for (it = 1; it < max_iters; ++it) {
    const int act = nx - (int)nc;                                                           // soft locking
    if (T) {   // l.207 + l.213: W = T(R), active columns written straight into the basis
      if ((rc = block_apply(ctx, T, m, act, R + nc, nx, S + nx, nsmax))) return rc;
    } else {
      CK(cudaMemcpy2DAsync(S + nx, rowS, R + nc, rowX, sizeof(double) * act, m, cudaMemcpyDeviceToDevice, st));
    }
    int ns = 2 * nx - (int)nc;
    if (it > 1) {
      CK(cudaMemcpy2DAsync(S + ns, rowS, P + nc, rowX, sizeof(double) * act, m, cudaMemcpyDeviceToDevice, st));  // l.217
      ns = 3 * nx - 2 * (int)nc;
    }
    if ((rc = apply_A(ns, S, nsmax, AS, nsmax))) return rc;                                 // l.225
    const double *BSp = S;
    if (B) { if ((rc = block_apply(ctx, B, m, ns, S, nsmax, BS, nsmax))) return rc; BSp = BS; }   // l.226
    CK(launch_blk_gram(m, S, nsmax, ns, AS, nsmax, ns, partial, nb, GA, st));               // l.229
    CK(launch_blk_gram(m, S, nsmax, ns, BSp, nsmax, ns, partial, nb, GB, st));              // l.230
    ctx->launches += 4;
    if ((rc = allreduce(GA, ns * ns))) return rc;
    if ((rc = allreduce(GB, ns * ns))) return rc;
    if ((rc = rayleigh_ritz(ctx, ns, GA, GB, EA, EB, D, theta, C, work, lwork, info_dev))) return rc;   // l.233
    CK(launch_blk_update(m, S, nsmax, ns, nx, C, ns, S2, nsmax, P, nx, nb, st));   // l.239 X = S C(:, 1:nx) -> next basis; l.249 P
    ctx->launches += 1;
    std::swap(S, S2);                                                                       // X = S[:, 0:nx] from here on
    if ((rc = apply_A(nx, S, nsmax, AX, nx))) return rc;                                    // l.242
    const double *BXq = S;
    int ldbx = nsmax;
    if (B) { if ((rc = block_apply(ctx, B, m, nx, S, nsmax, BX, nx))) return rc; BXq = BX; ldbx = nx; }   // l.243
    CK(launch_blk_residual(m, nx, AX, BXq, ldbx, S, nsmax, theta, R, partial, nb, norms2, st));   // l.246, 254
    ctx->launches += 2;
    if ((rc = allreduce(norms2, 2 * nx))) return rc;
    CK(cudaMemcpyAsync(h.data(), norms2, sizeof(double) * 2 * nx, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(th.data(), theta, sizeof(double) * nx, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (nc = 0; nc < nev; ++nc) {                                                          // l.257-269
      const double r = std::sqrt(h[nc]), xn = std::sqrt(h[nx + nc]);
      const double tol = tau * (A2normest + B2normest * std::fabs(th[nc])) * xn;
      if (!(r <= tol)) break;
    }
    if (nc == nev) break;                                                                   // l.277
  }
  CK(cudaMemcpy2DAsync(X, rowX, S, rowS, rowX, m, cudaMemcpyDeviceToDevice, st));            // hand the eigenvector block back
  CK(cudaMemcpyAsync(th.data(), theta, sizeof(double) * nx, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (uint64_t i = 0; i < nev; ++i) theta_host[i] = th[i];
  *num_iters = it;
  *num_converged = nc;
  return OB200_OK;
}
