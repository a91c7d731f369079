// C-ABI glue: error state, launch bookkeeping, argument checks and the composed
// "acquire" entry points (pairs -> posterior -> acquisition) with host or device buffers.
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <cstdio>
#include <cstring>
#include "sot_common.cuh"
#include "launch.cuh"

namespace bosot {

static thread_local char t_error[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* msg) {
  std::snprintf(t_error, sizeof(t_error), "%s", msg ? msg : "");
  return code;
}

int cuda_check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return BOSOT_OK;
  std::snprintf(t_error, sizeof(t_error), "%s: %s", what, cudaGetErrorString(e));
  return BOSOT_ECUDA;
}

int after_launch(const char* kernel_name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cuda_check(cudaGetLastError(), kernel_name);
}

int check_grammar(const bosot_grammar* g) {
  if (g->n_symbols < 1 || g->n_symbols > BOSOT_MAX_SYMBOLS) return set_error(BOSOT_EINVAL, "grammar: n_symbols out of range");
  if (g->n_blocks < 1 || g->n_blocks > BOSOT_MAX_BLOCKS) return set_error(BOSOT_EINVAL, "grammar: n_blocks out of range");
  if (g->n_dim_cols < 1 || g->n_dim_cols > BOSOT_MAX_DIM_COLS) return set_error(BOSOT_EINVAL, "grammar: n_dim_cols out of range");
  for (int s = 0; s < g->n_symbols; ++s) {
    if (g->sym_block[s] >= g->n_blocks || g->sym_col[s] >= g->n_dim_cols) return set_error(BOSOT_EINVAL, "grammar: symbol table out of range");
    if ((g->sym_block[s] < 0) != (g->sym_col[s] < 0)) return set_error(BOSOT_EINVAL, "grammar: block/column must both be set or both be -1");
  }
  for (int b = 0; b < g->n_blocks; ++b)
    if (g->block_null_col[b] < 0 || g->block_null_col[b] >= g->n_dim_cols) return set_error(BOSOT_EINVAL, "grammar: NULL column out of range");
  return BOSOT_OK;
}

// defined in sot_pairs.cu / posterior.cu
int launch_pairs(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar, const void* d_eval_blob,
                 int n_eval, const double weights[3], double variance, double lengthscale, double* d_K, double* d_D,
                 double* d_Df, int64_t ld, int32_t* d_status, void* d_scratch, size_t scratch_bytes, cudaStream_t stream);
size_t pairs_scratch_bytes(int64_t n_cand);
int launch_posterior(const double* d_Kstar, int64_t n_cand, int64_t ld, int n_eval, const double* d_Linv, int ldl,
                     const double* d_beta, double mean_c, double variance, int acq_kind, const double* d_current_max,
                     double xi, double ucb_beta, double* d_mu, double* d_sigma, double* d_acq, cudaStream_t stream,
                     double* const* d_peers = nullptr, int n_peers = 0, int64_t peer_off = 0);

// work buffer of the composed calls:  K*[n_cand][ld] | trees[n_cand][64] | mu | sigma | acq | status | scan records
struct WorkLayout {
  int64_t ld;
  size_t off_K, off_trees, off_mu, off_sigma, off_acq, off_status, off_recs, total;
};
static WorkLayout work_layout(int64_t n_cand, int n_eval) {
  WorkLayout W;
  W.ld = (n_eval + 7) & ~7;  // 64-byte aligned rows
  size_t o = 0;
  W.off_K = o;      o = align16(o + sizeof(double) * (size_t)n_cand * W.ld);
  W.off_trees = o;  o = align16(o + (size_t)n_cand * kTreeBytes);
  W.off_mu = o;     o = align16(o + sizeof(double) * (size_t)n_cand);
  W.off_sigma = o;  o = align16(o + sizeof(double) * (size_t)n_cand);
  W.off_acq = o;    o = align16(o + sizeof(double) * (size_t)n_cand);
  W.off_status = o; o = align16(o + sizeof(int32_t) * (size_t)n_cand);
  W.off_recs = o;   o = align16(o + pairs_scratch_bytes(n_cand) + 16 * 16);  // + one 16-byte scratch head per chunk of the pipelined path
  W.total = o;
  return W;
}

}  // namespace bosot

using namespace bosot;

extern "C" const char* bosot_last_error(void) { return t_error; }
extern "C" int bosot_abi_version(void) { return BOSOT_ABI_VERSION; }
extern "C" int64_t bosot_kernel_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

extern "C" int bosot_posterior_acq(const double* d_Kstar, int64_t n_cand, int64_t ld, int n_eval,
                                   const double* d_Linv, int ldl, const double* d_beta, double mean_c, double variance,
                                   int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                                   double* d_mu, double* d_sigma, double* d_acq, void* stream) {
  return launch_posterior(d_Kstar, n_cand, ld, n_eval, d_Linv, ldl, d_beta, mean_c, variance, acq_kind, d_current_max,
                          xi, ucb_beta, d_mu, d_sigma, d_acq, (cudaStream_t)stream);
}

extern "C" int bosot_posterior_acq_scatter(const double* d_Kstar, int64_t n_cand, int64_t ld, int n_eval,
                                           const double* d_Linv, int ldl, const double* d_beta, double mean_c, double variance,
                                           int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                                           double* d_mu, double* d_sigma, double* d_acq,
                                           const void* d_peer_ptrs, int n_peers, int64_t peer_offset, void* stream) {
  if (!d_peer_ptrs) return set_error(BOSOT_EINVAL, "bosot_posterior_acq_scatter: null peer pointer table");
  return launch_posterior(d_Kstar, n_cand, ld, n_eval, d_Linv, ldl, d_beta, mean_c, variance, acq_kind, d_current_max,
                          xi, ucb_beta, d_mu, d_sigma, d_acq, (cudaStream_t)stream,
                          static_cast<double* const*>(d_peer_ptrs), n_peers, peer_offset);
}

extern "C" size_t bosot_acquire_work_bytes(int64_t n_cand, int n_eval) {
  if (n_cand < 0 || n_eval < 1 || n_eval > BOSOT_MAX_EVAL) return 0;
  return work_layout(n_cand, n_eval).total;
}

// helper streams / events for the pipelined host entry point (created lazily, one set per device, never destroyed)
namespace bosot {
constexpr int kMaxChunks = 16;
static_assert(kMaxChunks == 16, "work_layout() reserves one 16-byte scratch head per chunk: 16 * 16 bytes");
struct HostPipe {
  bool ready = false;
  cudaStream_t h2d = nullptr, d2h = nullptr, alt = nullptr;  // alt: second compute stream (odd chunks)
  cudaEvent_t start = nullptr, done = nullptr, copied[kMaxChunks] = {}, computed[kMaxChunks] = {};
};
static HostPipe g_pipe[64];
static std::mutex g_pipe_mutex;  // the helper events are per device: serialise the (short) enqueue section

static int host_pipe(int dev, HostPipe** out) {
  if (dev < 0 || dev >= 64) return set_error(BOSOT_EINVAL, "bosot_acquire_host: device index out of range");
  HostPipe& P = g_pipe[dev];
  if (!P.ready) {
    if (int rc = cuda_check(cudaStreamCreateWithFlags(&P.h2d, cudaStreamNonBlocking), "cudaStreamCreate(h2d)")) return rc;
    if (int rc = cuda_check(cudaStreamCreateWithFlags(&P.d2h, cudaStreamNonBlocking), "cudaStreamCreate(d2h)")) return rc;
    if (int rc = cuda_check(cudaStreamCreateWithFlags(&P.alt, cudaStreamNonBlocking), "cudaStreamCreate(alt)")) return rc;
    if (int rc = cuda_check(cudaEventCreateWithFlags(&P.start, cudaEventDisableTiming), "cudaEventCreate")) return rc;
    if (int rc = cuda_check(cudaEventCreateWithFlags(&P.done, cudaEventDisableTiming), "cudaEventCreate")) return rc;
    for (int k = 0; k < kMaxChunks; ++k) {
      if (int rc = cuda_check(cudaEventCreateWithFlags(&P.copied[k], cudaEventDisableTiming), "cudaEventCreate")) return rc;
      if (int rc = cuda_check(cudaEventCreateWithFlags(&P.computed[k], cudaEventDisableTiming), "cudaEventCreate")) return rc;
    }
    P.ready = true;
  }
  *out = &P;
  return BOSOT_OK;
}
}  // namespace bosot

extern "C" int bosot_acquire_device(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar,
                                    const void* d_eval_blob, int n_eval, const double weights[3],
                                    double variance, double lengthscale,
                                    const double* d_Linv, int ldl, const double* d_beta, double mean_c,
                                    int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                                    void* d_work, size_t work_bytes,
                                    double* d_mu, double* d_sigma, double* d_acq, void* stream) {
  if (!d_work) return set_error(BOSOT_EINVAL, "bosot_acquire_device: null work buffer");
  if (n_cand < 0 || n_eval < 1 || n_eval > BOSOT_MAX_EVAL) return set_error(BOSOT_EINVAL, "bosot_acquire_device: bad sizes");
  const WorkLayout W = work_layout(n_cand, n_eval);
  if (work_bytes < W.total) return set_error(BOSOT_EINVAL, "bosot_acquire_device: work buffer too small");
  uint8_t* w = static_cast<uint8_t*>(d_work);
  double* K = reinterpret_cast<double*>(w + W.off_K);
  int32_t* status = reinterpret_cast<int32_t*>(w + W.off_status);
  cudaStream_t st = (cudaStream_t)stream;
  // one set of launches: cutting a device-resident pool into chunks on two streams was measured slower (4.07 vs 3.95 ms)
  if (int rc = launch_pairs(d_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance, lengthscale, K, nullptr,
                            nullptr, W.ld, status, w + W.off_recs, pairs_scratch_bytes(n_cand), st)) return rc;
  return launch_posterior(K, n_cand, W.ld, n_eval, d_Linv, ldl, d_beta, mean_c, variance, acq_kind, d_current_max, xi,
                          ucb_beta, d_mu, d_sigma, d_acq, st);
}

extern "C" int bosot_acquire_device_scatter(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar,
                                            const void* d_eval_blob, int n_eval, const double weights[3],
                                            double variance, double lengthscale,
                                            const double* d_Linv, int ldl, const double* d_beta, double mean_c,
                                            int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                                            void* d_work, size_t work_bytes,
                                            double* d_mu, double* d_sigma, double* d_acq,
                                            const void* d_peer_ptrs, int n_peers, int64_t peer_offset, void* stream) {
  if (!d_work || !d_peer_ptrs) return set_error(BOSOT_EINVAL, "bosot_acquire_device_scatter: null argument");
  if (n_cand < 0 || n_eval < 1 || n_eval > BOSOT_MAX_EVAL) return set_error(BOSOT_EINVAL, "bosot_acquire_device_scatter: bad sizes");
  const WorkLayout W = work_layout(n_cand, n_eval);
  if (work_bytes < W.total) return set_error(BOSOT_EINVAL, "bosot_acquire_device_scatter: work buffer too small");
  uint8_t* w = static_cast<uint8_t*>(d_work);
  double* K = reinterpret_cast<double*>(w + W.off_K);
  int32_t* status = reinterpret_cast<int32_t*>(w + W.off_status);
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = launch_pairs(d_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance, lengthscale, K, nullptr,
                            nullptr, W.ld, status, w + W.off_recs, pairs_scratch_bytes(n_cand), st)) return rc;
  return launch_posterior(K, n_cand, W.ld, n_eval, d_Linv, ldl, d_beta, mean_c, variance, acq_kind, d_current_max, xi,
                          ucb_beta, d_mu, d_sigma, d_acq, st, static_cast<double* const*>(d_peer_ptrs), n_peers, peer_offset);
}

// Host buffers in, host buffers out.  Large pools are cut into chunks that flow through a three-stage
// pipeline on three streams -- H2D of chunk k+1 (copy engine) | kernels of chunk k (caller's stream) |
// D2H of chunk k-1 (second copy engine) -- joined back into the caller's stream with events, so the call
// keeps its contract: asynchronous, ordered on `stream`, the caller synchronises `stream`.
extern "C" int bosot_acquire_host(const uint8_t* h_trees, int64_t n_cand, const bosot_grammar* grammar,
                                  const void* d_eval_blob, int n_eval, const double weights[3],
                                  double variance, double lengthscale,
                                  const double* d_Linv, int ldl, const double* d_beta, double mean_c,
                                  int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                                  void* d_work, size_t work_bytes,
                                  double* h_mu, double* h_sigma, double* h_acq, void* stream) {
  if (!h_trees || !d_work) return set_error(BOSOT_EINVAL, "bosot_acquire_host: null argument");
  if (n_cand < 0 || n_eval < 1 || n_eval > BOSOT_MAX_EVAL) return set_error(BOSOT_EINVAL, "bosot_acquire_host: bad sizes");
  const WorkLayout W = work_layout(n_cand, n_eval);
  if (work_bytes < W.total) return set_error(BOSOT_EINVAL, "bosot_acquire_host: work buffer too small");
  if (n_cand == 0) return BOSOT_OK;
  uint8_t* w = static_cast<uint8_t*>(d_work);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* d_trees = w + W.off_trees;
  double* d_K = reinterpret_cast<double*>(w + W.off_K);
  double* d_mu = reinterpret_cast<double*>(w + W.off_mu);
  double* d_sigma = reinterpret_cast<double*>(w + W.off_sigma);
  double* d_acq = reinterpret_cast<double*>(w + W.off_acq);
  int32_t* d_status = reinterpret_cast<int32_t*>(w + W.off_status);
  uint8_t* d_recs = w + W.off_recs;

  if (n_cand < 400000) {
    // small pool: nothing to overlap, keep everything on the caller's stream (no cross-stream event hops)
    if (int rc = cuda_check(cudaMemcpyAsync(d_trees, h_trees, (size_t)n_cand * kTreeBytes, cudaMemcpyHostToDevice, st), "H2D trees")) return rc;
    if (int rc = launch_pairs(d_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance, lengthscale, d_K, nullptr, nullptr, W.ld,
                              d_status, d_recs, pairs_scratch_bytes(n_cand), st)) return rc;
    if (int rc = launch_posterior(d_K, n_cand, W.ld, n_eval, d_Linv, ldl, d_beta, mean_c, variance, acq_kind, d_current_max, xi, ucb_beta,
                                  h_mu ? d_mu : nullptr, h_sigma ? d_sigma : nullptr, h_acq ? d_acq : nullptr, st)) return rc;
    const size_t nb = sizeof(double) * (size_t)n_cand;
    if (h_mu) if (int rc = cuda_check(cudaMemcpyAsync(h_mu, d_mu, nb, cudaMemcpyDeviceToHost, st), "D2H mu")) return rc;
    if (h_sigma) if (int rc = cuda_check(cudaMemcpyAsync(h_sigma, d_sigma, nb, cudaMemcpyDeviceToHost, st), "D2H sigma")) return rc;
    if (h_acq) if (int rc = cuda_check(cudaMemcpyAsync(h_acq, d_acq, nb, cudaMemcpyDeviceToHost, st), "D2H acq")) return rc;
    return BOSOT_OK;
  }
  int dev = 0;
  if (int rc = cuda_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  std::lock_guard<std::mutex> lock(g_pipe_mutex);
  HostPipe* P = nullptr;
  if (int rc = host_pipe(dev, &P)) return rc;

  // fork: both copy streams start after everything already queued on the caller's stream
  if (int rc = cuda_check(cudaEventRecord(P->start, st), "cudaEventRecord")) return rc;
  if (int rc = cuda_check(cudaStreamWaitEvent(P->h2d, P->start, 0), "cudaStreamWaitEvent")) return rc;
  if (int rc = cuda_check(cudaStreamWaitEvent(P->d2h, P->start, 0), "cudaStreamWaitEvent")) return rc;
  if (int rc = cuda_check(cudaStreamWaitEvent(P->alt, P->start, 0), "cudaStreamWaitEvent")) return rc;
  // Chunks alternate between the caller's stream and a second compute stream: the persistent kernels of chunk k+1
  // take over the SMs that the tail of chunk k has already left (every chunk has its own slice of the work buffer).
  // (measured at 1M x 200: 4.39 ms on one compute stream, 4.27 ms on two)
  // Chunk sizes grow geometrically: a small first chunk (a twentieth of the pool) so that the kernels start early, then
  // every chunk three times the one before -- the copy engine delivers trees several times faster than the kernels
  // consume them, so the next, larger chunk is resident before the current one is done -- which keeps the number of
  // launches (and their tails) small.  Measured at 1M x 200: growth 2 / 3 / 4 = 4.22 / 4.19 / 4.26 ms (equal chunks of
  // 200k behind a quarter-size first chunk: 4.23 ms).
  constexpr int kGrow = 3;
  int64_t bounds[kMaxChunks + 1];
  int n_parts = 0;
  {
    int64_t len = ((n_cand / 20) + 63) & ~(int64_t)63, c = 0;
    while (c < n_cand && n_parts < kMaxChunks - 1) {
      bounds[n_parts++] = c;
      c += len;
      len = (len * kGrow + 63) & ~(int64_t)63;
    }
    if (c < n_cand) bounds[n_parts++] = c;  // whatever is left goes into the last chunk
    bounds[n_parts] = n_cand;
  }
  for (int k = 0; k < n_parts; ++k) {
    const int64_t c0 = bounds[k];
    const int64_t n = (bounds[k + 1] < n_cand ? bounds[k + 1] : n_cand) - c0;
    if (n <= 0) break;
    if (int rc = cuda_check(cudaMemcpyAsync(d_trees + c0 * kTreeBytes, h_trees + c0 * kTreeBytes, (size_t)n * kTreeBytes,
                                            cudaMemcpyHostToDevice, P->h2d), "H2D trees")) return rc;
    if (int rc = cuda_check(cudaEventRecord(P->copied[k], P->h2d), "cudaEventRecord")) return rc;
    cudaStream_t cs = (k & 1) ? P->alt : st;
    if (int rc = cuda_check(cudaStreamWaitEvent(cs, P->copied[k], 0), "cudaStreamWaitEvent")) return rc;
    if (int rc = launch_pairs(d_trees + c0 * kTreeBytes, n, grammar, d_eval_blob, n_eval, weights, variance, lengthscale,
                              d_K + c0 * W.ld, nullptr, nullptr, W.ld, d_status + c0, d_recs + (size_t)c0 * kScanBytes + 16 * (size_t)k,
                              pairs_scratch_bytes(n), cs)) return rc;  // chunk k: its own scratch head, then its records
    if (int rc = launch_posterior(d_K + c0 * W.ld, n, W.ld, n_eval, d_Linv, ldl, d_beta, mean_c, variance, acq_kind, d_current_max,
                                  xi, ucb_beta, h_mu ? d_mu + c0 : nullptr, h_sigma ? d_sigma + c0 : nullptr,
                                  h_acq ? d_acq + c0 : nullptr, cs)) return rc;
    if (int rc = cuda_check(cudaEventRecord(P->computed[k], cs), "cudaEventRecord")) return rc;
    if (int rc = cuda_check(cudaStreamWaitEvent(P->d2h, P->computed[k], 0), "cudaStreamWaitEvent")) return rc;
    const size_t nb = sizeof(double) * (size_t)n;
    if (h_mu) if (int rc = cuda_check(cudaMemcpyAsync(h_mu + c0, d_mu + c0, nb, cudaMemcpyDeviceToHost, P->d2h), "D2H mu")) return rc;
    if (h_sigma) if (int rc = cuda_check(cudaMemcpyAsync(h_sigma + c0, d_sigma + c0, nb, cudaMemcpyDeviceToHost, P->d2h), "D2H sigma")) return rc;
    if (h_acq) if (int rc = cuda_check(cudaMemcpyAsync(h_acq + c0, d_acq + c0, nb, cudaMemcpyDeviceToHost, P->d2h), "D2H acq")) return rc;
  }
  // join: the caller's stream completes only after the last D2H
  if (int rc = cuda_check(cudaEventRecord(P->done, P->d2h), "cudaEventRecord")) return rc;
  return cuda_check(cudaStreamWaitEvent(st, P->done, 0), "cudaStreamWaitEvent");
}
