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

bool pdl_enabled() {
  static const bool on = std::getenv("BOSOT_NO_PDL") == nullptr;
  return on;
}
// BOSOT_EXPERIMENT=<bit mask>: measurement switches for A/B runs (0 = product path).  Bits in use (csrc/sot_pairs.cu):
//   2  two-launch path (scan_kernel + pair kernel) also for pools up to the fused-scan limit
//   4  general exp_nonpos route even when the hyper-parameters would allow the table-driven exp
//   8  no block-vector cache of the DIM_WISE distances
// None of them changes a result beyond the last bits of exp (bit 4); bits 2 and 8 are bit-identical to the product path.
int experiment_flag() {
  static const int v = std::getenv("BOSOT_EXPERIMENT") ? std::atoi(std::getenv("BOSOT_EXPERIMENT")) : 0;
  return v;
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
                 double* d_Df, int64_t ld, int32_t* d_status, void* d_scratch, size_t scratch_bytes, cudaStream_t stream,
                 int32_t* d_bad, bool zero_bad, bool f32 = false);
size_t pairs_scratch_bytes(int64_t n_cand);
int launch_posterior(const double* d_Kstar, int64_t n_cand, int64_t ld, int n_eval, const double* d_Linv, int ldl,
                     const double* d_beta, double mean_c, double variance, int acq_kind, const double* d_current_max,
                     double xi, double ucb_beta, double* d_mu, double* d_sigma, double* d_acq, cudaStream_t stream,
                     double* const* d_peers = nullptr, int n_peers = 0, int64_t peer_off = 0, bool f32 = false);

// work buffer of the composed calls:  K*[n_cand][ld] | trees[n_cand][64] | mu | sigma | acq | status | bad count | scan records
struct WorkLayout {
  int64_t ld;
  size_t off_K, off_trees, off_mu, off_sigma, off_acq, off_status, off_bad, off_recs, total;
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
  W.off_bad = o;    o += 16;
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

extern "C" int bosot_posterior_acq_f32(const float* d_Kstar, int64_t n_cand, int64_t ld, int n_eval,
                                       const double* d_Linv, int ldl, const double* d_beta, double mean_c, double variance,
                                       int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                                       float* d_mu, float* d_sigma, float* d_acq, void* stream) {
  return launch_posterior(reinterpret_cast<const double*>(d_Kstar), n_cand, ld, n_eval, d_Linv, ldl, d_beta, mean_c, variance,
                          acq_kind, d_current_max, xi, ucb_beta, reinterpret_cast<double*>(d_mu), reinterpret_cast<double*>(d_sigma),
                          reinterpret_cast<double*>(d_acq), (cudaStream_t)stream, nullptr, 0, 0, true);
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

// Helper streams / events of the pipelined host entry points: the ONLY state this library keeps besides caller-owned
// buffers.  One set per device, created on the first large-pool bosot_acquire_host* call on that device, guarded by a
// per-device mutex (calls on one device are serialised for the short time they ENQUEUE work; they stay asynchronous),
// released by bosot_shutdown().
namespace bosot {
constexpr int kMaxChunks = 16;
static_assert(kMaxChunks == 16, "work_layout() reserves one 16-byte scratch head per chunk: 16 * 16 bytes");
constexpr int kMaxDevices = 64;
struct HostPipe {
  bool ready = false;
  cudaStream_t h2d = nullptr, d2h = nullptr, alt = nullptr;  // alt: second compute stream (odd chunks)
  cudaEvent_t start = nullptr, done = nullptr, alt_done = nullptr, h2d_done = nullptr, copied[kMaxChunks] = {}, computed[kMaxChunks] = {};
  std::mutex mutex;
};
static HostPipe g_pipe[kMaxDevices];
// tuning knobs of the host pipeline (bosot_host_pipeline_config): smallest pool that is chunked, smallest first chunk
static std::atomic<long long> g_pipe_min_cand{100000}, g_pipe_first_chunk{16384};

static void host_pipe_destroy(HostPipe& P) {
  if (P.h2d) cudaStreamDestroy(P.h2d);
  if (P.d2h) cudaStreamDestroy(P.d2h);
  if (P.alt) cudaStreamDestroy(P.alt);
  cudaEvent_t* single[] = {&P.start, &P.done, &P.alt_done, &P.h2d_done};
  for (cudaEvent_t* e : single) if (*e) { cudaEventDestroy(*e); *e = nullptr; }
  for (int k = 0; k < kMaxChunks; ++k) {
    if (P.copied[k]) { cudaEventDestroy(P.copied[k]); P.copied[k] = nullptr; }
    if (P.computed[k]) { cudaEventDestroy(P.computed[k]); P.computed[k] = nullptr; }
  }
  P.h2d = P.d2h = P.alt = nullptr;
  P.ready = false;
}

// caller holds P.mutex
static int host_pipe_init(HostPipe& P) {
  if (P.ready) return BOSOT_OK;
  int rc = BOSOT_OK;
  auto ev = [&](cudaEvent_t* e) { if (!rc) rc = cuda_check(cudaEventCreateWithFlags(e, cudaEventDisableTiming), "cudaEventCreate"); };
  auto st = [&](cudaStream_t* s, const char* what) { if (!rc) rc = cuda_check(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking), what); };
  st(&P.h2d, "cudaStreamCreate(h2d)"); st(&P.d2h, "cudaStreamCreate(d2h)"); st(&P.alt, "cudaStreamCreate(alt)");
  ev(&P.start); ev(&P.done); ev(&P.alt_done); ev(&P.h2d_done);
  for (int k = 0; k < kMaxChunks; ++k) { ev(&P.copied[k]); ev(&P.computed[k]); }
  if (rc) { host_pipe_destroy(P); return rc; }  // nothing half-built is kept
  P.ready = true;
  return BOSOT_OK;
}
}  // namespace bosot

extern "C" int bosot_enable_peer_access(int peer_device) {
  int dev = 0, can = 0;
  if (int rc = cuda_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (peer_device == dev) return BOSOT_OK;
  if (int rc = cuda_check(cudaDeviceCanAccessPeer(&can, dev, peer_device), "cudaDeviceCanAccessPeer")) return rc;
  if (!can) return set_error(BOSOT_EINVAL, "bosot_enable_peer_access: the current device cannot map the peer's memory");
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return BOSOT_OK; }
  return cuda_check(e, "cudaDeviceEnablePeerAccess");
}

extern "C" int bosot_host_pipeline_config(int64_t min_candidates, int64_t first_chunk) {
  if (min_candidates < 0 || first_chunk < 0) return set_error(BOSOT_EINVAL, "bosot_host_pipeline_config: negative argument");
  if (min_candidates > 0) g_pipe_min_cand.store(min_candidates, std::memory_order_relaxed);
  if (first_chunk > 0) g_pipe_first_chunk.store(first_chunk, std::memory_order_relaxed);
  return BOSOT_OK;
}

extern "C" int bosot_shutdown(void) {
  int keep = 0;
  const bool have_dev = cudaGetDevice(&keep) == cudaSuccess;
  for (int d = 0; d < kMaxDevices; ++d) {
    HostPipe& P = g_pipe[d];
    std::lock_guard<std::mutex> lock(P.mutex);
    if (!P.ready) continue;
    if (cudaSetDevice(d) == cudaSuccess) {
      cudaStreamSynchronize(P.h2d); cudaStreamSynchronize(P.d2h); cudaStreamSynchronize(P.alt);
    }
    host_pipe_destroy(P);
  }
  if (have_dev) cudaSetDevice(keep);
  cudaGetLastError();
  return BOSOT_OK;
}

namespace bosot {
// pairs -> posterior (-> peer scatter) on one stream for candidates resident on the device
static int acquire_on_device(const char* who, const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar,
                             const void* d_eval_blob, int n_eval, const double weights[3], double variance, double lengthscale,
                             const double* d_Linv, int ldl, const double* d_beta, double mean_c, int acq_kind,
                             const double* d_current_max, double xi, double ucb_beta, void* d_work, size_t work_bytes,
                             double* d_mu, double* d_sigma, double* d_acq, int32_t* d_n_bad,
                             double* const* d_peers, int n_peers, int64_t peer_off, cudaStream_t st, bool f32 = false) {
  char msg[128];
  if (!d_work) { std::snprintf(msg, sizeof(msg), "%s: null work buffer", who); return set_error(BOSOT_EINVAL, msg); }
  if (n_cand < 0 || n_eval < 1 || n_eval > BOSOT_MAX_EVAL) { std::snprintf(msg, sizeof(msg), "%s: bad sizes", who); return set_error(BOSOT_EINVAL, msg); }
  const WorkLayout W = work_layout(n_cand, n_eval);
  if (work_bytes < W.total) { std::snprintf(msg, sizeof(msg), "%s: work buffer too small", who); return set_error(BOSOT_EINVAL, msg); }
  uint8_t* w = static_cast<uint8_t*>(d_work);
  double* K = reinterpret_cast<double*>(w + W.off_K);
  int32_t* status = reinterpret_cast<int32_t*>(w + W.off_status);
  if (n_cand == 0 && d_n_bad) return cuda_check(cudaMemsetAsync(d_n_bad, 0, sizeof(int32_t), st), "cudaMemsetAsync");
  // one set of launches: cutting a device-resident pool into chunks on two streams was measured slower (4.07 vs 3.95 ms)
  if (int rc = launch_pairs(d_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance, lengthscale, K, nullptr,
                            nullptr, W.ld, status, w + W.off_recs, pairs_scratch_bytes(n_cand), st, d_n_bad, true, f32)) return rc;
  return launch_posterior(K, n_cand, W.ld, n_eval, d_Linv, ldl, d_beta, mean_c, variance, acq_kind, d_current_max, xi,
                          ucb_beta, d_mu, d_sigma, d_acq, st, d_peers, n_peers, peer_off, f32);
}
}  // namespace bosot

extern "C" int bosot_acquire_device(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar,
                                    const void* d_eval_blob, int n_eval, const double weights[3],
                                    double variance, double lengthscale,
                                    const double* d_Linv, int ldl, const double* d_beta, double mean_c,
                                    int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                                    void* d_work, size_t work_bytes,
                                    double* d_mu, double* d_sigma, double* d_acq, int32_t* d_n_bad, void* stream) {
  return acquire_on_device("bosot_acquire_device", d_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance, lengthscale,
                           d_Linv, ldl, d_beta, mean_c, acq_kind, d_current_max, xi, ucb_beta, d_work, work_bytes,
                           d_mu, d_sigma, d_acq, d_n_bad, nullptr, 0, 0, (cudaStream_t)stream);
}

// fp32 variant: the Gram panel in the work buffer and the three outputs are float arrays; the evaluated-set side
// (L^-1, beta, current_max) and the accumulation of the posterior stay fp64
extern "C" int bosot_acquire_device_f32(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar,
                                        const void* d_eval_blob, int n_eval, const double weights[3],
                                        double variance, double lengthscale,
                                        const double* d_Linv, int ldl, const double* d_beta, double mean_c,
                                        int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                                        void* d_work, size_t work_bytes,
                                        float* d_mu, float* d_sigma, float* d_acq, int32_t* d_n_bad, void* stream) {
  return acquire_on_device("bosot_acquire_device_f32", d_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance, lengthscale,
                           d_Linv, ldl, d_beta, mean_c, acq_kind, d_current_max, xi, ucb_beta, d_work, work_bytes,
                           reinterpret_cast<double*>(d_mu), reinterpret_cast<double*>(d_sigma), reinterpret_cast<double*>(d_acq),
                           d_n_bad, nullptr, 0, 0, (cudaStream_t)stream, true);
}

extern "C" int bosot_acquire_device_scatter(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar,
                                            const void* d_eval_blob, int n_eval, const double weights[3],
                                            double variance, double lengthscale,
                                            const double* d_Linv, int ldl, const double* d_beta, double mean_c,
                                            int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                                            void* d_work, size_t work_bytes,
                                            double* d_mu, double* d_sigma, double* d_acq, int32_t* d_n_bad,
                                            const void* d_peer_ptrs, int n_peers, int64_t peer_offset, void* stream) {
  if (!d_peer_ptrs) return set_error(BOSOT_EINVAL, "bosot_acquire_device_scatter: null peer pointer table");
  return acquire_on_device("bosot_acquire_device_scatter", d_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance,
                           lengthscale, d_Linv, ldl, d_beta, mean_c, acq_kind, d_current_max, xi, ucb_beta, d_work, work_bytes,
                           d_mu, d_sigma, d_acq, d_n_bad, static_cast<double* const*>(d_peer_ptrs), n_peers, peer_offset,
                           (cudaStream_t)stream);
}

// Host buffers in, host buffers out.  Large pools are cut into chunks that flow through a three-stage
// pipeline on three streams -- H2D of chunk k+1 (copy engine) | kernels of chunk k (caller's stream) |
// D2H of chunk k-1 (second copy engine) -- joined back into the caller's stream with events, so the call
// keeps its contract: asynchronous, ordered on `stream`, the caller synchronises `stream`.
namespace bosot {
static int acquire_from_host(const char* who, const uint8_t* h_trees, int64_t n_cand, const bosot_grammar* grammar,
                             const void* d_eval_blob, int n_eval, const double weights[3], double variance, double lengthscale,
                             const double* d_Linv, int ldl, const double* d_beta, double mean_c, int acq_kind,
                             const double* d_current_max, double xi, double ucb_beta, void* d_work, size_t work_bytes,
                             double* h_mu, double* h_sigma, double* h_acq, int32_t* h_n_bad,
                             double* const* d_peers, int n_peers, int64_t peer_off, cudaStream_t st) {
  char msg[128];
  if (!h_trees || !d_work) { std::snprintf(msg, sizeof(msg), "%s: null argument", who); return set_error(BOSOT_EINVAL, msg); }
  if (n_cand < 0 || n_eval < 1 || n_eval > BOSOT_MAX_EVAL) { std::snprintf(msg, sizeof(msg), "%s: bad sizes", who); return set_error(BOSOT_EINVAL, msg); }
  const WorkLayout W = work_layout(n_cand, n_eval);
  if (work_bytes < W.total) { std::snprintf(msg, sizeof(msg), "%s: work buffer too small", who); return set_error(BOSOT_EINVAL, msg); }
  if (n_cand == 0) { if (h_n_bad) *h_n_bad = 0; return BOSOT_OK; }
  uint8_t* w = static_cast<uint8_t*>(d_work);
  uint8_t* d_trees = w + W.off_trees;
  double* d_K = reinterpret_cast<double*>(w + W.off_K);
  double* d_mu = reinterpret_cast<double*>(w + W.off_mu);
  double* d_sigma = reinterpret_cast<double*>(w + W.off_sigma);
  double* d_acq = reinterpret_cast<double*>(w + W.off_acq);
  int32_t* d_status = reinterpret_cast<int32_t*>(w + W.off_status);
  int32_t* d_bad = reinterpret_cast<int32_t*>(w + W.off_bad);
  uint8_t* d_recs = w + W.off_recs;
  const bool want_acq = h_acq != nullptr || d_peers != nullptr;

  if (n_cand < g_pipe_min_cand.load(std::memory_order_relaxed)) {
    // small pool: nothing to overlap, keep everything on the caller's stream (no cross-stream event hops)
    if (int rc = cuda_check(cudaMemcpyAsync(d_trees, h_trees, (size_t)n_cand * kTreeBytes, cudaMemcpyHostToDevice, st), "H2D trees")) return rc;
    if (int rc = launch_pairs(d_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance, lengthscale, d_K, nullptr, nullptr, W.ld,
                              d_status, d_recs, pairs_scratch_bytes(n_cand), st, d_bad, true)) return rc;
    if (int rc = launch_posterior(d_K, n_cand, W.ld, n_eval, d_Linv, ldl, d_beta, mean_c, variance, acq_kind, d_current_max, xi, ucb_beta,
                                  h_mu ? d_mu : nullptr, h_sigma ? d_sigma : nullptr, want_acq ? d_acq : nullptr, st,
                                  d_peers, n_peers, peer_off)) return rc;
    const size_t nb = sizeof(double) * (size_t)n_cand;
    if (h_mu) if (int rc = cuda_check(cudaMemcpyAsync(h_mu, d_mu, nb, cudaMemcpyDeviceToHost, st), "D2H mu")) return rc;
    if (h_sigma) if (int rc = cuda_check(cudaMemcpyAsync(h_sigma, d_sigma, nb, cudaMemcpyDeviceToHost, st), "D2H sigma")) return rc;
    if (h_acq) if (int rc = cuda_check(cudaMemcpyAsync(h_acq, d_acq, nb, cudaMemcpyDeviceToHost, st), "D2H acq")) return rc;
    if (h_n_bad) if (int rc = cuda_check(cudaMemcpyAsync(h_n_bad, d_bad, sizeof(int32_t), cudaMemcpyDeviceToHost, st), "D2H bad count")) return rc;
    return BOSOT_OK;
  }
  int dev = 0;
  if (int rc = cuda_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (dev < 0 || dev >= kMaxDevices) return set_error(BOSOT_EINVAL, "bosot_acquire_host: device index out of range");
  HostPipe& P = g_pipe[dev];
  std::lock_guard<std::mutex> lock(P.mutex);
  if (int rc = host_pipe_init(P)) return rc;

  // the chunks share one count of malformed records: zeroed once, before the fork
  if (int rc = cuda_check(cudaMemsetAsync(d_bad, 0, sizeof(int32_t), st), "cudaMemsetAsync")) return rc;
  // fork: the helper streams start after everything already queued on the caller's stream
  if (int rc = cuda_check(cudaEventRecord(P.start, st), "cudaEventRecord")) return rc;
  // Chunks alternate between the caller's stream and a second compute stream: the persistent kernels of chunk k+1
  // take over the SMs that the tail of chunk k has already left (every chunk has its own slice of the work buffer).
  // (measured at 1M x 200: 4.39 ms on one compute stream, 4.27 ms on two)
  // Chunk sizes grow geometrically: a small first chunk (a twentieth of the pool) so that the kernels start early, then
  // every chunk three times the one before -- the copy engine delivers trees several times faster than the kernels
  // consume them, so the next, larger chunk is resident before the current one is done -- which keeps the number of
  // launches (and their tails) small.  Measured at 1M x 200: growth 2 / 3 / 4 = 4.22 / 4.19 / 4.26 ms (equal chunks of
  // 200k behind a quarter-size first chunk: 4.23 ms).  Round 2 (profiles/r2c_pipe.log): shards of 125k / 250k / 500k candidates
  // (the 8 / 4 / 2-GPU runs) un-chunked 1.33 / 1.45 / 2.73 ms, chunked from 100k up with a first chunk of >= 16384 candidates
  // 0.77 / 1.17 / 2.21 ms (first chunk 8192: 0.94 ms at 125k -- too many small launches).
  constexpr int kGrow = 3;
  int64_t bounds[kMaxChunks + 1];
  int n_parts = 0;
  {
    const int64_t first = g_pipe_first_chunk.load(std::memory_order_relaxed);
    int64_t len = n_cand / 20 > first ? n_cand / 20 : first, c = 0;
    len = (len + 63) & ~(int64_t)63;
    while (c < n_cand && n_parts < kMaxChunks - 1) {
      bounds[n_parts++] = c;
      c += len;
      len = (len * kGrow + 63) & ~(int64_t)63;
    }
    if (c < n_cand) bounds[n_parts++] = c;  // whatever is left goes into the last chunk
    bounds[n_parts] = n_cand;
  }
  // Everything between fork and join: a failure in the middle must still join the helper streams back into the
  // caller's stream, or later work on that stream could overtake chunks that are still using d_work.
  auto enqueue = [&]() -> int {
    if (int rc = cuda_check(cudaStreamWaitEvent(P.h2d, P.start, 0), "cudaStreamWaitEvent")) return rc;
    if (int rc = cuda_check(cudaStreamWaitEvent(P.d2h, P.start, 0), "cudaStreamWaitEvent")) return rc;
    if (int rc = cuda_check(cudaStreamWaitEvent(P.alt, P.start, 0), "cudaStreamWaitEvent")) return rc;
    for (int k = 0; k < n_parts; ++k) {
      const int64_t c0 = bounds[k];
      const int64_t n = (bounds[k + 1] < n_cand ? bounds[k + 1] : n_cand) - c0;
      if (n <= 0) break;
      if (int rc = cuda_check(cudaMemcpyAsync(d_trees + c0 * kTreeBytes, h_trees + c0 * kTreeBytes, (size_t)n * kTreeBytes,
                                              cudaMemcpyHostToDevice, P.h2d), "H2D trees")) return rc;
      if (int rc = cuda_check(cudaEventRecord(P.copied[k], P.h2d), "cudaEventRecord")) return rc;
      cudaStream_t cs = (k & 1) ? P.alt : st;
      if (int rc = cuda_check(cudaStreamWaitEvent(cs, P.copied[k], 0), "cudaStreamWaitEvent")) return rc;
      if (int rc = launch_pairs(d_trees + c0 * kTreeBytes, n, grammar, d_eval_blob, n_eval, weights, variance, lengthscale,
                                d_K + c0 * W.ld, nullptr, nullptr, W.ld, d_status + c0, d_recs + (size_t)c0 * kScanBytes + 16 * (size_t)k,
                                pairs_scratch_bytes(n), cs, d_bad, false)) return rc;  // chunk k: its own scratch head, then its records
      if (int rc = launch_posterior(d_K + c0 * W.ld, n, W.ld, n_eval, d_Linv, ldl, d_beta, mean_c, variance, acq_kind, d_current_max,
                                    xi, ucb_beta, h_mu ? d_mu + c0 : nullptr, h_sigma ? d_sigma + c0 : nullptr,
                                    want_acq ? d_acq + c0 : nullptr, cs, d_peers, n_peers, peer_off + c0)) return rc;
      if (int rc = cuda_check(cudaEventRecord(P.computed[k], cs), "cudaEventRecord")) return rc;
      if (int rc = cuda_check(cudaStreamWaitEvent(P.d2h, P.computed[k], 0), "cudaStreamWaitEvent")) return rc;
      const size_t nb = sizeof(double) * (size_t)n;
      if (h_mu) if (int rc = cuda_check(cudaMemcpyAsync(h_mu + c0, d_mu + c0, nb, cudaMemcpyDeviceToHost, P.d2h), "D2H mu")) return rc;
      if (h_sigma) if (int rc = cuda_check(cudaMemcpyAsync(h_sigma + c0, d_sigma + c0, nb, cudaMemcpyDeviceToHost, P.d2h), "D2H sigma")) return rc;
      if (h_acq) if (int rc = cuda_check(cudaMemcpyAsync(h_acq + c0, d_acq + c0, nb, cudaMemcpyDeviceToHost, P.d2h), "D2H acq")) return rc;
    }
    if (h_n_bad) if (int rc = cuda_check(cudaMemcpyAsync(h_n_bad, d_bad, sizeof(int32_t), cudaMemcpyDeviceToHost, P.d2h), "D2H bad count")) return rc;
    return BOSOT_OK;
  };
  const int rc_work = enqueue();
  char first_error[sizeof(t_error)];
  if (rc_work) std::snprintf(first_error, sizeof(first_error), "%s", t_error);
  // join: the caller's stream completes only after everything the helper streams were given (d2h waits on every
  // computed[k], so on success it alone would do; after a failure all three are joined)
  int rc_join = BOSOT_OK;
  auto join = [&](cudaEvent_t e, cudaStream_t from) {
    int rc = cuda_check(cudaEventRecord(e, from), "cudaEventRecord");
    if (!rc) rc = cuda_check(cudaStreamWaitEvent(st, e, 0), "cudaStreamWaitEvent");
    if (rc && !rc_join) rc_join = rc;
  };
  join(P.done, P.d2h);
  join(P.alt_done, P.alt);
  join(P.h2d_done, P.h2d);
  if (rc_work) { set_error(rc_work, first_error); return rc_work; }
  return rc_join;
}
}  // namespace bosot

extern "C" int bosot_acquire_host(const uint8_t* h_trees, int64_t n_cand, const bosot_grammar* grammar,
                                  const void* d_eval_blob, int n_eval, const double weights[3],
                                  double variance, double lengthscale,
                                  const double* d_Linv, int ldl, const double* d_beta, double mean_c,
                                  int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                                  void* d_work, size_t work_bytes,
                                  double* h_mu, double* h_sigma, double* h_acq, int32_t* h_n_bad, void* stream) {
  return acquire_from_host("bosot_acquire_host", h_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance, lengthscale,
                           d_Linv, ldl, d_beta, mean_c, acq_kind, d_current_max, xi, ucb_beta, d_work, work_bytes,
                           h_mu, h_sigma, h_acq, h_n_bad, nullptr, 0, 0, (cudaStream_t)stream);
}

extern "C" int bosot_acquire_host_scatter(const uint8_t* h_trees, int64_t n_cand, const bosot_grammar* grammar,
                                          const void* d_eval_blob, int n_eval, const double weights[3],
                                          double variance, double lengthscale,
                                          const double* d_Linv, int ldl, const double* d_beta, double mean_c,
                                          int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                                          void* d_work, size_t work_bytes,
                                          double* h_mu, double* h_sigma, double* h_acq, int32_t* h_n_bad,
                                          const void* d_peer_ptrs, int n_peers, int64_t peer_offset, void* stream) {
  if (!d_peer_ptrs) return set_error(BOSOT_EINVAL, "bosot_acquire_host_scatter: null peer pointer table");
  return acquire_from_host("bosot_acquire_host_scatter", h_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance,
                           lengthscale, d_Linv, ldl, d_beta, mean_c, acq_kind, d_current_max, xi, ucb_beta, d_work, work_bytes,
                           h_mu, h_sigma, h_acq, h_n_bad, static_cast<double* const*>(d_peer_ptrs), n_peers, peer_offset,
                           (cudaStream_t)stream);
}
