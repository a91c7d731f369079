// C-ABI glue: error state, launch bookkeeping, argument checks and the composed
// "acquire" entry points (pairs -> posterior -> acquisition) with host or device buffers.
#include <atomic>
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
                     double xi, double ucb_beta, double* d_mu, double* d_sigma, double* d_acq, cudaStream_t stream);

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
  W.off_recs = o;   o = align16(o + pairs_scratch_bytes(n_cand));
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

extern "C" size_t bosot_acquire_work_bytes(int64_t n_cand, int n_eval) {
  if (n_cand < 0 || n_eval < 1 || n_eval > BOSOT_MAX_EVAL) return 0;
  return work_layout(n_cand, n_eval).total;
}

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
  if (int rc = launch_pairs(d_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance, lengthscale, K, nullptr,
                            nullptr, W.ld, status, w + W.off_recs, pairs_scratch_bytes(n_cand), st)) return rc;
  return launch_posterior(K, n_cand, W.ld, n_eval, d_Linv, ldl, d_beta, mean_c, variance, acq_kind, d_current_max, xi,
                          ucb_beta, d_mu, d_sigma, d_acq, st);
}

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
  double* d_mu = reinterpret_cast<double*>(w + W.off_mu);
  double* d_sigma = reinterpret_cast<double*>(w + W.off_sigma);
  double* d_acq = reinterpret_cast<double*>(w + W.off_acq);
  if (int rc = cuda_check(cudaMemcpyAsync(d_trees, h_trees, (size_t)n_cand * kTreeBytes, cudaMemcpyHostToDevice, st), "H2D trees")) return rc;
  if (int rc = bosot_acquire_device(d_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance, lengthscale, d_Linv,
                                    ldl, d_beta, mean_c, acq_kind, d_current_max, xi, ucb_beta, d_work, work_bytes,
                                    h_mu ? d_mu : nullptr, h_sigma ? d_sigma : nullptr, h_acq ? d_acq : nullptr, stream)) return rc;
  const size_t nb = sizeof(double) * (size_t)n_cand;
  if (h_mu) if (int rc = cuda_check(cudaMemcpyAsync(h_mu, d_mu, nb, cudaMemcpyDeviceToHost, st), "D2H mu")) return rc;
  if (h_sigma) if (int rc = cuda_check(cudaMemcpyAsync(h_sigma, d_sigma, nb, cudaMemcpyDeviceToHost, st), "D2H sigma")) return rc;
  if (h_acq) if (int rc = cuda_check(cudaMemcpyAsync(h_acq, d_acq, nb, cudaMemcpyDeviceToHost, st), "D2H acq")) return rc;
  return BOSOT_OK;
}
