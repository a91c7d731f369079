/*
 * bosot_b200 -- C-ABI of the B200-native SOT kernel-kernel hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * Every entry point below replaces one step of the reference's Python hot path
 * (boschresearch/bosot; paths relative to bosot/ in the reference tree):
 *
 *   flat tree encoding ........ stands in for List[BaseKernelGrammarExpression]
 *                               (kernels/kernel_grammar/kernel_grammar.py:44-180,337-342)
 *   bosot_featurize ........... OptimalTransportKernelKernel.internal_transform_X
 *                               (kernels/kernel_kernel_grammar_tree.py:164-185) ->
 *                               get_subtree_dict / get_elementary_path_dict
 *                               (kernel_grammar.py:599-610,683-696) and
 *                               DimWiseWeightedDistanceExtractor (optimal_transport_mappings.py:88-108)
 *   bosot_eval_build .......... create_hash_array_mapping + feature_matrix for the
 *                               evaluated set (kernel_kernel_grammar_tree.py:203-230),
 *                               re-expressed as an inverted index (no global dictionary)
 *   bosot_sot_pairs ........... wasserstein_distance / K(X, X2) / get_distance_matrix
 *                               (kernel_kernel_grammar_tree.py:110-113,121-162) and
 *                               manhatten_distance (utils/utils.py:168-170)
 *   bosot_posterior_acq ....... GPR_with_posterior.predict_f as reached from
 *                               GPModel.predictive_dist (models/gp_model.py:290-306) and the
 *                               EI / GP-UCB formulas of BayesianOptimizerObjects.acquisition_function
 *                               (bayesian_optimization/bayesian_optimizer_objects.py:159-168)
 *   bosot_max_f64 ............. current_max = np.max(pred_mu_data) (same file, :165)
 *   bosot_acquire_host / _device  the whole acquisition_function call (same file, :159-168) on flat trees
 *   bosot_*_scatter ........... the same for a pool that is row-sharded over the GPUs of one node (no reference
 *                               counterpart: the reference is single-process)
 *   bosot_topk_f64, bosot_argmax_f64  EvolutionaryOptimizerObjects.select / np.argmax on the device
 *                               (bayesian_optimization/evolutionary_optimizer_objects.py:42-44,57-64)
 *   bosot_gp_nll_grad ......... GPR.training_loss + gradient for the ML-II fit (models/gp_model.py:162-199)
 *   bosot_neighbours, bosot_host_choices(_cb), bosot_host_recursive_dataset(_cb)
 *                               the candidate producers (kernels/kernel_grammar/kernel_grammar_search_spaces.py:97-151,
 *                               kernel_grammar_candidate_generator.py:85-111)
 *   bosot_hellinger_* ......... the alternative kernel-kernel (kernels/kernel_kernel_hellinger.py:104-259)
 *
 * All device pointers must belong to the current CUDA device of the calling thread.
 * Every call is asynchronous on the given stream (a cudaStream_t passed as void*; blocking and
 * non-blocking streams alike: nothing is uploaded behind the caller's back, all tables are static
 * data of the library) and returns 0 on success or a negative BOSOT_E* code; it never throws or
 * exits.  One call in flight per stream.
 * State: none besides caller-owned buffers, with ONE exception -- bosot_acquire_host /
 * bosot_acquire_host_scatter keep, per device, three helper streams and a few events for their
 * chunked copy/compute pipeline (created on the first call with >= 100000 candidates on that
 * device, guarded by a per-device mutex that is held only while work is ENQUEUED, released by
 * bosot_shutdown()).
 * fp64 throughout ("f64" path); the reference computes in float64 as well
 * (kernel_kernel_grammar_tree.py:35).  The *_f32 entry points are the optional reduced-precision variants.
 */
#ifndef BOSOT_B200_H
#define BOSOT_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BOSOT_ABI_VERSION 2

/* ---- flat tree encoding ("one tree = one 64-byte record") -------------------
 * byte 0      : n = number of nodes (odd, 1..63)   => at most 32 leaves
 * bytes 1..n  : tokens in PRE-ORDER (node, left subtree, right subtree) -- the order
 *               in which the reference enumerates subtrees (kernel_grammar.py:469)
 *               leaf  : symbol index 0..n_symbols-1  (< 0x80) = index of the base
 *                       kernel in search_space.base_kernel_config_list
 *                       (kernel_grammar_search_spaces.py:176-227)
 *               0x80  : ADD          0x81 : MULTIPLY   (KernelGrammarOperator, kernel_grammar.py:39-41)
 * bytes n+1.. : 0xFF padding
 */
#define BOSOT_TREE_BYTES 64
#define BOSOT_MAX_NODES 63
#define BOSOT_MAX_LEAVES 32
#define BOSOT_MAX_INTERNAL 31
#define BOSOT_OP_ADD 0x80
#define BOSOT_OP_MUL 0x81
#define BOSOT_PAD 0xFF

#define BOSOT_MAX_SYMBOLS 64
#define BOSOT_MAX_BLOCKS 16
#define BOSOT_MAX_DIM_COLS 80
#define BOSOT_PATH_CODES 64 /* reduced path code: 0 = empty, else 1 + 2*(len-1) + (nearest op == MULTIPLY) */

/* Search-space symbol table: what DimWiseWeightedDistanceExtractor.create_dim_class_dict
 * (optimal_transport_mappings.py:45-86) derives from generator_name/input_dimension. */
typedef struct bosot_grammar {
  int32_t n_symbols;                       /* B */
  int32_t n_blocks;                        /* differentiated dimensions */
  int32_t n_dim_cols;                      /* sum over blocks of (1 + kinds) */
  int32_t reserved;
  int8_t sym_block[BOSOT_MAX_SYMBOLS];     /* block of a symbol, -1 = not counted */
  int8_t sym_col[BOSOT_MAX_SYMBOLS];       /* dim-wise column of a symbol, -1 = none */
  int8_t block_null_col[BOSOT_MAX_BLOCKS]; /* column of NULL_i */
} bosot_grammar;

/* error codes */
#define BOSOT_OK 0
#define BOSOT_EINVAL (-1)  /* bad argument (null pointer, size out of range) */
#define BOSOT_ECUDA (-2)   /* CUDA runtime error, see bosot_last_error() */
#define BOSOT_ETREE (-3)   /* malformed tree record */
#define BOSOT_ERAW (-4)    /* bosot_host_*: the block of raw generator outputs is used up, call again with a longer one */

const char* bosot_last_error(void); /* thread-local message of the last failing call */
int bosot_abi_version(void);
/* Synchronises and destroys the helper streams / events of the host-buffer pipeline on every device (see "State"
 * above).  Optional; the library can be used again afterwards (the pipeline is rebuilt on demand). */
int bosot_shutdown(void);
/* Tuning knobs of that pipeline (process-wide; 0 keeps a value): pools of fewer than min_candidates candidates run
 * un-chunked on the caller's stream (default 100000); the first chunk holds max(n_cand / 20, first_chunk) candidates
 * (default 16384), every later chunk three times the one before. */
int bosot_host_pipeline_config(int64_t min_candidates, int64_t first_chunk);

/* ---- featurisation (ordered sparse histograms) ---------------------------------
 * Per tree i (all arrays row-major, caller-allocated on the device):
 *   n_nodes[i], n_leaves[i]
 *   sub_n[i], sub_id[i][63], sub_cnt[i][63]   SUBTREES classes in pre-order first-occurrence
 *                                              order; id = 64-bit canonical class id
 *                                              (leaf: symbol+1; internal: bit 63 set)
 *   path_n[i], path_id[i][32], path_cnt[i][32] REDUCED_ELEMENTARY_PATHS in the reference's
 *                                              insertion order; id = symbol*64 + path code
 *   sym_cnt[i][n_symbols]                      raw leaf counts per symbol
 *   dimvec[i][n_dim_cols]                      DIM_WISE feature vector (fp64)
 *   status[i]                                  0 ok, else a BOSOT_TREE_* defect code
 * Any output pointer may be NULL (skipped) except status.
 */
#define BOSOT_TREE_BAD_LENGTH 1
#define BOSOT_TREE_BAD_TOKEN 2
#define BOSOT_TREE_BAD_SHAPE 3
#define BOSOT_NEIGHBOUR_BAD_INDEX 4
#define BOSOT_NEIGHBOUR_TOO_BIG 5 /* the move would exceed 63 nodes */
#define BOSOT_HELLINGER_TOO_MANY_PARAMS 6 /* more hyper-parameters than columns of the quantile table */

int bosot_featurize(const uint8_t* d_trees, int64_t n_trees, const bosot_grammar* grammar,
                    uint8_t* d_n_nodes, uint8_t* d_n_leaves,
                    uint8_t* d_sub_n, uint64_t* d_sub_id, uint8_t* d_sub_cnt,
                    uint8_t* d_path_n, uint16_t* d_path_id, uint8_t* d_path_cnt,
                    uint8_t* d_sym_cnt, double* d_dimvec, int32_t* d_status, void* stream);

/* ---- evaluated-set state (inverted index over the E evaluated trees) --------------
 * bosot_eval_state_bytes: size of the opaque device blob for E trees.
 * bosot_eval_build: featurises the E trees and builds, inside the blob, a hash table
 * canonical-class-id -> postings (e, count, total), direct tables for leaf symbols and
 * reduced paths, the transposed dim-wise matrix and the per-block tables of distinct dim-wise
 * vectors ("states").  d_status: int32[E].
 */
#define BOSOT_MAX_EVAL 4095
size_t bosot_eval_state_bytes(int n_eval);
int bosot_eval_build(const uint8_t* d_trees, int n_eval, const bosot_grammar* grammar,
                     void* d_blob, size_t blob_bytes, int32_t* d_status, void* stream);
/* Diagnostics of a built blob (synchronises the stream): h_info[0] = 1 if the pair kernel evaluates DIM_WISE
 * through the block-state tables (every block of input dimensions takes <= 32 distinct vectors over the
 * evaluated trees), 0 = dense column loop; h_info[1] = dim-wise columns that belong to a block;
 * h_info[2] = largest number of distinct vectors in one block (33 = more than 32); h_info[3] = n_eval. */
int bosot_eval_info(const void* d_blob, int n_eval, int32_t h_info[4], void* stream);

/* ---- SOT pair kernel ------------------------------------------------------------
 * For every candidate tree c (row) and evaluated tree e (column):
 *   Df[f][c][e] = L1 distance of the f-th feature histograms, f = 0 DIM_WISE, 1 PATHS, 2 SUBTREES
 *   D[c][e]     = sum_f weights[f] * Df[f][c][e]        (weights = alpha_f / sum(alpha), 0 = feature off)
 *   K[c][e]     = variance * exp(-D / lengthscale^2)
 * Outputs are [n_cand][ld] fp64 with leading dimension ld >= n_eval (Df: 3 planes of
 * n_cand*ld); any of d_K / d_D / d_Df may be NULL.  d_status: int32[n_cand] tree defects (may be NULL).
 * d_scratch: device scratch of bosot_sot_pairs_scratch_bytes(n_cand) bytes, 16-byte aligned
 * (the work counter of the pair launch, then the compact per-tree records that the scan launch hands to it).
 */
size_t bosot_sot_pairs_scratch_bytes(int64_t n_cand);
int bosot_sot_pairs(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar,
                    const void* d_eval_blob, int n_eval, const double weights[3],
                    double variance, double lengthscale,
                    double* d_K, double* d_D, double* d_Df, int64_t ld,
                    int32_t* d_status, void* d_scratch, size_t scratch_bytes, void* stream);

/* ---- GP posterior + acquisition -----------------------------------------------------
 * Inputs: K*[n_cand][ld] (rows = candidates), LinvF = inverse of the lower Cholesky factor of
 * K_EE + noise*I in the zero-padded fragment-major layout written by bosot_pack_linv (order ldl >= n_eval,
 * a multiple of 16): LinvF[(rb * ldl/4 + ks) * 32 + 4 g + t] = Linv[8 rb + g][4 ks + t], i.e. one coalesced
 * 256-byte A fragment of the fp64 tensor-core tile per (row block, k-step);
 * beta[n_eval] = (K_EE + noise*I)^-1 (y - c).
 *   mu    = K* . beta + mean_c
 *   var   = variance - || L^-1 k* ||^2          (K_diag == variance: d(x,x) = 0 exactly)
 *   sigma = sqrt(var)                           (NaN if var < 0, as in the reference)
 *   acq   = EI:  d*Phi(d/sigma) + sigma*phi(d/sigma),  d = mu - *d_current_max - xi
 *           UCB: mu + sqrt(ucb_beta)*sigma
 * d_current_max is a device scalar (ignored for UCB / NONE).  Any of mu/sigma/acq may be NULL.
 * When ld > n_eval the padding columns [n_eval, ld) of K* must hold finite values (bosot_sot_pairs
 * writes zeros there).  Supported n_eval: up to about 800 (shared-memory tiles).
 */
#define BOSOT_ACQ_NONE 0
#define BOSOT_ACQ_EI 1
#define BOSOT_ACQ_UCB 2

int bosot_pack_linv(const double* d_Linv /* row-major [n][ld_src], lower triangular */, int n, int ld_src,
                    double* d_LinvF /* [ldl*ldl] */, int ldl, void* stream);
int bosot_posterior_acq(const double* d_Kstar, int64_t n_cand, int64_t ld, int n_eval,
                        const double* d_LinvF, int ldl, const double* d_beta, double mean_c, double variance,
                        int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                        double* d_mu, double* d_sigma, double* d_acq, void* stream);

/* What bosot_posterior_acq needs from the evaluated set, in ONE launch, for n_eval <= bosot_gp_fit_max_eval():
 * Cholesky of K_EE + noise_variance * I fused with the inversion of the factor, LinvF in the packed layout above
 * (order ldl, a multiple of 16, >= n_eval) and beta = (K_EE + noise_variance * I)^-1 err (err = y - mean_c).
 * d_K: [n_eval][ld] Gram of the evaluated set (lower triangle read).  *info (device memory, or host memory the device
 * can write: pinned) = 0, or k + 1 when the leading minor k is not positive definite -- the outputs are then undefined,
 * like a failed tf.linalg.cholesky in the reference (models/gp_model.py:183-186 retries).  Larger evaluated sets take
 * cuSOLVER / cuBLAS and bosot_pack_linv. */
int bosot_gp_factorize(const double* d_K, int n_eval, int64_t ld, double noise_variance, const double* d_err,
                       double* d_LinvF, int ldl, double* d_beta, int32_t* info, void* stream);

/* max of n doubles into a device scalar (NaNs ignored like fmax) */
int bosot_max_f64(const double* d_x, int64_t n, double* d_out, void* stream);

/* ---- host-buffer convenience (the end-to-end call) --------------------------------------
 * Candidates in (preferably pinned) host memory -> acquisition values in host memory.
 * Copies h_trees into d_work, runs pairs + posterior, copies acq (and mu/sigma when non-NULL) back, ordered
 * on `stream`; the caller synchronises the stream.  Pools of >= 100000 candidates flow through a chunked
 * H2D | kernels | D2H pipeline on helper streams that is joined back into `stream` before the call returns
 * (also when it returns an error); with pageable host memory the copies (and so the call) block.
 * d_work: device scratch of bosot_acquire_work_bytes(n_cand, n_eval) bytes.
 * h_n_bad / d_n_bad (int32, may be NULL): number of malformed candidate records of this call.  Their rows are
 * NaN in every output (K* row poisoned), all other rows are unaffected; a caller that cannot rule malformed
 * records out must check the count (0 = every value is valid) once the stream has been synchronised.
 */
size_t bosot_acquire_work_bytes(int64_t n_cand, int n_eval);
int bosot_acquire_host(const uint8_t* h_trees, int64_t n_cand, const bosot_grammar* grammar,
                       const void* d_eval_blob, int n_eval, const double weights[3],
                       double variance, double lengthscale,
                       const double* d_LinvF, int ldl, const double* d_beta, double mean_c,
                       int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                       void* d_work, size_t work_bytes,
                       double* h_mu, double* h_sigma, double* h_acq, int32_t* h_n_bad, void* stream);

/* same, candidates already on the device, outputs on the device */
int bosot_acquire_device(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar,
                         const void* d_eval_blob, int n_eval, const double weights[3],
                         double variance, double lengthscale,
                         const double* d_LinvF, int ldl, const double* d_beta, double mean_c,
                         int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                         void* d_work, size_t work_bytes,
                         double* d_mu, double* d_sigma, double* d_acq, int32_t* d_n_bad, void* stream);

/* ---- fp32 variants (SURVEY.md 8b "fp32 variants", 8d "fp32 mode: replace 8 by 4") -------------------------
 * Same contracts with float arrays for everything that scales with the number of candidates: the Gram / distance
 * outputs of the pair kernel (K, D, Df: [n_cand][ld] floats), K* read by the posterior, and mu / sigma / acq.
 * The integer minima and the per-feature distances are formed exactly as in the fp64 entry points and rounded once;
 * the Gram entry is variance * exp(x) with the hardware ex2 approximation (2 ulp of fp32).  Everything that belongs
 * to the evaluated set (LinvF, beta, current_max, the hyper-parameters) and the accumulation of the posterior stay
 * fp64: sigma^2 = variance - ||L^-1 k*||^2 cancels, and a condition number of 1e5 times fp32 rounding would not meet
 * any tolerance.  Stated bar (tests/test_gpu_f32.py): 1e-4 relative on D, K, mu, sigma and EI against the fp64 oracle
 * (sigma and EI: plus 1e-4 * sqrt(variance) absolute, the share of the fp32 rounding of K* that L^-1 amplifies).
 * The reference itself computes in float64 (kernels/kernel_kernel_grammar_tree.py:35); these entry points trade
 * accuracy for half the Gram bytes and are never the default of the Python layer. */
int bosot_sot_pairs_f32(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar,
                        const void* d_eval_blob, int n_eval, const double weights[3],
                        double variance, double lengthscale,
                        float* d_K, float* d_D, float* d_Df, int64_t ld,
                        int32_t* d_status, void* d_scratch, size_t scratch_bytes, void* stream);
int bosot_posterior_acq_f32(const float* d_Kstar, int64_t n_cand, int64_t ld, int n_eval,
                            const double* d_LinvF, int ldl, const double* d_beta, double mean_c, double variance,
                            int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                            float* d_mu, float* d_sigma, float* d_acq, void* stream);
/* d_work: bosot_acquire_work_bytes(n_cand, n_eval) bytes as for the fp64 call (the Gram panel uses half of its share) */
int bosot_acquire_device_f32(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar,
                             const void* d_eval_blob, int n_eval, const double weights[3],
                             double variance, double lengthscale,
                             const double* d_LinvF, int ldl, const double* d_beta, double mean_c,
                             int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                             void* d_work, size_t work_bytes,
                             float* d_mu, float* d_sigma, float* d_acq, int32_t* d_n_bad, void* stream);

/* ---- row-sharded pools: posterior + acquisition fused with the exchange of the result ----
 * The reference has no multi-process path; here the candidate pool is split into row blocks over the GPUs of one
 * node and every GPU needs the whole acquisition vector (the evolutionary optimiser selects on it,
 * bosot/bayesian_optimization/evolutionary_optimizer_objects.py:57-71).  Instead of an all-gather after the kernel,
 * the epilogue of the posterior kernel stores every acquisition value straight into the result vector of EVERY GPU:
 *     peer[r][peer_offset + c] = acq[c]     for r = 0 .. n_peers-1 (the own buffer included), c = 0 .. n_cand-1
 * d_peer_ptrs: device array of n_peers (<= 64) device pointers to fp64 buffers that are peer-mapped on this GPU
 * (NVLink / NVSwitch P2P, e.g. a torch symmetric-memory allocation); peer_offset = first global row of this shard.
 * d_acq (the local copy) is still written and must not be NULL.  The caller orders the buffers with a cross-GPU
 * barrier after the call (symmetric-memory barrier) before any rank reads its vector. */
/* Maps the memory of peer_device into the current device (cudaDeviceEnablePeerAccess; "already enabled" is not an
 * error), for callers that allocate the peer vectors themselves instead of through a symmetric-memory allocator. */
int bosot_enable_peer_access(int peer_device);
int bosot_posterior_acq_scatter(const double* d_Kstar, int64_t n_cand, int64_t ld, int n_eval,
                                const double* d_LinvF, int ldl, const double* d_beta, double mean_c, double variance,
                                int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                                double* d_mu, double* d_sigma, double* d_acq,
                                const void* d_peer_ptrs, int n_peers, int64_t peer_offset, void* stream);
int bosot_acquire_device_scatter(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar,
                                 const void* d_eval_blob, int n_eval, const double weights[3],
                                 double variance, double lengthscale,
                                 const double* d_LinvF, int ldl, const double* d_beta, double mean_c,
                                 int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                                 void* d_work, size_t work_bytes,
                                 double* d_mu, double* d_sigma, double* d_acq, int32_t* d_n_bad,
                                 const void* d_peer_ptrs, int n_peers, int64_t peer_offset, void* stream);
/* The same from HOST buffers: this rank's shard of the pool runs through the pipeline of bosot_acquire_host (H2D of the
 * shard | kernels | D2H of the shard's values into h_mu / h_sigma / h_acq, any of which may be NULL) while the posterior
 * epilogue of every chunk stores the acquisition values into all peers at peer_offset + row. */
int bosot_acquire_host_scatter(const uint8_t* h_trees, int64_t n_cand, const bosot_grammar* grammar,
                               const void* d_eval_blob, int n_eval, const double weights[3],
                               double variance, double lengthscale,
                               const double* d_LinvF, int ldl, const double* d_beta, double mean_c,
                               int acq_kind, const double* d_current_max, double xi, double ucb_beta,
                               void* d_work, size_t work_bytes,
                               double* h_mu, double* h_sigma, double* h_acq, int32_t* h_n_bad,
                               const void* d_peer_ptrs, int n_peers, int64_t peer_offset, void* stream);

/* ---- device-side selection for the evolutionary optimiser -------------------------
 * EvolutionaryOptimizerObjects.select keeps the last k = int(pop / (offspring + 1)) entries of np.argsort(values)
 * (bayesian_optimization/evolutionary_optimizer_objects.py:57-64) and maximize takes np.argmax (:42-44).
 * bosot_topk_f64: d_out_index[j] (int64, + index_offset) / d_out_value[j] (may be NULL), j = 0 .. k-1, are the last k
 * entries of the ascending order by (value, index) = np.argsort(values, kind="stable")[n-k:]; NaN sorts above +inf
 * like NumPy, -0.0 == +0.0.  1 <= k <= n < 2^31.  d_work: bosot_topk_work_bytes(n, k) bytes, 16-byte aligned
 * (not touched when n <= 4096: one CTA sorts in shared memory).  No host synchronisation.
 * bosot_argmax_f64: d_out2[0] = max value, d_out2[1] = its FIRST index (+ index_offset) as a double (np.argmax: a NaN
 * counts as the maximum).  d_work: bosot_argmax_work_bytes() bytes (used when n > 65536). */
size_t bosot_topk_work_bytes(int64_t n, int64_t k);
int bosot_topk_f64(const double* d_values, int64_t n, int64_t k, int64_t index_offset, int64_t* d_out_index,
                   double* d_out_value, void* d_work, size_t work_bytes, void* stream);
size_t bosot_argmax_work_bytes(void);
int bosot_argmax_f64(const double* d_values, int64_t n, int64_t index_offset, double* d_out2, void* d_work,
                     size_t work_bytes, void* stream);

/* ---- hyper-parameter fit of the GP over kernels ----------------------------------
 * One launch = one evaluation of GPR's training loss (negative log marginal likelihood; reference
 * bosot/models/gp_model.py:195-199) and its analytic gradient, for K = variance*exp(-D/ls^2) + noise*I,
 * D = sum_f weights[f]*Df[f], residual y - mean_c.  d_Df: [3][n_eval][ld] per-feature distance matrices of
 * the evaluated set (bosot_sot_pairs with d_Df), constants of the fit.  d_out9: [0] loss, [1..3] dloss/dw_f
 * (weights treated as independent), [4] dloss/dlengthscale, [5] dloss/dvariance, [6] dloss/dnoise,
 * [7] dloss/dmean_c, [8] info (0, or k+1 when the leading minor k is not positive definite: then [0..7] = NaN).
 * Single CTA, matrices in shared memory: n_eval <= bosot_gp_fit_max_eval(). */
int bosot_gp_fit_max_eval(void);
int bosot_gp_nll_grad(const double* d_Df, int n_eval, int64_t ld, const double* d_y, const double weights[3],
                      double lengthscale, double variance, double noise_variance, double mean_c,
                      double* d_out9, void* stream);

/* ---- candidate production on the flat encoding --------------------------------
 * bosot_neighbours: out[i] = neighbour number index[i] of trees[i] in the enumeration of
 * BaseKernelGrammarSearchSpace.get_neighbour_at_index (kernel_grammar_search_spaces.py:108-142):
 * a tree with L leaves and n nodes has L*B + 2*n*B neighbours (leaf swaps first, then extensions of
 * every sub-expression, root first, by (ADD | MULTIPLY, base kernel)).  d_status (may be NULL): 0 or a
 * BOSOT_TREE_* / BOSOT_NEIGHBOUR_* code; on error the input tree is copied through unchanged. */
int bosot_neighbours(const uint8_t* d_trees, int64_t n_trees, const int64_t* d_index, int n_symbols,
                     uint8_t* d_out, int32_t* d_status, void* stream);

/* Host side of the producers, for seeds that must propose the same kernels as the reference: the reference draws
 * every random index with np.random.choice(n), one Python call per draw (kernel_grammar_search_spaces.py:144-151,
 * kernel_grammar_candidate_generator.py:85-111).  These two functions replay exactly those draws from h_raw, a block
 * of raw 32-bit outputs of the same generator (np.random.randint(0, 2**32, size, dtype=uint32)):
 * RandomState.choice(n) is a masked rejection on 32-bit outputs (no output is used for n == 1).
 * *h_consumed = outputs used; the caller restores the generator state and advances it by that many outputs.
 * Return BOSOT_ERAW when h_raw is too short (nothing is to be kept from such a call).
 *   bosot_host_choices           out[i] = choice(n[i]), i = 0 .. count-1 in order
 *   bosot_host_recursive_dataset get_dataset_recursivly_generated(n_data, n_per_step) on flat records: the B roots,
 *                                then repeatedly a random neighbour of a randomly chosen earlier element;
 *                                h_pool: [max(n_data, n_symbols)][64] */
int bosot_host_choices(const uint32_t* h_raw, int64_t n_raw, const int64_t* h_n, int64_t count, int64_t* h_out,
                       int64_t* h_consumed);
int bosot_host_recursive_dataset(const uint32_t* h_raw, int64_t n_raw, int n_symbols, int64_t n_data, int n_per_step,
                                 uint8_t* h_pool, int64_t* h_consumed);
/* The same two producers pulling their raw outputs straight from the generator: next_u32(rng_state) must return the
 * generator's next raw 32-bit output (NumPy: g = numpy.random.get_bit_generator().ctypes; next_u32 = g.next_uint32,
 * rng_state = g.state).  The generator is left exactly where the reference's np.random.choice calls would have left
 * it -- no block of outputs, no state to restore; *h_consumed = outputs drawn.  A validation failure draws nothing. */
typedef uint32_t (*bosot_next_u32_fn)(void* rng_state);
int bosot_host_choices_cb(bosot_next_u32_fn next_u32, void* rng_state, const int64_t* h_n, int64_t count, int64_t* h_out,
                          int64_t* h_consumed);
int bosot_host_recursive_dataset_cb(bosot_next_u32_fn next_u32, void* rng_state, int n_symbols, int64_t n_data, int n_per_step,
                                    uint8_t* h_pool, int64_t* h_consumed);

/* ---- Hellinger kernel-kernel --------------------------------------------------
 * The alternative kernel-kernel behind the same BaseObjectKernel API (reference
 * kernels/kernel_kernel_hellinger.py:104-259): every tree is turned into S Gram matrices of its composed
 * kernel on V virtual points, one per hyper-parameter sample (sample_over_hps, :192-212); the distance of two
 * trees is the mean over samples of 1 - exp(logdet(K1)/4 + logdet(K2)/4 - logdet((K1+K2)/2 + jitter)/2)
 * (hellinger_distance / expected_hellinger_distance, :160-190), samples whose term is NaN or negative
 * are left out; K = variance * exp(-0.5 * d / lengthscale^2) (:128-132).  jitter = 1e-6 (:126).
 *
 * bosot_hellinger_space: what the reference reads from the leaf kernels.  Symbols are those of bosot_grammar.
 * Parameters of a tree are numbered leaf by leaf, left to right, inside a leaf in GPflow's
 * trainable_parameters order: RBF (lengthscales[nd], variance), LINEAR (offset, variance),
 * PERIODIC (period, lengthscales[nd], variance), RQ (alpha, lengthscales[nd], variance); nd = 1 for a kernel on
 * one dimension, input_dimension otherwise.  Parameter number j of sample i takes the value
 * d_quantiles[(prior * n_samples + i) * n_sobol_dims + j] = quantile of its prior at Sobol coordinate (i, j)
 * (tfd.Gamma(a, b).quantile(tf.math.sobol_sample(n_params, S)[i, j]), :199-207); the table is built by the host.
 */
#define BOSOT_HELLINGER_MAX_VIRTUAL 20
#define BOSOT_HELLINGER_MAX_DIM 16
#define BOSOT_HK_RBF 0
#define BOSOT_HK_LINEAR 1
#define BOSOT_HK_PERIODIC 2
#define BOSOT_HK_RQ 3
typedef struct bosot_hellinger_space {
  int32_t n_symbols;
  int32_t input_dimension;               /* <= BOSOT_HELLINGER_MAX_DIM */
  int32_t n_virtual;                     /* V <= BOSOT_HELLINGER_MAX_VIRTUAL */
  int32_t n_samples;                     /* S = num_param_samples */
  int32_t n_sobol_dims;                  /* columns of the quantile table; >= parameters of any tree passed */
  int32_t n_priors;                      /* planes of the quantile table */
  uint8_t sym_kind[BOSOT_MAX_SYMBOLS];   /* BOSOT_HK_* */
  int8_t sym_dim[BOSOT_MAX_SYMBOLS];     /* active dimension, -1 = all dimensions */
  uint8_t prior_of[4][3];                /* quantile plane of (kind, parameter slot in the order above) */
} bosot_hellinger_space;

/* Doubles per (tree, sample) record of the Gram store for V virtual points.  A record is opaque to callers:
 * K + 2 jitter I, lower triangle, in the column-cyclic 4-lane order the pair kernel factorises in (V rounded up
 * to a multiple of 4 with an identity block; entry (slot m, row i) of lane r = K[i][4m + r], two consecutive
 * entries of a lane adjacent), padded so that consecutive records fall into different shared-memory banks. */
int bosot_hellinger_record_doubles(int n_virtual);
/* Gram store of n_trees trees: d_G [S][n_trees][record] and d_logdet [S][n_trees] = log|det(K + jitter I)|.
 * d_virtual_X: [V][input_dimension] row-major.  d_status: int32[n_trees], 0 or a BOSOT_TREE_* code (then NaN Grams). */
int bosot_hellinger_grams(const uint8_t* d_trees, int64_t n_trees, const bosot_hellinger_space* space,
                          const double* d_virtual_X, const double* d_quantiles,
                          double* d_G, double* d_logdet, int32_t* d_status, void* stream);
/* K[a][b] (and / or D[a][b]), a < n_a rows, b < n_b columns, leading dimension ld >= n_b, from two Gram stores
 * (the same store twice gives the symmetric K(X, X)).  d_info: int64[2] (zeroed by the call): [0] (pair, sample)
 * terms left out as NaN / negative, [1] pairs without any valid sample (their D and K are NaN; the reference
 * divides by zero there). */
int bosot_hellinger_pairs(const double* d_Ga, const double* d_logdet_a, int64_t n_a,
                          const double* d_Gb, const double* d_logdet_b, int64_t n_b,
                          int n_virtual, int n_samples, double variance, double lengthscale,
                          double* d_K, double* d_D, int64_t ld, int64_t* d_info, void* stream);

/* ---- measurement helpers -----------------------------------------------------------------
 * bosot_fp64_fma_probe / bosot_fp64_mma_probe: run `iters` x 8 independent DFMA (resp. fp64
 * mma.sync m8n8k4) per thread (resp. warp) on a full grid so bench.py can measure this GPU's
 * fp64 ceilings (the secondary roofline of the posterior kernel).
 * bosot_kernel_launch_count: number of kernels this library has launched in this process. */
int bosot_fp64_fma_probe(double* d_out, int n_blocks, int n_threads, int iters, void* stream);
int bosot_fp64_mma_probe(double* d_out, int n_blocks, int n_threads, int iters, void* stream);
int bosot_fp64_mma_probe2(double* d_out, int n_blocks, int n_threads, int iters, void* stream); /* 2 A x 8 B fragments, 16 accumulator tiles */
int64_t bosot_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BOSOT_B200_H */
