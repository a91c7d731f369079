// GP-over-kernels marginal likelihood and its analytic gradient in ONE launch.
//
// What the reference evaluates through GPflow/TF autodiff on every L-BFGS step of the hyper-parameter
// fit (bosot/models/gp_model.py:162-199 -> GPR.training_loss = -log_marginal_likelihood):
//     K    = var * exp(-D / ls^2) + noise * I,   D = sum_f w_f * Df[f]      (kernel_kernel_grammar_tree.py:110-113,159-161)
//     loss = 0.5 r^T K^-1 r + sum_i log L_ii + 0.5 E log(2 pi),  r = y - c,  K = L L^T
//     dloss/dtheta = 0.5 * sum_ij (Kinv_ij - alpha_i alpha_j) dK_ij/dtheta,  alpha = K^-1 r
// The three E x E per-feature distance matrices Df are constants of the fit (they come from the pair
// kernel once per model update), E is small (tens to ~150), so the whole evaluation is one CTA:
// Cholesky in shared memory fused with the inversion of the factor (one sweep of E steps, one barrier each,
// L^-1 written into the unused upper triangle), alpha from two triangular matrix-vector products,
// and Kinv_ij formed on the fly as row dot products while the five weighted sums are accumulated.
// Gradients are returned with respect to the CONSTRAINED quantities (w_f treated as independent, ls,
// var, noise, c); the host applies the chain rule of the sigmoid / softplus transforms.
#include "sot_common.cuh"
#include <atomic>
#include "launch.cuh"

namespace bosot {

// one CTA of TILE x TILE threads: 16 x 16 up to 48 evaluated kernels, 32 x 32 beyond (measured: at E = 26 / 80 / 160 the
// small tile takes 21 / 106 / 473 us, the large one 25 / 85 / 313 us)
constexpr int kFitMaxE = 160;

struct FitParams {
  const double* Df;  // [3][E][ld]
  int E;
  int64_t ld;
  const double* y;   // [E]
  double w[3], ls, var, noise, c;
  double* out;       // [0] loss, [1..3] dL/dw_f, [4] dL/dls, [5] dL/dvar, [6] dL/dnoise, [7] dL/dc, [8] info (0 ok, k+1 = minor k not PD)
};

template <int kFitThreads>
__device__ __forceinline__ double block_sum(double v, double* red) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < kFitThreads / 32; ++w) s += red[w];  // fixed order: reproducible
  return s;
}

template <int kFitTile>
__global__ void __launch_bounds__(kFitTile * kFitTile)
gp_nll_grad_kernel(FitParams p) {
  constexpr int kFitThreads = kFitTile * kFitTile;
  extern __shared__ __align__(16) uint8_t smem[];
  const int E = p.E, lda = E | 1;
  double* A = reinterpret_cast<double*>(smem);  // [E][lda]: lower = L, strict upper = (L^-1)^T
  double* r = A + (size_t)E * lda;              // residual, then forward-solve result
  double* alpha = r + E;
  double* dinv = alpha + E;                     // diagonal of L^-1
  double* red = dinv + E;                       // [kFitThreads / 32]
  __shared__ int s_info;
  const int tid = threadIdx.x;
  const double inv_ls2 = 1.0 / (p.ls * p.ls);
  const size_t plane = (size_t)E * p.ld;
  if (tid == 0) s_info = 0;

  // K (lower triangle incl. diagonal), Y = 0 in the strict upper triangle, and the residual
  const int tx = tid & (kFitTile - 1), ty = tid / kFitTile;  // thread tile of the two-dimensional sweeps
  for (int i = ty; i < E; i += kFitTile) {
    for (int j = tx; j < E; j += kFitTile) {
      double v = 0.0;
      if (j <= i) {
        const size_t o = (size_t)i * p.ld + j;
        const double D = p.w[0] * p.Df[o] + p.w[1] * p.Df[plane + o] + p.w[2] * p.Df[2 * plane + o];
        v = p.var * exp(-D * inv_ls2) + (i == j ? p.noise : 0.0);
      }
      A[i * lda + j] = v;
    }
  }
  for (int i = tid; i < E; i += kFitThreads) r[i] = p.y[i] - p.c;
  __syncthreads();

  // ONE sweep of E steps, one barrier each, factorises K and inverts the factor.  Step k reads the pivot
  // d_k = A[k][k] and the (unscaled) column k of the lower triangle -- neither is touched again -- and applies
  //   (a) the trailing Cholesky update        A[i][j] -= A[i][k] A[j][k] / d_k            k < j <= i
  //   (b) the forward substitution of L Y = I  Y[i][j] -= A[i][k] Y[k][j] / d_k           j <= k < i, Y[k][k] = 1
  // with Y[i][j] (i > j) kept transposed in the strict upper triangle, A[j][i].  In scaled terms L[i][k] =
  // A[i][k] / sqrt(d_k) and L^-1[i][j] = Y[i][j] / sqrt(d_i): no thread ever waits for one that takes a square root.
  for (int k = 0; k < E; ++k) {
    const double dk = A[k * lda + k];
    if (!(dk > 0.0)) {  // uniform: every thread reads the same value
      if (tid == 0) s_info = k + 1;
      break;
    }
    const double inv = 1.0 / dk;
    for (int i = k + 1 + ty; i < E; i += kFitTile) {
      const double lik = A[i * lda + k] * inv;
      for (int j = k + 1 + tx; j <= i; j += kFitTile) A[i * lda + j] -= lik * A[j * lda + k];
    }
    for (int j = ty; j <= k; j += kFitTile) {
      const double c = (j == k ? 1.0 : A[j * lda + k]) * inv;
      for (int i = k + 1 + tx; i < E; i += kFitTile) A[j * lda + i] -= A[i * lda + k] * c;
    }
    __syncthreads();
  }
  __syncthreads();
  if (s_info != 0) {  // not positive definite: report like a failed Cholesky, the caller retries
    if (tid == 0) {
      const double qnan = __longlong_as_double(0x7ff8000000000000LL);
      for (int q = 0; q < 8; ++q) p.out[q] = qnan;
      p.out[8] = (double)s_info;
    }
    return;
  }
  double logdet = 0.0;
  for (int k = tid; k < E; k += kFitThreads) {
    const double dk = A[k * lda + k];
    dinv[k] = 1.0 / sqrt(dk);  // diagonal of L^-1
    logdet += 0.5 * log(dk);   // log L[k][k]
  }
  logdet = block_sum<kFitThreads>(logdet, red);  // (two barriers: dinv is visible afterwards)
  // strict upper triangle: A[j][i] = L^-1[i][j] = Y[i][j] / L[i][i], i > j
  for (int j = ty; j < E; j += kFitTile)
    for (int i = j + 1 + tx; i < E; i += kFitTile) A[j * lda + i] *= dinv[i];
  __syncthreads();
  // alpha = K^-1 r = L^-T (L^-1 r): two triangular matrix-vector products, no sequential substitution
  for (int i = tid; i < E; i += kFitThreads) {
    double z = dinv[i] * r[i];
    for (int j = 0; j < i; ++j) z += A[j * lda + i] * r[j];
    alpha[i] = z;  // z = L^-1 r, kept in alpha's storage until the second product
  }
  __syncthreads();
  double quad = 0.0;
  for (int i = tid; i < E; i += kFitThreads) quad += alpha[i] * alpha[i];
  quad = block_sum<kFitThreads>(quad, red);
  for (int j = tid; j < E; j += kFitThreads) {
    double a = dinv[j] * alpha[j];
    for (int i = j + 1; i < E; ++i) a += A[j * lda + i] * alpha[i];
    r[j] = a;
  }
  __syncthreads();
  for (int j = tid; j < E; j += kFitThreads) alpha[j] = r[j];
  __syncthreads();

  // Kinv_ij = sum_{k >= i} U[i][k] U[j][k] (i >= j), U = (L^-1)^T upper triangular (diag in dinv); weighted sums
  double s_var = 0.0, s_f0 = 0.0, s_f1 = 0.0, s_f2 = 0.0, s_noise = 0.0;
  const int n_pairs = E * (E + 1) / 2;
  for (int idx = tid; idx < n_pairs; idx += kFitThreads) {
    // idx -> (i, j), j <= i, row-major over the lower triangle
    int i = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= idx) ++i;
    while (i * (i + 1) / 2 > idx) --i;
    const int j = idx - i * (i + 1) / 2;
    const double* ui = A + (size_t)i * lda;
    const double* uj = A + (size_t)j * lda;
    double kinv = dinv[i] * (i == j ? dinv[i] : uj[i]);
    for (int k = i + 1; k < E; ++k) kinv += ui[k] * uj[k];
    const double q = kinv - alpha[i] * alpha[j];
    const size_t o = (size_t)i * p.ld + j;
    const double d0 = p.Df[o], d1 = p.Df[plane + o], d2 = p.Df[2 * plane + o];
    const double G = exp(-(p.w[0] * d0 + p.w[1] * d1 + p.w[2] * d2) * inv_ls2);
    const double qg = (i == j ? 1.0 : 2.0) * q * G;  // symmetric: every off-diagonal pair counts twice
    s_var += qg;
    s_f0 += qg * d0;
    s_f1 += qg * d1;
    s_f2 += qg * d2;
    if (i == j) s_noise += q;
  }
  s_var = block_sum<kFitThreads>(s_var, red);
  s_f0 = block_sum<kFitThreads>(s_f0, red);
  s_f1 = block_sum<kFitThreads>(s_f1, red);
  s_f2 = block_sum<kFitThreads>(s_f2, red);
  s_noise = block_sum<kFitThreads>(s_noise, red);
  double s_alpha = 0.0;
  for (int i = tid; i < E; i += kFitThreads) s_alpha += alpha[i];
  s_alpha = block_sum<kFitThreads>(s_alpha, red);

  if (tid == 0) {
    p.out[0] = 0.5 * quad + logdet + 0.5 * E * 1.8378770664093454836;  // log(2 pi)
    const double cf = -0.5 * p.var * inv_ls2;                          // dK/dw_f = -var/ls^2 * G * Df[f]
    p.out[1] = cf * s_f0;
    p.out[2] = cf * s_f1;
    p.out[3] = cf * s_f2;
    p.out[4] = p.var / (p.ls * p.ls * p.ls) * (p.w[0] * s_f0 + p.w[1] * s_f1 + p.w[2] * s_f2);  // 0.5 * sum q * var G D * 2/ls^3
    p.out[5] = 0.5 * s_var;
    p.out[6] = 0.5 * s_noise;
    p.out[7] = -s_alpha;
    p.out[8] = 0.0;
  }
}

// ---- factorisation for the posterior: L^-1 (packed for the posterior kernel) and beta in ONE launch ------------
// What sot.gp_factorize otherwise takes from cuSOLVER / cuBLAS in six launches and a synchronisation (potrf, a
// triangular solve against the identity, the fragment-major packing, two triangular solves for beta): for evaluated
// sets that fit the shared-memory Cholesky of the fit kernel (E <= kFitMaxE) the same sweep -- factorisation fused
// with the inversion of the factor -- runs once per model update here, and the results leave the kernel in the
// layout the posterior kernel reads (bosot_pack_linv).  Larger sets keep the library route.
struct FactorParams {
  const double* K;   // [E][ld] Gram of the evaluated set (lower triangle read)
  int E;
  int64_t ld;
  double noise;
  const double* err; // [E] y - c
  double* LinvF;     // [ldl * ldl] fragment-major, zero padded
  int ldl;
  double* beta;      // [E] (K + noise I)^-1 err
  int32_t* info;     // 0, or k + 1 when the leading minor k is not positive definite (device or mapped host memory)
};

template <int kFitTile>
__global__ void __launch_bounds__(kFitTile * kFitTile)
gp_factor_kernel(FactorParams p) {
  constexpr int kFitThreads = kFitTile * kFitTile;
  extern __shared__ __align__(16) uint8_t smem[];
  const int E = p.E, lda = E | 1;
  double* A = reinterpret_cast<double*>(smem);  // [E][lda]: lower = L, strict upper = (L^-1)^T
  double* r = A + (size_t)E * lda;
  double* z = r + E;
  double* dinv = z + E;
  __shared__ int s_info;
  const int tid = threadIdx.x;
  if (tid == 0) s_info = 0;
  const int tx = tid & (kFitTile - 1), ty = tid / kFitTile;
  for (int i = ty; i < E; i += kFitTile)
    for (int j = tx; j < E; j += kFitTile)
      A[i * lda + j] = j <= i ? p.K[(size_t)i * p.ld + j] + (i == j ? p.noise : 0.0) : 0.0;
  for (int i = tid; i < E; i += kFitThreads) r[i] = p.err[i];
  __syncthreads();
  // the sweep of gp_nll_grad_kernel: trailing Cholesky update + forward substitution of L Y = I, one barrier per column
  for (int k = 0; k < E; ++k) {
    const double dk = A[k * lda + k];
    if (!(dk > 0.0)) {
      if (tid == 0) s_info = k + 1;
      break;
    }
    const double inv = 1.0 / dk;
    for (int i = k + 1 + ty; i < E; i += kFitTile) {
      const double lik = A[i * lda + k] * inv;
      for (int j = k + 1 + tx; j <= i; j += kFitTile) A[i * lda + j] -= lik * A[j * lda + k];
    }
    for (int j = ty; j <= k; j += kFitTile) {
      const double c = (j == k ? 1.0 : A[j * lda + k]) * inv;
      for (int i = k + 1 + tx; i < E; i += kFitTile) A[j * lda + i] -= A[i * lda + k] * c;
    }
    __syncthreads();
  }
  __syncthreads();
  if (s_info != 0) {
    if (tid == 0) *p.info = s_info;
    return;
  }
  for (int k = tid; k < E; k += kFitThreads) dinv[k] = 1.0 / sqrt(A[k * lda + k]);
  __syncthreads();
  for (int j = ty; j < E; j += kFitTile)
    for (int i = j + 1 + tx; i < E; i += kFitTile) A[j * lda + i] *= dinv[i];  // A[j][i] = L^-1[i][j], i > j
  __syncthreads();
  for (int i = tid; i < E; i += kFitThreads) {  // z = L^-1 err
    double v = dinv[i] * r[i];
    for (int j = 0; j < i; ++j) v += A[j * lda + i] * r[j];
    z[i] = v;
  }
  __syncthreads();
  for (int j = tid; j < E; j += kFitThreads) {  // beta = L^-T z
    double a = dinv[j] * z[j];
    for (int i = j + 1; i < E; ++i) a += A[j * lda + i] * z[i];
    p.beta[j] = a;
  }
  // fragment-major packing (bosot_pack_linv): LinvF[(rb * ldl/4 + ks) * 32 + 4 g + t] = Linv[8 rb + g][4 ks + t]
  const int KS = p.ldl >> 2;
  for (int x = tid; x < p.ldl * p.ldl; x += kFitThreads) {
    const int l = x & 31, f = x >> 5;
    const int rb = f / KS, ks = f - rb * KS;
    const int i = 8 * rb + (l >> 2), j = 4 * ks + (l & 3);
    double v = 0.0;
    if (i < E && j < E && j <= i) v = i == j ? dinv[i] : A[j * lda + i];
    p.LinvF[x] = v;
  }
  if (tid == 0) *p.info = 0;
}

}  // namespace bosot

extern "C" int bosot_gp_fit_max_eval(void) { return bosot::kFitMaxE; }

extern "C" int bosot_gp_factorize(const double* d_K, int n_eval, int64_t ld, double noise_variance, const double* d_err,
                                  double* d_LinvF, int ldl, double* d_beta, int32_t* info, void* stream) {
  using namespace bosot;
  if (!d_K || !d_err || !d_LinvF || !d_beta || !info) return set_error(BOSOT_EINVAL, "bosot_gp_factorize: null argument");
  if (n_eval < 1 || n_eval > kFitMaxE || ld < n_eval) return set_error(BOSOT_EINVAL, "bosot_gp_factorize: n_eval out of range (see bosot_gp_fit_max_eval)");
  if (ldl < n_eval || (ldl % 16) != 0) return set_error(BOSOT_EINVAL, "bosot_gp_factorize: ldl must be a multiple of 16 and >= n_eval");
  FactorParams p;
  p.K = d_K; p.E = n_eval; p.ld = ld; p.noise = noise_variance; p.err = d_err;
  p.LinvF = d_LinvF; p.ldl = ldl; p.beta = d_beta; p.info = info;
  const size_t smem = sizeof(double) * ((size_t)n_eval * (n_eval | 1) + 3 * (size_t)n_eval);
  static std::atomic<bool> attr_set[64];  // zero-initialised; the (idempotent) set-up may run twice under a race, never zero times
  int dev = 0;
  if (int rc = cuda_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (!(dev < 64 && attr_set[dev])) {
    if (int rc = cuda_check(cudaFuncSetAttribute(gp_factor_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024),
                            "cudaFuncSetAttribute(gp_factor_kernel<16>)")) return rc;
    if (int rc = cuda_check(cudaFuncSetAttribute(gp_factor_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024),
                            "cudaFuncSetAttribute(gp_factor_kernel<32>)")) return rc;
    if (dev < 64) attr_set[dev] = true;
  }
  if (n_eval <= 48) gp_factor_kernel<16><<<1, 256, smem, (cudaStream_t)stream>>>(p);
  else gp_factor_kernel<32><<<1, 1024, smem, (cudaStream_t)stream>>>(p);
  return after_launch("gp_factor_kernel");
}

extern "C" int bosot_gp_nll_grad(const double* d_Df, int n_eval, int64_t ld, const double* d_y, const double weights[3],
                                 double lengthscale, double variance, double noise_variance, double mean_c,
                                 double* d_out9, void* stream) {
  using namespace bosot;
  if (!d_Df || !d_y || !weights || !d_out9) return set_error(BOSOT_EINVAL, "bosot_gp_nll_grad: null argument");
  if (n_eval < 1 || n_eval > kFitMaxE || ld < n_eval) return set_error(BOSOT_EINVAL, "bosot_gp_nll_grad: n_eval out of range (see bosot_gp_fit_max_eval)");
  FitParams p;
  p.Df = d_Df; p.E = n_eval; p.ld = ld; p.y = d_y;
  p.w[0] = weights[0]; p.w[1] = weights[1]; p.w[2] = weights[2];
  p.ls = lengthscale; p.var = variance; p.noise = noise_variance; p.c = mean_c; p.out = d_out9;
  const size_t smem = sizeof(double) * ((size_t)n_eval * (n_eval | 1) + 3 * (size_t)n_eval + 32);
  static std::atomic<bool> attr_set[64];  // zero-initialised; the (idempotent) set-up may run twice under a race, never zero times
  int dev = 0;
  if (int rc = cuda_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (!(dev < 64 && attr_set[dev])) {
    if (int rc = cuda_check(cudaFuncSetAttribute(gp_nll_grad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024),
                            "cudaFuncSetAttribute(gp_nll_grad_kernel<16>)")) return rc;
    if (int rc = cuda_check(cudaFuncSetAttribute(gp_nll_grad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024),
                            "cudaFuncSetAttribute(gp_nll_grad_kernel<32>)")) return rc;
    if (dev < 64) attr_set[dev] = true;
  }
  if (n_eval <= 48) gp_nll_grad_kernel<16><<<1, 256, smem, (cudaStream_t)stream>>>(p);
  else gp_nll_grad_kernel<32><<<1, 1024, smem, (cudaStream_t)stream>>>(p);
  return after_launch("gp_nll_grad_kernel");
}
