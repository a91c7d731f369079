// Hellinger kernel-kernel: expected Hellinger distance between the GP priors of two
// kernel-grammar trees (reference bosot/kernels/kernel_kernel_hellinger.py:104-259).
//
// Two launches:
//   hellinger_gram_kernel<VP>  one CTA per tree: for every hyper-parameter sample the Gram of the composed
//                              kernel on the V virtual points (sample_over_hps, :192-212) and its log|det|;
//   hellinger_pair_kernel<VP>  one CTA per 32 x 16 tile of (row tree, column tree) pairs, looping over the
//                              samples: log|det| of K1 + K2 for every (pair, sample) -> mean distance -> K.
// The pair kernel is the hot one: S * n_a * n_b symmetric V x V determinants (V = 20: 1330 fused
// multiply-adds each), fp64-pipe bound.  A matrix is factorised (LDL^T, no pivoting: the matrices are
// Gram + jitter) by FOUR lanes that keep the lower triangle in registers, column j on lane j mod 4; the
// pivot column travels by warp shuffle, so no shared-memory traffic besides reading the two operands once.
// Operands are staged per sample by TMA bulk copies (double buffered) in exactly the order the lanes
// consume them: element (slot m, row i, lane r) = K[i][4m + r] at ((base_m + i - 4m) * 4 + r), records
// padded to a stride = 4 (mod 16) doubles so that the four groups of a half-warp hit distinct banks.
//
// logdet((K1 + K2) / 2 + jitter I) = logdet(K1' + K2') - V log 2 with K' = K + 2 jitter I: the store keeps
// K' (what the reference caches is K + jitter I; its log|det| is computed here from K' - jitter I).
#include "sot_common.cuh"
#include "launch.cuh"

namespace bosot {

constexpr double kHJitter = 1e-6;  // kernel_kernel_hellinger.py:126
constexpr int kHMaxParamsSmem = 160 * 1024;

template <int VP>
struct HRec {
  static constexpr int kSlots = VP / 4;
  static constexpr int kT = kSlots * VP - 2 * kSlots * (kSlots - 1);  // per-lane entries: sum_m (VP - 4m)
  static constexpr int kDoubles = 4 * kT;
  static constexpr int kStride = ((kDoubles + 11) / 16) * 16 + 4;     // smallest value = 4 (mod 16) >= kDoubles
  __host__ __device__ static constexpr int base(int m) { return m * VP - 2 * m * (m - 1); }
};

__host__ __device__ inline int hellinger_vp(int v) { return (v + 3) & ~3; }

// ---- LDL^T of one symmetric matrix spread over 4 lanes; returns the product of the pivots in two halves ----
// a[base(m) + i - 4m] = A[i][4m + r], i >= 4m.  Entries above the diagonal of a column and columns that are
// already eliminated receive harmless updates (uniform trip counts: no divergence inside the warp).
template <int VP>
__device__ __forceinline__ void ldl_pivot_products(double (&a)[HRec<VP>::kT], int r, int lane0, double& p_lo, double& p_hi) {
  using R = HRec<VP>;
  p_lo = 1.0;
  p_hi = 1.0;
#pragma unroll
  for (int k = 0; k < VP; ++k) {
    const int q = k >> 2, rk = k & 3;
    double c[VP];
#pragma unroll
    for (int i = k; i < VP; ++i) c[i] = __shfl_sync(0xffffffffu, a[R::base(q) + i - 4 * q], lane0 + rk);
    const double d = c[k];
    if (k < VP / 2) p_lo *= d; else p_hi *= d;
    if (k + 1 < VP) {
      const double inv = 1.0 / d;
#pragma unroll
      for (int m = q; m < R::kSlots; ++m) {
        // l_j = A[j][k] / d for this lane's column j = 4m + r (0 when the column is already eliminated)
        double cj = 0.0;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
          if (4 * m + rr > k && rr == r) cj = c[4 * m + rr];
        const double lj = cj * inv;
#pragma unroll
        for (int i = (4 * m > k + 1 ? 4 * m : k + 1); i < VP; ++i)
          a[R::base(m) + i - 4 * m] = fma(-c[i], lj, a[R::base(m) + i - 4 * m]);
      }
    }
  }
}

// =====================================================================================================
// Gram store
// =====================================================================================================
struct GramParams {
  const uint8_t* trees;
  int64_t n_trees;
  bosot_hellinger_space sp;
  const double* X;
  const double* quant;
  double* G;
  double* logdet;
  int32_t* status;
  int n_warps;
};

__device__ __forceinline__ double hellinger_leaf(const bosot_hellinger_space& sp, uint32_t sym, const double* th, const double* xi,
                                                 const double* xj) {
  const int kind = sp.sym_kind[sym];
  const int dim = sp.sym_dim[sym];
  const int lo = dim < 0 ? 0 : dim;
  const int nd = dim < 0 ? sp.input_dimension : 1;
  switch (kind) {
    case BOSOT_HK_RBF: {  // th = 1/lengthscales[nd], variance
      double r2 = 0.0;
      for (int a = 0; a < nd; ++a) { const double t = (xi[lo + a] - xj[lo + a]) * th[a]; r2 = fma(t, t, r2); }
      return th[nd] * exp(-0.5 * r2);
    }
    case BOSOT_HK_LINEAR: {  // th = offset, variance
      double s = 0.0;
      for (int a = 0; a < nd; ++a) s = fma(xi[lo + a] * th[1], xj[lo + a], s);
      return s + th[0];
    }
    case BOSOT_HK_PERIODIC: {  // th = 1/period, 1/lengthscales[nd], variance
      double r2 = 0.0;
      for (int a = 0; a < nd; ++a) {
        const double t = sin(3.14159265358979323846 * (xi[lo + a] - xj[lo + a]) * th[0]) * th[1 + a];
        r2 = fma(t, t, r2);
      }
      return th[1 + nd] * exp(-0.5 * r2);
    }
    default: {  // RQ: th = alpha, 1/lengthscales[nd], variance
      double r2 = 0.0;
      for (int a = 0; a < nd; ++a) { const double t = (xi[lo + a] - xj[lo + a]) * th[1 + a]; r2 = fma(t, t, r2); }
      return th[1 + nd] * pow(1.0 + 0.5 * r2 / th[0], -th[0]);
    }
  }
}

template <int VP>
__global__ void __launch_bounds__(256) hellinger_gram_kernel(GramParams p) {
  using R = HRec<VP>;
  extern __shared__ __align__(16) uint8_t smem[];
  const int V = p.sp.n_virtual, d = p.sp.input_dimension, S = p.sp.n_samples, J = p.sp.n_sobol_dims;
  double* sX = reinterpret_cast<double*>(smem);                                     // [V][d]
  double* sTheta = sX + ((V * d + 1) & ~1);                                         // [n_warps][8][J]
  uint16_t* sOff = reinterpret_cast<uint16_t*>(sTheta + (size_t)p.n_warps * 8 * J); // [64] first parameter of the leaf at a position
  uint8_t* sTok = reinterpret_cast<uint8_t*>(sOff + 64);                            // [64]
  uint8_t* sPlane = sTok + 64;                                                      // [J] bit 7 = store the reciprocal
  uint8_t* sPosM = sPlane + ((J + 15) & ~15);                                       // [kT] slot of a per-lane entry
  uint8_t* sPosI = sPosM + R::kT;                                                   // [kT] row of a per-lane entry
  __shared__ int s_nparams, s_defect;

  const int64_t tree = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint8_t* rec = p.trees + tree * kTreeBytes;
  const int n = rec[0];
  if (tid < 64) sTok[tid] = tid < 63 ? rec[1 + tid] : 0xFF;
  for (int i = tid; i < V * d; i += blockDim.x) sX[i] = p.X[i];
  for (int t = tid; t < R::kT; t += blockDim.x) {
    int m = 0;
    while (m + 1 < R::kSlots && R::base(m + 1) <= t) ++m;
    sPosM[t] = (uint8_t)m;
    sPosI[t] = (uint8_t)(4 * m + t - R::base(m));
  }
  __syncthreads();
  if (tid == 0) {  // structure check + parameter numbering, leaves left to right
    int defect = 0, cursor = 0;
    if (n < 1 || n > kMaxNodes || !(n & 1)) defect = BOSOT_TREE_BAD_LENGTH;
    int pending = 1;
    for (int q = 0; q < n && !defect; ++q) {
      const uint32_t t = sTok[q];
      if (pending <= 0) { defect = BOSOT_TREE_BAD_SHAPE; break; }
      if (t == BOSOT_OP_ADD || t == BOSOT_OP_MUL) { ++pending; continue; }
      if (t >= (uint32_t)p.sp.n_symbols) { defect = BOSOT_TREE_BAD_TOKEN; break; }
      --pending;
      const int kind = p.sp.sym_kind[t], nd = p.sp.sym_dim[t] < 0 ? d : 1;
      const int count = (kind == BOSOT_HK_RBF ? 1 : 2) + (kind == BOSOT_HK_LINEAR ? 0 : nd);
      if (cursor + count > J) { defect = BOSOT_HELLINGER_TOO_MANY_PARAMS; break; }
      sOff[q] = (uint16_t)cursor;
      int slot = 0;
      if (kind != BOSOT_HK_RBF) {  // offset | period | alpha
        sPlane[cursor++] = p.sp.prior_of[kind][slot++] | (kind == BOSOT_HK_PERIODIC ? 0x80 : 0);
      }
      if (kind != BOSOT_HK_LINEAR) {
        for (int a = 0; a < nd; ++a) sPlane[cursor++] = p.sp.prior_of[kind][slot] | 0x80;
        ++slot;
      }
      sPlane[cursor++] = p.sp.prior_of[kind][slot];  // variance
    }
    if (!defect && pending != 0) defect = BOSOT_TREE_BAD_SHAPE;
    s_defect = defect;
    s_nparams = cursor;
    p.status[tree] = defect;
  }
  __syncthreads();
  const int defect = s_defect, n_par = s_nparams;
  const size_t rec_stride = R::kStride;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  if (defect) {
    for (int s = warp; s < S; s += p.n_warps) {
      double* out = p.G + ((size_t)s * p.n_trees + tree) * rec_stride;
      for (int t = lane; t < R::kStride; t += 32) out[t] = qnan;
      if (lane == 0) p.logdet[(size_t)s * p.n_trees + tree] = qnan;
    }
    return;
  }
  double* theta = sTheta + (size_t)warp * 8 * J;
  const int g = lane >> 2, r = lane & 3;
  for (int s0 = warp * 8; s0 < S; s0 += p.n_warps * 8) {
    __syncwarp();
    for (int idx = lane; idx < 8 * n_par; idx += 32) {
      const int gg = idx / n_par, j = idx - gg * n_par;
      const int s = min(s0 + gg, S - 1);
      const uint32_t pl = sPlane[j];
      const double v = p.quant[((size_t)(pl & 0x7f) * S + s) * J + j];
      theta[gg * J + j] = (pl & 0x80) ? 1.0 / v : v;
    }
    __syncwarp();
    // the 8 records of this warp: entry (slot m, row i, lane rr) = K[i][4m + rr] (+ 2 jitter on the diagonal)
    for (int idx = lane; idx < 8 * R::kDoubles; idx += 32) {
      const int gg = idx / R::kDoubles, pos = idx - gg * R::kDoubles;
      const int s = s0 + gg;
      if (s >= S) break;
      const int t = pos >> 2, rr = pos & 3;
      const int i = sPosI[t], j = 4 * sPosM[t] + rr;
      double val = 0.0;
      if (i >= j) {
        if (i >= V) {
          val = i == j ? 1.0 : 0.0;  // padding up to a multiple of 4: identity block
        } else {
          double st[kMaxLeaves];
          int sp = 0;
          const double* th = theta + gg * J;
          for (int q = n - 1; q >= 0; --q) {  // reversed pre-order = postfix with swapped operands
            const uint32_t tk = sTok[q];
            if (tk >= 0x80) {
              const double lhs = st[sp - 1], rhs = st[sp - 2];
              --sp;
              st[sp - 1] = tk == BOSOT_OP_ADD ? lhs + rhs : lhs * rhs;
            } else {
              st[sp++] = hellinger_leaf(p.sp, tk, th + sOff[q], sX + i * d, sX + j * d);
            }
          }
          val = st[0] + (i == j ? 2.0 * kHJitter : 0.0);
        }
      }
      p.G[((size_t)s * p.n_trees + tree) * rec_stride + pos] = val;
    }
    __syncwarp();
    // log|det(K + jitter I)|: group g of this warp factorises sample s0 + g
    {
      const int s = min(s0 + g, S - 1);
      const double* in = p.G + ((size_t)s * p.n_trees + tree) * rec_stride;
      double a[R::kT];
#pragma unroll
      for (int t = 0; t < R::kT; ++t) a[t] = in[4 * t + r];
#pragma unroll
      for (int m = 0; m < R::kSlots; ++m)
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
          if (rr == r && 4 * m + rr < V) a[R::base(m) + rr] -= kHJitter;
      double p_lo, p_hi;
      ldl_pivot_products<VP>(a, r, lane & ~3, p_lo, p_hi);
      if (r == 0 && s0 + g < S) p.logdet[(size_t)s * p.n_trees + tree] = log(fabs(p_lo)) + log(fabs(p_hi));
    }
  }
  for (int s = warp; s < S; s += p.n_warps) {  // tail of every record: finite padding for the bulk copies
    double* out = p.G + ((size_t)s * p.n_trees + tree) * rec_stride;
    for (int t = R::kDoubles + lane; t < R::kStride; t += 32) out[t] = 0.0;
  }
}

// =====================================================================================================
// pairs
// =====================================================================================================
constexpr int kPairThreads = 256;
constexpr int kPairTA = 32, kPairTB = 16;

struct PairParams {
  const double* Ga;
  const double* lda;
  int64_t n_a;
  const double* Gb;
  const double* ldb;
  int64_t n_b;
  int S;
  double variance, ls2;
  double* K;
  double* D;
  int64_t ld;
  unsigned long long* info;
  int tiles_b;
};

__device__ __forceinline__ uint32_t hl_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hl_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hl_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void hl_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hl_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hl_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(hl_smem_u32(dst)), "l"(src), "r"(bytes), "r"(hl_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void hl_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(hl_smem_u32(bar)), "r"(parity) : "memory");
}

template <int VP>
__global__ void __launch_bounds__(kPairThreads, 1) hellinger_pair_kernel(PairParams p) {
  using R = HRec<VP>;
  constexpr int kStageDoubles = (kPairTA + kPairTB) * R::kStride;
  extern __shared__ __align__(16) uint8_t smem[];
  double* stage0 = reinterpret_cast<double*>(smem);
  double* sAcc = stage0 + 2 * kStageDoubles;                         // [512]
  int* sCnt = reinterpret_cast<int*>(sAcc + kPairTA * kPairTB);     // [512]
  uint64_t* full = reinterpret_cast<uint64_t*>(sCnt + kPairTA * kPairTB);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, r = lane & 3;
  const int64_t a0 = (int64_t)(blockIdx.x / p.tiles_b) * kPairTA;
  const int64_t b0 = (int64_t)(blockIdx.x % p.tiles_b) * kPairTB;
  const int na = (int)min((int64_t)kPairTA, p.n_a - a0), nb = (int)min((int64_t)kPairTB, p.n_b - b0);
  const uint32_t bytes_a = (uint32_t)(na * R::kStride * sizeof(double)), bytes_b = (uint32_t)(nb * R::kStride * sizeof(double));

  if (tid == 0) {
    hl_mbar_init(&full[0], 1);
    hl_mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < kPairTA * kPairTB; i += kPairThreads) { sAcc[i] = 0.0; sCnt[i] = 0; }
  __syncthreads();
  auto issue = [&](int s) {
    double* dst = stage0 + (s & 1) * kStageDoubles;
    hl_mbar_expect_tx(&full[s & 1], bytes_a + bytes_b);
    hl_bulk_g2s(dst, p.Ga + ((size_t)s * p.n_a + a0) * R::kStride, bytes_a, &full[s & 1]);
    hl_bulk_g2s(dst + kPairTA * R::kStride, p.Gb + ((size_t)s * p.n_b + b0) * R::kStride, bytes_b, &full[s & 1]);
  };
  if (tid == 0) issue(0);

  for (int s = 0; s < p.S; ++s) {
    if (tid == 0 && s + 1 < p.S) issue(s + 1);  // the other buffer was released by the barrier closing sample s - 1
    hl_mbar_wait(&full[s & 1], (s >> 1) & 1);
    const double* sA = stage0 + (s & 1) * kStageDoubles;
    const double* sB = sA + kPairTA * R::kStride;
#pragma unroll 1
    for (int u = 0; u < 8; ++u) {
      // half-warp = 4 groups: the same row record (broadcast) and 4 consecutive column records (stride = 4 mod 16 banks)
      const int bi = (g & 3) + 4 * (u & 3), ai = w + 8 * (g >> 2) + 16 * (u >> 2);
      const bool valid = ai < na && bi < nb;
      const double* Ra = sA + min(ai, na - 1) * R::kStride + r;
      const double* Rb = sB + min(bi, nb - 1) * R::kStride + r;
      double a[R::kT];
#pragma unroll
      for (int t = 0; t < R::kT; ++t) a[t] = Ra[4 * t] + Rb[4 * t];
      double p_lo, p_hi;
      ldl_pivot_products<VP>(a, r, lane & ~3, p_lo, p_hi);
      if (r == 0 && valid) {
        const double ld_avg = log(fabs(p_lo)) + log(fabs(p_hi)) - VP * 0.69314718055994530942;
        const double l1 = p.lda[(size_t)s * p.n_a + a0 + ai], l2 = p.ldb[(size_t)s * p.n_b + b0 + bi];
        const double hd = 1.0 - exp(0.25 * l1 + 0.25 * l2 - 0.5 * ld_avg);  // hellinger_distance, :185-190
        if (hd >= 0.0) {                                                     // false for NaN as well (:176-178)
          const int slot = ai * kPairTB + bi;
          sAcc[slot] += hd;
          sCnt[slot] += 1;
        }
      }
    }
    __syncthreads();
  }
  unsigned long long dropped = 0, empty = 0;
  for (int slot = tid; slot < kPairTA * kPairTB; slot += kPairThreads) {
    const int ai = slot / kPairTB, bi = slot - ai * kPairTB;
    if (ai < na && bi < nb) {
      const int cnt = sCnt[slot];
      const double dist = cnt ? sAcc[slot] * (1.0 / cnt) : __longlong_as_double(0x7ff8000000000000LL);
      dropped += (unsigned long long)(p.S - cnt);
      empty += cnt == 0;
      const size_t o = (size_t)(a0 + ai) * p.ld + b0 + bi;
      if (p.D) p.D[o] = dist;
      if (p.K) p.K[o] = p.variance * exp(-0.5 * (dist / p.ls2));  // K, :128-132
    }
  }
  if (dropped) atomicAdd(p.info, dropped);
  if (empty) atomicAdd(p.info + 1, empty);
}

template <int VP>
int launch_grams(GramParams& p, cudaStream_t st) {
  using R = HRec<VP>;
  const int J = p.sp.n_sobol_dims, V = p.sp.n_virtual, d = p.sp.input_dimension;
  int n_warps = (p.sp.n_samples + 7) / 8;
  if (n_warps > 8) n_warps = 8;  // 256 threads: the factorisation keeps 60 doubles per lane in registers
  const size_t per_warp = (size_t)8 * J * sizeof(double);
  while (n_warps > 1 && n_warps * per_warp > (size_t)kHMaxParamsSmem) --n_warps;
  if (per_warp > (size_t)kHMaxParamsSmem) return set_error(BOSOT_EINVAL, "bosot_hellinger_grams: n_sobol_dims too large for shared memory");
  p.n_warps = n_warps;
  const size_t smem = sizeof(double) * (((V * d + 1) & ~1) + (size_t)n_warps * 8 * J) + 64 * 2 + 64 + ((J + 15) & ~15) + 2 * R::kT + 16;
  if (int rc = cuda_check(cudaFuncSetAttribute(hellinger_gram_kernel<VP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                          "cudaFuncSetAttribute(hellinger_gram_kernel)")) return rc;
  hellinger_gram_kernel<VP><<<(unsigned)p.n_trees, n_warps * 32, smem, st>>>(p);
  return after_launch("hellinger_gram_kernel");
}

template <int VP>
int launch_pairs(PairParams& p, cudaStream_t st) {
  using R = HRec<VP>;
  const size_t smem = sizeof(double) * (2 * (size_t)(kPairTA + kPairTB) * R::kStride + kPairTA * kPairTB) + sizeof(int) * kPairTA * kPairTB + 16;
  if (int rc = cuda_check(cudaFuncSetAttribute(hellinger_pair_kernel<VP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(hellinger_pair_kernel)")) return rc;
  const int64_t tiles_a = (p.n_a + kPairTA - 1) / kPairTA;
  p.tiles_b = (int)((p.n_b + kPairTB - 1) / kPairTB);
  const int64_t tiles = tiles_a * p.tiles_b;
  if (tiles > 0x7fffffffLL) return set_error(BOSOT_EINVAL, "bosot_hellinger_pairs: too many tiles for one launch");
  hellinger_pair_kernel<VP><<<(unsigned)tiles, kPairThreads, smem, st>>>(p);
  return after_launch("hellinger_pair_kernel");
}

}  // namespace bosot

extern "C" int bosot_hellinger_record_doubles(int n_virtual) {
  using namespace bosot;
  switch (hellinger_vp(n_virtual)) {
    case 4: return HRec<4>::kStride;
    case 8: return HRec<8>::kStride;
    case 12: return HRec<12>::kStride;
    case 16: return HRec<16>::kStride;
    case 20: return HRec<20>::kStride;
    default: return set_error(BOSOT_EINVAL, "bosot_hellinger_record_doubles: n_virtual must be 1..20");
  }
}

extern "C" int bosot_hellinger_grams(const uint8_t* d_trees, int64_t n_trees, const bosot_hellinger_space* space,
                                     const double* d_virtual_X, const double* d_quantiles,
                                     double* d_G, double* d_logdet, int32_t* d_status, void* stream) {
  using namespace bosot;
  if (!d_trees || !space || !d_virtual_X || !d_quantiles || !d_G || !d_logdet || !d_status)
    return set_error(BOSOT_EINVAL, "bosot_hellinger_grams: null argument");
  if (n_trees < 0 || n_trees > 0x7fffffffLL) return set_error(BOSOT_EINVAL, "bosot_hellinger_grams: n_trees out of range");
  if (space->n_symbols < 1 || space->n_symbols > BOSOT_MAX_SYMBOLS || space->input_dimension < 1 ||
      space->input_dimension > BOSOT_HELLINGER_MAX_DIM || space->n_virtual < 1 || space->n_virtual > BOSOT_HELLINGER_MAX_VIRTUAL ||
      space->n_samples < 1 || space->n_sobol_dims < 1 || space->n_sobol_dims > 65535 || space->n_priors < 1 || space->n_priors > 127)
    return set_error(BOSOT_EINVAL, "bosot_hellinger_grams: space out of range");
  for (int s = 0; s < space->n_symbols; ++s) {
    if (space->sym_kind[s] > BOSOT_HK_RQ || space->sym_dim[s] >= space->input_dimension)
      return set_error(BOSOT_EINVAL, "bosot_hellinger_grams: bad symbol kind / dimension");
    for (int k = 0; k < 3; ++k)
      if (space->prior_of[space->sym_kind[s]][k] >= space->n_priors) return set_error(BOSOT_EINVAL, "bosot_hellinger_grams: prior plane out of range");
  }
  if (n_trees == 0) return BOSOT_OK;
  GramParams p;
  p.trees = d_trees; p.n_trees = n_trees; p.sp = *space; p.X = d_virtual_X; p.quant = d_quantiles;
  p.G = d_G; p.logdet = d_logdet; p.status = d_status; p.n_warps = 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (hellinger_vp(space->n_virtual)) {
    case 4: return launch_grams<4>(p, st);
    case 8: return launch_grams<8>(p, st);
    case 12: return launch_grams<12>(p, st);
    case 16: return launch_grams<16>(p, st);
    default: return launch_grams<20>(p, st);
  }
}

extern "C" int bosot_hellinger_pairs(const double* d_Ga, const double* d_logdet_a, int64_t n_a,
                                     const double* d_Gb, const double* d_logdet_b, int64_t n_b,
                                     int n_virtual, int n_samples, double variance, double lengthscale,
                                     double* d_K, double* d_D, int64_t ld, int64_t* d_info, void* stream) {
  using namespace bosot;
  if (!d_Ga || !d_logdet_a || !d_Gb || !d_logdet_b || !d_info || (!d_K && !d_D))
    return set_error(BOSOT_EINVAL, "bosot_hellinger_pairs: null argument");
  if (n_a < 0 || n_b < 0 || ld < n_b || n_samples < 1 || n_virtual < 1 || n_virtual > BOSOT_HELLINGER_MAX_VIRTUAL)
    return set_error(BOSOT_EINVAL, "bosot_hellinger_pairs: size out of range");
  if (!(lengthscale > 0.0)) return set_error(BOSOT_EINVAL, "bosot_hellinger_pairs: lengthscale must be positive");
  if ((reinterpret_cast<uintptr_t>(d_Ga) | reinterpret_cast<uintptr_t>(d_Gb)) & 15)
    return set_error(BOSOT_EINVAL, "bosot_hellinger_pairs: Gram stores must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = cuda_check(cudaMemsetAsync(d_info, 0, 2 * sizeof(int64_t), st), "cudaMemsetAsync(info)")) return rc;
  if (n_a == 0 || n_b == 0) return BOSOT_OK;
  PairParams p;
  p.Ga = d_Ga; p.lda = d_logdet_a; p.n_a = n_a; p.Gb = d_Gb; p.ldb = d_logdet_b; p.n_b = n_b; p.S = n_samples;
  p.variance = variance; p.ls2 = lengthscale * lengthscale; p.K = d_K; p.D = d_D; p.ld = ld;
  p.info = reinterpret_cast<unsigned long long*>(d_info); p.tiles_b = 0;
  switch (hellinger_vp(n_virtual)) {
    case 4: return launch_pairs<4>(p, st);
    case 8: return launch_pairs<8>(p, st);
    case 12: return launch_pairs<12>(p, st);
    case 16: return launch_pairs<16>(p, st);
    default: return launch_pairs<20>(p, st);
  }
}
