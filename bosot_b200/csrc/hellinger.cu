// Hellinger kernel-kernel: expected Hellinger distance between the GP priors of two
// kernel-grammar trees (reference bosot/kernels/kernel_kernel_hellinger.py:104-259).
//
// Three kernels:
//   hellinger_gram_kernel<VP>    one CTA per tree: for every hyper-parameter sample the Gram of the composed
//                                kernel on the V virtual points (sample_over_hps, :192-212);
//   hellinger_logdet_kernel<VP>  log|det| of every stored Gram (get_slog_det, :214-218);
//   hellinger_pair_kernel<VP>    one CTA per 36 x 8 tile of (row tree, column tree) pairs, looping over the
//                                samples: det(K1 + K2) for every (pair, sample) -> mean distance -> K.
// The pair kernel is the hot one: S * n_a * n_b symmetric V x V determinants (V = 20: 1330 fused
// multiply-adds each).  A matrix is factorised (LDL^T, no pivoting: Gram + jitter) by FOUR lanes that keep
// the lower triangle in registers, column j on lane j mod 4; per step the pivot column goes through a small
// per-group shared-memory scratch (one 16-byte store per two rows by the owner, 16-byte broadcast loads by
// all four lanes: half the shared-memory wavefronts of 64-bit warp shuffles, which bounded the first version).
// Operands are staged per sample by TMA bulk copies (double buffered) in exactly the order the lanes consume
// them: per-lane entry t = (slot m, row i) of lane r = K[i][4m + r] at ((t / 2) * 4 + r) * 2 + t % 2, records
// padded to a stride = 8 (mod 16) doubles so that the groups of a quarter-warp hit distinct banks.  All eight
// groups of a warp work on the same row tree: its record is read as a broadcast.
//
// logdet((K1 + K2) / 2 + jitter I) = logdet(K1' + K2') - V log 2 with K' = K + 2 jitter I: the store keeps
// K' (what the reference caches is K + jitter I; its log|det| is computed from K' - jitter I).  The distance
// 1 - exp(ld1 / 4 + ld2 / 4 - ld_avg / 2) is evaluated as 1 - det1^(1/4) det2^(1/4) / sqrt(det_avg).
#include "sot_common.cuh"
#include "launch.cuh"
#include "fast_math.cuh"

namespace bosot {

constexpr double kHJitter = 1e-6;  // kernel_kernel_hellinger.py:126
constexpr int kHMaxParamsSmem = 160 * 1024;

template <int VP>
struct HRec {
  static constexpr int kSlots = VP / 4;
  static constexpr int kT = kSlots * VP - 2 * kSlots * (kSlots - 1);  // per-lane entries: sum_m (VP - 4m), even
  static constexpr int kDoubles = 4 * kT;
  static constexpr int kStride = ((kDoubles + 7) / 16) * 16 + 8;      // smallest value = 8 (mod 16) >= kDoubles
  __host__ __device__ static constexpr int base(int m) { return m * VP - 2 * m * (m - 1); }
  // per-lane entry t of lane r sits at pos(t, r): two consecutive entries of a lane form one 16-byte load
  __host__ __device__ static constexpr int pos(int t, int r) { return ((t >> 1) * 4 + r) * 2 + (t & 1); }
};

__host__ __device__ inline int hellinger_vp(int v) { return (v + 3) & ~3; }

__device__ __forceinline__ double2 lds128(const double* p) { return *reinterpret_cast<const double2*>(p); }

// ---- LDL^T of one symmetric matrix spread over 4 lanes; returns the product of the pivots in two halves ----
// a[base(m) + i - 4m] = A[i][4m + r], i >= 4m.  Step k: the lane owning column k hands it to the other three by
// warp shuffles, four rows at a time in ascending order (l_j of a slot is known when its rows arrive, so only
// four column entries are live at once), and every lane applies the rank-1 update to its own columns.  Entries
// above the diagonal of a column and columns that are already eliminated receive harmless updates (uniform
// trip counts: no divergence inside the warp).  Shared-memory bandwidth is the ceiling of this scheme: a
// 64-bit shuffle costs two wavefronts, exactly what a 16-byte broadcast load per two rows would cost, without
// the stores (measured: the shared-memory variant was 20 % slower).
template <int VP>
__device__ __forceinline__ void ldl_pivot_products(double (&a)[HRec<VP>::kT], int r, int lane0, double& p_lo, double& p_hi) {
  using R = HRec<VP>;
  p_lo = 1.0;
  p_hi = 1.0;
#pragma unroll
  for (int k = 0; k < VP; ++k) {
    const int q = k >> 2, rk = k & 3;
    const double d = __shfl_sync(0xffffffffu, a[R::base(q) + k - 4 * q], lane0 + rk);
    if (k < VP / 2) p_lo *= d; else p_hi *= d;
    if (k + 1 < VP) {
      const double inv = fast_rcp(d);
      double lj[R::kSlots];
#pragma unroll
      for (int mb = q; mb < R::kSlots; ++mb) {
        double c[4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
          if (4 * mb + rr > k) c[rr] = __shfl_sync(0xffffffffu, a[R::base(q) + 4 * mb + rr - 4 * q], lane0 + rk);
        double cj = 0.0;  // A[j][k] for this lane's column j = 4 mb + r (0 when that column is already eliminated)
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
          if (4 * mb + rr > k && rr == r) cj = c[rr];
        lj[mb] = cj * inv;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int i = 4 * mb + rr;
          if (i > k) {
#pragma unroll
            for (int m = (rk == 3 ? q + 1 : q); m <= mb; ++m)
              a[R::base(m) + i - 4 * m] = fma(-c[rr], lj[m], a[R::base(m) + i - 4 * m]);
          }
        }
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
      return th[nd] * exp_nonpos(-0.5 * r2);
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
      return th[1 + nd] * exp_nonpos(-0.5 * r2);
    }
    default: {  // RQ: th = alpha, 1/lengthscales[nd], variance
      double r2 = 0.0;
      for (int a = 0; a < nd; ++a) { const double t = (xi[lo + a] - xj[lo + a]) * th[1 + a]; r2 = fma(t, t, r2); }
      return th[1 + nd] * exp_nonpos(-th[0] * log_ge1(1.0 + 0.5 * r2 * fast_rcp(th[0])));  // (1 + r2 / (2 alpha))^-alpha
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
  if (defect) {  // NaN records: every distance involving this tree drops all its samples
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int s = warp; s < S; s += p.n_warps) {
      double* out = p.G + ((size_t)s * p.n_trees + tree) * rec_stride;
      for (int t = lane; t < R::kStride; t += 32) out[t] = qnan;
    }
    return;
  }
  double* theta = sTheta + (size_t)warp * 8 * J;
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
    // the 8 records of this warp: per-lane entry t of lane rr = K[i][4m + rr] (+ 2 jitter on the diagonal)
    for (int idx = lane; idx < 8 * R::kStride; idx += 32) {
      const int gg = idx / R::kStride, pos = idx - gg * R::kStride;
      const int s = s0 + gg;
      if (s >= S) break;
      double val = 0.0;  // also the tail padding of the record
      if (pos < R::kDoubles) {
        const int t = ((pos >> 3) << 1) | (pos & 1), rr = (pos >> 1) & 3;
        const int i = sPosI[t], j = 4 * sPosM[t] + rr;
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
      }
      p.G[((size_t)s * p.n_trees + tree) * rec_stride + pos] = val;
    }
  }
}

// log|det(K + jitter I)| of every record of a store: one 4-lane group per (sample, tree)
constexpr int kDetThreads = 256;
template <int VP>
__global__ void __launch_bounds__(kDetThreads) hellinger_logdet_kernel(const double* __restrict__ G, int64_t n_records, int V,
                                                                       double* __restrict__ logdet) {
  using R = HRec<VP>;
  const int tid = threadIdx.x, r = tid & 3;
  const int64_t rec = (int64_t)blockIdx.x * (kDetThreads / 4) + (tid >> 2);
  const double* in = G + (size_t)min(rec, n_records - 1) * R::kStride;
  double a[R::kT];
#pragma unroll
  for (int t = 0; t < R::kT; t += 2) {
    const double2 v = *reinterpret_cast<const double2*>(in + R::pos(t, r));
    a[t] = v.x;
    a[t + 1] = v.y;
  }
#pragma unroll
  for (int m = 0; m < R::kSlots; ++m)
#pragma unroll
    for (int rr = 0; rr < 4; ++rr)
      if (rr == r && 4 * m + rr < V) a[R::base(m) + rr] -= kHJitter;  // the store carries two jitters, the reference's Gram one
  double p_lo, p_hi;
  ldl_pivot_products<VP>(a, r, tid & 28, p_lo, p_hi);
  if (r == 0 && rec < n_records) logdet[rec] = log(fabs(p_lo)) + log(fabs(p_hi));
}

// =====================================================================================================
// pairs
// =====================================================================================================
constexpr int kPairThreads = 384;                   // 12 warps = 96 four-lane groups
constexpr int kPairTB = 8;                          // the 8 groups of a warp: one row tree x 8 column trees
constexpr int kPairPasses = 3;
constexpr int kPairTA = (kPairThreads / 32) * kPairPasses;  // 36 row trees per tile

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
struct PairSmem {
  using R = HRec<VP>;
  static constexpr int kRecs = kPairTA + kPairTB;
  static constexpr size_t kStageDoubles = (size_t)kRecs * R::kStride;
  static constexpr size_t oRoot = 2 * kStageDoubles;                                  // doubles; [2][kRecs] det^(1/4)
  static constexpr size_t oAcc = oRoot + 2 * kRecs;                                   // [TA*TB]
  static constexpr size_t oCnt = oAcc + kPairTA * kPairTB;                            // int[TA*TB]
  static constexpr size_t oBar = oCnt + (kPairTA * kPairTB + 1) / 2;                  // 2 mbarriers
  static constexpr size_t kBytes = (oBar + 2) * sizeof(double);
};

template <int VP>
__global__ void __launch_bounds__(kPairThreads, 1) hellinger_pair_kernel(PairParams p) {
  using R = HRec<VP>;
  using L = PairSmem<VP>;
  extern __shared__ __align__(16) uint8_t smem[];
  double* stage0 = reinterpret_cast<double*>(smem);
  double* sRoot = stage0 + L::oRoot;
  double* sAcc = stage0 + L::oAcc;
  int* sCnt = reinterpret_cast<int*>(stage0 + L::oCnt);
  uint64_t* full = reinterpret_cast<uint64_t*>(stage0 + L::oBar);

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
  auto issue = [&](int s) {  // operands of sample s: two TMA bulk copies into buffer s & 1
    double* dst = stage0 + (s & 1) * L::kStageDoubles;
    hl_mbar_expect_tx(&full[s & 1], bytes_a + bytes_b);
    hl_bulk_g2s(dst, p.Ga + ((size_t)s * p.n_a + a0) * R::kStride, bytes_a, &full[s & 1]);
    hl_bulk_g2s(dst + kPairTA * R::kStride, p.Gb + ((size_t)s * p.n_b + b0) * R::kStride, bytes_b, &full[s & 1]);
  };
  auto roots = [&](int s) {  // det(K + jitter I)^(1/4) of the tile's trees for sample s
    if (tid < na) sRoot[(s & 1) * L::kRecs + tid] = exp(0.25 * p.lda[(size_t)s * p.n_a + a0 + tid]);
    else if (tid >= kPairTA && tid < kPairTA + nb) sRoot[(s & 1) * L::kRecs + tid] = exp(0.25 * p.ldb[(size_t)s * p.n_b + b0 + tid - kPairTA]);
  };
  __syncthreads();
  if (tid == 0) issue(0);
  roots(0);
  __syncthreads();

  const double scale = exp2(0.5 * VP);  // det((K1' + K2') / 2) = det(K1' + K2') / 2^VP
  for (int s = 0; s < p.S; ++s) {
    if (s + 1 < p.S) {  // the other buffer was released by the barrier closing sample s - 1
      if (tid == 0) issue(s + 1);
      roots(s + 1);
    }
    hl_mbar_wait(&full[s & 1], (s >> 1) & 1);
    const double* sA = stage0 + (s & 1) * L::kStageDoubles;
    const double* sB = sA + kPairTA * R::kStride;
    const double* root = sRoot + (s & 1) * L::kRecs;
    const int bi = g;
    const double* Rb = sB + min(bi, nb - 1) * R::kStride;
#pragma unroll 1
    for (int u = 0; u < kPairPasses; ++u) {
      const int ai = w + (kPairThreads / 32) * u;  // the whole warp reads one row record (broadcast), 8 column records
      const double* Ra = sA + min(ai, na - 1) * R::kStride;
      double a[R::kT];
#pragma unroll
      for (int t = 0; t < R::kT; t += 2) {
        const double2 x = lds128(Ra + R::pos(t, r)), y = lds128(Rb + R::pos(t, r));
        a[t] = x.x + y.x;
        a[t + 1] = x.y + y.y;
      }
      double p_lo, p_hi;
      ldl_pivot_products<VP>(a, r, lane & 28, p_lo, p_hi);
      if (r == 0 && ai < na && bi < nb) {
        // 1 - exp(logdet1 / 4 + logdet2 / 4 - logdet_avg / 2)  (hellinger_distance, :185-190) without the logarithms
        const double hd = 1.0 - root[ai] * root[kPairTA + bi] * scale * rsqrt(fabs(p_lo * p_hi));
        if (hd >= 0.0) {  // false for NaN as well (:176-178)
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
  if (n_warps > 8) n_warps = 8;
  const size_t per_warp = (size_t)8 * J * sizeof(double);
  while (n_warps > 1 && n_warps * per_warp > (size_t)kHMaxParamsSmem) --n_warps;
  if (per_warp > (size_t)kHMaxParamsSmem) return set_error(BOSOT_EINVAL, "bosot_hellinger_grams: n_sobol_dims too large for shared memory");
  p.n_warps = n_warps;
  const size_t smem = sizeof(double) * (((V * d + 1) & ~1) + (size_t)n_warps * 8 * J) + 64 * 2 + 64 + ((J + 15) & ~15) + 2 * R::kT + 16;
  if (int rc = cuda_check(cudaFuncSetAttribute(hellinger_gram_kernel<VP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                          "cudaFuncSetAttribute(hellinger_gram_kernel)")) return rc;
  hellinger_gram_kernel<VP><<<(unsigned)p.n_trees, n_warps * 32, smem, st>>>(p);
  if (int rc = after_launch("hellinger_gram_kernel")) return rc;
  const int64_t n_records = (int64_t)p.sp.n_samples * p.n_trees;
  const int64_t blocks = (n_records + kDetThreads / 4 - 1) / (kDetThreads / 4);
  if (blocks > 0x7fffffffLL) return set_error(BOSOT_EINVAL, "bosot_hellinger_grams: too many records for one launch");
  hellinger_logdet_kernel<VP><<<(unsigned)blocks, kDetThreads, 0, st>>>(p.G, n_records, V, p.logdet);
  return after_launch("hellinger_logdet_kernel");
}

template <int VP>
int launch_pairs(PairParams& p, cudaStream_t st) {
  const size_t smem = PairSmem<VP>::kBytes;
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
