// GP posterior over kernel space + acquisition, for a block of candidates.
//
// Restates GPflow 2.4.0 GPR_with_posterior.predict_f (base_conditional) as reached from the
// reference's GPModel.predictive_dist (bosot/models/gp_model.py:290-306; PREDICT_F) with the
// ObjectConstant mean (bosot/models/object_mean_functions.py:40-48), and the EI / GP-UCB formulas
// of BayesianOptimizerObjects.acquisition_function
// (bosot/bayesian_optimization/bayesian_optimizer_objects.py:159-168):
//     A = Lm^-1 Kmn ; fvar = Knn - colsum(A*A) ; fmean = Kmn^T (Kmm + s^2 I)^-1 (y - c) + c
//     Knn = K_diag == variance  (the reference builds the whole N x N distance matrix for its
//     diagonal, which is exactly 0, kernel_kernel_grammar_tree.py:115-119)
// The E x E Cholesky stays with cuSOLVER (host side); this kernel consumes L^-1 (transposed)
// and beta = (Kmm + s^2 I)^-1 (y - c).
//
// Shape of the work: V[i][c] = sum_{j<=i} Linv[i][j] * K*[c][j]  -- a lower-triangular E x E
// matrix times an E x N panel in fp64, reduced on the fly to sum_i V[i][c]^2; V is never stored.
// This IS a dense fp64 contraction, so it runs on the fp64 tensor path: mma.sync m8n8k4 (DMMA;
// tcgen05 has no fp64 kind).  CTA = 64 candidates (K* tile staged in shared memory, B operand),
// L^-1 streamed from L2 exactly once per CTA (A operand): the 8 warps pull 16-row blocks of L^-1
// from a shared work counter, largest block first, so the triangular cost profile balances itself;
// partial sums go to a per-block slot and are added in block order => bitwise reproducible.
#include "sot_common.cuh"
#include <atomic>
#include "launch.cuh"

namespace bosot {

constexpr int kPostThreads = 256;
constexpr int kPostWarps = kPostThreads / 32;
constexpr int kPostCand = 64;            // candidates per CTA tile = 8 n-blocks of 8
constexpr int kNB = kPostCand / 8;
constexpr int kRowBlk = 16;              // rows of L^-1 per work unit = 2 m-blocks of 8

struct PostParams {
  const double* Kstar;
  int64_t n_cand;
  int64_t ld;
  int E;
  int ldl;  // padded order of L^-1: multiple of 16
  const double* Linv;  // FRAGMENT-MAJOR: LinvF[(rb * ldl/4 + ks) * 32 + 4 g + t] = Linv[8 rb + g][4 ks + t] (bosot_pack_linv)
  const double* beta;
  double mean_c, variance;
  int acq_kind;
  const double* cur_max;
  double xi, sqrt_ucb_beta;
  double* mu;
  double* sigma;
  double* acq;
  // optional scatter of the acquisition value into the row-sharded result vector of every GPU of the node
  // (peer memory over NVLink / NVSwitch): peers[r][peer_off + candidate] = acq, r = 0 .. n_peers-1
  double* const* peers;
  int n_peers;
  int64_t peer_off;
};

// K* tile row stride (doubles): == 4 or 12 (mod 16) so that the B-fragment read
// Ks[8*nb + lane/4][4*ks + lane%4] hits 16 distinct 8-byte banks per half warp
__host__ __device__ inline int post_tile_ld(int E) {
  const int e4 = (E + 3) & ~3;
  return (e4 % 8 == 4) ? e4 : e4 + 4;
}

__device__ __forceinline__ double norm_cdf(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }
__device__ __forceinline__ double norm_pdf(double z) { return 0.39894228040143267794 * exp(-0.5 * z * z); }

// D(8x8) += A(8x4, row) * B(4x8, col), fp64.  Lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// KT = double, or float for the fp32 entry points: K*, mu, sigma and acq are then float arrays (PostParams keeps
// them as double pointers), the tile is widened to fp64 while it is staged and everything else is unchanged.
template <typename KT>
__global__ void __launch_bounds__(kPostThreads, 2)
posterior_kernel(PostParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int E = p.E;
  const int tld = post_tile_ld(E);
  const int n_blk = (E + kRowBlk - 1) / kRowBlk;
  const int ks_max = (E + 3) >> 2;                            // k-steps that hold real columns
  double* Ks = reinterpret_cast<double*>(smem);              // [64][tld]
  double* red = Ks + (size_t)kPostCand * tld;                // [n_blk][64] per-block sums of v^2
  double* mus = red + (size_t)n_blk * kPostCand;             // [64]
  int* work = reinterpret_cast<int*>(mus + kPostCand);       // shared work counter
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  pdl_wait();  // K* may come from the launch directly in front of this one

  for (int64_t tile = blockIdx.x; tile * kPostCand < p.n_cand; tile += gridDim.x) {
    const int64_t c0 = tile * kPostCand;
    const int in_tile = (int)min((int64_t)kPostCand, p.n_cand - c0);
    __syncthreads();
    if (threadIdx.x == 0) *work = 0;
    // stage the K* tile (rows = candidates); rows beyond the tile and columns >= E are zero
    for (int x = threadIdx.x; x < kPostCand * tld; x += kPostThreads) {
      const int c = x / tld, j = x - c * tld;
      Ks[x] = (c < in_tile && j < E) ? (double)__ldg(reinterpret_cast<const KT*>(p.Kstar) + (c0 + c) * p.ld + j) : 0.0;
    }
    __syncthreads();

    // mean: 4 threads per candidate, each a quarter of the dot product with beta
    {
      const int c = threadIdx.x >> 2, q = threadIdx.x & 3;
      double s = 0.0;
      for (int j = q; j < E; j += 4) s = fma(Ks[c * tld + j], __ldg(&p.beta[j]), s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (q == 0) mus[c] = s + p.mean_c;
    }

    // variance: pull 16-row blocks of L^-1, largest first
    while (true) {
      int m = 0;
      if (lane == 0) m = atomicAdd(work, 1);
      m = n_blk - 1 - __shfl_sync(0xffffffffu, m, 0);
      if (m < 0) break;
      const int i0 = m * kRowBlk;
      const int nks = min(i0 / 4 + 4, ks_max);     // columns j < i0 + 16 (and < E4)
      const int nks_up = min(i0 / 4 + 2, ks_max);  // upper 8 rows only reach j < i0 + 8
      const bool lower_live = i0 + 8 < E;          // lower 8 rows may be pure padding in the last block
      double up[kNB][2], lo[kNB][2];
#pragma unroll
      for (int nb = 0; nb < kNB; ++nb) { up[nb][0] = up[nb][1] = 0.0; lo[nb][0] = lo[nb][1] = 0.0; }
      const int KS = p.ldl >> 2;
      const double* a_up = p.Linv + (size_t)(2 * m) * KS * 32 + lane;  // one coalesced 256-byte fragment per k-step
      const double* a_lo = a_up + (size_t)KS * 32;
      const double* bp = Ks + g * tld + t;
      double an_up = __ldg(a_up), an_lo = lower_live ? __ldg(a_lo) : 0.0;
      for (int ks = 0; ks < nks; ++ks) {
        const double au = an_up, al = an_lo;
        if (ks + 1 < nks) {  // prefetch the next A fragments (L2 latency) behind this step's 16 DMMAs
          an_up = (ks + 1 < nks_up) ? __ldg(a_up + 32 * (ks + 1)) : 0.0;
          an_lo = lower_live ? __ldg(a_lo + 32 * (ks + 1)) : 0.0;
        }
        double b[kNB];
#pragma unroll
        for (int nb = 0; nb < kNB; ++nb) b[nb] = bp[(size_t)nb * 8 * tld + 4 * ks];
        if (ks < nks_up) {
#pragma unroll
          for (int nb = 0; nb < kNB; ++nb) dmma8x8x4(up[nb][0], up[nb][1], au, b[nb]);
        }
        if (lower_live) {
#pragma unroll
          for (int nb = 0; nb < kNB; ++nb) dmma8x8x4(lo[nb][0], lo[nb][1], al, b[nb]);
        }
      }
      // sum of squares over the 16 rows: own rows first, then across the 8 row groups (lane bits 2..4)
#pragma unroll
      for (int nb = 0; nb < kNB; ++nb) {
        double s0 = fma(up[nb][0], up[nb][0], lo[nb][0] * lo[nb][0]);
        double s1 = fma(up[nb][1], up[nb][1], lo[nb][1] * lo[nb][1]);
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (g == 0) {
          red[m * kPostCand + nb * 8 + 2 * t] = s0;
          red[m * kPostCand + nb * 8 + 2 * t + 1] = s1;
        }
      }
    }
    __syncthreads();

    if (threadIdx.x < in_tile) {
      const int c = threadIdx.x;
      double ss = 0.0;
      for (int m = 0; m < n_blk; ++m) ss += red[m * kPostCand + c];
      const double var = p.variance - ss;
      const double sg = sqrt(var);
      const double mval = mus[c];
      if (p.mu) reinterpret_cast<KT*>(p.mu)[c0 + c] = (KT)mval;
      if (p.sigma) reinterpret_cast<KT*>(p.sigma)[c0 + c] = (KT)sg;
      if (p.acq) {
        double a;
        if (p.acq_kind == BOSOT_ACQ_EI) {
          const double d = mval - *p.cur_max - p.xi;
          const double z = d / sg;
          a = d * norm_cdf(z) + sg * norm_pdf(z);
        } else if (p.acq_kind == BOSOT_ACQ_UCB) {
          a = mval + p.sqrt_ucb_beta * sg;
        } else {
          a = mval;
        }
        reinterpret_cast<KT*>(p.acq)[c0 + c] = (KT)a;
        if (p.peers) {
#pragma unroll 1
          for (int r = 0; r < p.n_peers; ++r) p.peers[r][p.peer_off + c0 + c] = a;
        }
      }
    }
  }
}

// ---- warp-specialised variant -----------------------------------------------------------------
// One persistent CTA of 8 warps per SM (2 per scheduler: room for ~190 registers per thread).  K* tiles (TC
// candidates) live in a 2-stage shared-memory ring filled by TMA bulk row copies completing on "full" mbarriers.
// The warps pull (tile, unit) work items from ONE shared queue -- units of a tile = its 16-row blocks of L^-1,
// largest first, then the mean dot products -- so nobody ever waits at a CTA-wide barrier.  The warp that
// completes the last unit of a tile immediately issues the copies of the tile after next into the stage it just
// freed (no producer warp, no "empty" barrier) and runs the tile's epilogue behind that copy; units of the later
// tile check a flag before they overwrite the partial sums / means the epilogue is still reading.
constexpr int kWsConsumers = 8;
constexpr int kWsThreads = kWsConsumers * 32;

__device__ __forceinline__ uint32_t ws_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ws_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ws_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ws_tma_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ws_smem_u32(dst)), "l"(src), "r"(bytes), "r"(ws_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ws_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WS_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WS_WAIT_%=;\n"
      "}\n" ::"r"(ws_smem_u32(bar)), "r"(parity) : "memory");
}

struct WsLayout {
  size_t off_ks, off_red, off_mus, total;
  int tld, n_blk;
};
template <int TC>
__host__ __device__ inline WsLayout ws_layout(int E) {
  WsLayout W;
  W.tld = post_tile_ld(E);
  W.n_blk = (E + kRowBlk - 1) / kRowBlk;
  size_t o = 64;  // full[2], next_work, done[2], epi[2], issued[2]
  W.off_ks = o;   o += sizeof(double) * 2 * (size_t)TC * W.tld;
  W.off_red = o;  o += sizeof(double) * 2 * (size_t)W.n_blk * TC;
  W.off_mus = o;  o += sizeof(double) * 2 * TC;
  W.total = o;
  return W;
}

template <int TC>
__global__ void __launch_bounds__(kWsThreads, 1)
posterior_ws_kernel(PostParams p) {
  constexpr int NB = TC / 8;
  extern __shared__ __align__(16) uint8_t smem[];
  const int E = p.E;
  const WsLayout W = ws_layout<TC>(E);
  const int tld = W.tld, n_blk = W.n_blk, n_units = n_blk + 1;
  const int ks_max = (E + 3) >> 2;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  int* next_work = reinterpret_cast<int*>(smem + 32);
  int* done = next_work + 1;  // [2]
  volatile int* epi = done + 2;  // [2] set while the epilogue of the tile that used the stage before still reads its sums / means
  // issued[s] = 1 + number of the CTA tile whose copies were last issued into stage s.  full[s] alternates its phase
  // parity per use, and try_wait.parity on the parity of the phase BEFORE the current one returns true at once, so a
  // unit of tile i + 2 that reaches the barrier while tile i has not even landed (few units per tile: E <= 96, eight
  // warps claim units of four consecutive tiles) would sail through.  A unit therefore first waits until ITS tile
  // has been issued -- at that point every earlier phase of the barrier is complete and the parity wait is exact.
  volatile int* issued = epi + 2;  // [2]
  double* Ks = reinterpret_cast<double*>(smem + W.off_ks);
  double* red = reinterpret_cast<double*>(smem + W.off_red);
  double* mus = reinterpret_cast<double*>(smem + W.off_mus);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;

  const int64_t n_tiles = (p.n_cand + TC - 1) / TC;
  const int n_my = blockIdx.x < n_tiles ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  if (threadIdx.x == 0) {
    ws_mbar_init(&full[0], 1); ws_mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *next_work = 0; done[0] = 0; done[1] = 0; epi[0] = 0; epi[1] = 0; issued[0] = 0; issued[1] = 0;
  }
  for (int x = threadIdx.x; x < (int)((W.total - W.off_ks) / sizeof(double)); x += kWsThreads) Ks[x] = 0.0;  // everything finite
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");             // generic-proxy zeros before async-proxy writes
  __syncthreads();
  pdl_wait();  // everything above overlapped the tail of the launch in front (the pair kernel); K* is complete from here on

  // one warp issues the TMA row copies of this CTA's tile number i into stage i & 1 (every lane its share of the rows)
  const uint32_t row_bytes = (uint32_t)((E + 1) & ~1) * 8u;  // whole 16-byte units; for odd E one (finite) padding column comes along
  auto issue_tile = [&](int i) {
    const int s = i & 1;
    const int64_t c0 = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * TC;
    const int rows = (int)min((int64_t)TC, p.n_cand - c0);
    if (lane == 0) ws_mbar_expect_tx(&full[s], (uint32_t)rows * row_bytes);
    __syncwarp();
    double* dst = Ks + (size_t)s * TC * tld;
    for (int c = lane; c < rows; c += 32) ws_tma_g2s(dst + (size_t)c * tld, p.Kstar + (c0 + c) * p.ld, row_bytes, &full[s]);
    __syncwarp();
    if (lane == 0) { __threadfence_block(); issued[s] = i + 1; }  // after expect_tx: the barrier is in tile i's phase now
  };
  if (warp == 0) {
    if (n_my > 0) issue_tile(0);
    if (n_my > 1) issue_tile(1);
  }

  // ---------------- consumers
  while (true) {
    int w = 0;
    if (lane == 0) w = atomicAdd(next_work, 1);
    w = __shfl_sync(0xffffffffu, w, 0);
    const int i = w / n_units, u = w - i * n_units;
    if (i >= n_my) break;
    const int s = i & 1;
    while (issued[s] < i + 1) {}  // see issued[] above
    ws_mbar_wait(&full[s], (i >> 1) & 1);
    const double* Kt = Ks + (size_t)s * TC * tld;
    double* red_s = red + (size_t)s * n_blk * TC;
    double* mus_s = mus + (size_t)s * TC;

    if (u < n_blk) {
      const int m = n_blk - 1 - u;  // largest block first
      const int i0 = m * kRowBlk;
      const int nks = min(i0 / 4 + 4, ks_max);
      const int nks_up = min(i0 / 4 + 2, ks_max);
      const bool lower_live = i0 + 8 < E;
      double up[NB][2], lo[NB][2];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) { up[nb][0] = up[nb][1] = 0.0; lo[nb][0] = lo[nb][1] = 0.0; }
      // Software pipeline: A fragments (L^-1, from L2) run kPD k-steps ahead in a register ring, B fragments
      // (K* tile, shared memory) one k-step ahead, so the 16 DMMAs of a step never wait on a load of that step.
      // The k range is padded to a multiple of kPD; the extra steps only load, they issue no DMMA.
      constexpr int kPD = 4;
      const int KS = p.ldl >> 2;
      const double* a_up = p.Linv + (size_t)(2 * m) * KS * 32 + lane;  // one coalesced 256-byte fragment per k-step
      const double* a_lo = a_up + (size_t)KS * 32;
      const double* bp = Kt + g * tld + t;
      const int nks_pad = (nks + kPD - 1) & ~(kPD - 1);
      // L^-1 fragments are read once per tile and shared by nobody on this SM: ld.global.cg (L2 only) keeps them out of the
      // small L1 that is left beside 210 KB of shared memory (1.644 -> 1.621 ms at 1M x 200 against ld.global.nc, on-box A/B)
      double ra_up[kPD], ra_lo[kPD], bf[2][NB];  // B fragments ping-pong between two register sets (kPD is even)
#pragma unroll
      for (int u = 0; u < kPD; ++u) {
        ra_up[u] = __ldcg(a_up + 32 * u);
        ra_lo[u] = lower_live ? __ldcg(a_lo + 32 * u) : 0.0;
      }
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) bf[0][nb] = bp[(size_t)nb * 8 * tld];
      // whole groups of kPD steps below nks run unguarded; the last, partly padded group guards its DMMAs
      const int n_full = nks & ~(kPD - 1);
      for (int ks0 = 0; ks0 < n_full; ks0 += kPD) {
#pragma unroll
        for (int u = 0; u < kPD; ++u) {
          const int ks = ks0 + u;
          const double au = ra_up[u], al = ra_lo[u];
          if (ks + kPD < nks_pad) {
            ra_up[u] = __ldcg(a_up + 32 * (ks + kPD));
            ra_lo[u] = lower_live ? __ldcg(a_lo + 32 * (ks + kPD)) : 0.0;
          }
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) bf[(u + 1) & 1][nb] = bp[(size_t)nb * 8 * tld + 4 * (ks + 1)];
          if (ks < nks_up) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) dmma8x8x4(up[nb][0], up[nb][1], au, bf[u & 1][nb]);
          }
          if (lower_live) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) dmma8x8x4(lo[nb][0], lo[nb][1], al, bf[u & 1][nb]);
          }
        }
      }
      if (n_full < nks_pad) {
#pragma unroll
        for (int u = 0; u < kPD; ++u) {
          const int ks = n_full + u;
          // the B fragments of the padding steps (ks >= nks) may lie past the row, in the next candidate's row, which
          // holds NaN when that record was malformed: they are loaded (uniform pipeline) but never multiplied
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) bf[(u + 1) & 1][nb] = bp[(size_t)nb * 8 * tld + 4 * (ks + 1)];
          if (ks < nks_up) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) dmma8x8x4(up[nb][0], up[nb][1], ra_up[u], bf[u & 1][nb]);
          }
          if (lower_live && ks < nks) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) dmma8x8x4(lo[nb][0], lo[nb][1], ra_lo[u], bf[u & 1][nb]);
          }
        }
      }
      while (epi[s]) {}  // the previous tile's epilogue may still be reading this stage's sums
      // Column sums of squares over the 16 rows: own rows first, then across the 8 row groups (lane bits 2..4) by a
      // halving butterfly -- at every level a lane keeps one half of its values and receives the partner's half,
      // so 2 NB values per lane shrink to 2 NB / 8 fully reduced ones with 2 NB - 2 NB / 8 exchanges.
      double v[2 * NB];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        v[2 * nb] = fma(up[nb][0], up[nb][0], lo[nb][0] * lo[nb][0]);
        v[2 * nb + 1] = fma(up[nb][1], up[nb][1], lo[nb][1] * lo[nb][1]);
      }
      int n = 2 * NB, base = 0;
      bool writer = true;
#pragma unroll
      for (int lvl = 0; lvl < 3; ++lvl) {
        const int o = 4 << lvl;
        const bool hi = (lane & o) != 0;
        if (n >= 2) {
          n >>= 1;
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            if (j < n) {
              const double send = hi ? v[j] : v[j + n];
              const double keep = hi ? v[j + n] : v[j];
              v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          base += hi ? n : 0;
        } else {
          v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
          writer = writer && !hi;
        }
      }
      // the n values left on this lane are original indices base .. base + n - 1 (index = 2 nb + column parity)
      if (n == 2) {
        *reinterpret_cast<double2*>(&red_s[m * TC + (base >> 1) * 8 + 2 * t]) = make_double2(v[0], v[1]);
      } else if (writer) {
        red_s[m * TC + (base >> 1) * 8 + 2 * t + (base & 1)] = v[0];
      }
    } else {
      // mean unit: K* . beta for the TC candidates of the tile (same 4-way split per candidate as the generic kernel)
      while (epi[s]) {}
      for (int c = lane >> 2; c < TC; c += 8) {
        const int q = lane & 3;
        double sacc = 0.0;
        for (int j = q; j < E; j += 4) sacc = fma(Kt[c * tld + j], __ldg(&p.beta[j]), sacc);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
        if (q == 0) mus_s[c] = sacc + p.mean_c;
      }
    }

    // publish this unit; the warp that completes the tile runs its epilogue and frees the stage
    __threadfence_block();
    int d = 0;
    if (lane == 0) d = atomicAdd(&done[s], 1);
    d = __shfl_sync(0xffffffffu, d, 0);
    if (d == n_units - 1) {
      __threadfence_block();
      // the K* tile of this stage is dead: refill it NOW (tile after next) and run the epilogue behind the copy; units
      // of that later tile wait on epi[s] before they overwrite the sums / means the epilogue is reading
      if (lane == 0) { epi[s] = 1; done[s] = 0; }
      __threadfence_block();
      __syncwarp();
      if (i + 2 < n_my) issue_tile(i + 2);
      const int64_t c0 = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * TC;
      const int in_tile = (int)min((int64_t)TC, p.n_cand - c0);
      for (int c = lane; c < in_tile; c += 32) {
        double ss = 0.0;
        for (int m = 0; m < n_blk; ++m) ss += red_s[m * TC + c];
        const double var = p.variance - ss;
        const double sg = sqrt(var);
        const double mval = mus_s[c];
        if (p.mu) p.mu[c0 + c] = mval;
        if (p.sigma) p.sigma[c0 + c] = sg;
        if (p.acq) {
          double a;
          if (p.acq_kind == BOSOT_ACQ_EI) {
            const double dd = mval - *p.cur_max - p.xi;
            const double z = dd / sg;
            a = dd * norm_cdf(z) + sg * norm_pdf(z);
          } else if (p.acq_kind == BOSOT_ACQ_UCB) {
            a = mval + p.sqrt_ucb_beta * sg;
          } else {
            a = mval;
          }
          p.acq[c0 + c] = a;
          if (p.peers) {
#pragma unroll 1
            for (int r = 0; r < p.n_peers; ++r) p.peers[r][p.peer_off + c0 + c] = a;
          }
        }
      }
      __syncwarp();
      __threadfence_block();
      __syncwarp();
      if (lane == 0) epi[s] = 0;
    }
  }
}

// row-major L^-1 (n x n, leading dimension ld_src) -> zero-padded fragment-major copy of order ldl
__global__ void pack_linv_kernel(const double* __restrict__ src, int n, int ld_src, double* __restrict__ dst, int ldl) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= ldl * ldl) return;
  const int KS = ldl >> 2;
  const int l = x & 31, f = x >> 5;       // lane within fragment, fragment index = rb * KS + ks
  const int rb = f / KS, ks = f - rb * KS;
  const int i = 8 * rb + (l >> 2), j = 4 * ks + (l & 3);
  dst[x] = (i < n && j < n && j <= i) ? src[(size_t)i * ld_src + j] : 0.0;
}

__global__ void max_f64_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ out) {
  __shared__ double part[32];
  double m = -INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, x[i]);
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : -INFINITY;
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) *out = m;
  }
}

__global__ void fp64_fma_probe_kernel(double* out, int iters) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  const double m = 1.0 - 1e-12, c = 1e-13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keep the chain alive
}

__global__ void fp64_mma_probe_kernel(double* out, int iters) {
  double d[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k) { d[k][0] = 0.0; d[k][1] = 0.0; }
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) dmma8x8x4(d[k][0], d[k][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += d[k][0] + d[k][1];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// register-pattern probes: NA A-fragments x NB B-fragments -> NA*NB accumulator tiles per step (no loads).
// A_OUTER: for a { for b { mma(acc[a][b], A[a], B[b]) } }  (A stationary), else B stationary.
template <int NA, int NBF, bool A_OUTER>
__global__ void fp64_mma_pattern_kernel(double* out, int iters) {
  double acc[NA][NBF][2], a[NA], b[NBF];
#pragma unroll
  for (int i = 0; i < NA; ++i) { a[i] = 1.0 - 1e-9 * (threadIdx.x + i);
#pragma unroll
    for (int j = 0; j < NBF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0; }
#pragma unroll
  for (int j = 0; j < NBF; ++j) b[j] = 1.0 + 1e-9 * (threadIdx.x + j);
  for (int it = 0; it < iters; ++it) {
    if (A_OUTER) {
#pragma unroll
      for (int i = 0; i < NA; ++i)
#pragma unroll
        for (int j = 0; j < NBF; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    } else {
#pragma unroll
      for (int j = 0; j < NBF; ++j)
#pragma unroll
        for (int i = 0; i < NA; ++i) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < NA; ++i) a[i] += 1e-12;
#pragma unroll
    for (int j = 0; j < NBF; ++j) b[j] += 1e-13;
  }
  double sacc = 0.0;
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < NBF; ++j) sacc += acc[i][j][0] + acc[i][j][1];
  if (sacc == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = sacc;
}

// even warps: DMMA chains, odd warps: DFMA chains -- do the two share the fp64 execution units?
__global__ void fp64_mixed_probe_kernel(double* out, int iters) {
  const bool tensor = ((threadIdx.x >> 5) & 1) == 0;
  double d[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k) { d[k][0] = 1.0 + 1e-9 * k; d[k][1] = 0.5; }
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x, c = 1e-13;
  if (tensor) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dmma8x8x4(d[k][0], d[k][1], a, b);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { d[k][0] = fma(d[k][0], b, c); d[k][1] = fma(d[k][1], b, c); }
#pragma unroll
      for (int k = 0; k < 8; ++k) { d[k][0] = fma(d[k][0], b, c); d[k][1] = fma(d[k][1], b, c); }
#pragma unroll
      for (int k = 0; k < 8; ++k) { d[k][0] = fma(d[k][0], b, c); d[k][1] = fma(d[k][1], b, c); }
#pragma unroll
      for (int k = 0; k < 8; ++k) { d[k][0] = fma(d[k][0], b, c); d[k][1] = fma(d[k][1], b, c); }
    }
  }
  double sacc = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) sacc += d[k][0] + d[k][1];
  if (sacc == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = sacc;
}

int launch_posterior(const double* d_Kstar, int64_t n_cand, int64_t ld, int n_eval, const double* d_Linv, int ldl,
                     const double* d_beta, double mean_c, double variance, int acq_kind, const double* d_current_max,
                     double xi, double ucb_beta, double* d_mu, double* d_sigma, double* d_acq, cudaStream_t stream,
                     double* const* d_peers, int n_peers, int64_t peer_off, bool f32) {
  if (!d_Kstar || !d_Linv || !d_beta || n_cand < 0) return set_error(BOSOT_EINVAL, "bosot_posterior_acq: null argument");
  if (n_eval < 1 || n_eval > BOSOT_MAX_EVAL || ld < n_eval) return set_error(BOSOT_EINVAL, "bosot_posterior_acq: bad n_eval / ld");
  if (ldl < n_eval || (ldl % kRowBlk) != 0)
    return set_error(BOSOT_EINVAL, "bosot_posterior_acq: Linv must be [ldl][ldl] with ldl a multiple of 16 and >= n_eval");
  if (acq_kind == BOSOT_ACQ_EI && d_acq && !d_current_max) return set_error(BOSOT_EINVAL, "bosot_posterior_acq: EI needs current_max");
  if (f32 && d_peers) return set_error(BOSOT_EINVAL, "bosot_posterior_acq: the peer scatter is fp64 only");
  if (d_peers && (!d_acq || n_peers < 1 || n_peers > 64 || peer_off < 0))
    return set_error(BOSOT_EINVAL, "bosot_posterior_acq: the peer scatter needs d_acq, 1 <= n_peers <= 64 and a non-negative offset");
  if (n_cand == 0) return BOSOT_OK;

  PostParams p;
  p.Kstar = d_Kstar; p.n_cand = n_cand; p.ld = ld; p.E = n_eval; p.ldl = ldl;
  p.Linv = d_Linv; p.beta = d_beta; p.mean_c = mean_c; p.variance = variance;
  p.acq_kind = acq_kind; p.cur_max = d_current_max; p.xi = xi; p.sqrt_ucb_beta = sqrt(ucb_beta);
  p.mu = d_mu; p.sigma = d_sigma; p.acq = d_acq;
  p.peers = d_peers; p.n_peers = d_peers ? n_peers : 0; p.peer_off = peer_off;

  int dev = 0, sms = 0;
  if (int rc = cuda_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (int rc = cuda_check(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) return rc;
  const size_t smem_cap = 226 * 1024;
  const bool early = n_cand <= 32768;  // programmatic dependent launch only where launch latency matters (see launch_pairs)

  // warp-specialised TMA variant: needs 16-byte aligned rows of whole 16-byte units
  const bool tma_ok = !f32 && (ld % 2 == 0) && (n_eval % 2 == 0 || ld > n_eval) && !(reinterpret_cast<uintptr_t>(d_Kstar) & 15);
  if (tma_ok) {
    static std::atomic<bool> ws_attr[64];  // zero-initialised; the (idempotent) set-up may run twice under a race, never zero times
    if (!(dev < 64 && ws_attr[dev])) {
      if (int rc = cuda_check(cudaFuncSetAttribute(posterior_ws_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap),
                              "cudaFuncSetAttribute(posterior_ws_kernel<64>)")) return rc;
      if (int rc = cuda_check(cudaFuncSetAttribute(posterior_ws_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap),
                              "cudaFuncSetAttribute(posterior_ws_kernel<32>)")) return rc;
      if (int rc = cuda_check(cudaFuncSetAttribute(posterior_ws_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap),
                              "cudaFuncSetAttribute(posterior_ws_kernel<16>)")) return rc;
      if (dev < 64) ws_attr[dev] = true;
    }
    // Tile width: 64 candidates per CTA tile is the efficient shape (every A fragment of L^-1 feeds 16 DMMAs), but a
    // small pool must still put work on every SM: take the widest tile that fits shared memory and still gives every
    // SM two tiles, else the narrowest one (10k candidates: 313 tiles of 32; the reference's own call size of 100
    // candidates: 7 tiles of 16 instead of 2 of 64).
    const size_t smem_tc[3] = {ws_layout<64>(n_eval).total, ws_layout<32>(n_eval).total, ws_layout<16>(n_eval).total};
    const int width[3] = {64, 32, 16};
    int pick = -1;
    for (int i = 0; i < 3 && pick < 0; ++i)
      if (smem_tc[i] <= smem_cap && (n_cand + width[i] - 1) / width[i] >= 2LL * sms) pick = i;
    for (int i = 2; i >= 0 && pick < 0; --i)
      if (smem_tc[i] <= smem_cap) pick = i;
    if (pick >= 0) {
      const int64_t tiles = (n_cand + width[pick] - 1) / width[pick];
      const dim3 grid((unsigned)(tiles < sms ? tiles : sms)), block(kWsThreads);
      cudaError_t e = cudaSuccess;
      if (pick == 0) e = launch_pdl(early, posterior_ws_kernel<64>, grid, block, smem_tc[0], stream, p);
      else if (pick == 1) e = launch_pdl(early, posterior_ws_kernel<32>, grid, block, smem_tc[1], stream, p);
      else e = launch_pdl(early, posterior_ws_kernel<16>, grid, block, smem_tc[2], stream, p);
      if (int rc = cuda_check(e, "cudaLaunchKernelEx(posterior_ws_kernel)")) return rc;
      return after_launch("posterior_ws_kernel");
    }
  }

  const int n_blk = (n_eval + kRowBlk - 1) / kRowBlk;
  const size_t smem = sizeof(double) * ((size_t)kPostCand * post_tile_ld(n_eval) + (size_t)n_blk * kPostCand + kPostCand) + 16;
  if (smem > smem_cap) return set_error(BOSOT_EINVAL, "bosot_posterior_acq: n_eval too large for shared memory");
  static std::atomic<bool> attr_set[64];  // zero-initialised; the (idempotent) set-up may run twice under a race, never zero times
  if (!(dev < 64 && attr_set[dev])) {
    if (int rc = cuda_check(cudaFuncSetAttribute(posterior_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap),
                            "cudaFuncSetAttribute(posterior_kernel)")) return rc;
    if (int rc = cuda_check(cudaFuncSetAttribute(posterior_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap),
                            "cudaFuncSetAttribute(posterior_kernel)")) return rc;
    if (dev < 64) attr_set[dev] = true;
  }
  void (*generic)(PostParams) = f32 ? posterior_kernel<float> : posterior_kernel<double>;
  int per_sm = 1;
  if (int rc = cuda_check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, generic, kPostThreads, smem),
                          "cudaOccupancyMaxActiveBlocksPerMultiprocessor")) return rc;
  if (per_sm < 1) per_sm = 1;
  const int64_t tiles = (n_cand + kPostCand - 1) / kPostCand;
  const int64_t wave = (int64_t)sms * per_sm;
  const int64_t grid = tiles < wave ? tiles : wave;
  if (int rc = cuda_check(launch_pdl(early, generic, dim3((unsigned)grid), dim3(kPostThreads), smem, stream, p), "cudaLaunchKernelEx(posterior_kernel)")) return rc;
  return after_launch("posterior_kernel");
}

}  // namespace bosot

using namespace bosot;

extern "C" int bosot_pack_linv(const double* d_Linv, int n, int ld_src, double* d_LinvF, int ldl, void* stream) {
  if (!d_Linv || !d_LinvF || n < 1 || ld_src < n || ldl < n || (ldl % 16) != 0) return set_error(BOSOT_EINVAL, "bosot_pack_linv: bad argument");
  pack_linv_kernel<<<(ldl * ldl + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_Linv, n, ld_src, d_LinvF, ldl);
  return after_launch("pack_linv_kernel");
}

extern "C" int bosot_max_f64(const double* d_x, int64_t n, double* d_out, void* stream) {
  if (!d_x || !d_out || n < 1) return set_error(BOSOT_EINVAL, "bosot_max_f64: bad argument");
  max_f64_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_x, n, d_out);
  return after_launch("max_f64_kernel");
}

extern "C" int bosot_fp64_mma_probe(double* d_out, int n_blocks, int n_threads, int iters, void* stream) {
  if (!d_out || n_blocks < 1 || n_threads < 32 || n_threads > 1024 || iters < 1) return set_error(BOSOT_EINVAL, "bosot_fp64_mma_probe: bad argument");
  fp64_mma_probe_kernel<<<n_blocks, n_threads, 0, (cudaStream_t)stream>>>(d_out, iters);
  return after_launch("fp64_mma_probe_kernel");
}

extern "C" int bosot_fp64_mma_probe2(double* d_out, int n_blocks, int n_threads, int iters, void* stream) {
  if (!d_out || n_blocks < 1 || n_threads < 32 || n_threads > 1024 || iters < 1) return set_error(BOSOT_EINVAL, "bosot_fp64_mma_probe2: bad argument");
  // `iters` low 4 bits select the pattern (measurement tool only)
  const int pat = iters & 15;
  const int it = iters >> 4;
  cudaStream_t st = (cudaStream_t)stream;
  switch (pat) {
    case 0: fp64_mma_pattern_kernel<2, 8, true><<<n_blocks, n_threads, 0, st>>>(d_out, it); break;
    case 1: fp64_mma_pattern_kernel<2, 8, false><<<n_blocks, n_threads, 0, st>>>(d_out, it); break;
    case 2: fp64_mma_pattern_kernel<4, 4, true><<<n_blocks, n_threads, 0, st>>>(d_out, it); break;
    case 3: fp64_mma_pattern_kernel<4, 4, false><<<n_blocks, n_threads, 0, st>>>(d_out, it); break;
    case 4: fp64_mma_pattern_kernel<8, 2, true><<<n_blocks, n_threads, 0, st>>>(d_out, it); break;
    case 5: fp64_mma_pattern_kernel<8, 2, false><<<n_blocks, n_threads, 0, st>>>(d_out, it); break;
    case 6: fp64_mma_pattern_kernel<1, 8, true><<<n_blocks, n_threads, 0, st>>>(d_out, it); break;
    case 7: fp64_mma_pattern_kernel<8, 1, false><<<n_blocks, n_threads, 0, st>>>(d_out, it); break;
    case 8: fp64_mma_pattern_kernel<1, 16, true><<<n_blocks, n_threads, 0, st>>>(d_out, it); break;
    case 9: fp64_mma_pattern_kernel<16, 1, false><<<n_blocks, n_threads, 0, st>>>(d_out, it); break;
    case 15: fp64_mixed_probe_kernel<<<n_blocks, n_threads, 0, st>>>(d_out, it); break;
    default: return set_error(BOSOT_EINVAL, "bosot_fp64_mma_probe2: unknown pattern");
  }
  return after_launch("fp64_mma_pattern_kernel");
}

extern "C" int bosot_fp64_fma_probe(double* d_out, int n_blocks, int n_threads, int iters, void* stream) {
  if (!d_out || n_blocks < 1 || n_threads < 32 || n_threads > 1024 || iters < 1) return set_error(BOSOT_EINVAL, "bosot_fp64_fma_probe: bad argument");
  fp64_fma_probe_kernel<<<n_blocks, n_threads, 0, (cudaStream_t)stream>>>(d_out, iters);
  return after_launch("fp64_fma_probe_kernel");
}
