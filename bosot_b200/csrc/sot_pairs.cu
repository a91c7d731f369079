// Fused SOT pair kernel: flat candidate trees -> per-feature L1 distances -> weighted SOT
// distance -> Gram entries, against the inverted index of the evaluated set.
//
// Restates, for every (candidate, evaluated) pair, the reference's
//   wasserstein_distance(X, X2)   bosot/kernels/kernel_kernel_grammar_tree.py:143-161
//   manhatten_distance            bosot/utils/utils.py:168-170
//   K(X, X2)                      bosot/kernels/kernel_kernel_grammar_tree.py:110-113
// without ever forming the dense [N, F] / [N, M, F] tensors: for histograms normalised to
// sum 1, sum_k |a_k - b_k| = 2 - 2 * sum_k min(a_k, b_k), and with a_k = ca/Ta, b_k = cb/Tb
// the sum of minima is the exact integer M = sum_k min(ca*Tb, cb*Ta) over Ta*Tb.  Only
// features present in BOTH trees contribute, so one candidate costs O(matches), found
// through the evaluated set's postings, instead of O(E * F).
//
// Two launches per candidate batch (HBM is nearly idle on this path, so the 320-byte compact
// record per tree that connects them costs ~0.1 ms per 10^6 trees and buys twice the occupancy):
//   scan_kernel      thread-per-tree : scan_tree() -> class ids, leaf symbols, reduced path codes,
//                                      written as compact records (coalesced) to global memory
//   sot_pair_kernel  warp-per-tree   : phase B  dedupe with __match_any_sync; operator classes go
//                                      through the hash table, reduced paths through a direct table,
//                                      and their (short) posting lists are walked into shared-memory
//                                      accumulators;
//                                      phase C  every lane owns R evaluated trees in registers:
//                                      leaf-symbol classes (dense: most symbols occur in most trees)
//                                      against the transposed count table, DIM_WISE dense L1,
//                                      combine, exp, store.  The next tree's record is prefetched
//                                      into registers while the current one is processed.
#include "sot_common.cuh"
#include "launch.cuh"
#include "fast_math.cuh"

namespace bosot {

constexpr int kMaxPairWarps = 32;   // one persistent CTA of up to 32 warps per SM
constexpr int kScanThreads = 128;
constexpr int kMaxR = 8;            // evaluated trees per lane per register tile (E <= 256 in one pass, else tiles of 256)

__constant__ double c_recip[64];  // 1/T for T = 1..63 ([0] unused)

struct PairParams {
  const uint8_t* trees;
  int64_t n_cand;
  const uint8_t* blob;
  EvalLayout L;
  double w_dim, w_path, w_sub;
  double variance, neg_inv_lsq;  // -1 / lengthscale^2
  double* K;
  double* D;
  double* Df;
  int64_t ld;
  const uint8_t* recs;  // compact records written by scan_kernel: [n_cand][kScanBytes]
};

// per-warp shared memory: candidate dim-wise vector, accumulators, symbol counts
__host__ __device__ inline size_t pair_warp_smem(int e_pad) {
  return sizeof(double) * BOSOT_MAX_DIM_COLS + sizeof(uint32_t) * (size_t)e_pad +
         sizeof(uint32_t) * (BOSOT_MAX_SYMBOLS + BOSOT_MAX_BLOCKS);
}

// CTA-wide staged copy of the evaluated set's dense tables (section sizes are multiples of 16 bytes)
struct StageLayout {
  size_t off_dimT, off_rsub, off_rpath, off_symT, off_nn, off_nl, total;
  uint32_t b_dimT, b_r, b_symT, b_n;
};
__host__ __device__ inline StageLayout stage_layout(int e_pad, int n_dim_cols, int n_symbols) {
  StageLayout S;
  S.b_dimT = (uint32_t)(sizeof(double) * (size_t)n_dim_cols * e_pad);
  S.b_r = (uint32_t)(sizeof(double) * (size_t)e_pad);
  S.b_symT = (uint32_t)((size_t)n_symbols * e_pad);
  S.b_n = (uint32_t)e_pad;
  size_t o = 16;  // mbarrier
  S.off_dimT = o;  o += S.b_dimT;
  S.off_rsub = o;  o += S.b_r;
  S.off_rpath = o; o += S.b_r;
  S.off_symT = o;  o += S.b_symT;
  S.off_nn = o;    o += S.b_n;
  S.off_nl = o;    o += S.b_n;
  S.total = align16(o);
  return S;
}

// ---- TMA bulk copy (global -> shared) completing on an mbarrier -----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---- launch 1: thread-per-tree scan -> compact records ------------------------------------
__global__ void __launch_bounds__(kScanThreads)
scan_kernel(const uint8_t* __restrict__ trees, int64_t n_trees, int n_symbols, uint8_t* __restrict__ recs_out,
            int32_t* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int64_t base = (int64_t)blockIdx.x * kScanThreads;
  const int in_block = (int)min((int64_t)kScanThreads, n_trees - base);
  for (int w = threadIdx.x; w < in_block * 4; w += kScanThreads) {  // coalesced: 16 B per thread
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(trees + base * kTreeBytes) + w);
    uint2* dst = reinterpret_cast<uint2*>(smem + (w >> 2) * kRecBytes + kRecTok + (w & 3) * 16);  // 8-byte aligned
    dst[0] = make_uint2(v.x, v.y);
    dst[1] = make_uint2(v.z, v.w);
  }
  __syncthreads();
  if ((int)threadIdx.x < in_block) {
    const int st = scan_tree(smem + threadIdx.x * kRecBytes, n_symbols);
    if (status) status[base + threadIdx.x] = st;
  }
  __syncthreads();
  // coalesced copy-out of the compact part (320 B = 40 x 8 B per record)
  uint2* out = reinterpret_cast<uint2*>(recs_out + base * kScanBytes);
  for (int w = threadIdx.x; w < in_block * (kScanBytes / 8); w += kScanThreads) {
    const int t = w / (kScanBytes / 8), q = w % (kScanBytes / 8);
    out[w] = *reinterpret_cast<const uint2*>(smem + t * kRecBytes + q * 8);
  }
}

// ---- launch 2: warp-per-tree pairs ----------------------------------------------------------
template <int R, bool kExtra>  // R: evaluated trees per lane per tile; kExtra: also write D and/or Df
__global__ void __launch_bounds__(kMaxPairWarps * 32, 1)
sot_pair_kernel(PairParams p, bosot_grammar g) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int n_warps = blockDim.x >> 5;
  const int E = p.L.n_eval;
  const int e_pad = p.L.e_pad;
  const int F = g.n_dim_cols;
  const unsigned full = 0xffffffffu;

  // ---- stage the evaluated set's dense tables once per (persistent) CTA: TMA bulk copies -> mbarrier
  const StageLayout S = stage_layout(e_pad, F, g.n_symbols);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, S.b_dimT + 2 * S.b_r + S.b_symT + 2 * S.b_n);
    tma_bulk_g2s(smem + S.off_dimT, p.blob + p.L.off_dimT, S.b_dimT, bar);
    tma_bulk_g2s(smem + S.off_rsub, p.blob + p.L.off_rsub, S.b_r, bar);
    tma_bulk_g2s(smem + S.off_rpath, p.blob + p.L.off_rpath, S.b_r, bar);
    tma_bulk_g2s(smem + S.off_symT, p.blob + p.L.off_symT, S.b_symT, bar);
    tma_bulk_g2s(smem + S.off_nn, p.blob + p.L.off_nnodes, S.b_n, bar);
    tma_bulk_g2s(smem + S.off_nl, p.blob + p.L.off_nleaves, S.b_n, bar);
  }
  const double* dimT = reinterpret_cast<const double*>(smem + S.off_dimT);
  const double* e_rsub = reinterpret_cast<const double*>(smem + S.off_rsub);
  const double* e_rpath = reinterpret_cast<const double*>(smem + S.off_rpath);
  const uint8_t* symT = smem + S.off_symT;
  const uint8_t* e_nn = smem + S.off_nn;
  const uint8_t* e_nl = smem + S.off_nl;

  uint8_t* wbase = smem + S.total + (size_t)warp * pair_warp_smem(e_pad);
  double* avec = reinterpret_cast<double*>(wbase);                         // candidate DIM_WISE vector
  uint32_t* acc = reinterpret_cast<uint32_t*>(avec + BOSOT_MAX_DIM_COLS);  // M_sub(ops) | M_path << 16 per e
  uint32_t* scnt = acc + e_pad;                                            // leaf count per symbol
  uint32_t* btot = scnt + BOSOT_MAX_SYMBOLS;                               // leaf count per block

  const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(p.blob + p.L.off_keys);
  const uint2* range = reinterpret_cast<const uint2*>(p.blob + p.L.off_range);
  const uint32_t* post = reinterpret_cast<const uint32_t*>(p.blob + p.L.off_post);
  const int hash_cap = p.L.hash_cap;
  const size_t plane = (size_t)p.n_cand * p.ld;

  // one lane-slice of a compact record: class id of operator node `lane`, symbol / path code of leaf `lane`
  struct Slice { unsigned long long id; uint32_t meta, sym, code; };
  auto load_slice = [&](int64_t t) {
    Slice s;
    const uint8_t* r = p.recs + (size_t)t * kScanBytes;
    s.meta = __ldg(reinterpret_cast<const uint32_t*>(r + kRecMeta));
    s.id = lane < kMaxInternal ? __ldg(reinterpret_cast<const unsigned long long*>(r + kRecIid) + lane) : 0ULL;
    s.sym = __ldg(r + kRecLeafSym + lane);
    s.code = __ldg(r + kRecLeafCode + lane);
    return s;
  };

  const int64_t stride = (int64_t)gridDim.x * n_warps;
  int64_t t = (int64_t)blockIdx.x * n_warps + warp;
  Slice nxt;
  if (t < p.n_cand) nxt = load_slice(t);
  mbar_wait(bar, 0);  // tables landed (the copy overlapped the first record load)

  for (; t < p.n_cand; t += stride) {
    const Slice cur = nxt;
    if (t + stride < p.n_cand) nxt = load_slice(t + stride);  // prefetch behind this tree's work

    const int nn = cur.meta & 0xff, nl = (cur.meta >> 8) & 0xff, ni = (cur.meta >> 16) & 0xff;
    const bool bad = (cur.meta >> 24) != 0;
    const size_t row = (size_t)t * p.ld;

    if (bad) {  // malformed record: poison the row, status already written by the scan (cold path: keep it small)
      const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll 1
      for (int e = lane; e < E; e += 32) {
        if (p.K) p.K[row + e] = qnan;
        if (p.D) p.D[row + e] = qnan;
        if (p.Df) { p.Df[row + e] = qnan; p.Df[plane + row + e] = qnan; p.Df[2 * plane + row + e] = qnan; }
      }
      continue;
    }

    __syncwarp();
#pragma unroll 1
    for (int e = lane; e < e_pad; e += 32) acc[e] = 0u;
#pragma unroll 1
    for (int s = lane; s < BOSOT_MAX_SYMBOLS + BOSOT_MAX_BLOCKS; s += 32) scnt[s] = 0u;
#pragma unroll 1
    for (int c = lane; c < F; c += 32) avec[c] = 0.0;
    __syncwarp();

    // leaves: symbol multiplicities (SUBTREES leaf classes, DIM_WISE) and PATHS (key = symbol * 64 + code)
    const bool is_leaf = lane < nl;
    const uint32_t sym = is_leaf ? cur.sym : 0u;
    const uint32_t pid = is_leaf ? sym * kPathCodes + cur.code : 0u;
    const unsigned m_sym = __match_any_sync(full, is_leaf ? sym : 0x1000u + lane);
    const unsigned m_pid = __match_any_sync(full, is_leaf ? pid : 0x10000u + lane);
    const bool lead_sym = is_leaf && (__ffs(m_sym) - 1 == lane);
    const bool lead_pid = is_leaf && (__ffs(m_pid) - 1 == lane);
    const uint32_t cnt_sym = (uint32_t)__popc(m_sym);
    uint2 r_pid = make_uint2(0u, 0u);
    if (lead_sym) {
      scnt[sym] = cnt_sym;
      if (g.sym_block[sym] >= 0) atomicAdd(&btot[g.sym_block[sym]], cnt_sym);
    }
    if (lead_pid) r_pid = __ldg(&range[hash_cap + BOSOT_MAX_SYMBOLS + pid]);
    const unsigned sym_leads = __ballot_sync(full, lead_sym);

    // operator nodes: SUBTREES classes through the hash table
    const bool is_int = lane < ni;
    const unsigned long long id = is_int ? cur.id : (unsigned long long)lane;
    const unsigned m_id = __match_any_sync(full, id);
    uint2 r_int = make_uint2(0u, 0u);
    if (is_int && (__ffs(m_id) - 1 == lane)) {
      uint32_t s = hash_slot(id, hash_cap);
      while (true) {
        const unsigned long long k = __ldg(&keys[s]);
        if (k == id) { r_int = __ldg(&range[s]); break; }
        if (k == 0ULL) break;
        s = (s + 1) & (uint32_t)(hash_cap - 1);
      }
    }
    __syncwarp();

    // candidate DIM_WISE vector from the symbol counts (optimal_transport_mappings.py:88-108)
#pragma unroll 1
    for (int s = lane; s < g.n_symbols; s += 32) {
      const int b = g.sym_block[s], c = g.sym_col[s];
      if (b >= 0 && c >= 0) {
        const uint32_t tot = btot[b];
        avec[c] = tot ? (double)scnt[s] / (double)tot : 0.0;
      }
    }
    if (lane < g.n_blocks) avec[g.block_null_col[lane]] = btot[lane] == 0u ? 1.0 : 0.0;

    // Walk the postings of all matched features as ONE flattened list: an exclusive scan of the list
    // lengths over the lanes gives every feature its segment; in each round lane j takes posting
    // q = 32 * round + j and finds its feature by a 5-step binary search over the segment starts
    // (register shuffles).  Full rounds instead of one mostly empty round per feature; postings of
    // different features may hit the same evaluated tree, hence the shared-memory atomic.
    auto walk = [&](uint2 r, uint32_t cnt_a, uint32_t Ta, int shift) {
      uint32_t inc = r.y;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(full, inc, o);
        if (lane >= o) inc += v;
      }
      const uint32_t total = __shfl_sync(full, inc, 31);
      const uint32_t start = inc - r.y;
#pragma unroll 1
      for (uint32_t q0 = 0; q0 < total; q0 += 32) {
        const uint32_t q = q0 + lane;
        int k = 0;  // largest lane whose segment starts at or before q
#pragma unroll
        for (int step = 16; step; step >>= 1) {
          const uint32_t sk = __shfl_sync(full, start, k + step);
          if (sk <= q) k += step;
        }
        const uint32_t beg = __shfl_sync(full, r.x, k);
        const uint32_t st = __shfl_sync(full, start, k);
        const uint32_t ca = __shfl_sync(full, cnt_a, k);
        if (q < total) {
          const uint32_t pq = __ldg(&post[beg + (q - st)]);
          const uint32_t e = pq & 0xffffu, cb = (pq >> 16) & 0xffu, Tb = pq >> 24;
          atomicAdd(&acc[e], min(ca * Tb, cb * Ta) << shift);
        }
      }
    };
    walk(r_int, (uint32_t)__popc(m_id), (uint32_t)nn, 0);
    walk(r_pid, (uint32_t)__popc(m_pid), (uint32_t)nl, 16);
    __syncwarp();

    // ---- phase C: register tiles of R evaluated trees per lane (tree e = e0 + lane + 32 r)
    const double rTa_sub = c_recip[nn], rTa_path = c_recip[nl];
    for (int e0 = 0; e0 < E; e0 += 32 * R) {
      const int eb = e0 + lane;  // e_pad is a whole number of tiles: eb + 32 r < e_pad, padded arrays stay in bounds
      uint32_t msub[R];
      double ddim[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        msub[r] = 0u;
        ddim[r] = 0.0;
      }
      // SUBTREES, leaf classes: sum over the candidate's distinct symbols of min(ca * Tb, cb * Ta)
      {
        const uint8_t* nnp = e_nn + eb;
        for (unsigned todo = sym_leads; todo;) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const uint32_t s = __shfl_sync(full, sym, src);
          const uint32_t ca = __shfl_sync(full, cnt_sym, src);
          const uint8_t* col = symT + s * (uint32_t)e_pad + eb;
#pragma unroll
          for (int r = 0; r < R; ++r) msub[r] += min(ca * (uint32_t)nnp[32 * r], (uint32_t)col[32 * r] * (uint32_t)nn);
        }
      }
      // DIM_WISE: dense L1 over the d * (kinds + 1) columns, column-outer so the candidate value is read once
      {
        const double* col = dimT + eb;
#pragma unroll 2
        for (int c = 0; c < F; ++c, col += e_pad) {
          const double a = avec[c];
#pragma unroll
          for (int r = 0; r < R; ++r) ddim[r] += fabs(a - col[32 * r]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int e = eb + 32 * r;
        if (e < E) {
          const uint32_t a = acc[e];
          const uint32_t m_sub = msub[r] + (a & 0xffffu), m_path = a >> 16;
          const uint32_t tb_sub = e_nn[e], tb_path = e_nl[e];
          // 2 - 2*M/(Ta*Tb); exactly 0 when the histograms coincide; (1/Ta)*(1/Tb) first keeps D(a,b) == D(b,a) bitwise
          const double d_sub = (m_sub == (uint32_t)nn * tb_sub)
                                   ? 0.0 : fma(-2.0, (double)m_sub * (rTa_sub * e_rsub[e]), 2.0);
          const double d_path = (m_path == (uint32_t)nl * tb_path)
                                    ? 0.0 : fma(-2.0, (double)m_path * (rTa_path * e_rpath[e]), 2.0);
          // sum_f w_f * D_f in feature_type_list order (kernel_kernel_grammar_tree.py:159-161)
          const double d = __dadd_rn(__dadd_rn(__dmul_rn(p.w_dim, ddim[r]), __dmul_rn(p.w_path, d_path)),
                                     __dmul_rn(p.w_sub, d_sub));
          if (kExtra) {
            if (p.Df) { p.Df[row + e] = ddim[r]; p.Df[plane + row + e] = d_path; p.Df[2 * plane + row + e] = d_sub; }
            if (p.D) p.D[row + e] = d;
          }
          if (p.K) p.K[row + e] = p.variance * exp_nonpos(d * p.neg_inv_lsq);
        } else if (p.K && e < p.ld) {
          p.K[row + e] = 0.0;  // padding columns of the Gram row stay finite (the posterior kernel may read them)
        }
      }
    }
  }
}

static bool g_recip_ready[64] = {false};  // per device

static int ensure_recip_table() {
  int dev = 0;
  if (int rc = cuda_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (dev < 64 && g_recip_ready[dev]) return BOSOT_OK;
  double h[64];
  h[0] = 0.0;
  for (int t = 1; t < 64; ++t) h[t] = 1.0 / (double)t;
  if (int rc = cuda_check(cudaMemcpyToSymbol(c_recip, h, sizeof(h)), "cudaMemcpyToSymbol(c_recip)")) return rc;
  if (dev < 64) g_recip_ready[dev] = true;
  return BOSOT_OK;
}

size_t pairs_scratch_bytes(int64_t n_cand) { return align16((size_t)n_cand * kScanBytes); }

int launch_pairs(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar, const void* d_eval_blob,
                 int n_eval, const double weights[3], double variance, double lengthscale, double* d_K, double* d_D,
                 double* d_Df, int64_t ld, int32_t* d_status, void* d_scratch, size_t scratch_bytes, cudaStream_t stream) {
  if (!grammar || !d_eval_blob || !weights || n_cand < 0 || (n_cand > 0 && (!d_trees || !d_scratch)))
    return set_error(BOSOT_EINVAL, "bosot_sot_pairs: null argument");
  if (n_eval < 1 || n_eval > BOSOT_MAX_EVAL) return set_error(BOSOT_EINVAL, "bosot_sot_pairs: n_eval out of range");
  if (ld < n_eval) return set_error(BOSOT_EINVAL, "bosot_sot_pairs: ld < n_eval");
  if (scratch_bytes < pairs_scratch_bytes(n_cand) || (reinterpret_cast<uintptr_t>(d_scratch) & 15))
    return set_error(BOSOT_EINVAL, "bosot_sot_pairs: scratch too small or not 16-byte aligned (bosot_sot_pairs_scratch_bytes)");
  if (int rc = check_grammar(grammar)) return rc;
  if (n_cand == 0) return BOSOT_OK;
  if (int rc = ensure_recip_table()) return rc;

  PairParams p;
  p.trees = d_trees;
  p.n_cand = n_cand;
  p.blob = static_cast<const uint8_t*>(d_eval_blob);
  p.L = eval_layout(n_eval);
  p.w_dim = weights[0];
  p.w_path = weights[1];
  p.w_sub = weights[2];
  p.variance = variance;
  p.neg_inv_lsq = -1.0 / (lengthscale * lengthscale);
  p.K = d_K;
  p.D = d_D;
  p.Df = d_Df;
  p.ld = ld;
  p.recs = static_cast<const uint8_t*>(d_scratch);

  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  if (int rc = cuda_check(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) return rc;

  // launch 1: scan
  static bool scan_attr[64] = {false};
  if (!(dev < 64 && scan_attr[dev])) {
    if (int rc = cuda_check(cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kScanThreads * kRecBytes),
                            "cudaFuncSetAttribute(scan_kernel)")) return rc;
    if (dev < 64) scan_attr[dev] = true;
  }
  const int64_t scan_blocks = (n_cand + kScanThreads - 1) / kScanThreads;
  if (scan_blocks > 0x7fffffffLL) return set_error(BOSOT_EINVAL, "bosot_sot_pairs: too many candidates");
  scan_kernel<<<(unsigned)scan_blocks, kScanThreads, kScanThreads * kRecBytes, stream>>>(
      d_trees, n_cand, grammar->n_symbols, static_cast<uint8_t*>(d_scratch), d_status);
  if (int rc = after_launch("scan_kernel")) return rc;

  // launch 2: pairs.  One persistent CTA per SM; small pools shrink the CTA so that every SM gets warps.
  int warps = kMaxPairWarps;
  while (warps > 1 && (n_cand + warps - 1) / warps < sms) warps >>= 1;
  const StageLayout S = stage_layout(p.L.e_pad, grammar->n_dim_cols, grammar->n_symbols);
  const size_t smem_cap = 220 * 1024;
  size_t smem = S.total + (size_t)warps * pair_warp_smem(p.L.e_pad);
  while (smem > smem_cap && warps > 1) {
    warps >>= 1;
    smem = S.total + (size_t)warps * pair_warp_smem(p.L.e_pad);
  }
  if (smem > smem_cap) return set_error(BOSOT_EINVAL, "bosot_sot_pairs: n_eval too large for shared memory");

  const bool extra = d_D != nullptr || d_Df != nullptr;
  const int rounds = p.L.e_pad / 32;
  const int R = rounds < kMaxR ? rounds : kMaxR;
  using KernelFn = void (*)(PairParams, bosot_grammar);
  static const KernelFn table[kMaxR][2] = {
      {sot_pair_kernel<1, false>, sot_pair_kernel<1, true>}, {sot_pair_kernel<2, false>, sot_pair_kernel<2, true>},
      {sot_pair_kernel<3, false>, sot_pair_kernel<3, true>}, {sot_pair_kernel<4, false>, sot_pair_kernel<4, true>},
      {sot_pair_kernel<5, false>, sot_pair_kernel<5, true>}, {sot_pair_kernel<6, false>, sot_pair_kernel<6, true>},
      {sot_pair_kernel<7, false>, sot_pair_kernel<7, true>}, {sot_pair_kernel<8, false>, sot_pair_kernel<8, true>}};
  KernelFn kernel = table[R - 1][extra ? 1 : 0];
  static bool attr_set[64] = {false};
  if (!(dev < 64 && attr_set[dev])) {
    for (int r = 0; r < kMaxR; ++r)
      for (int x = 0; x < 2; ++x)
        if (int rc = cuda_check(cudaFuncSetAttribute(table[r][x], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap),
                                "cudaFuncSetAttribute(sot_pair_kernel)")) return rc;
    if (dev < 64) attr_set[dev] = true;
  }
  int per_sm = 1;
  if (int rc = cuda_check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem),
                          "cudaOccupancyMaxActiveBlocksPerMultiprocessor")) return rc;
  if (per_sm < 1) per_sm = 1;
  if (per_sm * warps > kMaxPairWarps) per_sm = kMaxPairWarps / warps;
  const int64_t want = (n_cand + warps - 1) / warps;
  const int64_t wave = (int64_t)sms * per_sm;
  const int64_t grid = want < wave ? want : wave;
  kernel<<<(unsigned)grid, warps * 32, smem, stream>>>(p, *grammar);
  return after_launch("sot_pair_kernel");
}

}  // namespace bosot

extern "C" size_t bosot_sot_pairs_scratch_bytes(int64_t n_cand) { return n_cand < 0 ? 0 : bosot::pairs_scratch_bytes(n_cand); }

extern "C" int bosot_sot_pairs(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar,
                               const void* d_eval_blob, int n_eval, const double weights[3],
                               double variance, double lengthscale,
                               double* d_K, double* d_D, double* d_Df, int64_t ld,
                               int32_t* d_status, void* d_scratch, size_t scratch_bytes, void* stream) {
  return bosot::launch_pairs(d_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance, lengthscale, d_K, d_D,
                             d_Df, ld, d_status, d_scratch, scratch_bytes, (cudaStream_t)stream);
}
