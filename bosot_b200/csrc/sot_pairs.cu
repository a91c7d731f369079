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
//                                      and their (short) posting lists are walked as one flattened list
//                                      (segment of a posting found by a warp OR-reduction, no search)
//                                      into shared-memory accumulators; the candidate's DIM_WISE vector
//                                      and its L1 distance to every distinct block vector ("state") of
//                                      the evaluated set;
//                                      phase C  every lane owns R evaluated trees in registers:
//                                      leaf-symbol classes (dense: most symbols occur in most trees)
//                                      against the packed transposed count table (two trees per 32-bit
//                                      lane), DIM_WISE as one table lookup per (pair, block) or the dense
//                                      column loop, then combine, exp and store in groups of four pairs.
//                                      The next tree's record is prefetched into registers while the
//                                      current one is processed.
#include <atomic>
#include <type_traits>
#include "sot_common.cuh"
#include "launch.cuh"
#include "fast_math.cuh"
#include "exp_table.cuh"

namespace bosot {

int experiment_flag();

// One persistent CTA of 32 warps per SM (64 registers per thread).  Round 1 measured 20 / 24 / 28 / 32 warps = 2.68 / 2.52 /
// 2.43 / 2.49 ms at 1M x 200 and kept 28; with the leaner round-2 kernel 24 / 28 / 32 warps = 2.119 / 2.023 / 1.976 ms, and
// the kernels with eight evaluated trees per lane (E > 224, a few spilled bytes either way) 0.495 / 0.480 ms at 200k x 256
// for 28 / 32 warps (profiles/r3l_ab_warps.log).
__host__ __device__ constexpr int pair_max_warps(int /*R*/) { return 32; }
constexpr int kScanThreads = 128;
// Block-vector cache of the DIM_WISE distances (state mode, blocks of at most two kinds): the vector of a candidate in
// one block of input dimensions is fixed by the leaf counts (c1, c2) of the block's kinds, so L1(block vector, state)
// is tabulated once per CTA for all c1, c2 <= kBvMaxCount and a candidate only forms its index per block
constexpr int kBvMaxCount = 5, kBvRadix = kBvMaxCount + 1, kBvEntries = kBvRadix * kBvRadix;
constexpr int kBvStates = 20;  // states per cache row (the 200 x d = 5 benchmark set has 19 at most in a block)
constexpr int kMaxR = 8;            // evaluated trees per lane per register tile (E <= 256 in one pass, else tiles of 256)
constexpr int64_t kFusedScanMax = 16384;  // pools up to this size take the fused-scan path (no scan launch); measured
                                          // 10k x 50: 48 against 56 us per step, 100 x 50: 21 against 34 us, 30k x 80: 110 against 101 us (profiles/r2i_ab.log)

// Small tables, initialised STATICALLY (constant-folded by the host compiler, so correctly rounded like Python's
// float(a) / float(b)): nothing is uploaded at run time, hence nothing to order against the caller's stream.
struct RecipTable {  // v[T] = 1 / T for T = 1..63 ([0] unused)
  double v[64];
  constexpr RecipTable() : v{} { for (int t = 1; t < 64; ++t) v[t] = 1.0 / (double)t; }
};
struct FracTable {   // v[c * 33 + t] = (double)c / (double)t (optimal_transport_mappings.py:104-106), t = 0 -> 0
  double v[33 * 33];
  constexpr FracTable() : v{} {
    for (int c = 0; c < 33; ++c)
      for (int t = 1; t < 33; ++t) v[c * 33 + t] = (double)c / (double)t;
  }
};
__constant__ RecipTable c_recip = RecipTable();
__device__ const FracTable g_frac = FracTable();

struct PairParams {
  const uint8_t* trees;
  int64_t n_cand;
  const uint8_t* blob;
  EvalLayout L;
  double w_dim, w_path, w_sub;
  double variance, neg_inv_lsq;  // -1 / lengthscale^2
  double a_dim, a_path, a_sub;   // w_f * (-1 / lengthscale^2): the Gram argument is formed as one fused sum (kFast)
  double* K;
  double* D;
  double* Df;
  int64_t ld;
  const uint8_t* recs;  // compact records written by scan_kernel: [n_cand][kScanBytes]
  unsigned long long* next;  // work counter (zeroed by the scan launch): the warps claim candidates one at a time
  int32_t* bad;              // optional: number of malformed candidate records met (their rows are poisoned)
  int flags;                 // experiment switches (0 = product path)
  int32_t* status;           // fused-scan path only: per-tree defect codes (may be null)
  int n_symbols;
  int bv_states;             // rows of the block-vector cache hold this many states (0 = no cache in this launch)
};

// per-warp shared memory: candidate dim-wise vector, block-state distance table, walk segments, accumulators, symbol counts
struct WarpLayout {
  size_t off_tab, off_slot, off_acc, off_scnt, off_scan, total;
};
// scratch of the warp-cooperative scan (fused-scan path): idpos u64[64] | lvl_lo u32[34] | lvl_hi u32[34] | oprec u32[32] |
// endcnt u32[64] | leafrec u16[32]; the cold path (malformed record) reuses it as one scan_tree() working record
constexpr int kWsIdpos = 0, kWsLvlLo = 512, kWsLvlHi = 648, kWsOprec = 784, kWsEndcnt = 912, kWsLeafrec = 1168, kWarpScanBytes = 1232;
static_assert(kWarpScanBytes >= kRecBytes, "the cold path runs scan_tree() in the same scratch");
__host__ __device__ inline WarpLayout warp_layout(int e_pad, int n_dim_cols, int n_blocks, bool fused_scan = false) {
  WarpLayout W;
  size_t o = sizeof(double) * (size_t)((n_dim_cols + 1) & ~1);           // avec [n_pcols]
  W.off_tab = o;   o += sizeof(double) * (size_t)kMaxStates * n_blocks;  // tab [n_blocks][kMaxStates]
  W.off_slot = o;  o += sizeof(uint2) * 32;                              // segments of the flattened postings walk
  W.off_acc = o;   o += sizeof(uint32_t) * (size_t)e_pad;                // M_sub(ops) | M_path << 16 per e
  W.off_scnt = o;  o += sizeof(uint32_t) * (BOSOT_MAX_SYMBOLS + BOSOT_MAX_BLOCKS);
  o = align16(o);
  W.off_scan = o;  o += fused_scan ? kWarpScanBytes : 0;
  W.total = align16(o);
  return W;
}

// CTA-wide staged copy of the evaluated set's dense tables (section sizes are multiples of 16 bytes).
// The DIM_WISE section holds either the block-state tables (stateP + sid8T) or, when a block has more than
// kMaxStates distinct vectors, the dense transposed matrix; it is sized for the larger of the two.
struct StageLayout {
  size_t off_meta, off_dim, off_sid, off_flat, off_rsub, off_rpath, off_symT2, off_nn2, off_nn, off_nl, off_exp, off_bvw, off_bv, total;
  uint32_t b_meta, b_state, b_sid, b_flat, b_dimT, b_r, b_symT2, b_nn2, b_n, b_exp;
};
__host__ __device__ inline StageLayout stage_layout(int e_pad, int n_dim_cols, int n_symbols, int n_blocks, bool fast_exp, int bv_states = 0) {
  StageLayout S;
  S.b_meta = (uint32_t)sizeof(DimMeta);
  S.b_state = (uint32_t)(sizeof(double) * (size_t)n_dim_cols * kMaxStates);
  S.b_sid = (uint32_t)((size_t)n_blocks * e_pad);
  S.b_flat = (uint32_t)(sizeof(uint32_t) * (size_t)n_blocks * kMaxStates);
  S.b_dimT = (uint32_t)(sizeof(double) * (size_t)n_dim_cols * e_pad);
  S.b_r = (uint32_t)(sizeof(double) * (size_t)e_pad);
  S.b_symT2 = (uint32_t)(sizeof(uint32_t) * (size_t)n_symbols * (e_pad / 2));
  S.b_nn2 = (uint32_t)(sizeof(uint32_t) * (size_t)(e_pad / 2));
  S.b_n = (uint32_t)e_pad;
  size_t o = 16;  // mbarrier
  S.off_meta = o;  o += S.b_meta;
  S.off_dim = o;   S.off_sid = o + S.b_state;   S.off_flat = S.off_sid + S.b_sid;
  o += (S.b_state + S.b_sid + S.b_flat > S.b_dimT) ? (S.b_state + S.b_sid + S.b_flat) : S.b_dimT;
  S.off_rsub = o;  o += S.b_r;
  S.off_rpath = o; o += S.b_r;
  S.off_symT2 = o; o += S.b_symT2;
  S.off_nn2 = o;   o += S.b_nn2;
  S.off_nn = o;    o += S.b_n;
  S.off_nl = o;    o += S.b_n;
  o = align16(o);
  S.b_exp = fast_exp ? (uint32_t)(sizeof(double) * kExpTableSize) : 0u;
  S.off_exp = o;   o += S.b_exp;
  S.off_bvw = o;   o += bv_states ? BOSOT_MAX_SYMBOLS : 0;                                          // index weight per symbol
  S.off_bv = o;    o += sizeof(double) * (size_t)n_blocks * kBvEntries * bv_states;                 // [block][entry][state]
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
            int32_t* __restrict__ status, unsigned long long* __restrict__ next, int32_t* __restrict__ bad_zero) {
  extern __shared__ __align__(16) uint8_t smem[];
  pdl_launch_dependents();  // the pair launch may stage its tables while the scan runs (it waits before reading the records)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *next = 0ULL;  // work counter of the pair launch that follows on the stream
    if (bad_zero) *bad_zero = 0;  // ... and its count of malformed records
  }
  const int64_t base = (int64_t)blockIdx.x * kScanThreads;
  const int in_block = (int)min((int64_t)kScanThreads, n_trees - base);
  for (int w = threadIdx.x; w < in_block * 4; w += kScanThreads) {  // coalesced: 16 B per thread
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(trees + base * kTreeBytes) + w);
    uint2* dst = reinterpret_cast<uint2*>(smem + (w >> 2) * kRecBytes + kRecTok + (w & 3) * 16);  // 8-byte aligned
    dst[0] = make_uint2(v.x, v.y);
    dst[1] = make_uint2(v.z, v.w);
  }
  __syncthreads();
  // Threads of a warp run their trees in lock step (the passes of scan_tree are branch-free per token), so a warp
  // costs as much as its LONGEST tree: hand the trees out sorted by token count (counting sort over the CTA's
  // records) and every warp gets trees of similar length.
  __shared__ int bin_pos[64];
  __shared__ uint8_t order[kScanThreads];
  if (threadIdx.x < 64) bin_pos[threadIdx.x] = 0;
  __syncthreads();
  int my_bin = 0;
  if ((int)threadIdx.x < in_block) {
    const int n = smem[threadIdx.x * kRecBytes + kRecTok];
    my_bin = n < 64 ? n : 0;  // malformed lengths fail fast in scan_tree
    atomicAdd(&bin_pos[my_bin], 1);
  }
  __syncthreads();
  if (threadIdx.x < 32) {  // exclusive scan of the 64 bin counts: two bins per lane
    const int a = bin_pos[2 * threadIdx.x], b = bin_pos[2 * threadIdx.x + 1];
    int inc = a + b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int)threadIdx.x >= o) inc += v;
    }
    bin_pos[2 * threadIdx.x] = inc - a - b;
    bin_pos[2 * threadIdx.x + 1] = inc - b;
  }
  __syncthreads();
  if ((int)threadIdx.x < in_block) order[atomicAdd(&bin_pos[my_bin], 1)] = (uint8_t)threadIdx.x;
  __syncthreads();
  if ((int)threadIdx.x < in_block) {
    const int mine = order[threadIdx.x];
    const int st = scan_tree(smem + mine * kRecBytes, n_symbols);
    if (status) status[base + mine] = st;
  }
  __syncthreads();
  // coalesced copy-out of the compact part (320 B = 40 x 8 B per record)
  uint2* out = reinterpret_cast<uint2*>(recs_out + base * kScanBytes);
  for (int w = threadIdx.x; w < in_block * (kScanBytes / 8); w += kScanThreads) {
    const int t = w / (kScanBytes / 8), q = w % (kScanBytes / 8);
    out[w] = *reinterpret_cast<const uint2*>(smem + t * kRecBytes + q * 8);
  }
}

// ---- warp-cooperative scan of ONE tree (fused-scan path of small pools) ---------------------------------------
// scan_kernel is a chain of ~2000 dependent shared-memory operations per warp whatever the pool size (17 us at 10k
// candidates, a third of the whole step); a pool that small cannot fill the machine anyway, so here the 32 lanes of
// the warp that owns a candidate split its token positions (lane l: positions l and l + 32) and everything except the
// class ids is closed form on 64-bit position masks.  With w = +1 for an operator and -1 for a leaf and S[p] the
// inclusive prefix sum of w over the pre-order tokens (S[-1] = 0):
//   * the record is a tree iff S[p] >= 0 for p < n - 1 and S[n - 1] = -1;
//   * the subtree that starts at x ends at the first q >= x with S[q] = S[x - 1] - 1, so the right child of the operator
//     at p is 1 + (first q > p with S[q] = S[p] - 1);
//   * the parent of p is p - 1 if that is an operator, else the nearest operator b < p with S[b] = S[p - 1] + 1;
//   * an operator "changes" when it is the root or differs from its parent's operator; the reduced path of a leaf has
//     as many runs as it has changing ancestors = (changing operators before it) - (changing operators whose subtree
//     ended before it), and its nearest operator is its parent's  (reduce_operator_list, kernel_grammar.py:649-658).
// "First / nearest position with S = v" is one bit scan on the mask of positions of level v (match_any groups).
// Class ids go bottom-up: operator number r (pre-order rank) lives on lane r and fires once both children are ready.
// Any record the closed form rejects is handed to scan_tree() on lane 0, which yields the exact defect code.
struct Slice { unsigned long long id; uint32_t meta, sym, code; };
struct RawTree { uint32_t n, tk0, tk1; };

__device__ __forceinline__ RawTree load_raw(const uint8_t* __restrict__ rec, int lane) {
  RawTree r;
  r.n = __ldg(rec);
  r.tk0 = __ldg(rec + 1 + lane);
  r.tk1 = lane < 31 ? (uint32_t)__ldg(rec + 33 + lane) : (uint32_t)BOSOT_PAD;
  return r;
}

__device__ __forceinline__ unsigned long long mask_below(int p) { return p >= 64 ? ~0ULL : (1ULL << p) - 1ULL; }  // positions < p

__device__ __forceinline__ Slice warp_scan_tree(const RawTree raw, const uint8_t* __restrict__ rec_global, int n_symbols,
                                                uint8_t* scratch, int lane) {
  const unsigned full = 0xffffffffu;
  const int n = (int)raw.n;
  const int p0 = lane, p1 = lane + 32;
  const bool len_ok = n >= 1 && n <= kMaxNodes && (n & 1);
  const bool v0 = len_ok && p0 < n, v1 = len_ok && p1 < n;
  const uint32_t tk0 = raw.tk0, tk1 = raw.tk1;
  const bool op0 = v0 && tk0 >= 0x80u, op1 = v1 && tk1 >= 0x80u;
  const bool bt0 = v0 && (tk0 >= 0x80u ? tk0 > (uint32_t)BOSOT_OP_MUL : (int)tk0 >= n_symbols);
  const bool bt1 = v1 && (tk1 >= 0x80u ? tk1 > (uint32_t)BOSOT_OP_MUL : (int)tk1 >= n_symbols);
  const unsigned long long opmask = (unsigned long long)__ballot_sync(full, op0) | ((unsigned long long)__ballot_sync(full, op1) << 32);
  const unsigned long long mulmask = (unsigned long long)__ballot_sync(full, op0 && tk0 == (uint32_t)BOSOT_OP_MUL) |
                                     ((unsigned long long)__ballot_sync(full, op1 && tk1 == (uint32_t)BOSOT_OP_MUL) << 32);
  const int s0 = 2 * __popcll(opmask & mask_below(p0 + 1)) - (p0 + 1);
  const int s1 = 2 * __popcll(opmask & mask_below(p1 + 1)) - (p1 + 1);
  const bool sb0 = v0 && (p0 == n - 1 ? s0 != -1 : s0 < 0);
  const bool sb1 = v1 && (p1 == n - 1 ? s1 != -1 : s1 < 0);
  Slice out;
  if (!len_ok || __ballot_sync(full, bt0 | bt1 | sb0 | sb1) != 0u) {
    // cold path: the sequential scan decides (and names the defect); the slice is read back from its working record
    __syncwarp();
    if (lane < 16) reinterpret_cast<uint32_t*>(scratch + kRecTok)[lane] = __ldg(reinterpret_cast<const uint32_t*>(rec_global) + lane);
    __syncwarp();
    if (lane == 0) scan_tree(scratch, n_symbols);
    __syncwarp();
    out.meta = *reinterpret_cast<const uint32_t*>(scratch + kRecMeta);
    out.id = lane < kMaxInternal ? reinterpret_cast<const unsigned long long*>(scratch + kRecIid)[lane] : 0ULL;
    out.sym = scratch[kRecLeafSym + lane];
    out.code = scratch[kRecLeafCode + lane];
    __syncwarp();
    return out;
  }
  unsigned long long* idpos = reinterpret_cast<unsigned long long*>(scratch + kWsIdpos);
  uint32_t* lvl_lo = reinterpret_cast<uint32_t*>(scratch + kWsLvlLo);
  uint32_t* lvl_hi = reinterpret_cast<uint32_t*>(scratch + kWsLvlHi);
  uint32_t* oprec = reinterpret_cast<uint32_t*>(scratch + kWsOprec);
  uint32_t* endcnt = reinterpret_cast<uint32_t*>(scratch + kWsEndcnt);
  uint16_t* leafrec = reinterpret_cast<uint16_t*>(scratch + kWsLeafrec);
  lvl_lo[lane] = 0u; lvl_hi[lane] = 0u;
  if (lane < 2) { lvl_lo[32 + lane] = 0u; lvl_hi[32 + lane] = 0u; }
  endcnt[lane] = 0u; endcnt[lane + 32] = 0u;
  __syncwarp();
  // level masks: positions with the same S (index S + 1 in 0 .. 32)
  const unsigned m0 = __match_any_sync(full, v0 ? s0 + 1 : 64 + lane);
  const unsigned m1 = __match_any_sync(full, v1 ? s1 + 1 : 64 + lane);
  if (v0 && __ffs(m0) - 1 == lane) lvl_lo[s0 + 1] = m0;
  if (v1 && __ffs(m1) - 1 == lane) lvl_hi[s1 + 1] = m1;
  __syncwarp();
  auto level = [&](int sv) { return (unsigned long long)lvl_lo[sv + 1] | ((unsigned long long)lvl_hi[sv + 1] << 32); };
  const unsigned long long validmask = mask_below(n);
  // per position: end of the own subtree, right child (operators), parent, "changing" flag
  int par[2], eself[2];
  bool chg[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = h ? p1 : p0, sv = h ? s1 : s0;
    const bool valid = h ? v1 : v0, isop = h ? op1 : op0;
    const uint32_t tk = h ? tk1 : tk0;
    par[h] = -1; eself[h] = p; chg[h] = false;
    if (valid) {
      const unsigned long long after = ~mask_below(p + 1);
      if (p > 0) {
        if ((opmask >> (p - 1)) & 1ULL) par[h] = p - 1;
        else {
          const int sprev = isop ? sv - 1 : sv + 1;
          par[h] = 63 - __clzll((long long)(level(sprev + 1) & opmask & mask_below(p)));
        }
      }
      if (isop) {
        const int eol = __ffsll((long long)(level(sv - 1) & after)) - 1;   // end of the left subtree
        eself[h] = __ffsll((long long)(level(sv - 2) & after)) - 1;
        const bool mul = (mulmask >> p) & 1ULL;
        chg[h] = p == 0 || (((mulmask >> par[h]) & 1ULL) != (mul ? 1ULL : 0ULL));
        oprec[__popcll(opmask & mask_below(p))] = (uint32_t)p | ((uint32_t)(eol + 1) << 8) | ((mul ? 2u : 1u) << 16);
      } else {
        idpos[p] = leaf_class(tk);
      }
    }
  }
  const unsigned long long chgmask = (unsigned long long)__ballot_sync(full, chg[0]) | ((unsigned long long)__ballot_sync(full, chg[1]) << 32);
  if (chg[0]) atomicAdd(&endcnt[eself[0]], 1u);
  if (chg[1]) atomicAdd(&endcnt[eself[1]], 1u);
  __syncwarp();
  {  // inclusive prefix sums of endcnt over the 64 positions
    uint32_t a = endcnt[lane], b = endcnt[lane + 32];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t xa = __shfl_up_sync(full, a, o), xb = __shfl_up_sync(full, b, o);
      if (lane >= o) { a += xa; b += xb; }
    }
    b += __shfl_sync(full, a, 31);
    __syncwarp();
    endcnt[lane] = a; endcnt[lane + 32] = b;
  }
  __syncwarp();
  const unsigned long long leafmask = ~opmask & validmask;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = h ? p1 : p0;
    const bool valid = h ? v1 : v0, isop = h ? op1 : op0;
    if (valid && !isop) {
      const int runs = __popcll(chgmask & mask_below(p)) - (p ? (int)endcnt[p - 1] : 0);
      const uint32_t code = runs ? 1u + 2u * (uint32_t)(runs - 1) + (uint32_t)((mulmask >> par[h]) & 1ULL) : 0u;
      leafrec[__popcll(leafmask & mask_below(p))] = (uint16_t)((h ? tk1 : tk0) | (code << 8));
    }
  }
  __syncwarp();
  // class ids bottom-up: operator number `lane` fires when both children are ready
  const int n_int = __popcll(opmask), n_leaf = n - n_int;
  const bool mine = lane < n_int;
  const uint32_t rec = mine ? oprec[lane] : 0u;
  const int pos = rec & 0xff, right = (rec >> 8) & 0xff;
  unsigned long long ready = leafmask, my_id = 0ULL;
  bool done = !mine;
  while (true) {
    const bool go = !done && ((ready >> (pos + 1)) & 1ULL) && ((ready >> right) & 1ULL);
    if (go) {
      my_id = class_combine(rec >> 16, idpos[pos + 1], idpos[right]);
      idpos[pos] = my_id;
      done = true;
    }
    const uint32_t lo = __reduce_or_sync(full, (go && pos < 32) ? (1u << pos) : 0u);
    const uint32_t hi = __reduce_or_sync(full, (go && pos >= 32) ? (1u << (pos - 32)) : 0u);
    if ((lo | hi) == 0u) break;
    ready |= (unsigned long long)lo | ((unsigned long long)hi << 32);
    __syncwarp();
  }
  const uint32_t lr = leafrec[lane];
  out.id = my_id;
  out.meta = (uint32_t)n | ((uint32_t)n_leaf << 8) | ((uint32_t)n_int << 16);
  out.sym = lr & 0xffu;
  out.code = lr >> 8;
  __syncwarp();
  return out;
}

// zeroes the work counter (and the count of malformed records) of a fused-scan pair launch: what scan_kernel does on
// the two-launch path
__global__ void pair_init_kernel(unsigned long long* next, int32_t* bad_zero) {
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    *next = 0ULL;
    if (bad_zero) *bad_zero = 0;
  }
}

// ---- launch 2: warp-per-tree pairs ----------------------------------------------------------
// exp_nonpos (fast_math.cuh) on G arguments at once: the same operations in the same order (bit-identical
// results), but every coefficient is fetched once per group and the G Horner chains interleave.
template <int G>
__device__ __forceinline__ void exp_nonpos_group(double (&x)[G]) {
  double r[G], q[G];
  int n[G];
  {
    const double lo = c_exp[15], l2e = c_exp[0], magic = c_exp[1], ln2h = c_exp[2], ln2l = c_exp[3], c13 = c_exp[4];
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const double v = x[i] < lo ? lo : x[i];  // NaN stays NaN
      const double t = fma(v, l2e, magic);
      n[i] = __double2loint(t);
      const double nf = t - magic;
      r[i] = fma(nf, ln2l, fma(nf, ln2h, v));
      q[i] = c13;
    }
  }
#pragma unroll
  for (int k = 5; k <= 14; ++k) {
    const double ck = c_exp[k];
#pragma unroll
    for (int i = 0; i < G; ++i) q[i] = fma(q[i], r[i], ck);
  }
#pragma unroll
  for (int i = 0; i < G; ++i) {
    double v = fma(q[i], r[i], 0.5);
    v = fma(v, r[i], 1.0);
    v = fma(v, r[i], 1.0);
    const int n1 = n[i] >> 1, n2 = n[i] - n1;
    x[i] = (v * __hiloint2double((1023 + n1) << 20, 0)) * __hiloint2double((1023 + n2) << 20, 0);
  }
}

// Table-driven exp for the launches whose hyper-parameters bound the Gram argument (launch_pairs: -600 <= x <= 0, all
// weights >= 0, variance in [2^-100, 2^100]): x = (1024 n + j) ln2 / 1024 + r with |r| <= ln2 / 2048, so
//   variance * exp(x) = 2^n * (variance * 2^(j/1024)) * (1 + r + r^2/2 + r^3/6 + r^4/24)      (r^5/120 < 4e-20)
// with the 1024 values variance * 2^(j/1024) in shared memory (exp_table.cuh, scaled once per CTA).  No clamp and
// no two-step scaling: the result is a normal number, so 2^n is one integer add on the exponent field.  9 fp64
// operations per value instead of 21; at most 0.93 ulp off exp(x) before the variance (checked against 50-digit
// arithmetic, gen_exp_table.py), exactly `variance` at x = 0.
__constant__ double c_fexp[6] = {
    1477.3197218702985,              // [0] 1024 / ln 2
    6755399441055744.0,              // [1] 1.5 * 2^52
    -0x1.62e42fee00000p-11,          // [2] -ln2 / 1024, high part (32 significant bits: n * high is exact for |n| < 2^21)
    -0x1.a39ef35793c76p-43,          // [3] -ln2 / 1024, low part
    1.0 / 24.0, 1.0 / 6.0};
template <int G>
__device__ __forceinline__ void exp_table_group(double (&x)[G], const uint8_t* __restrict__ tab) {
  const double c = c_fexp[0], magic = c_fexp[1], hi = c_fexp[2], lo = c_fexp[3], c24 = c_fexp[4], c6 = c_fexp[5];
  double r[G], tj[G];
  int m[G];
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const double t = fma(x[i], c, magic);
    m[i] = __double2loint(t);
    const double nf = t - magic;
    r[i] = fma(nf, lo, fma(nf, hi, x[i]));
    tj[i] = *reinterpret_cast<const double*>(tab + ((m[i] << 3) & (8 * (kExpTableSize - 1))));
  }
#pragma unroll
  for (int i = 0; i < G; ++i) {
    double q = fma(r[i], c24, c6);
    q = fma(q, r[i], 0.5);
    q = fma(q, r[i], 1.0);
    const double v = fma(tj[i] * r[i], q, tj[i]);
    x[i] = __hiloint2double(__double2hiint(v) + ((m[i] & ~(kExpTableSize - 1)) << (20 - kExpTableBits)), __double2loint(v));
  }
}

// a << n with PTX semantics (shift amounts beyond 31 give 0; in C++ they are undefined)
__device__ __forceinline__ uint32_t shl_clamp(uint32_t a, uint32_t n) {
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(n));
  return r;
}

// exact uint32 -> double for m < 2^32 without the conversion pipe: 2^52 + m is exact, subtract 2^52
__device__ __forceinline__ double u32_to_f64(uint32_t m) {
  return __hiloint2double(0x43300000, (int)m) - 4503599627370496.0;
}

// R: evaluated trees per lane per tile; kExtra: also write D and/or Df; kFused: no scan launch in front, every warp scans
// its candidate itself (warp_scan_tree) straight from the flat records; kMode: 0 = general exp_nonpos route, 1 = bounded
// Gram argument (exp_table_group), 2 = fp32 outputs (K / D / Df point to float arrays of the same shape: the integer
// minima and the distances are formed as in the fp64 modes, the Gram entry is variance * ex2.approx in fp32)
template <typename T>
__device__ __forceinline__ void store_out(double* base, size_t idx, double v) {
  reinterpret_cast<T*>(base)[idx] = (T)v;
}

template <int R, bool kExtra, bool kFused, int kMode>
__global__ void __launch_bounds__(pair_max_warps(R) * 32, 1)
sot_pair_kernel(PairParams p, bosot_grammar g) {
  constexpr bool kFast = kMode == 1, kF32 = kMode == 2;
  using OutT = typename std::conditional<kF32, float, double>::type;
  constexpr int RP = (R + 1) / 2;  // packed pairs of evaluated trees per lane
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int n_warps = blockDim.x >> 5;
  const int E = p.L.n_eval;
  const int e_pad = p.L.e_pad;
  const int half_pad = e_pad >> 1;
  const unsigned full = 0xffffffffu;
  unsigned lt_mask, le_mask;  // special registers: one S2R where the compiler would rematerialise shift + subtract at every use
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
  asm("mov.u32 %0, %%lanemask_le;" : "=r"(le_mask));

  pdl_launch_dependents();  // a posterior launch behind this one may set up its shared memory early (it waits before reading K*)
  // ---- stage the evaluated set's dense tables once per (persistent) CTA: TMA bulk copies -> mbarrier
  // (the blob is the product of an OLDER launch than the scan in front of this one, so this may run before pdl_wait())
  const StageLayout S = stage_layout(e_pad, g.n_dim_cols, g.n_symbols, g.n_blocks, kFast, p.bv_states);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    const int states = __ldg(reinterpret_cast<const int*>(p.blob + p.L.off_meta));  // DimMeta::use_states
    mbar_expect_tx(bar, S.b_meta + (states ? S.b_state + S.b_sid + S.b_flat : S.b_dimT) + 2 * S.b_r + S.b_symT2 + S.b_nn2 + 2 * S.b_n + S.b_exp);
    tma_bulk_g2s(smem + S.off_meta, p.blob + p.L.off_meta, S.b_meta, bar);
    if (states) {
      tma_bulk_g2s(smem + S.off_dim, p.blob + p.L.off_stateP, S.b_state, bar);
      tma_bulk_g2s(smem + S.off_sid, p.blob + p.L.off_sid8T, S.b_sid, bar);
      tma_bulk_g2s(smem + S.off_flat, p.blob + p.L.off_flat, S.b_flat, bar);
    } else {
      tma_bulk_g2s(smem + S.off_dim, p.blob + p.L.off_dimT, S.b_dimT, bar);
    }
    tma_bulk_g2s(smem + S.off_rsub, p.blob + p.L.off_rsub, S.b_r, bar);
    tma_bulk_g2s(smem + S.off_rpath, p.blob + p.L.off_rpath, S.b_r, bar);
    tma_bulk_g2s(smem + S.off_symT2, p.blob + p.L.off_symT2, S.b_symT2, bar);
    tma_bulk_g2s(smem + S.off_nn2, p.blob + p.L.off_nn2, S.b_nn2, bar);
    tma_bulk_g2s(smem + S.off_nn, p.blob + p.L.off_nnodes, S.b_n, bar);
    tma_bulk_g2s(smem + S.off_nl, p.blob + p.L.off_nleaves, S.b_n, bar);
    if (kFast) tma_bulk_g2s(smem + S.off_exp, g_exp2_table, S.b_exp, bar);
  }
  const DimMeta* meta = reinterpret_cast<const DimMeta*>(smem + S.off_meta);
  const double* stateP = reinterpret_cast<const double*>(smem + S.off_dim);  // [n_pcols][kMaxStates]   (state mode)
  const double* dimT = stateP;                                               // [n_pcols][e_pad]        (dense mode)
  const uint8_t* sid8T = smem + S.off_sid;
  const uint32_t* flat = reinterpret_cast<const uint32_t*>(smem + S.off_flat);  // (block, state) list
  const double* e_rsub = reinterpret_cast<const double*>(smem + S.off_rsub);
  const double* e_rpath = reinterpret_cast<const double*>(smem + S.off_rpath);
  const uint32_t* symT2 = reinterpret_cast<const uint32_t*>(smem + S.off_symT2);
  const uint32_t* e_nn2 = reinterpret_cast<const uint32_t*>(smem + S.off_nn2);
  const uint8_t* e_nn = smem + S.off_nn;
  const uint8_t* e_nl = smem + S.off_nl;
  const uint8_t* exp_tab = smem + S.off_exp;  // kFast: variance * 2^(j/1024)

  const WarpLayout W = warp_layout(e_pad, g.n_dim_cols, g.n_blocks, kFused);
  uint8_t* wbase = smem + S.total + (size_t)warp * W.total;
  double* avec = reinterpret_cast<double*>(wbase);                   // candidate DIM_WISE vector, permuted column order
  double* tab = reinterpret_cast<double*>(wbase + W.off_tab);        // tab[b][s] = L1(candidate block b, state s)
  uint2* slot = reinterpret_cast<uint2*>(wbase + W.off_slot);        // (postings begin - flattened start, candidate count)
  uint32_t* acc = reinterpret_cast<uint32_t*>(wbase + W.off_acc);    // M_sub(ops) | M_path << 16 per e
  uint32_t* scnt = reinterpret_cast<uint32_t*>(wbase + W.off_scnt);  // leaf count per symbol
  uint32_t* btot = scnt + BOSOT_MAX_SYMBOLS;                         // leaf count per block

  const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(p.blob + p.L.off_keys);
  const uint2* range = reinterpret_cast<const uint2*>(p.blob + p.L.off_range);
  const uint32_t* post = reinterpret_cast<const uint32_t*>(p.blob + p.L.off_post);
  const int hash_cap = p.L.hash_cap;
  const size_t plane = (size_t)p.n_cand * p.ld;
  const int ld_i = p.ld < (int64_t)e_pad ? (int)p.ld : e_pad;  // columns this kernel may touch

  uint8_t* wscan = wbase + W.off_scan;                               // fused-scan path: scratch of warp_scan_tree

  // one lane-slice of a compact record: class id of operator node `lane`, symbol / path code of leaf `lane`
  auto load_slice = [&](int64_t t) {
    Slice s;
    const uint8_t* r = p.recs + (size_t)t * kScanBytes;
    s.meta = __ldg(reinterpret_cast<const uint32_t*>(r + kRecMeta));
    s.id = lane < kMaxInternal ? __ldg(reinterpret_cast<const unsigned long long*>(r + kRecIid) + lane) : 0ULL;
    s.sym = __ldg(r + kRecLeafSym + lane);
    s.code = __ldg(r + kRecLeafCode + lane);
    return s;
  };

  // Work distribution.  Trees differ in cost by an order of magnitude (1 .. 63 nodes), and with a static stride the warp
  // with the unluckiest share decided when the kernel ended, so candidates are claimed one at a time from a global
  // counter, one claim ahead of the work.  The FIRST candidate of every warp is its own index in the grid -- no atomic
  // on the way to the first tree: at small pools the first claims were a stampede of thousands of warps on one address
  // (30 % of all stall samples at 10k candidates, profiles/r2f) -- and the counter hands out the positions from W on.
  // (Claiming two ahead, so that the atomic is never waited for, measured slower: 42.0 against 37.9 us at 10k x 50,
  // no difference at 1M x 200, profiles/r2h_ab.log.)
  const int64_t n_warps_grid = (int64_t)gridDim.x * n_warps;
  pdl_wait();  // the scan launch has completed: work counter zeroed, compact records visible
  int64_t t = (int64_t)blockIdx.x + (int64_t)gridDim.x * warp;  // neighbouring warps of a CTA take trees far apart
  Slice nxt;
  RawTree nraw;
  if (t < p.n_cand) {
    if (kFused) nraw = load_raw(p.trees + (size_t)t * kTreeBytes, lane);
    else nxt = load_slice(t);
  }
  mbar_wait(bar, 0);  // tables landed (the copy overlapped the first record load)
  if (kFast) {        // fold the variance into the exp table (the entry of x = 0 becomes exactly `variance`)
    double* tv = reinterpret_cast<double*>(smem + S.off_exp);
    const int nt = blockDim.x;  // 32 .. 896 threads (a small pool runs one warp per CTA: eight entries in flight per thread)
#pragma unroll 1
    for (int i = threadIdx.x; i < kExpTableSize; i += 8 * nt) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (i + u * nt < kExpTableSize) ? tv[i + u * nt] : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) if (i + u * nt < kExpTableSize) tv[i + u * nt] = v[u] * p.variance;
    }
    __syncthreads();
  }
  const bool use_states = meta->use_states != 0;
  const int n_flat = meta->n_flat;

  // ---- block-vector cache (see kBvMaxCount): usable when every block is NULL + at most two kinds, every counted symbol
  // has its column, and no block has more states than a cache row holds; all threads reach the same verdict
  const int bvS = p.bv_states;
  const uint8_t* bvw = smem + S.off_bvw;
  const uint8_t* bvc = smem + S.off_bv;
  bool bv_on = bvS > 0 && use_states;
  if (bv_on) {
    int kinds_seen = 0, kinds_cols = 0;
#pragma unroll 1
    for (int b = 0; b < g.n_blocks; ++b) {
      const int width = meta->bbeg[b + 1] - meta->bbeg[b];
      bv_on = bv_on && width >= 1 && width <= 3 && meta->n_states[b] <= bvS;
      kinds_cols += width - 1;
    }
#pragma unroll 1
    for (int sy = 0; sy < g.n_symbols; ++sy) {
      if (g.sym_block[sy] >= 0) {
        bv_on = bv_on && meta->sym_pos[sy] >= 0;
        ++kinds_seen;
      }
    }
    bv_on = bv_on && kinds_seen == kinds_cols;
    if (bv_on) {
      uint8_t* w_out = smem + S.off_bvw;
      double* c_out = reinterpret_cast<double*>(smem + S.off_bv);
      for (int sy = threadIdx.x; sy < g.n_symbols; sy += blockDim.x) {
        const int b = g.sym_block[sy];
        int wgt = 0;
        if (b >= 0) {
          const int j = meta->sym_pos[sy];
          const int rank = (j - meta->bbeg[b]) - (meta->null_pos[b] < j ? 1 : 0);
          wgt = rank == 0 ? 1 : kBvRadix;
        }
        w_out[sy] = (uint8_t)wgt;
      }
      const int n_entries = g.n_blocks * kBvEntries * bvS;
#pragma unroll 1
      for (int x = threadIdx.x; x < n_entries; x += blockDim.x) {
        const int st = x % bvS, en = (x / bvS) % kBvEntries, b = x / (bvS * kBvEntries);
        if (st < meta->n_states[b]) {
          const uint32_t c1 = en % kBvRadix, c2 = en / kBvRadix, tot = c1 + c2;
          const int jn = meta->null_pos[b];
          double sum = 0.0;  // the same values in the same order as the per-candidate table build below
#pragma unroll 1
          for (int j = meta->bbeg[b]; j < meta->bbeg[b + 1]; ++j) {
            double a;
            if (j == jn) a = tot == 0u ? 1.0 : 0.0;
            else a = tot ? __ldg(&g_frac.v[(((j - meta->bbeg[b]) - (jn < j ? 1 : 0)) == 0 ? c1 : c2) * 33u + tot]) : 0.0;
            sum += fabs(a - stateP[j * kMaxStates + st]);
          }
          c_out[x] = sum;
        }
      }
    }
    __syncthreads();  // bv_on is the same for every thread of the CTA
  }

#pragma unroll 1
  for (int e = 2 * lane; e < e_pad; e += 64) *reinterpret_cast<uint2*>(acc + e) = make_uint2(0u, 0u);
  // (32-bit arithmetic on the low word of the counter while the pool and the grid together stay below 2^31 claims)
  const bool small_claims = p.n_cand + n_warps_grid < 0x7fffffffLL;
  auto claim = [&]() -> int64_t {
    if (small_claims) {
      uint32_t v = 0u;
      if (lane == 0) v = atomicAdd(reinterpret_cast<unsigned int*>(p.next), 1u);
      return (int64_t)((uint32_t)n_warps_grid + __shfl_sync(full, v, 0));
    }
    unsigned long long v = 0ULL;
    if (lane == 0) v = atomicAdd(p.next, 1ULL);
    return n_warps_grid + (int64_t)__shfl_sync(full, v, 0);
  };
  for (int64_t t1; t < p.n_cand; t = t1) {
    Slice cur;
    t1 = claim();
    if (kFused) {
      const RawTree craw = nraw;
      if (t1 < p.n_cand) nraw = load_raw(p.trees + (size_t)t1 * kTreeBytes, lane);  // prefetch behind this tree's work
      cur = warp_scan_tree(craw, p.trees + (size_t)t * kTreeBytes, p.n_symbols, wscan, lane);
      if (p.status && lane == 0) p.status[t] = (int32_t)(cur.meta >> 24);
    } else {
      cur = nxt;
      if (t1 < p.n_cand) nxt = load_slice(t1);  // prefetch the next record behind this tree's work
    }

    const int nn = cur.meta & 0xff, nl = (cur.meta >> 8) & 0xff, ni = (cur.meta >> 16) & 0xff;
    const bool bad = (cur.meta >> 24) != 0;
    const size_t row = (size_t)t * p.ld;

    if (bad) {  // malformed record: poison the row, status already written by the scan (cold path: keep it small)
      const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll 1
      for (int e = lane; e < ld_i; e += 32) {  // padding columns stay finite like those of every other row
        const double v = e < E ? qnan : 0.0;
        if (p.K) store_out<OutT>(p.K, row + e, v);
        if (e < E) {
          if (p.D) store_out<OutT>(p.D, row + e, v);
          if (p.Df) { store_out<OutT>(p.Df, row + e, v); store_out<OutT>(p.Df, plane + row + e, v); store_out<OutT>(p.Df, 2 * plane + row + e, v); }
        }
      }
      if (lane == 0 && p.bad) atomicAdd(p.bad, 1);
      continue;
    }

    __syncwarp();  // acc[] is all zero here: zeroed before the loop, and every entry is cleared again by the epilogue that reads it
    if (!bv_on) {
#pragma unroll 1
      for (int s = lane; s < g.n_symbols; s += 32) scnt[s] = 0u;  // only the grammar's symbols / blocks are ever read
    }
    if (lane < g.n_blocks) btot[lane] = 0u;  // leaf count per block, or (cache) the index of the block's vector
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
    const int blk = lead_sym ? (int)g.sym_block[sym] : -1;
    bool cached = bv_on;  // DIM_WISE of this candidate through the block-vector cache
    if (bv_on) {
      if (blk >= 0) atomicAdd(&btot[blk], cnt_sym * bvw[sym]);
      cached = !__any_sync(full, blk >= 0 && cnt_sym > (uint32_t)kBvMaxCount);
      if (!cached) {  // a kind with more leaves than the cache covers: this candidate builds its own table
        __syncwarp();
#pragma unroll 1
        for (int s = lane; s < g.n_symbols; s += 32) scnt[s] = 0u;
        if (lane < g.n_blocks) btot[lane] = 0u;
        __syncwarp();
      }
    }
    if (!cached && lead_sym) {
      scnt[sym] = cnt_sym;
      if (blk >= 0) atomicAdd(&btot[blk], cnt_sym);
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

    // candidate DIM_WISE vector from the symbol counts (optimal_transport_mappings.py:88-108); count / total comes
    // from the table of correctly rounded quotients (== float(a) / float(b)); every permuted column is written
    if (!cached) {
#pragma unroll 1
      for (int s = lane; s < g.n_symbols; s += 32) {
        const int b = g.sym_block[s], j = meta->sym_pos[s];
        if (b >= 0 && j >= 0) {
          const uint32_t tot = btot[b];
          avec[j] = tot ? __ldg(&g_frac.v[scnt[s] * 33u + tot]) : 0.0;
        }
      }
      if (lane < g.n_blocks) avec[meta->null_pos[lane]] = btot[lane] == 0u ? 1.0 : 0.0;
    }

    // Walk the postings of all matched features as ONE flattened list: an exclusive scan of the list lengths
    // over the lanes gives every feature its segment [start, start + len); the non-empty segments are compacted
    // into slot[0..) in lane order (= increasing start).  In each round lane j takes posting q = 32 * round + j;
    // its segment is found without a search: the lanes whose segment starts inside this round set bit
    // (start - q0) of a mask (warp OR-reduction), and the number of set bits at or below j, plus the number
    // of segments that started in earlier rounds, is the segment's ordinal.  Postings of different features
    // may hit the same evaluated tree, hence the shared-memory atomic.
    auto walk = [&](uint2 r, uint32_t cnt_a, uint32_t Ta, int shift) {
      uint32_t inc = r.y;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(full, inc, o);
        if (lane >= o) inc += v;
      }
      const uint32_t total = __shfl_sync(full, inc, 31);
      if (total == 0u) return;
      const uint32_t start = inc - r.y;
      const bool ne = r.y != 0u;
      const unsigned ne_mask = __ballot_sync(full, ne);
      if (ne) slot[__popc(ne_mask & lt_mask)] = make_uint2(r.x - start, cnt_a);
      __syncwarp();
      const uint32_t startx = ne ? start : 0x80000000u;  // empty segments never start inside a round
      uint32_t before = 0u;  // segments that start before this round
#pragma unroll 1
      for (uint32_t q0 = 0; q0 < total; q0 += 32) {
        const uint32_t m = __reduce_or_sync(full, shl_clamp(1u, startx - q0));  // shl.b32: 0 for shift amounts >= 32
        const uint32_t ord = before + (uint32_t)__popc(m & le_mask) - 1u;
        before += (uint32_t)__popc(m);
        const uint32_t q = q0 + lane;
        // the lanes past the end of the list (last round only) read on into the blob (the postings are followed by
        // the other tables of the evaluated set) and drop what they read: cheaper than a divergent branch per round
        const uint2 sl = slot[ord];
        const uint32_t pq = __ldg(&post[sl.x + q]);
        const uint32_t e = pq & 0xffffu, cb = (pq >> 16) & 0xffu, Tb = pq >> 24;
        const uint32_t contrib = min(sl.y * Tb, cb * Ta) << shift;
        if (q < total) atomicAdd(&acc[e], contrib);
      }
      __syncwarp();
    };
    walk(r_int, (uint32_t)__popc(m_id), (uint32_t)nn, 0);
    walk(r_pid, (uint32_t)__popc(m_pid), (uint32_t)nl, 16);

    // block-state distance table: one (block, state) pair per lane and round
    if (use_states && !cached) {
#pragma unroll 1
      for (int k = lane; k < n_flat; k += 32) {
        const uint32_t dsc = flat[k];
        const int j1 = (dsc >> 8) & 0xff, ts = dsc >> 16;
        const int j0 = dsc & 0xff;
        const double* sv = stateP + (ts & (kMaxStates - 1)) + j0 * kMaxStates;
        const double* av = avec + j0;
        double sum = 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)  // a block has 1 + kinds columns, four at most in the reference's search spaces
          if (j0 + u < j1) sum += fabs(av[u] - sv[u * kMaxStates]);
#pragma unroll 1
        for (int j = j0 + 4; j < j1; ++j) sum += fabs(avec[j] - stateP[(ts & (kMaxStates - 1)) + j * kMaxStates]);
        tab[ts] = sum;
      }
    }
    __syncwarp();

    // ---- phase C: register tiles of R evaluated trees per lane (tree e = e0 + lane + 32 r)
    const double rTa2_sub = 2.0 * c_recip.v[nn], rTa2_path = 2.0 * c_recip.v[nl];
    for (int e0 = 0; e0 < E; e0 += 32 * R) {
      const int eb = e0 + lane;  // e_pad is a whole number of tiles: eb + 32 r < e_pad, padded arrays stay in bounds
      uint32_t msub2[RP];
      double ddim[R];
#pragma unroll
      for (int q = 0; q < RP; ++q) msub2[q] = 0u;
#pragma unroll
      for (int r = 0; r < R; ++r) ddim[r] = 0.0;
      // SUBTREES, leaf classes: sum over the candidate's distinct symbols of min(ca * Tb, cb * Ta), two evaluated
      // trees per 32-bit word (all products and sums stay below 2^16: ca, cb <= 32, Ta, Tb <= 63)
      {
        const int wb = (e0 >> 1) + lane;
        uint32_t tb2[RP];
#pragma unroll
        for (int q = 0; q < RP; ++q) tb2[q] = e_nn2[wb + 32 * q];
        for (unsigned todo = sym_leads; todo;) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const uint32_t sc = __shfl_sync(full, sym | (cnt_sym << 8), src);  // symbol < 64, count <= 32
          const uint32_t s = sc & 0xffu, ca = sc >> 8;
          const uint32_t* col = symT2 + s * (uint32_t)half_pad + wb;
#pragma unroll
          for (int q = 0; q < RP; ++q) msub2[q] += __vminu2(ca * tb2[q], col[32 * q] * (uint32_t)nn);
        }
      }
      // DIM_WISE: per block, then across blocks (the same grouping in both modes, so they agree bit for bit)
      if (use_states) {
        const uint8_t* sp = sid8T + eb;
#pragma unroll 1
        for (int b = 0; b < g.n_blocks; ++b, sp += e_pad) {
          // row of L1(candidate block b, state): the cached row of the block's vector, or this candidate's own table
          const uint8_t* tb = cached ? bvc + (size_t)((b * kBvEntries + (int)btot[b]) * bvS) * 8
                                     : reinterpret_cast<const uint8_t*>(tab) + (size_t)b * 8 * kMaxStates;
#pragma unroll
          for (int r = 0; r < R; ++r) ddim[r] += *reinterpret_cast<const double*>(tb + sp[32 * r]);
        }
      } else {
#pragma unroll 1
        for (int b = 0; b < g.n_blocks; ++b) {
          double part[R];
#pragma unroll
          for (int r = 0; r < R; ++r) part[r] = 0.0;
#pragma unroll 1
          for (int j = meta->bbeg[b]; j < meta->bbeg[b + 1]; ++j) {
            const double a = avec[j];
            const double* col = dimT + (size_t)j * e_pad + eb;
#pragma unroll
            for (int r = 0; r < R; ++r) part[r] += fabs(a - col[32 * r]);
          }
#pragma unroll
          for (int r = 0; r < R; ++r) ddim[r] += part[r];
        }
      }
      // epilogue in groups of up to 4 trees per lane: distances, weighted sum, Gram entry
      uint32_t* accp = acc + eb;
      const uint8_t* nnp = e_nn + eb;
      const uint8_t* nlp = e_nl + eb;
      const double* rsp = e_rsub + eb;
      const double* rpp = e_rpath + eb;
#pragma unroll
      for (int r0 = 0; r0 < R; r0 += 4) {
        constexpr int kG = 4;
        double x[kG];
#pragma unroll
        for (int i = 0; i < kG; ++i) {
          const int r = r0 + i;
          if (r < R) {
            const uint32_t a = accp[32 * r];
            accp[32 * r] = 0u;  // ready for the next candidate
            const uint32_t m_sub = ((msub2[r >> 1] >> (16 * (r & 1))) & 0xffffu) + (a & 0xffffu), m_path = a >> 16;
            const uint32_t tb_sub = nnp[32 * r], tb_path = nlp[32 * r];
            // L1 = 2 - 2 M / (Ta Tb) = 2 (Ta Tb - M) / (Ta Tb) with the integer difference formed first: exactly 0
            // when the histograms coincide (like the reference), and (1/Ta)(1/Tb) first keeps D(a,b) == D(b,a) bitwise
            const double d_sub = u32_to_f64((uint32_t)nn * tb_sub - m_sub) * (rTa2_sub * rsp[32 * r]);
            const double d_path = u32_to_f64((uint32_t)nl * tb_path - m_path) * (rTa2_path * rpp[32 * r]);
            if (kExtra || kMode == 0) {
              // sum_f w_f * D_f in feature_type_list order (kernel_kernel_grammar_tree.py:159-161)
              const double d = __dadd_rn(__dadd_rn(__dmul_rn(p.w_dim, ddim[r]), __dmul_rn(p.w_path, d_path)),
                                         __dmul_rn(p.w_sub, d_sub));
              if (kExtra) {
                const int e = eb + 32 * r;
                if (e < E) {
                  if (p.Df) { store_out<OutT>(p.Df, row + e, ddim[r]); store_out<OutT>(p.Df, plane + row + e, d_path); store_out<OutT>(p.Df, 2 * plane + row + e, d_sub); }
                  if (p.D) store_out<OutT>(p.D, row + e, d);
                }
              }
              x[i] = d * p.neg_inv_lsq;
            }
            // bounded hyper-parameters: the Gram argument -d / l^2 as one fused sum of the three terms (same value for
            // (a, b) and (b, a), exactly -0 for coinciding histograms)
            if (kMode != 0) x[i] = fma(p.a_sub, d_sub, fma(p.a_path, d_path, p.a_dim * ddim[r]));
          } else {
            x[i] = 0.0;
          }
        }
        if (p.K && kF32) {
          // fp32 mode: one conversion, ex2.approx on the special-function unit (2 ulp of fp32), 4-byte stores
          const float varf = (float)p.variance;
#pragma unroll
          for (int i = 0; i < kG; ++i) {
            const int r = r0 + i, e = eb + 32 * r;
            if (r < R) {
              const float kv = varf * __expf((float)x[i]);
              float* Kf = reinterpret_cast<float*>(p.K);
              if (R < kMaxR && r < R - 1) Kf[row + e] = kv;
              else if (e < ld_i) Kf[row + e] = e < E ? kv : 0.0f;
            }
          }
        }
        if (p.K && !kF32) {
          if (kFast) exp_table_group<kG>(x, exp_tab);  // variance included
          else exp_nonpos_group<kG>(x);
#pragma unroll
          for (int i = 0; i < kG; ++i)  // pin the results here: otherwise every chain is sunk into its own guarded store
            if (r0 + i < R) asm volatile("" : "+d"(x[i]));
#pragma unroll
          for (int i = 0; i < kG; ++i) {
            const int r = r0 + i, e = eb + 32 * r;
            if (r < R) {
              const double kv = kFast ? x[i] : p.variance * x[i];
              // R = ceil(E / 32) when there is one tile (R < kMaxR): every row but the last lies inside E
              if (R < kMaxR && r < R - 1) p.K[row + e] = kv;
              // padding columns of the Gram row stay finite (the posterior kernel may read them)
              else if (e < ld_i) p.K[row + e] = e < E ? kv : 0.0;
            }
          }
        }
      }
    }
  }
}

// scratch of one call = 16 bytes (the work counter of the pair launch) + the compact records
constexpr size_t kScratchHead = 16;
size_t pairs_scratch_bytes(int64_t n_cand) { return align16(kScratchHead + (size_t)n_cand * kScanBytes); }

int launch_pairs(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar, const void* d_eval_blob,
                 int n_eval, const double weights[3], double variance, double lengthscale, double* d_K, double* d_D,
                 double* d_Df, int64_t ld, int32_t* d_status, void* d_scratch, size_t scratch_bytes, cudaStream_t stream,
                 int32_t* d_bad, bool zero_bad, bool f32) {
  if (!grammar || !d_eval_blob || !weights || n_cand < 0 || (n_cand > 0 && (!d_trees || !d_scratch)))
    return set_error(BOSOT_EINVAL, "bosot_sot_pairs: null argument");
  if (n_eval < 1 || n_eval > BOSOT_MAX_EVAL) return set_error(BOSOT_EINVAL, "bosot_sot_pairs: n_eval out of range");
  if (ld < n_eval) return set_error(BOSOT_EINVAL, "bosot_sot_pairs: ld < n_eval");
  if (scratch_bytes < pairs_scratch_bytes(n_cand) || (reinterpret_cast<uintptr_t>(d_scratch) & 15))
    return set_error(BOSOT_EINVAL, "bosot_sot_pairs: scratch too small or not 16-byte aligned (bosot_sot_pairs_scratch_bytes)");
  if (int rc = check_grammar(grammar)) return rc;
  if (n_cand == 0) return BOSOT_OK;

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
  p.a_dim = p.w_dim * p.neg_inv_lsq;
  p.a_path = p.w_path * p.neg_inv_lsq;
  p.a_sub = p.w_sub * p.neg_inv_lsq;
  // Bounded Gram argument?  Every per-feature distance lies in [0, 2] (DIM_WISE: in [0, 2] per block), so with
  // non-negative weights -d / l^2 >= -2 (n_blocks w_dim + w_path + w_sub) / l^2; when that is above -600 and the
  // variance is an ordinary number, exp needs neither a clamp nor a denormal path (exp_table_group).  Anything else
  // (negative or non-finite weights, extreme length scales) takes the general exp_nonpos route.
  const double x_min = 2.0 * (grammar->n_blocks * p.w_dim + p.w_path + p.w_sub) * p.neg_inv_lsq;
  const bool fast = p.w_dim >= 0.0 && p.w_path >= 0.0 && p.w_sub >= 0.0 && x_min >= -600.0 && x_min <= 0.0 &&
                    variance >= 0x1p-100 && variance <= 0x1p100 && !(experiment_flag() & 4) && !f32;
  p.K = d_K;
  p.D = d_D;
  p.Df = d_Df;
  p.ld = ld;
  p.next = static_cast<unsigned long long*>(d_scratch);
  p.bad = d_bad;
  p.flags = experiment_flag();
  p.status = d_status;
  p.n_symbols = grammar->n_symbols;
  p.recs = static_cast<const uint8_t*>(d_scratch) + kScratchHead;

  int dev = 0, sms = 0;
  if (int rc = cuda_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (int rc = cuda_check(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) return rc;

  // Small pools (and the evaluated set itself): no scan launch, every warp scans its own candidates
  // (warp_scan_tree); beyond kFusedScanMax candidates the thread-per-tree scan launch is cheaper in issue slots.
  const bool fused = n_cand <= kFusedScanMax && !(experiment_flag() & 2);
  if (fused) {
    pair_init_kernel<<<1, 32, 0, stream>>>(p.next, zero_bad ? d_bad : nullptr);
    if (int rc = after_launch("pair_init_kernel")) return rc;
  } else {
    // launch 1: scan
    static std::atomic<bool> scan_attr[64];  // zero-initialised; the (idempotent) set-up may run twice under a race, never zero times
    if (!(dev < 64 && scan_attr[dev])) {
      if (int rc = cuda_check(cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kScanThreads * kRecBytes),
                              "cudaFuncSetAttribute(scan_kernel)")) return rc;
      if (dev < 64) scan_attr[dev] = true;
    }
    const int64_t scan_blocks = (n_cand + kScanThreads - 1) / kScanThreads;
    if (scan_blocks > 0x7fffffffLL) return set_error(BOSOT_EINVAL, "bosot_sot_pairs: too many candidates");
    scan_kernel<<<(unsigned)scan_blocks, kScanThreads, kScanThreads * kRecBytes, stream>>>(
        d_trees, n_cand, grammar->n_symbols, static_cast<uint8_t*>(d_scratch) + kScratchHead, d_status, p.next, zero_bad ? d_bad : nullptr);
    if (int rc = after_launch("scan_kernel")) return rc;
  }

  // launch 2: pairs.  One persistent CTA per SM; small pools shrink the CTA so that every SM gets warps.
  const int rounds = (n_eval + 31) / 32;  // evaluated trees per lane; beyond 256 trees: tiles of 256 (e_pad is a multiple of 256)
  const int R = rounds < kMaxR ? rounds : kMaxR;
  const int max_warps = pair_max_warps(R);
  int warps = max_warps;
  while (warps > 1 && (n_cand + warps - 1) / warps < sms) warps >>= 1;
  StageLayout S = stage_layout(p.L.e_pad, grammar->n_dim_cols, grammar->n_symbols, grammar->n_blocks, fast);
  const size_t per_warp = warp_layout(p.L.e_pad, grammar->n_dim_cols, grammar->n_blocks, fused).total;
  const size_t smem_cap = 220 * 1024;
  size_t smem = S.total + (size_t)warps * per_warp;
  while (smem > smem_cap && warps > 1) {
    warps >>= 1;
    smem = S.total + (size_t)warps * per_warp;
  }
  if (smem > smem_cap) return set_error(BOSOT_EINVAL, "bosot_sot_pairs: n_eval too large for shared memory");
  // Block-vector cache of the DIM_WISE distances: worth its set-up (a few thousand table entries per CTA) only for pools
  // with hundreds of candidates per warp, and only when it fits beside the full set of warps.  Results are bit-identical
  // with and without it (same values added in the same order), so the choice may depend on the pool size.
  p.bv_states = 0;
  if (!fused && !(experiment_flag() & 8)) {
    const StageLayout Sbv = stage_layout(p.L.e_pad, grammar->n_dim_cols, grammar->n_symbols, grammar->n_blocks, fast, kBvStates);
    if (Sbv.total + (size_t)warps * per_warp <= smem_cap) {
      p.bv_states = kBvStates;
      S = Sbv;
      smem = S.total + (size_t)warps * per_warp;
    }
  }

  const bool extra = d_D != nullptr || d_Df != nullptr;
  using KernelFn = void (*)(PairParams, bosot_grammar);
#define BOSOT_PAIR_X(r, x, f) {sot_pair_kernel<r, x, f, 0>, sot_pair_kernel<r, x, f, 1>, sot_pair_kernel<r, x, f, 2>}
#define BOSOT_PAIR_ROW(r) {{BOSOT_PAIR_X(r, false, false), BOSOT_PAIR_X(r, false, true)}, {BOSOT_PAIR_X(r, true, false), BOSOT_PAIR_X(r, true, true)}}
  static const KernelFn table[kMaxR][2][2][3] = {BOSOT_PAIR_ROW(1), BOSOT_PAIR_ROW(2), BOSOT_PAIR_ROW(3), BOSOT_PAIR_ROW(4),
                                                 BOSOT_PAIR_ROW(5), BOSOT_PAIR_ROW(6), BOSOT_PAIR_ROW(7), BOSOT_PAIR_ROW(8)};
#undef BOSOT_PAIR_ROW
#undef BOSOT_PAIR_X
  KernelFn kernel = table[R - 1][extra ? 1 : 0][fused ? 1 : 0][f32 ? 2 : (fast ? 1 : 0)];
  static std::atomic<bool> attr_set[64];  // zero-initialised; the (idempotent) set-up may run twice under a race, never zero times
  if (!(dev < 64 && attr_set[dev])) {
    for (int r = 0; r < kMaxR; ++r)
      for (int x = 0; x < 12; ++x)
        if (int rc = cuda_check(cudaFuncSetAttribute(table[r][x / 6][(x / 3) & 1][x % 3], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap),
                                "cudaFuncSetAttribute(sot_pair_kernel)")) return rc;
    if (dev < 64) attr_set[dev] = true;
  }
  int per_sm = 1;
  if (int rc = cuda_check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem),
                          "cudaOccupancyMaxActiveBlocksPerMultiprocessor")) return rc;
  if (per_sm < 1) per_sm = 1;
  if (per_sm * warps > max_warps) per_sm = max_warps / warps;
  const int64_t want = (n_cand + warps - 1) / warps;
  const int64_t wave = (int64_t)sms * per_sm;
  const int64_t grid = want < wave ? want : wave;
  // programmatic dependent launch pays where launch latency matters (10k x 50: 54 against 58 us per step) and costs
  // 4 % at 1M x 200 (profiles/r2h_ab.log), so only the small-pool path chains its launches that way
  if (int rc = cuda_check(launch_pdl(fused, kernel, dim3((unsigned)grid), dim3(warps * 32), smem, stream, p, *grammar), "cudaLaunchKernelEx(sot_pair_kernel)")) return rc;
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
                             d_Df, ld, d_status, d_scratch, scratch_bytes, (cudaStream_t)stream, nullptr, false, false);
}

extern "C" int bosot_sot_pairs_f32(const uint8_t* d_trees, int64_t n_cand, const bosot_grammar* grammar,
                                   const void* d_eval_blob, int n_eval, const double weights[3],
                                   double variance, double lengthscale,
                                   float* d_K, float* d_D, float* d_Df, int64_t ld,
                                   int32_t* d_status, void* d_scratch, size_t scratch_bytes, void* stream) {
  return bosot::launch_pairs(d_trees, n_cand, grammar, d_eval_blob, n_eval, weights, variance, lengthscale,
                             reinterpret_cast<double*>(d_K), reinterpret_cast<double*>(d_D), reinterpret_cast<double*>(d_Df), ld,
                             d_status, d_scratch, scratch_bytes, (cudaStream_t)stream, nullptr, false, true);
}
