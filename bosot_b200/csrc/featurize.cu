// Ordered sparse featurisation of flat trees, and the evaluated-set index build.
//
// bosot_featurize restates OptimalTransportKernelKernel.internal_transform_X
// (bosot/kernels/kernel_kernel_grammar_tree.py:164-185 of the reference) for the three feature
// types, emitting every histogram in the reference's dict insertion order so that the host
// can rebuild the dense [N, F] matrices of feature_matrix (:213-230) column for column:
//   SUBTREES : pre-order first occurrence            (kernel_grammar.py:599-610)
//   PATHS    : leaves grouped by symbol in pre-order first appearance, then pre-order
//              inside a group                        (kernel_grammar.py:624-638,683-696)
//   DIM_WISE : per block NULL_i then the kinds       (optimal_transport_mappings.py:88-108)
// One thread per tree; this kernel serves the evaluated set (E <= a few hundred trees), the
// dense feature-matrix API and the parity tests -- the candidate stream of the hot path is
// featurised inside the fused pair kernel (sot_pairs.cu) with the same scan_tree().
#include "sot_common.cuh"
#include "launch.cuh"

namespace bosot {

constexpr int kFeatThreads = 64;

__global__ void __launch_bounds__(kFeatThreads)
featurize_kernel(const uint8_t* __restrict__ trees, int64_t n_trees, bosot_grammar g,
                 uint8_t* __restrict__ o_n_nodes, uint8_t* __restrict__ o_n_leaves,
                 uint8_t* __restrict__ o_sub_n, uint64_t* __restrict__ o_sub_id, uint8_t* __restrict__ o_sub_cnt,
                 uint8_t* __restrict__ o_path_n, uint16_t* __restrict__ o_path_id, uint8_t* __restrict__ o_path_cnt,
                 uint8_t* __restrict__ o_sym_cnt, double* __restrict__ o_dimvec, int32_t* __restrict__ o_status) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int64_t base = (int64_t)blockIdx.x * kFeatThreads;
  const int in_block = (int)min((int64_t)kFeatThreads, n_trees - base);

  // coalesced copy of the block's records into the per-tree working records
  for (int w = threadIdx.x; w < in_block * (kTreeBytes / 16); w += kFeatThreads) {
    const int t = w / (kTreeBytes / 16), q = w % (kTreeBytes / 16);
    const uint4 v = reinterpret_cast<const uint4*>(trees + (base + t) * kTreeBytes)[q];
    uint2* dst = reinterpret_cast<uint2*>(smem + t * kRecBytes + kRecTok + q * 16);  // records are 8-byte aligned
    dst[0] = make_uint2(v.x, v.y);
    dst[1] = make_uint2(v.z, v.w);
  }
  __syncthreads();
  if ((int)threadIdx.x >= in_block) return;

  const int64_t i = base + threadIdx.x;
  uint8_t* rec = smem + threadIdx.x * kRecBytes;
  const int status = scan_tree(rec, g.n_symbols);
  o_status[i] = status;
  const int n = status ? 0 : rec[kRecMeta + 0];
  const int n_leaf = status ? 0 : rec[kRecMeta + 1];
  const int n_int = status ? 0 : rec[kRecMeta + 2];
  if (o_n_nodes) o_n_nodes[i] = (uint8_t)n;
  if (o_n_leaves) o_n_leaves[i] = (uint8_t)n_leaf;

  const uint64_t* iid = reinterpret_cast<const uint64_t*>(rec + kRecIid);
  const uint8_t* tok = rec + kRecTok + 1;
  const uint8_t* lsym = rec + kRecLeafSym;
  const uint8_t* lcode = rec + kRecLeafCode;

  // SUBTREES: walk nodes in pre-order; the k-th operator in pre-order is entry n_int-1-k of iid
  if (o_sub_id && o_sub_cnt && o_sub_n) {
    uint64_t* ids = o_sub_id + i * kMaxNodes;
    uint8_t* cnt = o_sub_cnt + i * kMaxNodes;
    int m = 0, k = 0;
    for (int p = 0; p < n; ++p) {
      const uint32_t t = tok[p];
      const uint64_t id = (t >= 0x80) ? iid[n_int - 1 - (k++)] : leaf_class(t);
      int f = 0;
      while (f < m && ids[f] != id) ++f;
      if (f == m) { ids[m] = id; cnt[m] = 1; ++m; } else { cnt[f] += 1; }
    }
    for (int f = m; f < kMaxNodes; ++f) { ids[f] = 0; cnt[f] = 0; }
    o_sub_n[i] = (uint8_t)m;
  }

  // PATHS: symbol groups in first-appearance order, leaves of a group in pre-order
  if (o_path_id && o_path_cnt && o_path_n) {
    uint16_t* ids = o_path_id + i * kMaxLeaves;
    uint8_t* cnt = o_path_cnt + i * kMaxLeaves;
    int m = 0;
    for (int l = 0; l < n_leaf; ++l) {
      const uint32_t s = lsym[l];
      bool first = true;
      for (int q = 0; q < l; ++q) first = first && (lsym[q] != s);
      if (!first) continue;
      for (int q = l; q < n_leaf; ++q) {
        if (lsym[q] != s) continue;
        const uint16_t pid = (uint16_t)(s * kPathCodes + lcode[q]);
        int f = 0;
        while (f < m && ids[f] != pid) ++f;
        if (f == m) { ids[m] = pid; cnt[m] = 1; ++m; } else { cnt[f] += 1; }
      }
    }
    for (int f = m; f < kMaxLeaves; ++f) { ids[f] = 0xFFFF; cnt[f] = 0; }
    o_path_n[i] = (uint8_t)m;
  }

  // DIM_WISE: raw symbol counts, block totals, count / total (fp64 division, like float(a)/float(b))
  if (o_sym_cnt || o_dimvec) {
    uint8_t* sc = rec + kRecTok;  // 64 B scratch: the raw tokens are no longer needed below
    for (int s = 0; s < BOSOT_MAX_SYMBOLS; ++s) sc[s] = 0;
    for (int l = 0; l < n_leaf; ++l) sc[lsym[l]] += 1;
    if (o_sym_cnt)
      for (int s = 0; s < g.n_symbols; ++s) o_sym_cnt[i * g.n_symbols + s] = sc[s];
    if (o_dimvec) {
      double* v = o_dimvec + i * g.n_dim_cols;
      for (int c = 0; c < g.n_dim_cols; ++c) v[c] = 0.0;
      if (!status) {
        for (int b = 0; b < g.n_blocks; ++b) {
          int tot = 0;
          for (int s = 0; s < g.n_symbols; ++s) tot += (g.sym_block[s] == b) ? sc[s] : 0;
          v[g.block_null_col[b]] = tot == 0 ? 1.0 : 0.0;
          if (tot)
            for (int s = 0; s < g.n_symbols; ++s)
              if (g.sym_block[s] == b && g.sym_col[s] >= 0) v[g.sym_col[s]] = (double)sc[s] / (double)tot;
        }
      }
    }
  }
}

// ---- evaluated-set index ------------------------------------------------------------------
// Single CTA.  Input: the ordered histograms of the E evaluated trees (written by
// featurize_kernel into scratch at the tail of the blob).  Output: hash table / direct tables
// -> posting ranges, postings, totals and the transposed dim-wise matrix.
constexpr int kBuildThreads = 1024;

__global__ void __launch_bounds__(kBuildThreads)
eval_build_kernel(uint8_t* __restrict__ blob, EvalLayout L, bosot_grammar g,
                  const uint8_t* __restrict__ n_nodes, const uint8_t* __restrict__ n_leaves,
                  const uint8_t* __restrict__ sub_n, const uint64_t* __restrict__ sub_id, const uint8_t* __restrict__ sub_cnt,
                  const uint8_t* __restrict__ path_n, const uint16_t* __restrict__ path_id, const uint8_t* __restrict__ path_cnt,
                  const uint8_t* __restrict__ sym_cnt, const double* __restrict__ dimvec) {
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(blob + L.off_keys);
  uint2* range = reinterpret_cast<uint2*>(blob + L.off_range);
  uint32_t* post = reinterpret_cast<uint32_t*>(blob + L.off_post);
  uint8_t* b_nn = blob + L.off_nnodes;
  uint8_t* b_nl = blob + L.off_nleaves;
  double* dimT = reinterpret_cast<double*>(blob + L.off_dimT);
  uint32_t* cursor = reinterpret_cast<uint32_t*>(blob + L.off_cursor);
  const int E = L.n_eval;
  const int tid = threadIdx.x;
  const int n_dim_cols = g.n_dim_cols, n_symbols = g.n_symbols;

  // DIM_WISE column bookkeeping: block-major ("permuted") column order, see DimMeta
  __shared__ DimMeta meta;
  __shared__ uint8_t pcol[BOSOT_MAX_DIM_COLS];                 // permuted position -> reference column
  __shared__ double states[BOSOT_MAX_DIM_COLS * kMaxStates];   // [permuted position][state]
  if (tid == 0) {
    int8_t col_block[BOSOT_MAX_DIM_COLS], pos_of_col[BOSOT_MAX_DIM_COLS];
    for (int c = 0; c < BOSOT_MAX_DIM_COLS; ++c) { col_block[c] = -1; pos_of_col[c] = -1; }
    for (int b = 0; b < g.n_blocks; ++b) col_block[g.block_null_col[b]] = (int8_t)b;
    for (int s = 0; s < n_symbols; ++s)
      if (g.sym_block[s] >= 0 && g.sym_col[s] >= 0) col_block[g.sym_col[s]] = g.sym_block[s];
    int j = 0;
    for (int b = 0; b < g.n_blocks; ++b) {
      meta.bbeg[b] = (uint8_t)j;
      for (int c = 0; c < n_dim_cols; ++c)
        if (col_block[c] == b) { pcol[j] = (uint8_t)c; pos_of_col[c] = (int8_t)j; ++j; }
    }
    for (int b = g.n_blocks; b < BOSOT_MAX_BLOCKS + 4; ++b) meta.bbeg[b] = (uint8_t)j;
    meta.n_pcols = j;
    for (int s = 0; s < BOSOT_MAX_SYMBOLS; ++s)
      meta.sym_pos[s] = (s < n_symbols && g.sym_block[s] >= 0 && g.sym_col[s] >= 0) ? pos_of_col[g.sym_col[s]] : (int8_t)-1;
    for (int b = 0; b < BOSOT_MAX_BLOCKS; ++b) {
      meta.null_pos[b] = b < g.n_blocks ? pos_of_col[g.block_null_col[b]] : (int8_t)0;
      meta.n_states[b] = 0;
    }
    meta.n_flat = 0;
    int* hdr = reinterpret_cast<int*>(blob);
    hdr[0] = E; hdr[1] = L.e_pad; hdr[2] = L.hash_cap; hdr[3] = L.n_slots; hdr[4] = n_dim_cols;
  }
  for (int x = tid; x < BOSOT_MAX_DIM_COLS * kMaxStates; x += kBuildThreads) states[x] = 0.0;
  for (int s = tid; s < L.hash_cap; s += kBuildThreads) keys[s] = 0ULL;
  for (int s = tid; s < L.n_slots; s += kBuildThreads) { range[s] = make_uint2(0u, 0u); cursor[s] = 0u; }
  double* b_rsub = reinterpret_cast<double*>(blob + L.off_rsub);
  double* b_rpath = reinterpret_cast<double*>(blob + L.off_rpath);
  for (int e = tid; e < L.e_pad; e += kBuildThreads) {
    const int nn = e < E ? n_nodes[e] : 1, nl = e < E ? n_leaves[e] : 1;
    b_nn[e] = (uint8_t)nn;
    b_nl[e] = (uint8_t)nl;
    b_rsub[e] = 1.0 / (double)(nn ? nn : 1);
    b_rpath[e] = 1.0 / (double)(nl ? nl : 1);
  }
  // packed transposed tables: word 32 q + l holds tree 64 q + l (low half) and tree 64 q + 32 + l (high half)
  const int half = L.e_pad / 2;
  uint32_t* symT2 = reinterpret_cast<uint32_t*>(blob + L.off_symT2);
  uint32_t* nn2 = reinterpret_cast<uint32_t*>(blob + L.off_nn2);
  for (int x = tid; x < BOSOT_MAX_SYMBOLS * half; x += kBuildThreads) {
    const int s = x / half, w = x % half;
    const int e_lo = 64 * (w >> 5) + (w & 31), e_hi = e_lo + 32;
    const uint32_t lo = (s < n_symbols && e_lo < E) ? sym_cnt[(size_t)e_lo * n_symbols + s] : 0u;
    const uint32_t hi = (s < n_symbols && e_hi < E) ? sym_cnt[(size_t)e_hi * n_symbols + s] : 0u;
    symT2[x] = lo | (hi << 16);
  }
  for (int w = tid; w < half; w += kBuildThreads) {
    const int e_lo = 64 * (w >> 5) + (w & 31), e_hi = e_lo + 32;
    const uint32_t lo = e_lo < E ? n_nodes[e_lo] : 1u, hi = e_hi < E ? n_nodes[e_hi] : 1u;
    nn2[w] = lo | (hi << 16);
  }
  __syncthreads();
  for (int x = tid; x < BOSOT_MAX_DIM_COLS * L.e_pad; x += kBuildThreads) {
    const int j = x / L.e_pad, e = x % L.e_pad;
    dimT[x] = (j < meta.n_pcols && e < E) ? dimvec[(size_t)e * n_dim_cols + pcol[j]] : 0.0;
  }
  // distinct block vectors ("states") of the evaluated trees: warp b owns block b, lane s compares state s
  uint8_t* sid8T = blob + L.off_sid8T;
  for (int x = tid; x < BOSOT_MAX_BLOCKS * L.e_pad; x += kBuildThreads) sid8T[x] = 0;
  __syncthreads();
  {
    const int b = tid >> 5, lane = tid & 31;
    if (b < g.n_blocks) {
      const int j0 = meta.bbeg[b], j1 = meta.bbeg[b + 1];
      int ns = 0;
      for (int e = 0; e < E; ++e) {
        bool eq = lane < ns && lane < kMaxStates;
        for (int j = j0; j < j1; ++j)
          eq = eq && (__double_as_longlong(states[j * kMaxStates + lane]) ==
                      __double_as_longlong(dimvec[(size_t)e * n_dim_cols + pcol[j]]));
        const unsigned m = __ballot_sync(0xffffffffu, eq);
        int sid;
        if (m) {
          sid = __ffs(m) - 1;
        } else {
          sid = ns;
          if (ns < kMaxStates)
            for (int j = j0 + lane; j < j1; j += 32) states[j * kMaxStates + ns] = dimvec[(size_t)e * n_dim_cols + pcol[j]];
          if (ns <= kMaxStates) ++ns;  // saturates at kMaxStates + 1 = "too many"
          __syncwarp();
        }
        if (lane == 0) sid8T[(size_t)b * L.e_pad + e] = (uint8_t)(8 * (sid < kMaxStates ? sid : 0));
      }
      if (lane == 0) meta.n_states[b] = (uint8_t)ns;
    }
  }
  __syncthreads();
  if (tid == 0) {
    int ok = 1;
    for (int b = 0; b < g.n_blocks; ++b) ok = ok && (meta.n_states[b] <= kMaxStates);
    meta.use_states = ok;
    uint32_t* flat = reinterpret_cast<uint32_t*>(blob + L.off_flat);
    int k = 0;
    for (int b = 0; b < g.n_blocks; ++b)
      for (int s = 0; s < meta.n_states[b] && s < kMaxStates; ++s)
        flat[k++] = (uint32_t)meta.bbeg[b] | ((uint32_t)meta.bbeg[b + 1] << 8) | ((uint32_t)(b * kMaxStates + s) << 16);
    meta.n_flat = k;
    for (; k < BOSOT_MAX_BLOCKS * kMaxStates; ++k) flat[k] = 0u;
  }
  __syncthreads();
  {
    double* stateP = reinterpret_cast<double*>(blob + L.off_stateP);
    for (int x = tid; x < BOSOT_MAX_DIM_COLS * kMaxStates; x += kBuildThreads) stateP[x] = states[x];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&meta);
    uint32_t* dst = reinterpret_cast<uint32_t*>(blob + L.off_meta);
    for (int x = tid; x < (int)(sizeof(DimMeta) / 4); x += kBuildThreads) dst[x] = src[x];
  }

  // pass 1: slot of every (tree, feature); count per slot (range.y)
  auto slot_of_sub = [&](uint64_t id) -> uint32_t {  // operator classes only (bit 63 set)
    uint32_t s = hash_slot(id, L.hash_cap);
    while (true) {
      const unsigned long long prev = atomicCAS(&keys[s], 0ULL, (unsigned long long)id);
      if (prev == 0ULL || prev == id) return s;
      s = (s + 1) & (uint32_t)(L.hash_cap - 1);
    }
  };
  for (int x = tid; x < E * kMaxNodes; x += kBuildThreads) {
    const int e = x / kMaxNodes, f = x % kMaxNodes;
    if (f < sub_n[e] && (sub_id[x] >> 63)) atomicAdd(&range[slot_of_sub(sub_id[x])].y, 1u);  // leaf classes live in symT
  }
  for (int x = tid; x < E * kMaxLeaves; x += kBuildThreads) {
    const int e = x / kMaxLeaves, f = x % kMaxLeaves;
    if (f < path_n[e]) atomicAdd(&range[L.hash_cap + BOSOT_MAX_SYMBOLS + path_id[x]].y, 1u);
  }
  __syncthreads();

  // pass 2: exclusive scan of the counts over the unified slot space (chunk per thread + block scan)
  __shared__ uint32_t part[kBuildThreads];
  const int per = (L.n_slots + kBuildThreads - 1) / kBuildThreads;
  const int s0 = tid * per, s1 = min(L.n_slots, s0 + per);
  uint32_t sum = 0;
  for (int s = s0; s < s1; ++s) sum += range[s].y;
  part[tid] = sum;
  __syncthreads();
  for (int off = 1; off < kBuildThreads; off <<= 1) {
    const uint32_t v = tid >= off ? part[tid - off] : 0u;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  uint32_t run = part[tid] - sum;
  for (int s = s0; s < s1; ++s) { range[s].x = run; run += range[s].y; }
  __syncthreads();

  // pass 3: fill postings (order inside a posting list is irrelevant: one posting per tree)
  for (int x = tid; x < E * kMaxNodes; x += kBuildThreads) {
    const int e = x / kMaxNodes, f = x % kMaxNodes;
    if (f < sub_n[e] && (sub_id[x] >> 63)) {
      const uint32_t s = slot_of_sub(sub_id[x]);
      const uint32_t p = atomicAdd(&cursor[s], 1u);
      post[range[s].x + p] = pack_posting((uint32_t)e, sub_cnt[x], n_nodes[e]);
    }
  }
  for (int x = tid; x < E * kMaxLeaves; x += kBuildThreads) {
    const int e = x / kMaxLeaves, f = x % kMaxLeaves;
    if (f < path_n[e]) {
      const uint32_t s = L.hash_cap + BOSOT_MAX_SYMBOLS + path_id[x];
      const uint32_t p = atomicAdd(&cursor[s], 1u);
      post[range[s].x + p] = pack_posting((uint32_t)e, path_cnt[x], n_leaves[e]);
    }
  }
}

// scratch appended to the blob for the featurised evaluated set
struct EvalScratch {
  size_t n_nodes, n_leaves, sub_n, sub_id, sub_cnt, path_n, path_id, path_cnt, sym_cnt, dimvec, total;
};
static EvalScratch eval_scratch(int E, size_t base) {
  EvalScratch S;
  size_t o = base;
  S.sub_id = o;   o = align16(o + sizeof(uint64_t) * (size_t)E * kMaxNodes);
  S.dimvec = o;   o = align16(o + sizeof(double) * (size_t)E * BOSOT_MAX_DIM_COLS);
  S.path_id = o;  o = align16(o + sizeof(uint16_t) * (size_t)E * kMaxLeaves);
  S.sub_cnt = o;  o = align16(o + (size_t)E * kMaxNodes);
  S.path_cnt = o; o = align16(o + (size_t)E * kMaxLeaves);
  S.sym_cnt = o;  o = align16(o + (size_t)E * BOSOT_MAX_SYMBOLS);
  S.n_nodes = o;  o = align16(o + (size_t)E);
  S.n_leaves = o; o = align16(o + (size_t)E);
  S.sub_n = o;    o = align16(o + (size_t)E);
  S.path_n = o;   o = align16(o + (size_t)E);
  S.total = o;
  return S;
}

}  // namespace bosot

using namespace bosot;

extern "C" int bosot_featurize(const uint8_t* d_trees, int64_t n_trees, const bosot_grammar* grammar,
                               uint8_t* d_n_nodes, uint8_t* d_n_leaves,
                               uint8_t* d_sub_n, uint64_t* d_sub_id, uint8_t* d_sub_cnt,
                               uint8_t* d_path_n, uint16_t* d_path_id, uint8_t* d_path_cnt,
                               uint8_t* d_sym_cnt, double* d_dimvec, int32_t* d_status, void* stream) {
  if (!grammar || !d_status || n_trees < 0 || (n_trees > 0 && !d_trees)) return set_error(BOSOT_EINVAL, "bosot_featurize: null argument");
  if (int rc = check_grammar(grammar)) return rc;
  if (n_trees == 0) return BOSOT_OK;
  const int64_t blocks = (n_trees + kFeatThreads - 1) / kFeatThreads;
  if (blocks > 0x7fffffffLL) return set_error(BOSOT_EINVAL, "bosot_featurize: too many trees");
  featurize_kernel<<<(unsigned)blocks, kFeatThreads, kFeatThreads * kRecBytes, (cudaStream_t)stream>>>(
      d_trees, n_trees, *grammar, d_n_nodes, d_n_leaves, d_sub_n, d_sub_id, d_sub_cnt, d_path_n, d_path_id, d_path_cnt,
      d_sym_cnt, d_dimvec, d_status);
  return after_launch("featurize_kernel");
}

extern "C" size_t bosot_eval_state_bytes(int n_eval) {
  if (n_eval < 1 || n_eval > BOSOT_MAX_EVAL) return 0;
  const EvalLayout L = eval_layout(n_eval);
  return eval_scratch(n_eval, L.total).total;
}

extern "C" int bosot_eval_build(const uint8_t* d_trees, int n_eval, const bosot_grammar* grammar,
                                void* d_blob, size_t blob_bytes, int32_t* d_status, void* stream) {
  if (!d_trees || !grammar || !d_blob || !d_status) return set_error(BOSOT_EINVAL, "bosot_eval_build: null argument");
  if (n_eval < 1 || n_eval > BOSOT_MAX_EVAL) return set_error(BOSOT_EINVAL, "bosot_eval_build: n_eval out of range");
  if (int rc = check_grammar(grammar)) return rc;
  const EvalLayout L = eval_layout(n_eval);
  const EvalScratch S = eval_scratch(n_eval, L.total);
  if (blob_bytes < S.total) return set_error(BOSOT_EINVAL, "bosot_eval_build: blob too small");
  if (reinterpret_cast<uintptr_t>(d_blob) & 15) return set_error(BOSOT_EINVAL, "bosot_eval_build: blob must be 16-byte aligned");
  uint8_t* b = static_cast<uint8_t*>(d_blob);
  cudaStream_t st = (cudaStream_t)stream;
  // dim-wise scratch is written with the grammar's own column count
  const int blocks = (n_eval + kFeatThreads - 1) / kFeatThreads;
  featurize_kernel<<<blocks, kFeatThreads, kFeatThreads * kRecBytes, st>>>(
      d_trees, n_eval, *grammar, b + S.n_nodes, b + S.n_leaves, b + S.sub_n,
      reinterpret_cast<uint64_t*>(b + S.sub_id), b + S.sub_cnt, b + S.path_n,
      reinterpret_cast<uint16_t*>(b + S.path_id), b + S.path_cnt, b + S.sym_cnt,
      reinterpret_cast<double*>(b + S.dimvec), d_status);
  if (int rc = after_launch("featurize_kernel(eval)")) return rc;
  eval_build_kernel<<<1, kBuildThreads, 0, st>>>(
      b, L, *grammar, b + S.n_nodes, b + S.n_leaves, b + S.sub_n,
      reinterpret_cast<const uint64_t*>(b + S.sub_id), b + S.sub_cnt, b + S.path_n,
      reinterpret_cast<const uint16_t*>(b + S.path_id), b + S.path_cnt, b + S.sym_cnt,
      reinterpret_cast<const double*>(b + S.dimvec));
  return after_launch("eval_build_kernel");
}

extern "C" int bosot_eval_info(const void* d_blob, int n_eval, int32_t h_info[4], void* stream) {
  if (!d_blob || !h_info) return set_error(BOSOT_EINVAL, "bosot_eval_info: null argument");
  if (n_eval < 1 || n_eval > BOSOT_MAX_EVAL) return set_error(BOSOT_EINVAL, "bosot_eval_info: n_eval out of range");
  const EvalLayout L = eval_layout(n_eval);
  DimMeta m;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = cuda_check(cudaMemcpyAsync(&m, static_cast<const uint8_t*>(d_blob) + L.off_meta, sizeof(m), cudaMemcpyDeviceToHost, st),
                          "cudaMemcpyAsync(DimMeta)")) return rc;
  if (int rc = cuda_check(cudaStreamSynchronize(st), "cudaStreamSynchronize")) return rc;
  int most = 0;
  for (int b = 0; b < BOSOT_MAX_BLOCKS; ++b) most = m.n_states[b] > most ? m.n_states[b] : most;
  h_info[0] = m.use_states;
  h_info[1] = m.n_pcols;
  h_info[2] = most;
  h_info[3] = n_eval;
  return BOSOT_OK;
}
