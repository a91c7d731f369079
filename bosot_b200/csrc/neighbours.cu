// CKS neighbourhood moves on the flat encoding (the producer side of the hot path).
//
// Restates BaseKernelGrammarSearchSpace.get_neighbour_at_index
// (bosot/kernels/kernel_grammar/kernel_grammar_search_spaces.py:108-142 of the reference) for one
// neighbour number per tree.  With L leaves, n nodes and B base kernels a tree has L*B + n*2*B
// neighbours (:97-106):
//   k <  L*B             : leaf floor(k / B) (pre-order) is replaced by base kernel k mod B
//   k >= L*B, k' = k-L*B : sub-expression floor(k' / 2B) (pre-order, 0 = root) becomes
//                          (operator floor((k' mod 2B) / B); old sub-expression; base kernel k' mod B)
// The neighbour number is drawn on the host (np.random, like the reference), so a seeded run proposes
// the same kernels; the tree surgery runs here, one thread per tree.
#include "sot_common.cuh"
#include "launch.cuh"

namespace bosot {

__global__ void neighbour_kernel(const uint8_t* __restrict__ trees, int64_t n_trees, const int64_t* __restrict__ index,
                                 int n_symbols, uint8_t* __restrict__ out, int32_t* __restrict__ status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_trees) return;
  uint8_t rec[kTreeBytes], res[kTreeBytes];
  const uint4* src = reinterpret_cast<const uint4*>(trees + i * kTreeBytes);
#pragma unroll
  for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(rec)[q] = __ldg(src + q);
#pragma unroll
  for (int q = 0; q < kTreeBytes; ++q) res[q] = BOSOT_PAD;
  const int n = rec[0];
  const int B = n_symbols;
  int st = 0;
  if (n < 1 || n > kMaxNodes || !(n & 1)) {
    st = BOSOT_TREE_BAD_LENGTH;
  } else {
    const int L = (n + 1) >> 1;
    const int64_t k = index[i];
    const int64_t n_leaf_moves = (int64_t)L * B;
    if (k < 0 || k >= n_leaf_moves + (int64_t)n * 2 * B) {
      st = BOSOT_NEIGHBOUR_BAD_INDEX;
    } else if (k < n_leaf_moves) {
      const int which = (int)(k / B), sym = (int)(k % B);
      int seen = 0;
      for (int p = 1; p <= n; ++p) {
        uint8_t t = rec[p];
        if (t < 0x80) {
          if (seen == which) t = (uint8_t)sym;
          ++seen;
        }
        res[p] = t;
      }
      res[0] = (uint8_t)n;
      if (seen != L) st = BOSOT_TREE_BAD_SHAPE;
    } else if (n + 2 > kMaxNodes) {
      st = BOSOT_NEIGHBOUR_TOO_BIG;
    } else {
      const int64_t kk = k - n_leaf_moves;
      const int node = (int)(kk / (2 * B));  // pre-order position, 0 = root
      const int op = (int)((kk % (2 * B)) / B), sym = (int)(kk % B);
      int end = node, pending = 1;           // last token of the sub-expression rooted at `node`
      for (int p = node; p < n; ++p) {
        pending += (rec[1 + p] >= 0x80) ? 1 : -1;
        if (pending == 0) { end = p; break; }
      }
      if (pending != 0) {
        st = BOSOT_TREE_BAD_SHAPE;
      } else {
        int w = 1;
        for (int p = 0; p < node; ++p) res[w++] = rec[1 + p];
        res[w++] = (uint8_t)(BOSOT_OP_ADD + op);
        for (int p = node; p <= end; ++p) res[w++] = rec[1 + p];
        res[w++] = (uint8_t)sym;
        for (int p = end + 1; p < n; ++p) res[w++] = rec[1 + p];
        res[0] = (uint8_t)(n + 2);
      }
    }
  }
  if (st) {  // leave the input unchanged in the output slot so downstream kernels stay well defined
#pragma unroll
    for (int q = 0; q < kTreeBytes; ++q) res[q] = rec[q];
  }
  uint4* dst = reinterpret_cast<uint4*>(out + i * kTreeBytes);
#pragma unroll
  for (int q = 0; q < 4; ++q) dst[q] = reinterpret_cast<const uint4*>(res)[q];
  if (status) status[i] = st;
}

}  // namespace bosot

extern "C" int bosot_neighbours(const uint8_t* d_trees, int64_t n_trees, const int64_t* d_index, int n_symbols,
                                uint8_t* d_out, int32_t* d_status, void* stream) {
  using namespace bosot;
  if (n_trees < 0 || (n_trees > 0 && (!d_trees || !d_index || !d_out))) return set_error(BOSOT_EINVAL, "bosot_neighbours: null argument");
  if (n_symbols < 1 || n_symbols > BOSOT_MAX_SYMBOLS) return set_error(BOSOT_EINVAL, "bosot_neighbours: n_symbols out of range");
  if (n_trees == 0) return BOSOT_OK;
  const int64_t blocks = (n_trees + 127) / 128;
  if (blocks > 0x7fffffffLL) return set_error(BOSOT_EINVAL, "bosot_neighbours: too many trees");
  neighbour_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(d_trees, n_trees, d_index, n_symbols, d_out, d_status);
  return after_launch("neighbour_kernel");
}
