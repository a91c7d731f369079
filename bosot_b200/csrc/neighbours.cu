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

// one move on one record (host and device): res = neighbour number k of rec; returns 0 or a defect code
__host__ __device__ inline int neighbour_move(const uint8_t* rec, int64_t k, int B, uint8_t* res) {
  for (int q = 0; q < kTreeBytes; ++q) res[q] = BOSOT_PAD;
  const int n = rec[0];
  if (n < 1 || n > kMaxNodes || !(n & 1)) return BOSOT_TREE_BAD_LENGTH;
  const int L = (n + 1) >> 1;
  const int64_t n_leaf_moves = (int64_t)L * B;
  if (k < 0 || k >= n_leaf_moves + (int64_t)n * 2 * B) return BOSOT_NEIGHBOUR_BAD_INDEX;
  if (k < n_leaf_moves) {
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
    return seen != L ? BOSOT_TREE_BAD_SHAPE : 0;
  }
  if (n + 2 > kMaxNodes) return BOSOT_NEIGHBOUR_TOO_BIG;
  const int64_t kk = k - n_leaf_moves;
  const int node = (int)(kk / (2 * B));  // pre-order position, 0 = root
  const int op = (int)((kk % (2 * B)) / B), sym = (int)(kk % B);
  int end = node, pending = 1;           // last token of the sub-expression rooted at `node`
  for (int p = node; p < n; ++p) {
    pending += (rec[1 + p] >= 0x80) ? 1 : -1;
    if (pending == 0) { end = p; break; }
  }
  if (pending != 0) return BOSOT_TREE_BAD_SHAPE;
  int w = 1;
  for (int p = 0; p < node; ++p) res[w++] = rec[1 + p];
  res[w++] = (uint8_t)(BOSOT_OP_ADD + op);
  for (int p = node; p <= end; ++p) res[w++] = rec[1 + p];
  res[w++] = (uint8_t)sym;
  for (int p = end + 1; p < n; ++p) res[w++] = rec[1 + p];
  res[0] = (uint8_t)(n + 2);
  return 0;
}

__global__ void neighbour_kernel(const uint8_t* __restrict__ trees, int64_t n_trees, const int64_t* __restrict__ index,
                                 int n_symbols, uint8_t* __restrict__ out, int32_t* __restrict__ status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_trees) return;
  uint8_t rec[kTreeBytes], res[kTreeBytes];
  const uint4* src = reinterpret_cast<const uint4*>(trees + i * kTreeBytes);
#pragma unroll
  for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(rec)[q] = __ldg(src + q);
  const int st = neighbour_move(rec, index[i], n_symbols, res);
  if (st) {  // leave the input unchanged in the output slot so downstream kernels stay well defined
#pragma unroll
    for (int q = 0; q < kTreeBytes; ++q) res[q] = rec[q];
  }
  uint4* dst = reinterpret_cast<uint4*>(out + i * kTreeBytes);
#pragma unroll
  for (int q = 0; q < 4; ++q) dst[q] = reinterpret_cast<const uint4*>(res)[q];
  if (status) status[i] = st;
}

// ---- host side of the candidate producers: the reference draws every neighbour number with np.random.choice(n),
// one call per offspring, in a data-dependent order.  To propose the same kernels from the same seed the draws are
// replayed here from a block of raw 32-bit outputs of NumPy's global generator: RandomState.choice(n) = randint(0, n)
// = masked rejection on 32-bit outputs (mask = next power of two minus one, redraw while value > n - 1; no draw at
// all for n == 1).  The caller advances the global generator by the number of outputs consumed.
// The raw outputs come either from a block the caller drew in advance (RawBlock: the caller restores the generator
// and advances it by the number consumed) or straight from the generator through its C interface (RawCallback:
// numpy.random.get_bit_generator().ctypes.next_uint32 -- the generator ends up exactly where the reference's sequence
// of np.random.choice calls would have left it, no block, no save / restore).
struct RawBlock {
  const uint32_t* raw;
  int64_t n_raw, pos;
  __host__ bool next(uint32_t& v) {
    if (pos >= n_raw) return false;
    v = raw[pos++];
    return true;
  }
};
struct RawCallback {
  bosot_next_u32_fn fn;
  void* state;
  int64_t pos;
  __host__ bool next(uint32_t& v) {
    v = fn(state);
    ++pos;
    return true;
  }
};

template <class Src>
static inline bool legacy_choice(Src& src, int64_t n, int64_t& out) {
  const uint64_t rng = (uint64_t)n - 1;
  if (rng == 0) { out = 0; return true; }
  uint64_t mask = rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  while (true) {
    uint32_t raw;
    if (!src.next(raw)) return false;
    const uint64_t v = raw & mask;
    if (v <= rng) { out = (int64_t)v; return true; }
  }
}

static inline int64_t neighbour_count(const uint8_t* rec, int B) {  // kernel_grammar_search_spaces.py:97-106
  const int64_t n = rec[0];
  return ((n + 1) / 2) * B + n * 2 * B;
}

template <class Src>
static int host_choices(Src& src, const int64_t* h_n, int64_t count, int64_t* h_out, int64_t* h_consumed) {
  if (!h_n || !h_out || !h_consumed || count < 0) return set_error(BOSOT_EINVAL, "bosot_host_choices: bad argument");
  for (int64_t i = 0; i < count; ++i)
    if (h_n[i] < 1 || h_n[i] > 0x100000000LL) return set_error(BOSOT_EINVAL, "bosot_host_choices: population size out of range");
  for (int64_t i = 0; i < count; ++i)
    if (!legacy_choice(src, h_n[i], h_out[i])) return BOSOT_ERAW;
  *h_consumed = src.pos;
  return BOSOT_OK;
}

template <class Src>
static int host_recursive_dataset(Src& src, int n_symbols, int64_t n_data, int n_per_step, uint8_t* h_pool, int64_t* h_consumed) {
  if (!h_pool || !h_consumed || n_data < 0 || n_per_step < 1) return set_error(BOSOT_EINVAL, "bosot_host_recursive_dataset: bad argument");
  if (n_symbols < 1 || n_symbols > BOSOT_MAX_SYMBOLS) return set_error(BOSOT_EINVAL, "bosot_host_recursive_dataset: n_symbols out of range");
  const int B = n_symbols;
  int64_t len = 0;
  for (; len < B; ++len) {  // the roots: one leaf per base kernel
    uint8_t* r = h_pool + len * kTreeBytes;
    for (int q = 0; q < kTreeBytes; ++q) r[q] = BOSOT_PAD;
    r[0] = 1;
    r[1] = (uint8_t)len;
  }
  while (len < n_data) {
    int64_t c = 0;
    if (!legacy_choice(src, len, c)) return BOSOT_ERAW;
    const uint8_t* chosen = h_pool + c * kTreeBytes;
    for (int s = 0; s < n_per_step && len < n_data; ++s) {
      int64_t k = 0;
      if (!legacy_choice(src, neighbour_count(chosen, B), k)) return BOSOT_ERAW;
      const int st = neighbour_move(chosen, k, B, h_pool + len * kTreeBytes);
      if (st == BOSOT_NEIGHBOUR_TOO_BIG) return set_error(BOSOT_ETREE, "bosot_host_recursive_dataset: neighbour exceeds the 63-node flat record");
      if (st) return set_error(BOSOT_ETREE, "bosot_host_recursive_dataset: malformed record");
      ++len;
    }
  }
  *h_consumed = src.pos;
  return BOSOT_OK;
}

}  // namespace bosot

extern "C" int bosot_host_choices(const uint32_t* h_raw, int64_t n_raw, const int64_t* h_n, int64_t count, int64_t* h_out,
                                  int64_t* h_consumed) {
  if (!h_raw || n_raw < 0) return bosot::set_error(BOSOT_EINVAL, "bosot_host_choices: bad argument");
  bosot::RawBlock src{h_raw, n_raw, 0};
  return bosot::host_choices(src, h_n, count, h_out, h_consumed);
}

extern "C" int bosot_host_choices_cb(bosot_next_u32_fn next_u32, void* rng_state, const int64_t* h_n, int64_t count, int64_t* h_out,
                                     int64_t* h_consumed) {
  if (!next_u32) return bosot::set_error(BOSOT_EINVAL, "bosot_host_choices_cb: null generator callback");
  bosot::RawCallback src{next_u32, rng_state, 0};
  return bosot::host_choices(src, h_n, count, h_out, h_consumed);
}

extern "C" int bosot_host_recursive_dataset(const uint32_t* h_raw, int64_t n_raw, int n_symbols, int64_t n_data, int n_per_step,
                                            uint8_t* h_pool, int64_t* h_consumed) {
  if (!h_raw || n_raw < 0) return bosot::set_error(BOSOT_EINVAL, "bosot_host_recursive_dataset: bad argument");
  bosot::RawBlock src{h_raw, n_raw, 0};
  return bosot::host_recursive_dataset(src, n_symbols, n_data, n_per_step, h_pool, h_consumed);
}

extern "C" int bosot_host_recursive_dataset_cb(bosot_next_u32_fn next_u32, void* rng_state, int n_symbols, int64_t n_data, int n_per_step,
                                               uint8_t* h_pool, int64_t* h_consumed) {
  if (!next_u32) return bosot::set_error(BOSOT_EINVAL, "bosot_host_recursive_dataset_cb: null generator callback");
  bosot::RawCallback src{next_u32, rng_state, 0};
  return bosot::host_recursive_dataset(src, n_symbols, n_data, n_per_step, h_pool, h_consumed);
}

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
