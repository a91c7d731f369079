// Shared device code of the SOT hot path: flat-tree scans and canonical class ids.
//
// Reference semantics restated here (boschresearch/bosot, bosot/kernels/kernel_grammar/kernel_grammar.py):
//   * structural class of a subtree = (operator, unordered {class(left), class(right)}),
//     leaf class = its base-kernel symbol                       (get_hash, :268-271, :452-470)
//   * reduced leaf->root operator path: consecutive equal operators collapse, so a
//     path is fully described by (length, operator nearest to the leaf)
//                                                               (:624-658, :683-696, :216-220)
// The reference identifies classes by salted Python hash(); only the induced partition
// is portable, so classes are identified here by a 64-bit mix of the canonical triple.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/bosot_b200.h"

namespace bosot {

constexpr int kTreeBytes = BOSOT_TREE_BYTES;
constexpr int kMaxNodes = BOSOT_MAX_NODES;
constexpr int kMaxLeaves = BOSOT_MAX_LEAVES;
constexpr int kMaxInternal = BOSOT_MAX_INTERNAL;
constexpr int kPathCodes = BOSOT_PATH_CODES;

// Per-tree working record in shared memory (scan phase).  Stride 424 B = 8 * 53 (odd multiple of 8):
// 64-bit accesses of 32 threads working on 32 different trees spread over all banks.  The first
// kScanBytes bytes are the compact record that leaves the scan (class ids, meta, leaf symbols, reduced
// path codes); the rest is scratch of the two passes.
constexpr int kRecIid = 0;        // uint64 iid[31]   class ids of internal nodes (by reverse pre-order rank)
constexpr int kRecMeta = 248;     // uint8  n_nodes, n_leaves, n_internal, status (+4 pad)
constexpr int kRecLeafSym = 256;  // uint8  leaf_sym[32]   leaves in pre-order (left to right)
constexpr int kRecLeafCode = 288; // uint8  leaf_code[32]  reduced path code per leaf
constexpr int kScanBytes = 320;   // compact record = what the pair kernel consumes
constexpr int kRecTok = 320;      // uint8  raw record: n, tokens (64 B)
constexpr int kRecStack = 384;    // uint8  stack[32] scratch of the two scans
constexpr int kRecBytes = 424;

__device__ __forceinline__ uint64_t fmix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// Canonical id of an operator node from its operator (1 = ADD, 2 = MULTIPLY) and the ids of
// its children; symmetric in the children (commutative class), bit 63 always set so that
// it can never equal a leaf id (symbol + 1 <= 64).
__device__ __forceinline__ uint64_t class_combine(uint32_t op, uint64_t a, uint64_t b) {
  const uint64_t lo = a < b ? a : b;
  const uint64_t hi = a < b ? b : a;
  uint64_t x = fmix64(lo * 0x9E3779B97F4A7C15ULL + op);
  x = fmix64((x ^ hi) * 0xC2B2AE3D27D4EB4FULL + 0x165667B19E3779F9ULL);
  return x | 0x8000000000000000ULL;
}

__device__ __forceinline__ uint64_t leaf_class(uint32_t sym) { return (uint64_t)sym + 1; }

// Path-code transition: state of the children of an operator node with operator bit `o`
// (0 ADD, 1 MULTIPLY) whose own state is `cur` (0 = root / empty path).
__device__ __forceinline__ uint32_t path_child_code(uint32_t cur, uint32_t o) {
  if (cur != 0 && ((cur - 1) & 1) == o) return cur;  // same operator as the one above: run collapses
  const uint32_t len = cur == 0 ? 0 : ((cur - 1) >> 1) + 1;
  return 1 + 2 * len + o;
}

// One thread scans one tree record (already copied to rec + kRecTok).  Three passes, arranged so that the
// threads of a warp (32 different trees) diverge as little as possible:
//   1. backward over the pre-order tokens = stack evaluation of the STRUCTURE only: for every operator node
//      (numbered in reverse pre-order) the references of its two children -- a few instructions per token;
//   2. over the operator nodes in that order (children always come first): the 64-bit class id -- the
//      expensive mixing runs in lock step, every lane busy until its own tree runs out of operators;
//   3. forward = top-down propagation of the reduced-path state -> (symbol, code) of every leaf.
// Writes the meta bytes; returns 0 or a BOSOT_TREE_* defect code.
__device__ __forceinline__ int scan_tree(uint8_t* rec, int n_symbols) {
  uint64_t* iid = reinterpret_cast<uint64_t*>(rec + kRecIid);
  const uint8_t* tok = rec + kRecTok + 1;
  uint8_t* st = rec + kRecStack;
  uint8_t* lsym = rec + kRecLeafSym;
  uint8_t* lcode = rec + kRecLeafCode;
  const int n = rec[kRecTok];
  int status = 0;
  int n_int = 0, n_leaf = 0;
  if (n < 1 || n > kMaxNodes || !(n & 1)) {
    status = BOSOT_TREE_BAD_LENGTH;
  } else {
    // Pass 1, branch-free per token (leaf and operator lanes of a warp run the same instructions): an operator pops
    // its two children and pushes itself, a leaf pushes itself; defects only set flags, the stack pointer is kept in
    // range by clamping, and the verdict is formed after the loop.
    int sp = 0;
    int first = 0;  // the first defect met (scanning from the last token) wins, token checks before shape checks
    for (int i = n - 1; i >= 0; --i) {
      const uint32_t t = tok[i];
      const bool is_op = t >= 0x80;
      const bool bad_tok = is_op ? (t > BOSOT_OP_MUL) : ((int)t >= n_symbols);
      const bool bad_shape = is_op ? (sp < 2 || n_int >= kMaxInternal) : (sp >= kMaxLeaves);
      first = first ? first : (bad_tok ? BOSOT_TREE_BAD_TOKEN : (bad_shape ? BOSOT_TREE_BAD_SHAPE : 0));
      const uint32_t a = st[max(sp - 1, 0)], b = st[max(sp - 2, 0)];
      const int slot = min(n_int, kMaxInternal - 1);
      if (is_op) iid[slot] = (uint64_t)(a | (b << 8) | ((t - 0x7f) << 16));  // children + operator, replaced by the id in pass 2
      sp = is_op ? max(sp - 2, 0) : min(sp, kMaxLeaves - 1);
      st[sp++] = is_op ? (uint8_t)(0x80 | slot) : (uint8_t)t;
      n_int += is_op ? 1 : 0;
    }
    status = first ? first : (sp != 1 ? BOSOT_TREE_BAD_SHAPE : 0);
    if (status) n_int = 0;
    if (status == 0) {
      for (int r = 0; r < n_int; ++r) {
        const uint32_t packed = (uint32_t)iid[r];
        const uint32_t a = packed & 0xff, b = (packed >> 8) & 0xff, op = packed >> 16;
        const uint64_t ia = (a & 0x80) ? iid[a & 0x7f] : leaf_class(a);
        const uint64_t ib = (b & 0x80) ? iid[b & 0x7f] : leaf_class(b);
        iid[r] = class_combine(op, ia, ib);
      }
      // Pass 3, branch-free per token: an operator pushes the state its right child inherits, a leaf records
      // (symbol, state) and pops the state of the next pending right child.
      sp = 0;
      uint32_t cur = 0;
      for (int i = 0; i < n; ++i) {
        const uint32_t t = tok[i];
        const bool is_op = t >= 0x80;
        const uint32_t child = path_child_code(cur, t & 1);
        const uint32_t popped = st[max(sp - 1, 0)];
        lsym[n_leaf] = (uint8_t)t;       // overwritten by the next leaf when this token is an operator
        lcode[n_leaf] = (uint8_t)cur;
        n_leaf += is_op ? 0 : 1;
        if (is_op) st[sp] = (uint8_t)child;
        cur = is_op ? child : (sp > 0 ? popped : cur);
        sp += is_op ? 1 : (sp > 0 ? -1 : 0);
      }
    }
  }
  rec[kRecMeta + 0] = (uint8_t)n;
  rec[kRecMeta + 1] = (uint8_t)n_leaf;
  rec[kRecMeta + 2] = (uint8_t)n_int;
  rec[kRecMeta + 3] = (uint8_t)status;
  return status;
}

// ---- evaluated-set blob layout (all offsets 16-byte aligned) ---------------------------
// DIM_WISE bookkeeping written by eval_build_kernel.  The columns of the dim-wise vector are kept in
// BLOCK-MAJOR order ("permuted position" j): block 0's columns (ascending reference column), block 1's, ...;
// columns that belong to no block are identically zero for every tree and are dropped.  Inside a block the
// evaluated trees take only a few distinct vectors ("states": a block is NULL or a normalised count vector
// of its kinds); when every block has <= kMaxStates of them the pair kernel replaces the per-column L1 by
// one table lookup per (pair, block).
constexpr int kMaxStates = 32;
struct DimMeta {
  int32_t use_states;                    // 1: every block has <= kMaxStates distinct vectors
  int32_t n_pcols;                       // number of permuted columns
  uint8_t bbeg[BOSOT_MAX_BLOCKS + 4];    // block b = permuted positions bbeg[b] .. bbeg[b+1]-1
  int8_t sym_pos[BOSOT_MAX_SYMBOLS];     // permuted position of a symbol's column, -1 = none
  int8_t null_pos[BOSOT_MAX_BLOCKS];     // permuted position of NULL_b
  uint8_t n_states[BOSOT_MAX_BLOCKS];    // distinct vectors per block (saturates at kMaxStates + 1)
  int32_t n_flat;                        // entries of the flat (block, state) list, see off_flat
};
static_assert(sizeof(DimMeta) == 128, "DimMeta is staged with one 128-byte bulk copy");

struct EvalLayout {
  int n_eval;     // E
  int e_pad;      // E rounded up to 64 (lane l of a warp owns trees l + 32 r; two of them share a packed word);
                  // beyond 256 evaluated trees rounded up to 256 (whole register tiles)
  int hash_cap;   // power of two >= 2 * 31 * E (load factor <= 0.5)
  int n_slots;    // hash_cap + 64 + 64*64 : unified slot space (hash | leaf symbol | path id)
  int max_post;   // 63 * E
  size_t off_keys;     // uint64 [hash_cap]
  size_t off_range;    // uint2  [n_slots]   (begin, length) into postings
  size_t off_post;     // uint32 [max_post]  e | count << 16 | total << 24
  size_t off_nnodes;   // uint8  [e_pad]     SUBTREES total (nodes)
  size_t off_nleaves;  // uint8  [e_pad]     PATHS total (leaves)
  size_t off_rsub;     // double [e_pad]     1 / nodes
  size_t off_rpath;    // double [e_pad]     1 / leaves
  size_t off_symT2;    // uint32 [BOSOT_MAX_SYMBOLS][e_pad/2]  leaf count per symbol, transposed and packed:
                       //        word 32 q + l = count(tree 64 q + l) | count(tree 64 q + 32 + l) << 16
  size_t off_nn2;      // uint32 [e_pad/2]   node totals, packed the same way
  size_t off_meta;     // DimMeta
  size_t off_stateP;   // double [BOSOT_MAX_DIM_COLS][kMaxStates]  value of permuted column j in state s of its block
  size_t off_sid8T;    // uint8  [BOSOT_MAX_BLOCKS][e_pad]  8 * (state of tree e in block b)
  size_t off_flat;     // uint32 [BOSOT_MAX_BLOCKS * kMaxStates]  all (block, state) pairs in block order:
                       //        first column | (last column + 1) << 8 | (block * kMaxStates + state) << 16
  size_t off_dimT;     // double [BOSOT_MAX_DIM_COLS][e_pad]  transposed dim-wise matrix, permuted column order
  size_t off_cursor;   // uint32 [n_slots]   build scratch
  size_t total;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline EvalLayout eval_layout(int n_eval) {
  EvalLayout L;
  L.n_eval = n_eval;
  L.e_pad = n_eval <= 256 ? ((n_eval + 63) & ~63) : ((n_eval + 255) & ~255);
  int cap = 64;
  while (cap < 2 * kMaxInternal * n_eval) cap <<= 1;
  L.hash_cap = cap;
  L.n_slots = cap + BOSOT_MAX_SYMBOLS + BOSOT_MAX_SYMBOLS * kPathCodes;
  L.max_post = (kMaxInternal + kMaxLeaves) * n_eval;  // operator classes + reduced paths (leaf classes are dense)
  size_t o = 64;  // header
  L.off_keys = o;      o = align16(o + sizeof(uint64_t) * (size_t)cap);
  L.off_range = o;     o = align16(o + sizeof(uint2) * (size_t)L.n_slots);
  L.off_post = o;      o = align16(o + sizeof(uint32_t) * (size_t)L.max_post);
  L.off_nnodes = o;    o = align16(o + (size_t)L.e_pad);
  L.off_nleaves = o;   o = align16(o + (size_t)L.e_pad);
  L.off_rsub = o;      o = align16(o + sizeof(double) * (size_t)L.e_pad);
  L.off_rpath = o;     o = align16(o + sizeof(double) * (size_t)L.e_pad);
  L.off_symT2 = o;     o = align16(o + sizeof(uint32_t) * (size_t)BOSOT_MAX_SYMBOLS * (L.e_pad / 2));
  L.off_nn2 = o;       o = align16(o + sizeof(uint32_t) * (size_t)(L.e_pad / 2));
  L.off_meta = o;      o = align16(o + sizeof(DimMeta));
  L.off_stateP = o;    o = align16(o + sizeof(double) * (size_t)BOSOT_MAX_DIM_COLS * kMaxStates);
  L.off_sid8T = o;     o = align16(o + (size_t)BOSOT_MAX_BLOCKS * L.e_pad);
  L.off_flat = o;      o = align16(o + sizeof(uint32_t) * (size_t)BOSOT_MAX_BLOCKS * kMaxStates);
  L.off_dimT = o;      o = align16(o + sizeof(double) * (size_t)BOSOT_MAX_DIM_COLS * L.e_pad);
  L.off_cursor = o;    o = align16(o + sizeof(uint32_t) * (size_t)L.n_slots);
  L.total = o;
  return L;
}

__device__ __forceinline__ uint32_t hash_slot(uint64_t id, int cap) {
  return (uint32_t)(id ^ (id >> 32)) & (uint32_t)(cap - 1);
}

__device__ __forceinline__ uint32_t pack_posting(uint32_t e, uint32_t cnt, uint32_t total) {
  return e | (cnt << 16) | (total << 24);
}

}  // namespace bosot
