"""Flat tree encoding: the boundary between kernel-grammar expression objects and device memory.

One tree = one 64-byte record (include/bosot_b200.h): byte 0 = node count n, bytes 1..n =
PRE-ORDER tokens (leaf: symbol index, 0x80 ADD, 0x81 MULTIPLY), 0xFF padding.  Pre-order is the
order in which the reference enumerates subtrees (bosot/kernels/kernel_grammar/kernel_grammar.py:469),
so "first seen" in the reference = smallest token position here.

The symbol index of a leaf is the position of its base kernel in the search space's
``base_kernel_config_list`` (bosot/kernels/kernel_grammar/kernel_grammar_search_spaces.py:176-227:
kind-major, dimension-minor); leaf identity in the reference is the kernel ``name`` string
``<ConfigName>_on_<dim>`` (bosot/kernels/base_elementary_kernel.py:34-35).

``Grammar`` carries what ``DimWiseWeightedDistanceExtractor.create_dim_class_dict``
(bosot/kernels/kernel_grammar/optimal_transport_mappings.py:45-86) derives from
``generator_name`` / ``input_dimension``; unknown generators raise ``NotImplementedError`` as there.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

from . import _lib

RBF_NAME = "RBFWithPrior"
LIN_NAME = "LinearWithPrior"
PER_NAME = "PeriodicKernelWithPrior"
RQ_NAME = "RQWithPrior"

# generator name -> (kinds in search-space order, kinds in dim-wise feature order, per-dimension?)
_SPACES = {
    "CompositionalKernelSearchSpace": ([RBF_NAME, LIN_NAME, PER_NAME], [RBF_NAME, PER_NAME, LIN_NAME], True),
    "CKSWithRQSearchSpace": ([RBF_NAME, LIN_NAME, PER_NAME, RQ_NAME], [RBF_NAME, PER_NAME, LIN_NAME, RQ_NAME], True),
    "CKSHighDimSearchSpace": ([RBF_NAME, RQ_NAME], [RBF_NAME, RQ_NAME], True),
    "NDimFullKernelsSearchSpace": ([RBF_NAME, LIN_NAME, PER_NAME], [RBF_NAME, PER_NAME, LIN_NAME], False),
}


class Grammar:
    """Symbol table of one search space."""

    _cache: Dict[Tuple[str, int], "Grammar"] = {}

    def __init__(self, generator_name: str, input_dimension: int):
        if generator_name not in _SPACES:
            raise NotImplementedError(generator_name)
        space_kinds, dim_kinds, per_dim = _SPACES[generator_name]
        self.generator_name = generator_name
        self.input_dimension = int(input_dimension)
        if per_dim:
            self.symbols: List[str] = [k + "_on_" + str(i) for k in space_kinds for i in range(self.input_dimension)]
            self.n_blocks = self.input_dimension
        else:
            self.symbols = list(space_kinds)
            self.n_blocks = 1
        self.n_symbols = len(self.symbols)
        self.symbol_index: Dict[str, int] = {s: i for i, s in enumerate(self.symbols)}
        width = 1 + len(dim_kinds)
        self.n_dim_cols = self.n_blocks * width
        if self.n_symbols > _lib.MAX_SYMBOLS or self.n_blocks > _lib.MAX_BLOCKS or self.n_dim_cols > _lib.MAX_DIM_COLS:
            raise ValueError("search space too large for the flat encoding: %s(%d)" % (generator_name, input_dimension))
        # dim-wise column names in the reference's dict order: NULL_i, then the kinds of dimension i
        self.dim_col_names: List[str] = []
        self.sym_block = [-1] * self.n_symbols
        self.sym_col = [-1] * self.n_symbols
        self.block_null_col = []
        for b in range(self.n_blocks):
            self.block_null_col.append(len(self.dim_col_names))
            self.dim_col_names.append("NULL_" + str(b))
            for k in dim_kinds:
                name = k + "_on_" + str(b) if per_dim else k
                if name in self.symbol_index:
                    s = self.symbol_index[name]
                    self.sym_block[s] = b
                    self.sym_col[s] = len(self.dim_col_names)
                self.dim_col_names.append(name)
        g = _lib.BosotGrammar()
        g.n_symbols, g.n_blocks, g.n_dim_cols, g.reserved = self.n_symbols, self.n_blocks, self.n_dim_cols, 0
        for s in range(_lib.MAX_SYMBOLS):
            g.sym_block[s] = self.sym_block[s] if s < self.n_symbols else -1
            g.sym_col[s] = self.sym_col[s] if s < self.n_symbols else -1
        for b in range(_lib.MAX_BLOCKS):
            g.block_null_col[b] = self.block_null_col[b] if b < self.n_blocks else 0
        self.c_struct = g

    @classmethod
    def get(cls, generator_name: str, input_dimension: int) -> "Grammar":
        key = (generator_name, int(input_dimension))
        if key not in cls._cache:
            cls._cache[key] = cls(generator_name, input_dimension)
        return cls._cache[key]

    @classmethod
    def for_expressions(cls, X: Sequence[object]) -> "Grammar":
        """Like the reference, the search space is read off the first expression
        (kernel_kernel_grammar_tree.py:171-172)."""
        return cls.get(X[0].get_generator_name(), X[0].get_input_dimension())

    def __repr__(self):
        return "Grammar(%s, d=%d, B=%d)" % (self.generator_name, self.input_dimension, self.n_symbols)


# ------------------------------------------------------------------ encoders / decoders
def _tokens_of_expression(expr, symbol_index: Dict[str, int], out: List[int]) -> None:
    """Pre-order tokens of a (reference-shaped) expression object.  Duck-typed: operator nodes
    expose ``expression1`` / ``expression2`` / ``operator`` (kernel_grammar.py:338-342), leaves
    expose ``get_name()`` (kernel_grammar.py:259-260)."""
    stack = [expr]
    while stack:
        e = stack.pop()
        op = getattr(e, "operator", None)
        if op is None:
            out.append(symbol_index[e.get_name()])
        else:
            out.append(_lib.OP_ADD if op.name == "ADD" else _lib.OP_MUL)
            stack.append(e.expression2)
            stack.append(e.expression1)


def _tokens_of_tuple(tree, symbol_index: Dict[str, int], out: List[int]) -> None:
    stack = [tree]
    while stack:
        t = stack.pop()
        if isinstance(t, str):
            out.append(symbol_index[t])
        else:
            out.append(_lib.OP_ADD if t[0] == "ADD" else _lib.OP_MUL)
            stack.append(t[2])
            stack.append(t[1])


def _tokens_of_name(name: str, symbol_index: Dict[str, int], out: List[int]) -> None:
    """Tokens straight from the canonical name string ``str(expr)`` (kernel_grammar.py:379-383).
    In "(L OP R)" the operator of a node is only known after L: reserve its slot first."""
    slots: List[int] = []
    i, n = 0, len(name)
    while i < n:
        ch = name[i]
        if ch == "(":
            slots.append(len(out))
            out.append(-1)
            i += 1
        elif ch == ")":
            slots.pop()
            i += 1
        elif ch == " ":
            if name.startswith(" ADD ", i):
                out[slots[-1]] = _lib.OP_ADD
                i += 5
            elif name.startswith(" MULTIPLY ", i):
                out[slots[-1]] = _lib.OP_MUL
                i += 10
            else:
                raise ValueError("bad operator at %d in %r" % (i, name))
        else:
            j = i
            while j < n and name[j] not in " )":
                j += 1
            out.append(symbol_index[name[i:j]])
            i = j
    if slots or any(t < 0 for t in out):
        raise ValueError("unbalanced expression name %r" % name)


def encode(items: Iterable[object], grammar: Grammar) -> np.ndarray:
    """Encode expression objects, nested tuples ``(op, left, right)`` / leaf-name strings that
    contain no parenthesis, or canonical name strings into a uint8 ``[N, 64]`` array."""
    items = list(items)
    out = np.full((len(items), _lib.TREE_BYTES), _lib.PAD, dtype=np.uint8)
    sym = grammar.symbol_index
    toks: List[int] = []
    for r, it in enumerate(items):
        toks.clear()
        try:
            if isinstance(it, str):
                if it.startswith("("):
                    _tokens_of_name(it, sym, toks)
                else:
                    toks.append(sym[it])
            elif isinstance(it, tuple):
                _tokens_of_tuple(it, sym, toks)
            else:
                _tokens_of_expression(it, sym, toks)
        except KeyError as exc:
            raise ValueError("leaf %s is not a base kernel of %r" % (exc, grammar)) from None
        if len(toks) > _lib.MAX_NODES:
            raise ValueError("expression with %d nodes exceeds the %d-node flat record" % (len(toks), _lib.MAX_NODES))
        out[r, 0] = len(toks)
        out[r, 1 : 1 + len(toks)] = toks
    return out


def decode_to_tuple(record: np.ndarray, grammar: Grammar):
    """One flat record -> nested tuple tree (inverse of ``encode``)."""
    n = int(record[0])
    pos = 1

    def rec():
        nonlocal pos
        t = int(record[pos])
        pos += 1
        if t < 0x80:
            return grammar.symbols[t]
        left = rec()
        right = rec()
        return ("ADD" if t == _lib.OP_ADD else "MULTIPLY", left, right)

    tree = rec()
    if pos != n + 1:
        raise ValueError("malformed record")
    return tree


def decode_to_name(record: np.ndarray, grammar: Grammar) -> str:
    """One flat record -> canonical name string (kernel_grammar.py:379-383)."""

    def name(t):
        if isinstance(t, str):
            return t
        return "(" + name(t[1]) + " " + t[0] + " " + name(t[2]) + ")"

    return name(decode_to_tuple(record, grammar))


def validate(trees: np.ndarray, grammar: Grammar) -> None:
    """Host-side structural check of flat records (vectorised): odd length in range, tokens in
    range, and the pre-order 'pending subtrees' counter reaches 0 exactly at the last token."""
    if trees.dtype != np.uint8 or trees.ndim != 2 or trees.shape[1] != _lib.TREE_BYTES:
        raise ValueError("flat trees must be uint8 [N, 64]")
    n = trees[:, 0].astype(np.int64)
    if np.any((n < 1) | (n > _lib.MAX_NODES) | (n % 2 == 0)):
        raise ValueError("flat tree with bad node count")
    tok = trees[:, 1:].astype(np.int64)
    pos = np.arange(tok.shape[1])[None, :]
    live = pos < n[:, None]
    is_op = (tok == _lib.OP_ADD) | (tok == _lib.OP_MUL)
    is_leaf = tok < grammar.n_symbols
    if np.any(live & ~(is_op | is_leaf)):
        raise ValueError("flat tree with bad token")
    pending = 1 + np.cumsum(np.where(live, np.where(is_op, 1, -1), 0), axis=1)
    bad_mid = np.any(live & (pos < (n[:, None] - 1)) & (pending <= 0), axis=1)
    last = pending[np.arange(len(n)), n - 1]
    if np.any(bad_mid | (last != 0)):
        raise ValueError("flat tree is not a well-formed pre-order expression")


def random_stress_trees(n: int, grammar: Grammar, seed: int, max_depth: int = 5, p_leaf: float = 0.3) -> np.ndarray:
    """C5 stress trees generated directly in the flat encoding: a node at depth k
    is a leaf if k == max_depth or (k > 0 and U < p_leaf); operator and leaf symbol uniform.
    Vectorised level by level over all trees: pre-order position of the children is resolved
    with a per-tree explicit stack simulated by arrays."""
    rng = np.random.default_rng(seed)
    out = np.full((n, _lib.TREE_BYTES), _lib.PAD, dtype=np.uint8)
    # depth-first generation, all trees in lock step: stack of pending depths per tree
    stack = np.zeros((n, max_depth + 2), dtype=np.int8)
    sp = np.ones(n, dtype=np.int64)  # root pending at depth 0
    length = np.zeros(n, dtype=np.int64)
    active = np.arange(n)
    while active.size:
        d = stack[active, sp[active] - 1].astype(np.int64)
        sp[active] -= 1
        u = rng.random(active.size)
        leaf = (d == max_depth) | ((d > 0) & (u < p_leaf))
        sym = rng.integers(grammar.n_symbols, size=active.size)
        op = rng.integers(2, size=active.size)
        tok = np.where(leaf, sym, _lib.OP_ADD + op).astype(np.uint8)
        out[active, 1 + length[active]] = tok
        length[active] += 1
        grow = active[~leaf]
        dn = d[~leaf] + 1
        stack[grow, sp[grow]] = dn
        stack[grow, sp[grow] + 1] = dn
        sp[grow] += 2
        active = active[sp[active] > 0]
    out[:, 0] = length
    return out
