"""Candidate production and the evolutionary acquisition optimiser directly on flat tree records
(SURVEY.md 8f-2): the same moves, enumeration order and ``np.random`` consumption as the object API
(``kernels/kernel_grammar/kernel_grammar_search_spaces.py`` / ``kernel_grammar_candidate_generator.py`` /
``bayesian_optimization/evolutionary_optimizer_objects.py`` of the reference), without a Python object per
expression: the population lives on the GPU as a uint8 ``[N, 64]`` tensor, offspring are produced by the
``bosot_neighbours`` kernel, acquisition values by ``AcquisitionEngine.acquire_device``.  Neighbour numbers
are drawn on the host with ``np.random`` exactly like the reference, so a seeded run visits the same kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .encoding import Grammar, decode_to_name


def num_neighbours(trees: np.ndarray, n_symbols: int) -> np.ndarray:
    """L*B + n*2*B per record (kernel_grammar_search_spaces.py:97-106), from the node-count byte alone."""
    n = trees[:, 0].astype(np.int64)
    return ((n + 1) // 2) * n_symbols + n * 2 * n_symbols


def neighbour_at_index_host(record: np.ndarray, k: int, n_symbols: int) -> np.ndarray:
    """Host restatement of ``get_neighbour_at_index`` on one flat record (used where the generation is
    inherently sequential, and as the checker of the device kernel)."""
    n = int(record[0])
    toks = [int(t) for t in record[1 : 1 + n]]
    L, B = (n + 1) // 2, n_symbols
    if not (0 <= k < L * B + n * 2 * B):
        raise IndexError("neighbour index out of range")
    if k < L * B:
        which, sym, seen = k // B, k % B, 0
        for p, t in enumerate(toks):
            if t < 0x80:
                if seen == which:
                    toks[p] = sym
                seen += 1
    else:
        kk = k - L * B
        node, op, sym = kk // (2 * B), (kk % (2 * B)) // B, kk % B
        pending, end = 1, node
        for p in range(node, n):
            pending += 1 if toks[p] >= 0x80 else -1
            if pending == 0:
                end = p
                break
        toks = toks[:node] + [_lib.OP_ADD + op] + toks[node : end + 1] + [sym] + toks[end + 1 :]
        if len(toks) > _lib.MAX_NODES:
            raise ValueError("neighbour exceeds the %d-node flat record" % _lib.MAX_NODES)
    out = np.full(_lib.TREE_BYTES, _lib.PAD, dtype=np.uint8)
    out[0] = len(toks)
    out[1 : 1 + len(toks)] = toks
    return out


def neighbours_device(trees: torch.Tensor, index: torch.Tensor, grammar: Grammar, check: bool = True, status_out: Optional[list] = None) -> torch.Tensor:
    """``out[i]`` = neighbour number ``index[i]`` of ``trees[i]`` (C-ABI ``bosot_neighbours``).  ``check`` reads the per-row
    status back (one synchronisation); with ``check=False`` and a ``status_out`` list the status tensor is appended to it
    for the caller to inspect at its next synchronisation point (``raise_on_neighbour_status``)."""
    if not trees.is_cuda:
        raise RuntimeError("bosot_b200: neighbours_device needs CUDA tensors (no CPU implementation on the product path)")
    n = trees.shape[0]
    out = torch.empty_like(trees)
    status = torch.empty(n, dtype=torch.int32, device=trees.device)
    index = index.to(trees.device, torch.int64).contiguous()
    with torch.cuda.device(trees.device):
        rc = _lib.load().bosot_neighbours(trees.data_ptr(), n, index.data_ptr(), grammar.n_symbols, out.data_ptr(), status.data_ptr(),
                                          torch.cuda.current_stream(trees.device).cuda_stream)
    _lib.check(rc, "neighbours")
    if check:
        raise_on_neighbour_status([status])
    elif status_out is not None:
        status_out.append(status)
    return out


def raise_on_neighbour_status(statuses) -> None:
    """Raise if any neighbour move of the collected launches failed (a move that would exceed the 63-node flat record)."""
    for status in statuses:
        bad = torch.nonzero(status)
        if bad.numel():
            i = int(bad[0])
            raise ValueError("bosot_b200: neighbour move failed at row %d (code %d)" % (i, int(status[i])))


def _global_bitgen():
    """(next_uint32 function pointer, state pointer) of NumPy's GLOBAL generator through the bit generator's ctypes
    interface, or None when this NumPy cannot hand it out (then the block replay below is used)."""
    get = getattr(np.random, "get_bit_generator", None)
    if get is None:
        return None
    try:
        iface = get().ctypes
        return C.cast(iface.next_uint32, C.c_void_p), iface.state
    except Exception:
        return None


def _replay_global_rng(call, expected_outputs: int):
    """Run ``call(raw_ptr, n_raw, consumed_ref) -> rc`` on a block of raw 32-bit outputs of NumPy's GLOBAL generator and
    leave the generator exactly where the reference's sequence of ``np.random.choice`` calls would have left it:
    the state is saved, a block of outputs is drawn, the C side reports how many it used, the state is restored and
    advanced by that many outputs.  The block grows until it is long enough (BOSOT_ERAW)."""
    # a block only a little longer than what is expected to be used: drawing (and re-drawing, below) raw outputs is
    # the larger part of the host time of these producers (0.62 of 1.7 ms per 10k-candidate population with a block
    # four times too long); a block that turns out too short costs one retry
    n_raw = max(64, int(1.2 * expected_outputs) + 64)
    while True:
        state = np.random.get_state()
        raw = np.random.randint(0, 2**32, size=n_raw, dtype=np.uint32)
        consumed = C.c_int64(0)
        rc = call(raw.ctypes.data, n_raw, C.byref(consumed))
        np.random.set_state(state)
        if rc == _lib.ERAW:
            n_raw *= 4
            continue
        _lib.check(rc, "host producer")
        if consumed.value:
            np.random.randint(0, 2**32, size=consumed.value, dtype=np.uint32)
        return


def legacy_choices(sizes: np.ndarray) -> np.ndarray:
    """``[np.random.choice(n) for n in sizes]`` -- same values, same final generator state -- in one native call
    (C-ABI ``bosot_host_choices``)."""
    sizes = np.ascontiguousarray(sizes, dtype=np.int64)
    out = np.empty(len(sizes), dtype=np.int64)
    if len(sizes) == 0:
        return out
    lib = _lib.load()
    gen = _global_bitgen()
    if gen is not None:  # the draws come straight out of the global generator, which ends up where the reference's would
        used = C.c_int64(0)
        _lib.check(lib.bosot_host_choices_cb(gen[0], gen[1], sizes.ctypes.data, len(sizes), out.ctypes.data, C.byref(used)), "host_choices")
        return out
    # RandomState.choice(n) is a masked rejection on raw outputs: between 1 and 2 draws per choice on average
    _replay_global_rng(lambda raw, n_raw, used: lib.bosot_host_choices(raw, n_raw, sizes.ctypes.data, len(sizes), out.ctypes.data, used),
                       1.5 * len(sizes))
    return out


class FlatCandidateGenerator:
    """Flat-record counterpart of ``KernelGrammarCandidateGenerator`` (reference
    kernel_grammar_candidate_generator.py:27-117) for the two generators the EA uses."""

    def __init__(self, grammar: Grammar):
        self.grammar = grammar
        B = grammar.n_symbols
        self.roots = np.full((B, _lib.TREE_BYTES), _lib.PAD, dtype=np.uint8)
        self.roots[:, 0] = 1
        self.roots[:, 1] = np.arange(B)

    def get_dataset_recursivly_generated(self, n_data: int, n_per_step: int = 1) -> np.ndarray:
        """Roots first, then repeatedly a random neighbour of a randomly chosen earlier element
        (:98-111).  Sequential by construction, so it runs on the host, in native code (C-ABI
        ``bosot_host_recursive_dataset``) that replays the reference's ``np.random.choice`` draws from raw outputs of
        the global generator: the same trees from the same seed, the generator left in the same state."""
        B = self.grammar.n_symbols
        pool = np.empty((max(int(n_data), B), _lib.TREE_BYTES), dtype=np.uint8)
        lib = _lib.load()
        gen = _global_bitgen()
        if gen is not None:
            used = C.c_int64(0)
            _lib.check(lib.bosot_host_recursive_dataset_cb(gen[0], gen[1], B, int(n_data), int(n_per_step), pool.ctypes.data, C.byref(used)),
                       "host_recursive_dataset")
            return pool
        _replay_global_rng(lambda raw, n_raw, used: lib.bosot_host_recursive_dataset(raw, n_raw, B, int(n_data), int(n_per_step),
                                                                                      pool.ctypes.data, used), 3.1 * int(n_data))
        return pool

    def get_initial_for_evolutionary_opt(self, n_initial: int) -> np.ndarray:
        return self.get_dataset_recursivly_generated(n_initial, 1)

    def draw_offspring_indices(self, survivors: np.ndarray, n_around: int) -> Tuple[np.ndarray, np.ndarray]:
        """(parent row, neighbour number) for ``n_around`` random neighbours of every survivor, drawn in the
        reference's order (survivor-major, :85-93)."""
        counts = num_neighbours(survivors, self.grammar.n_symbols)
        parent = np.repeat(np.arange(len(survivors)), n_around)
        index = legacy_choices(counts[parent])  # one np.random.choice(count) per offspring in the reference
        return parent, index


class FlatEvolutionaryOptimizer:
    """``EvolutionaryOptimizerObjects`` (reference evolutionary_optimizer_objects.py:24-71) on flat records:
    select the top ``pop / (offspring + 1)`` by acquisition value, every survivor spawns ``offspring`` random
    neighbours, new population = offspring + survivors; the best individual ever seen is returned."""

    def __init__(self, population_size: int, num_offspring: int, generator: FlatCandidateGenerator, device: Optional[torch.device] = None):
        self.population_size = population_size
        self.num_offspring = num_offspring
        self.generator = generator
        self.survival_rate = 1 / (num_offspring + 1)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.current_maximizer: Optional[np.ndarray] = None
        self.current_maximizer_value = -np.inf

    def maximize(self, func: Callable[[torch.Tensor], torch.Tensor], steps: int):
        """``func`` maps a device tensor of flat records to a device tensor of acquisition values.

        Selection runs on the device (``sot.topk`` = the tail of the stable argsort, ``sot.argmax_first`` = np.argmax):
        per generation only the survivors' node-count bytes and 16 bytes of (best value, best index) go to the host,
        and the drawn (parent, neighbour number) pairs come back -- the acquisition vector never crosses PCIe."""
        from . import sot

        gen = self.generator
        population = torch.from_numpy(gen.get_initial_for_evolutionary_opt(self.population_size)).to(self.device)
        values = None
        statuses = []  # per-row status of every neighbour launch, inspected once at the end (no extra synchronisation per generation)
        best_pinned = torch.zeros(2, dtype=torch.float64).pin_memory()
        for step in range(steps):
            if step > 0:
                n_survive = int(self.population_size * self.survival_rate)
                if n_survive > 0:
                    surv_idx, _ = sot.topk(values, n_survive, want_values=False)  # ascending by (value, index), like the stable argsort
                    survivors = population[surv_idx]
                else:
                    survivors = population[:0]
                # the host draws the neighbour numbers (np.random replay); it needs the node count of every survivor only
                parent, index = gen.draw_offspring_indices(survivors[:, :1].cpu().numpy(), self.num_offspring)
                offspring = neighbours_device(survivors[torch.from_numpy(parent).to(self.device)].contiguous(),
                                              torch.from_numpy(index), gen.grammar, check=False, status_out=statuses)
                population = torch.cat([offspring, survivors])
            values = func(population).contiguous()
            sot.argmax_first(values, out=best_pinned)  # 16 bytes written by the kernel into pinned host memory
            torch.cuda.current_stream(self.device).synchronize()
            best = best_pinned.numpy()
            fittest = int(best[1])
            if self.current_maximizer_value < best[0]:
                self.current_maximizer_value = float(best[0])
                self.current_maximizer = population[fittest].cpu().numpy()
        raise_on_neighbour_status(statuses)
        return self.current_maximizer, self.current_maximizer_value

    def best_name(self) -> str:
        return decode_to_name(self.current_maximizer, self.generator.grammar)


def record_to_expression(record: np.ndarray, search_space):
    """Flat record -> expression objects of the mirror API (leaves built through the search space's own
    base-kernel configs, ``generator_name`` set like the reference's neighbour constructors do)."""
    from .encoding import decode_to_tuple
    from .kernels.kernel_factory import KernelFactory
    from .kernels.kernel_grammar.kernel_grammar import ElementaryKernelGrammarExpression, KernelGrammarExpression, KernelGrammarOperator

    grammar = Grammar.get(search_space.name, search_space.input_dimension)

    def build(t):
        if isinstance(t, str):
            e = ElementaryKernelGrammarExpression(KernelFactory.build(search_space.base_kernel_config_list[grammar.symbol_index[t]]))
        else:
            e = KernelGrammarExpression(build(t[1]), build(t[2]), KernelGrammarOperator[t[0]])
        e.set_generator_name(search_space.name)
        return e

    return build(decode_to_tuple(record, grammar))
