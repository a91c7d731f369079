"""Candidate production on the flat encoding: host restatement vs the object API (CPU),
device kernel vs host restatement and the flat EA vs the object EA (GPU)."""
import numpy as np
import pytest

from bosot_b200.configs.kernels.kernel_grammar_generators.generator_configs import CKSHighDimGeneratorConfig, CKSWithRQGeneratorConfig
from bosot_b200.encoding import Grammar, decode_to_name, encode, random_stress_trees
from bosot_b200.flat_search import FlatCandidateGenerator, neighbour_at_index_host, num_neighbours
from bosot_b200.kernels.kernel_grammar.generator_factory import GeneratorFactory


@pytest.mark.parametrize("cfg,d", [(CKSWithRQGeneratorConfig, 2), (CKSHighDimGeneratorConfig, 3)])
def test_host_moves_match_object_api(cfg, d):
    gen = GeneratorFactory.build(cfg(input_dimension=d))
    ss = gen.search_space
    gram = Grammar.get(ss.name, d)
    np.random.seed(9)
    exprs = gen.get_dataset_recursivly_generated(40, 1)[::3]
    flat = encode(exprs, gram)
    counts = num_neighbours(flat, gram.n_symbols)
    for e, rec, cnt in zip(exprs, flat, counts):
        assert cnt == ss.get_num_neighbour_expressions(e)
        for k in list(range(0, cnt, max(1, cnt // 17))) + [cnt - 1]:
            assert decode_to_name(neighbour_at_index_host(rec, k, gram.n_symbols), gram) == str(ss.get_neighbour_at_index(e, k))
    with pytest.raises(IndexError):
        neighbour_at_index_host(flat[0], int(counts[0]), gram.n_symbols)


def test_flat_generator_consumes_rng_like_object_generator():
    gen = GeneratorFactory.build(CKSWithRQGeneratorConfig(input_dimension=2))
    gram = Grammar.get("CKSWithRQSearchSpace", 2)
    np.random.seed(3)
    names = [str(x) for x in gen.get_dataset_recursivly_generated(100, 1)]
    after_obj = np.random.random()
    np.random.seed(3)
    flat = FlatCandidateGenerator(gram).get_dataset_recursivly_generated(100, 1)
    after_flat = np.random.random()
    assert [decode_to_name(r, gram) for r in flat] == names and after_obj == after_flat


def _python_recursive_dataset(grammar, n_data, n_per_step=1):
    """The reference's loop (kernel_grammar_candidate_generator.py:98-111) on flat records, one np.random.choice per draw."""
    B = grammar.n_symbols
    roots = np.full((B, 64), 0xFF, dtype=np.uint8)
    roots[:, 0] = 1
    roots[:, 1] = np.arange(B)
    pool = [r for r in roots]
    while len(pool) < n_data:
        chosen = pool[int(np.random.choice(len(pool)))]
        for _ in range(n_per_step):
            if len(pool) < n_data:
                k = int(np.random.choice(int(num_neighbours(chosen[None, :], B)[0])))
                pool.append(neighbour_at_index_host(chosen, k, B))
    return np.stack(pool[: max(n_data, B)])


@pytest.mark.parametrize("gen,d", [("CKSHighDimSearchSpace", 5), ("CKSWithRQSearchSpace", 2), ("CKSWithRQSearchSpace", 1)])
def test_native_producers_replay_numpy_choice(gen, d):
    """C-ABI bosot_host_recursive_dataset / bosot_host_choices draw what one np.random.choice call per draw would have
    drawn -- same records, same indices -- and leave NumPy's global generator in the same state."""
    from bosot_b200.flat_search import legacy_choices

    gram = Grammar.get(gen, d)
    g = FlatCandidateGenerator(gram)
    for n_data, per in ((1, 1), (gram.n_symbols, 1), (gram.n_symbols + 1, 1), (400, 1), (401, 3), (2500, 1)):
        np.random.seed(100 + n_data)
        ref = _python_recursive_dataset(gram, n_data, per)
        ref_next = np.random.randint(0, 2**32, dtype=np.uint32)
        np.random.seed(100 + n_data)
        got = g.get_dataset_recursivly_generated(n_data, per)
        assert np.array_equal(got, ref) and np.random.randint(0, 2**32, dtype=np.uint32) == ref_next
    sizes = np.concatenate([np.ones(5, dtype=np.int64), np.random.default_rng(0).integers(1, 4000, 3000), [2**31, 2**32 - 1, 2**32]])
    np.random.seed(3)
    ref = np.array([int(np.random.choice(int(n))) for n in sizes])
    ref_next = np.random.randint(0, 2**32, dtype=np.uint32)
    np.random.seed(3)
    got = legacy_choices(sizes)
    assert np.array_equal(got, ref) and np.random.randint(0, 2**32, dtype=np.uint32) == ref_next
    # offspring draws of the flat EA: survivor-major, one draw per offspring
    surv = _python_recursive_dataset(gram, 300)
    np.random.seed(4)
    cnt = num_neighbours(surv, gram.n_symbols)
    ref = np.array([int(np.random.choice(int(cnt[p]))) for p in np.repeat(np.arange(300), 4)])
    np.random.seed(4)
    parent, index = g.draw_offspring_indices(surv, 4)
    assert np.array_equal(index, ref) and np.array_equal(parent, np.repeat(np.arange(300), 4))


def test_native_producers_error_behaviour():
    """Return codes of the host entry points: BOSOT_ERAW when the block of raw outputs is too short (the wrapper then
    retries with a longer block and still leaves the generator where the reference would), BOSOT_EINVAL on bad sizes,
    BOSOT_ETREE when a move would exceed the 63-node record."""
    import ctypes as C

    from bosot_b200 import _lib
    from bosot_b200.flat_search import _replay_global_rng

    lib = _lib.load()
    raw = np.arange(1, dtype=np.uint32)
    sizes = np.full(8, 1000, dtype=np.int64)
    out = np.empty(8, dtype=np.int64)
    used = C.c_int64(0)
    assert lib.bosot_host_choices(raw.ctypes.data, 1, sizes.ctypes.data, 8, out.ctypes.data, C.byref(used)) == _lib.ERAW
    bad = np.zeros(1, dtype=np.int64)
    assert lib.bosot_host_choices(raw.ctypes.data, 1, bad.ctypes.data, 1, out.ctypes.data, C.byref(used)) == -1  # BOSOT_EINVAL
    # the wrapper grows the block: start far too small on purpose
    sizes = np.random.default_rng(1).integers(2, 5000, 4000)
    np.random.seed(8)
    ref = np.array([int(np.random.choice(int(n))) for n in sizes])
    ref_next = np.random.randint(0, 2**32, dtype=np.uint32)
    np.random.seed(8)
    got = np.empty(len(sizes), dtype=np.int64)
    sz = np.ascontiguousarray(sizes, dtype=np.int64)
    _replay_global_rng(lambda r, n, u: lib.bosot_host_choices(r, n, sz.ctypes.data, len(sz), got.ctypes.data, u), 1)
    assert np.array_equal(got, ref) and np.random.randint(0, 2**32, dtype=np.uint32) == ref_next
    pool = np.empty((64, 64), dtype=np.uint8)
    raw = np.zeros(16, dtype=np.uint32)
    assert lib.bosot_host_recursive_dataset(raw.ctypes.data, len(raw), 1, -5, 1, pool.ctypes.data, C.byref(used)) == -1
    # a walk that outgrows the record: always take the newest element and its last neighbour (an extension, +2 nodes);
    # a raw output equal to n - 1 is accepted at once, so the stream can be written down in advance
    B, raw, grown = 3, [], [np.array([1, 0] + [0xFF] * 62, dtype=np.uint8)]
    n_pool = B
    for _ in range(40):
        raw.append(n_pool - 1)  # choice(len(pool)) -> the newest element (the first pick is root B - 1: same size as root 0)
        count = int(num_neighbours(grown[-1][None, :], B)[0])
        raw.append(count - 1)
        if grown[-1][0] + 2 > 63:
            break
        grown.append(neighbour_at_index_host(grown[-1], count - 1, B))
        n_pool += 1
    raw = np.array(raw, dtype=np.uint32)
    pool = np.empty((200, 64), dtype=np.uint8)
    assert lib.bosot_host_recursive_dataset(raw.ctypes.data, len(raw), B, 200, 1, pool.ctypes.data, C.byref(used)) == -3  # BOSOT_ETREE
    assert b"63-node" in lib.bosot_last_error()


def test_generator_callback_and_block_replay_agree(monkeypatch):
    """The producers pull their raw outputs straight from NumPy's global bit generator (bosot_host_*_cb, the default) or,
    when this NumPy does not hand out the generator's C interface, replay a block drawn in advance: same records, same
    indices, same generator state either way; a rejected call draws nothing."""
    import ctypes as C

    from bosot_b200 import _lib, flat_search as fs

    assert fs._global_bitgen() is not None  # NumPy >= 1.24 here: the default route is the callback
    gram = Grammar.get("CKSHighDimSearchSpace", 5)
    g = FlatCandidateGenerator(gram)
    sizes = np.random.default_rng(2).integers(1, 900, 5000)

    def run():
        np.random.seed(21)
        pool = g.get_dataset_recursivly_generated(3000, 2)
        idx = fs.legacy_choices(sizes)
        return pool, idx, np.random.randint(0, 2**32, dtype=np.uint32)

    a = run()
    monkeypatch.setattr(fs, "_global_bitgen", lambda: None)
    b = run()
    monkeypatch.undo()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]

    lib = _lib.load()
    out, used = np.empty(3, dtype=np.int64), C.c_int64(0)
    ok = np.array([5, 7, 9], dtype=np.int64)
    assert lib.bosot_host_choices_cb(None, None, ok.ctypes.data, 3, out.ctypes.data, C.byref(used)) == -1  # null callback
    fn, st = fs._global_bitgen()
    bad = np.array([5, 0, 9], dtype=np.int64)
    np.random.seed(1)
    before = np.random.get_state()[2]
    assert lib.bosot_host_choices_cb(fn, st, bad.ctypes.data, 3, out.ctypes.data, C.byref(used)) == -1  # a size of 0: nothing is drawn
    assert np.random.get_state()[2] == before
    pool = np.empty((8, 64), dtype=np.uint8)
    assert lib.bosot_host_recursive_dataset_cb(None, None, 3, 8, 1, pool.ctypes.data, C.byref(used)) == -1


@pytest.mark.gpu
def test_device_moves_match_host():
    torch = pytest.importorskip("torch")
    from bosot_b200 import sot
    from bosot_b200.flat_search import neighbours_device

    gram = Grammar.get("CKSHighDimSearchSpace", 5)
    trees = random_stress_trees(3000, gram, seed=21, max_depth=4)
    rng = np.random.default_rng(0)
    counts = num_neighbours(trees, gram.n_symbols)
    index = (rng.random(len(trees)) * counts).astype(np.int64)
    index[:4] = [0, counts[1] - 1, (trees[2, 0] + 1) // 2 * gram.n_symbols, (trees[3, 0] + 1) // 2 * gram.n_symbols - 1]
    out = neighbours_device(sot.as_device_trees(trees, "cuda:0"), torch.from_numpy(index), gram).cpu().numpy()
    ref = np.stack([neighbour_at_index_host(r, int(k), gram.n_symbols) for r, k in zip(trees, index)])
    assert np.array_equal(out, ref)
    # errors: index out of range, tree that would outgrow the record
    big = random_stress_trees(2000, gram, seed=5)
    full = big[big[:, 0] == 63][:1]
    if len(full):
        with pytest.raises(ValueError, match="code 5"):
            neighbours_device(sot.as_device_trees(full, "cuda:0"), torch.tensor([32 * gram.n_symbols]), gram)
    with pytest.raises(ValueError, match="code 4"):
        neighbours_device(sot.as_device_trees(trees[:1], "cuda:0"), torch.tensor([int(counts[0])]), gram)


@pytest.mark.gpu
def test_flat_ea_visits_the_same_kernels_as_the_object_ea():
    torch = pytest.importorskip("torch")
    from bosot_b200 import _lib, sot
    from bosot_b200.bayesian_optimization.evolutionary_optimizer_objects import EvolutionaryOptimizerObjects
    from bosot_b200.flat_search import FlatEvolutionaryOptimizer
    from bosot_b200.oracles.structural_oracle import StructuralOracle

    d, space = 2, "CKSWithRQSearchSpace"
    gram = Grammar.get(space, d)
    gen = GeneratorFactory.build(CKSWithRQGeneratorConfig(input_dimension=d))
    oracle = StructuralOracle(gram.symbols, seed=3)
    np.random.seed(1)
    x = gen.get_random_canditates(16)
    y = np.array([oracle.query(e)[0] for e in x])
    dev = torch.device("cuda:0")
    eng = sot.AcquisitionEngine(sot.as_device_trees(encode(x, gram), dev), torch.from_numpy(y), gram, (0.3, 0.3, 0.4), 1.3, 0.8, 1e-3, -0.5,
                                acq_kind=_lib.ACQ_UCB)

    def acq_objects(pop):
        return eng.acquire_device(sot.as_device_trees(encode(pop, gram), dev))[2].cpu().numpy()

    np.random.seed(77)
    best_obj, val_obj = EvolutionaryOptimizerObjects(60, 4, gen).maximize(acq_objects, 4)
    np.random.seed(77)
    flat_ea = FlatEvolutionaryOptimizer(60, 4, FlatCandidateGenerator(gram), dev)
    best_flat, val_flat = flat_ea.maximize(lambda t: eng.acquire_device(t)[2], 4)
    assert flat_ea.best_name() == str(best_obj)
    assert val_flat == float(val_obj)
