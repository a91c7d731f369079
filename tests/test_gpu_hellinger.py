"""GPU parity of the Hellinger kernel-kernel against oracle/ref_hellinger.py, through the C-ABI
(``bosot_hellinger_grams`` / ``bosot_hellinger_pairs``).

Bars: Gram entries rel 1e-12 (different but equivalent evaluation order of the same fp64 formulas);
log-determinants abs 1e-6 and expected distances abs 1e-7: LDL^T without pivoting here, LAPACK LU in the
reference, on matrices whose condition number is set by the 1e-6 jitter (up to ~1e9); K rel 1e-6."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from bosot_b200 import _lib, hellinger, sot  # noqa: E402
from bosot_b200.encoding import Grammar, decode_to_tuple, encode, random_stress_trees  # noqa: E402
from oracle import ref_hellinger as rh  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _dense(record: np.ndarray, V: int) -> np.ndarray:
    """One store record (column-cyclic 4-lane order, two consecutive entries of a lane adjacent) -> dense symmetric matrix."""
    VP = (V + 3) // 4 * 4
    M = np.zeros((VP, VP))
    t = 0
    for m in range(VP // 4):
        for i in range(4 * m, VP):
            for r in range(4):
                j = 4 * m + r
                if i >= j:
                    M[i, j] = M[j, i] = record[((t >> 1) * 4 + r) * 2 + (t & 1)]
            t += 1
    return M


def _case(space_name, d, n, V, S, seed, max_depth=3):
    gram = Grammar.get(space_name, d)
    trees = random_stress_trees(n, gram, seed=seed, max_depth=max_depth)
    X = np.random.default_rng(seed).random((V, d))
    space = hellinger.HellingerSpace(gram, X, S, torch.device(DEV))
    tuples = [decode_to_tuple(r, gram) for r in trees]
    return gram, trees, X, space, tuples


@pytest.mark.parametrize("space_name,d,V,S", [("CKSWithRQSearchSpace", 2, 20, 100), ("CKSHighDimSearchSpace", 5, 20, 32),
                                               ("NDimFullKernelsSearchSpace", 3, 20, 16), ("CompositionalKernelSearchSpace", 2, 7, 13),
                                               ("CKSWithRQSearchSpace", 1, 12, 9)])
def test_gram_store_vs_oracle(space_name, d, V, S):
    gram, trees, X, space, tuples = _case(space_name, d, 12, V, S, seed=3)
    store = hellinger.grams(sot.as_device_trees(trees, DEV), space)
    G, ld = store.G.cpu().numpy(), store.logdet.cpu().numpy()
    for k, t in enumerate(tuples):
        g_ref, ld_ref = rh.sample_over_hps(t, X, S, d)
        for s in range(S):
            M = _dense(G[s, k], V)
            np.testing.assert_allclose(M[:V, :V], g_ref[s] + np.eye(V) * 1e-6, rtol=1e-12, atol=1e-15)  # the store carries 2 jitters
            assert np.array_equal(M[V:, V:], np.eye(M.shape[0] - V)) and not M[V:, :V].any()
        np.testing.assert_allclose(ld[:, k], ld_ref, rtol=0, atol=1e-6)


@pytest.mark.parametrize("space_name,d,V,S", [("CKSWithRQSearchSpace", 2, 20, 100), ("CKSHighDimSearchSpace", 4, 20, 24),
                                               ("NDimFullKernelsSearchSpace", 2, 16, 10), ("CompositionalKernelSearchSpace", 3, 9, 17)])
def test_distance_and_K_vs_oracle(space_name, d, V, S):
    gram, trees, X, space, tuples = _case(space_name, d, 40, V, S, seed=11)
    dev_trees = sot.as_device_trees(trees, DEV)
    store = hellinger.grams(dev_trees, space)
    out = hellinger.pairs(store, store, space, 1.3, 0.6, want=("K", "D"))
    D, K = out["D"].cpu().numpy(), out["K"].cpu().numpy()
    D_ref = rh.distance_matrix(tuples, None, X, S, d)
    np.testing.assert_allclose(D, D_ref, rtol=0, atol=1e-7)
    np.testing.assert_allclose(K, 1.3 * np.exp(-0.5 * (D_ref / 0.36)), rtol=1e-6)
    assert np.array_equal(D, D.T) and np.array_equal(K, K.T)  # commutative sums: bitwise symmetric like the reference's cache
    assert out["info"].cpu().tolist() == [0, 0]
    # rectangular call, rows chunked so that several transient stores are used
    sub = hellinger.grams(dev_trees[:7], space)
    rect = hellinger.cross(dev_trees, sub, space, 1.3, 0.6, want=("K", "D"), max_store_bytes=space.store_bytes(9))
    assert torch.equal(rect["D"], out["D"][:, :7]) and torch.equal(rect["K"], out["K"][:, :7])


def test_tile_edges_and_large_rectangles():
    """Sizes around the 32 x 16 tile of the pair kernel, checked through separability: D[a][b] only depends on the
    two trees, so any sub-rectangle must reproduce the entries of the full matrix bit for bit."""
    gram, trees, X, space, _ = _case("CKSWithRQSearchSpace", 2, 75, 20, 8, seed=5, max_depth=4)
    dev_trees = sot.as_device_trees(trees, DEV)
    full_store = hellinger.grams(dev_trees, space)
    full = hellinger.pairs(full_store, full_store, space, 1.0, 1.0, want=("D",))["D"]
    for na, nb in [(1, 1), (31, 15), (32, 16), (33, 17), (64, 49), (75, 1)]:
        a, b = hellinger.grams(dev_trees[:na], space), hellinger.grams(dev_trees[75 - nb:], space)
        part = hellinger.pairs(a, b, space, 1.0, 1.0, want=("D",))["D"]
        assert torch.equal(part, full[:na, 75 - nb:]), (na, nb)


def test_deep_trees_and_duplicates():
    gram, trees, X, space, tuples = _case("CKSWithRQSearchSpace", 2, 10, 20, 12, seed=21, max_depth=5)
    trees[4] = trees[1]
    tuples[4] = tuples[1]
    store = hellinger.grams(sot.as_device_trees(trees, DEV), space)
    D = hellinger.pairs(store, store, space, 1.0, 1.0, want=("D",))["D"].cpu().numpy()
    np.testing.assert_allclose(D, rh.distance_matrix(tuples, None, X, 12, 2), rtol=0, atol=1e-7)
    assert D[1, 4] == D[1, 1] == D[4, 4] and D[1, 1] > 0  # a kernel is not at distance 0 from itself (second jitter)


def test_malformed_trees_and_bad_arguments():
    gram, trees, X, space, _ = _case("CKSWithRQSearchSpace", 2, 4, 20, 4, seed=1)
    bad = trees.copy()
    bad[2, 0] = 4  # even node count
    with pytest.raises(ValueError):
        hellinger.grams(sot.as_device_trees(bad, DEV), space)
    store = hellinger.grams(sot.as_device_trees(bad, DEV), space, check=False)
    assert torch.isnan(store.logdet[:, 2]).all() and not torch.isnan(store.logdet[:, [0, 1, 3]]).any()
    out = hellinger.pairs(store, store, space, 1.0, 1.0, want=("D",))
    D = out["D"].cpu().numpy()
    assert np.isnan(D[2]).all() and np.isnan(D[:, 2]).all() and not np.isnan(np.delete(np.delete(D, 2, 0), 2, 1)).any()
    assert out["info"].cpu().tolist() == [7 * 4, 7]  # every sample of the 7 pairs touching tree 2 is dropped
    with pytest.raises(ValueError):
        hellinger.HellingerSpace(gram, np.zeros((21, 2)), 4, torch.device(DEV))
    with pytest.raises(ValueError):
        hellinger.pairs(store, store, space, 1.0, 0.0)
    lib = _lib.load()
    assert lib.bosot_hellinger_record_doubles(20) == 248 and lib.bosot_hellinger_record_doubles(21) < 0
