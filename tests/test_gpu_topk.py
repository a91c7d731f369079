"""Device-side selection (C-ABI bosot_topk_f64 / bosot_argmax_f64) against NumPy: the tail of the STABLE argsort and the
first index of the maximum -- what EvolutionaryOptimizerObjects.select / maximize compute
(bosot/bayesian_optimization/evolutionary_optimizer_objects.py:42-44,57-64).  Index lists must be identical."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from bosot_b200 import sot  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _check(values: np.ndarray, k: int, offset: int = 0):
    v = torch.from_numpy(values).to(DEV)
    idx, vals = sot.topk(v, k, index_offset=offset)
    ref = np.argsort(values, kind="stable")[len(values) - k:]
    got = idx.cpu().numpy()
    assert np.array_equal(got, ref + offset), (len(values), k)
    assert np.array_equal(vals.cpu().numpy(), values[ref], equal_nan=True)
    idx2, none = sot.topk(v, k, index_offset=offset, want_values=False)
    assert none is None and torch.equal(idx2, idx)
    best = sot.argmax_first(v, index_offset=offset).cpu().numpy()
    r = int(np.argmax(values))
    assert int(best[1]) == r + offset and (best[0] == values[r] or (np.isnan(best[0]) and np.isnan(values[r])))


@pytest.mark.parametrize("n", [1, 2, 31, 100, 2048, 5000, 16384, 16385, 20_000, 100_000, 1_000_000])
def test_topk_matches_stable_argsort(n):
    rng = np.random.default_rng(n)
    smooth = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 3, size=n)          # wide dynamic range, both signs
    ties = np.round(rng.standard_normal(n), 1)                                     # heavy ties (a few dozen distinct values)
    ei_like = np.abs(rng.standard_normal(n)) * np.exp(-40.0 * rng.random(n))       # EI-shaped: non-negative, far tails
    ei_like[rng.random(n) < 0.3] = ei_like[0]                                      # duplicates of one candidate
    for values in (smooth, ties, ei_like):
        for k in sorted({1, min(n, 2), max(1, n // 5), max(1, n // 2), max(1, n - 1), n}):
            _check(values, k, offset=int(rng.integers(0, 1000)))


def test_topk_special_values_and_tie_boundaries():
    rng = np.random.default_rng(0)
    for n in (1000, 40_000):
        v = rng.standard_normal(n)
        v[rng.choice(n, 50, replace=False)] = np.nan
        v[rng.choice(n, 50, replace=False)] = np.inf
        v[rng.choice(n, 50, replace=False)] = -np.inf
        v[rng.choice(n, 60, replace=False)] = 0.0
        v[rng.choice(n, 60, replace=False)] = -0.0
        v[rng.choice(n, 10, replace=False)] = 5e-324
        for k in (1, 25, 50, 51, 75, 100, 101, n // 5, n - 60, n):
            _check(v, k)
        c = np.full(n, 0.25)  # all equal: the survivors are the LAST k indices
        for k in (1, 7, n // 5, n):
            _check(c, k)
        two = np.where(np.arange(n) % 3 == 0, 1.0, 2.0)  # the threshold falls inside a large tie group
        for k in (1, n // 3, 2 * n // 3, 2 * n // 3 + 1, n - 1):
            _check(two, k)


def test_topk_argument_checks():
    v = torch.zeros(10, dtype=torch.float64, device=DEV)
    for bad_k in (0, 11, -1):
        with pytest.raises(ValueError):
            sot.topk(v, bad_k)
    with pytest.raises(ValueError):
        sot.topk(v.float(), 3)
    with pytest.raises(RuntimeError, match="no CPU"):
        sot.topk(v.cpu(), 3)
