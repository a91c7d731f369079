"""CPU checks of the Hellinger oracle (oracle/ref_hellinger.py).

The pair half (Grams + log-dets -> expected distance -> matrix) is compared with the LIVE reference
methods (kernel_kernel_hellinger.py:138-190,214-218) imported through the shim; the sampling half restates
TF / GPflow / TFP behaviour that cannot run here (parity unpinned, see the oracle's header) and is checked
for its documented properties only."""
import math

import numpy as np
import pytest

from oracle import ref_hellinger as rh
from oracle import ref_shim

TREES = [
    "RBFWithPrior_on_0",
    "LinearWithPrior_on_1",
    ("ADD", "RBFWithPrior_on_0", "LinearWithPrior_on_1"),
    ("MULTIPLY", ("ADD", "PeriodicKernelWithPrior_on_0", "RQWithPrior_on_1"), "RBFWithPrior_on_1"),
]


def test_sobol_points_are_the_gray_code_sequence_without_the_origin():
    u = rh.sobol_sample(3, 4)
    assert np.array_equal(u, [[0.5, 0.5, 0.5], [0.75, 0.25, 0.25], [0.25, 0.75, 0.75], [0.375, 0.375, 0.625]])
    # coordinate j does not depend on the number of dimensions drawn (what lets one table serve every tree)
    assert np.array_equal(rh.sobol_sample(7, 50)[:, :3], rh.sobol_sample(3, 50))


def test_gamma_quantile_inverts_the_cdf():
    from scipy.stats import gamma

    u = np.array([0.1, 0.5, 0.9])
    q = rh.gamma_quantile(2.0, 3.0, u)
    np.testing.assert_allclose(gamma.cdf(q, a=2.0, scale=1.0 / 3.0), u, rtol=1e-12)


def test_parameter_numbering():
    assert rh.num_parameters("RBFWithPrior_on_0", 2) == 2
    assert rh.num_parameters("RBFWithPrior", 3) == 4  # a lengthscale per dimension
    assert rh.num_parameters(TREES[3], 2) == 3 + 3 + 2
    assert rh.leaf_spec("PeriodicKernelWithPrior_on_1", 2) == ("periodic", ("period", "lengthscales", "variance"), 1, 1)


def test_elementary_grams_closed_forms():
    X = np.array([[0.0, 1.0], [0.5, 0.25]])
    th = {"lengthscales": np.array([2.0]), "variance": 3.0, "period": 0.7, "alpha": 1.5, "offset": 0.4}
    x = X[:, :1]
    d = 0.5
    np.testing.assert_allclose(rh.elementary_gram("rbf", x, th)[0, 1], 3.0 * math.exp(-0.5 * (d / 2.0) ** 2), rtol=1e-14)
    np.testing.assert_allclose(rh.elementary_gram("rq", x, th)[0, 1], 3.0 * (1 + 0.5 * (d / 2.0) ** 2 / 1.5) ** -1.5, rtol=1e-14)
    np.testing.assert_allclose(rh.elementary_gram("periodic", x, th)[0, 1], 3.0 * math.exp(-0.5 * (math.sin(math.pi * d / 0.7) / 2.0) ** 2), rtol=1e-14)
    np.testing.assert_allclose(rh.elementary_gram("linear", X, th)[0, 1], 3.0 * (0.0 * 0.5 + 1.0 * 0.25) + 0.4, rtol=1e-14)


def test_sum_and_product_fold_left_like_gpflow():
    X = rh.sobol_sample(2, 6)
    u = rh.sobol_sample(8, 3)[2]
    t = ("ADD", "RBFWithPrior_on_0", ("ADD", "LinearWithPrior_on_1", "RBFWithPrior_on_1"))
    a = rh.gram_of_tree("RBFWithPrior_on_0", X, u[0:2], 2)
    b = rh.gram_of_tree("LinearWithPrior_on_1", X, u[2:4], 2)
    c = rh.gram_of_tree("RBFWithPrior_on_1", X, u[4:6], 2)
    assert np.array_equal(rh.gram_of_tree(t, X, u, 2), (a + b) + c)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_pair_half_equals_live_reference():
    ref_shim.install()
    from bosot.kernels import kernel_kernel_hellinger as kh

    S, X = 24, rh.sobol_sample(2, 20)
    ref = object.__new__(kh.KernelKernelHellinger)  # __init__ needs TF; the methods under test are NumPy only
    ref.num_param_samples, ref.jitter, ref.virtual_X = S, 1e-6, X
    ref.expected_distance_cache = kh.ExpectedHellingerDistanceCache()
    wrapped, sampled = [], []
    for i, t in enumerate(TREES):
        g, d = rh.sample_over_hps(t, X, S, 2)
        Kj, ld = ref.get_slog_det(g[0] - np.eye(20) * 1e-6)
        assert np.array_equal(Kj, (g[0] - np.eye(20) * 1e-6) + np.eye(20) * 1e-6) and ld == rh.slog_det(g[0] - np.eye(20) * 1e-6)[1]
        wrapped.append(kh.EvaluatedKernelWrapper(list(g), list(d), "k%d" % i))
        sampled.append((g, d))
    D_ref = ref.get_hellinger_distance_matrix(wrapped)
    D = np.array([[rh.expected_hellinger_distance(*a, *b) for b in sampled] for a in sampled])
    assert np.array_equal(D_ref, D)
    assert np.array_equal(ref.get_hellinger_distance_matrix(wrapped, wrapped[:2]), D[:, :2])
    assert ref.hellinger_distance(sampled[0][0][3], sampled[0][1][3], sampled[2][0][3], sampled[2][1][3]) == rh.hellinger_distance(
        sampled[0][0][3], sampled[0][1][3], sampled[2][0][3], sampled[2][1][3])


def test_distance_matrix_properties():
    X = rh.sobol_sample(2, 20)
    D = rh.distance_matrix(TREES, None, X, 16, 2)
    assert np.array_equal(D, D.T) and np.all(D >= 0) and np.all(D <= 1)
    # the distance of a kernel to itself is NOT zero: the averaged Gram carries a second jitter (:185-187, :214-217)
    assert np.all(np.diag(D) > 0)
    K = rh.K(TREES, TREES[:2], X, 16, 2, 1.3, 0.5)
    np.testing.assert_allclose(K, 1.3 * np.exp(-0.5 * D[:, :2] / 0.25), rtol=1e-15)
