"""CPU: the NumPy restatement (oracle/ref_numpy.py) against the fixtures generated from the LIVE
reference featuriser (tests/golden, oracle/make_golden.py) and the hand-derived known answers of
SURVEY.md Appendix A.  Featurisation must be bit-exact; arithmetic within 1e-12 abs + 1e-9 rel."""
import numpy as np
import pytest

from oracle import ref_numpy as rn


def _trees(g, key):
    return [rn.parse_name(str(s)) for s in g[key]]


def test_names_roundtrip(golden):
    for key in ("eval_names", "cand_names"):
        for s in golden[key]:
            assert rn.tree_name(rn.parse_name(str(s))) == str(s)
    assert rn.base_kernel_names(str(golden["generator_name"]), int(golden["input_dimension"])) == [str(s) for s in golden["base_kernel_names"]]


def test_dense_features_bit_exact(golden):
    gen, d = str(golden["generator_name"]), int(golden["input_dimension"])
    xe, xc = _trees(golden, "eval_names"), _trees(golden, "cand_names")
    for f in rn.DEFAULT_FEATURE_TYPES:
        FE, FC = rn.dense_features(xe, xc, f, gen, d)
        assert np.array_equal(FE, golden["F_eval_joint_" + f]), f
        assert np.array_equal(FC, golden["F_cand_joint_" + f]), f
        FEE, _ = rn.dense_features(xe, None, f, gen, d)
        assert np.array_equal(FEE, golden["F_eval_self_" + f]), f


def test_distance_gram_posterior_ei(golden):
    gen, d = str(golden["generator_name"]), int(golden["input_dimension"])
    xe, xc = _trees(golden, "eval_names"), _trees(golden, "cand_names")
    h = rn.Hyper(**getattr(rn, str(golden["hyper_name"])))
    Df = rn.per_feature_distances(xe, xc, gen, d)
    np.testing.assert_allclose(Df, golden["Df_eval_cand"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(rn.K(xe, xc, h, gen, d), golden["K_eval_cand"], rtol=1e-13, atol=0)
    r = rn.acquisition_function(xe, golden["y"], xc, h, gen, d)
    np.testing.assert_allclose(r["mu_data"], golden["mu_eval"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r["mu"], golden["mu"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r["sigma"], golden["sigma"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r["acq"], golden["ei"], rtol=1e-7, atol=1e-14)
    # the literal K_diag (full N x N distance, then its diagonal) is exactly `variance`
    assert np.array_equal(rn.K_diag(xc[:40], h, gen, d, literal=True), np.full(40, h.variance))


def test_known_answers_appendix_a():
    SE0, PER0, LIN1, RQ1 = "RBFWithPrior_on_0", "PeriodicKernelWithPrior_on_0", "LinearWithPrior_on_1", "RQWithPrior_on_1"
    T1 = ("ADD", ("MULTIPLY", SE0, RQ1), PER0)
    T2 = ("ADD", PER0, ("MULTIPLY", RQ1, SE0))
    T3 = ("ADD", ("MULTIPLY", ("MULTIPLY", SE0, SE0), RQ1), PER0)
    T4 = LIN1
    T5 = ("ADD", ("ADD", SE0, SE0), ("MULTIPLY", SE0, LIN1))
    X = [T1, T2, T3, T4, T5]
    gen, d = "CKSWithRQSearchSpace", 2
    FD, _ = rn.dense_features(X, None, rn.DIM_WISE, gen, d)
    assert np.array_equal(FD[0], [0, 0.5, 0.5, 0, 0, 0, 0, 0, 0, 1.0])
    assert np.array_equal(FD[3], [1, 0, 0, 0, 0, 0, 0, 0, 1.0, 0])
    FP, _ = rn.dense_features(X, None, rn.PATHS, gen, d)
    counts = np.array([[1, 1, 1, 0, 0, 0], [1, 1, 1, 0, 0, 0], [2, 1, 1, 0, 0, 0], [0, 0, 0, 1, 0, 0], [1, 0, 0, 0, 2, 1]], dtype=float)
    assert np.array_equal(FP, counts / counts.sum(axis=1, keepdims=True))
    FS, _ = rn.dense_features(X, None, rn.SUBTREES, gen, d)
    sub = np.array([[1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0], [1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0], [0, 0, 2, 1, 1, 1, 1, 1, 0, 0, 0, 0],
                    [0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0], [0, 0, 3, 0, 0, 0, 0, 0, 1, 1, 1, 1]], dtype=float)
    assert np.array_equal(FS, sub / sub.sum(axis=1, keepdims=True))
    D = rn.per_feature_distances(X, None, gen, d)
    h = rn.Hyper()
    dist = rn.wasserstein_distance(X, None, h, gen, d)
    assert dist[0, 1] == 0.0
    np.testing.assert_allclose(D[:, 0, 2], [1 / 3, 1 / 3, 36 / 35], rtol=1e-15)
    np.testing.assert_allclose([dist[0, 2], dist[0, 3], dist[0, 4]], [178 / 315, 8 / 3, 61 / 30], rtol=1e-15)
    np.testing.assert_allclose([D[0, 2, 4], D[1, 2, 4], D[2, 2, 4]], [8 / 3, 3 / 2, 10 / 7], rtol=1e-15)
    np.testing.assert_allclose([D[0, 3, 4], D[1, 3, 4], D[2, 3, 4]], [2, 2, 12 / 7], rtol=1e-15)
    r = rn.acquisition_function([T1, T3, T4], np.array([0.3, -0.1, 0.5]), [T2, T5], h, gen, d)
    np.testing.assert_allclose(r["mu_data"], [0.299949493874, -0.099957867420, 0.499950586764], atol=1e-12)
    np.testing.assert_allclose(r["current_max"], 0.4999505867642072, rtol=1e-13)
    np.testing.assert_allclose(r["sigma"], [0.009999260597, 0.977054092246], atol=1e-12)
    np.testing.assert_allclose(r["acq"][1], 2.101167998194e-01, rtol=1e-12)
    np.testing.assert_allclose(r["acq"][0], 1.500882682270e-101, rtol=1e-6)


def test_invariants_commutativity_and_ranges():
    rng = np.random.default_rng(0)
    gen, d = "CKSHighDimSearchSpace", 5
    names = rn.base_kernel_names(gen, d)
    X = [rn.random_stress_tree(rng, names) for _ in range(40)]

    def swap(t):
        return t if isinstance(t, str) else (t[0], swap(t[2]), swap(t[1]))

    D = rn.per_feature_distances(X, [swap(t) for t in X], gen, d)
    assert np.all(np.diagonal(D, axis1=1, axis2=2) == 0.0)  # child swaps leave all three histograms unchanged
    Dxx = rn.per_feature_distances(X, None, gen, d)
    assert np.allclose(Dxx, np.transpose(Dxx, (0, 2, 1)), atol=1e-15)
    assert Dxx[0].max() <= 2 * d + 1e-12 and Dxx[1].max() <= 2 + 1e-12 and Dxx[2].max() <= 2 + 1e-12
    for f, s in ((rn.DIM_WISE, d), (rn.PATHS, 1), (rn.SUBTREES, 1)):
        F, _ = rn.dense_features(X, None, f, gen, d)
        np.testing.assert_allclose(F.sum(axis=1), s, rtol=1e-14)


def test_independent_restatements_agree(golden):
    """The arithmetic half cannot be pinned by executing TF / GPflow here, so the posterior (a9) and the marginal
    likelihood of the fit (f1) are restated twice through routes that share no factorisation -- Cholesky + triangular
    solves as GPflow's base_conditional does it, and a general solve / slogdet -- and must agree to 1e-10."""
    gen, d = str(golden["generator_name"]), int(golden["input_dimension"])
    xe, xc = _trees(golden, "eval_names"), _trees(golden, "cand_names")
    y = golden["y"]
    for h in (rn.Hyper(**rn.H0), rn.Hyper(**rn.H1), rn.Hyper(alphas=(0.9, 0.1, 0.6), lengthscale=2.5, variance=0.4, noise_variance=1e-2, mean_c=1.5)):
        K_EE, K_EN = rn.K(xe, None, h, gen, d), rn.K(xe, xc, h, gen, d)
        Knn = rn.K_diag(xc, h, gen, d, literal=False)
        mu, sigma = rn.gp_posterior_from_grams(K_EE, K_EN, Knn, y, h)
        mu2, var2 = rn.gp_posterior_direct(K_EE, K_EN, Knn, y, h)
        np.testing.assert_allclose(mu, mu2, rtol=1e-10, atol=1e-10)
        ok = np.isfinite(sigma)
        np.testing.assert_allclose(sigma[ok] ** 2, var2[ok], rtol=1e-10, atol=1e-10)
        assert np.all(var2[~ok] < 1e-10)  # sqrt of a rounding-negative variance: both routes see (numerically) zero
        a, b = rn.log_marginal_likelihood(K_EE, y, h), rn.log_marginal_likelihood_direct(K_EE, y, h)
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-10)
