"""Third anchor for the arithmetic half of the oracle (GP posterior, EI inputs, log marginal likelihood): scikit-learn's
GaussianProcessRegressor -- an independent, widely used implementation of the same textbook formulas -- fed with the
oracle's own SOT Gram through a precomputed-Gram kernel.

TensorFlow / GPflow cannot be installed here, so `oracle/ref_numpy.py` restates GPflow 2.4's
`base_conditional` and `GPR.log_marginal_likelihood` from their published definitions; `tests/test_oracle_golden.py`
cross-checks that restatement against a second route of our own (LU solve / slogdet).  This file checks it against code
we did not write: same Gram, same targets, same noise -> same mean, standard deviation and log marginal likelihood."""
import numpy as np
import pytest

sklearn_gp = pytest.importorskip("sklearn.gaussian_process")
from sklearn.gaussian_process.kernels import Kernel  # noqa: E402

from oracle import ref_numpy as rn  # noqa: E402

GEN, DIM = "CKSHighDimSearchSpace", 3


class IndexGram(Kernel):
    """k(i, j) = G[i, j] for integer 'inputs' i, j: lets scikit-learn's GPR run on a precomputed Gram."""

    def __init__(self, G):
        self.G = G

    def __call__(self, X, Y=None, eval_gradient=False):
        i = np.asarray(X)[:, 0].astype(int)
        j = i if Y is None else np.asarray(Y)[:, 0].astype(int)
        K = self.G[np.ix_(i, j)]
        return (K, np.empty((len(i), len(j), 0))) if eval_gradient else K

    def diag(self, X):
        i = np.asarray(X)[:, 0].astype(int)
        return self.G[i, i].copy()

    def is_stationary(self):
        return False


@pytest.mark.parametrize("hyper", [rn.H0, rn.H1])
def test_posterior_and_lml_match_scikit_learn(hyper):
    rng = np.random.default_rng(12)
    names = rn.base_kernel_names(GEN, DIM)
    trees = [rn.random_stress_tree(rng, names, max_depth=4) for _ in range(75)]
    ev, ca = trees[:30], trees[30:]
    hyp = rn.Hyper(**hyper)
    y = rng.standard_normal(len(ev))
    G = rn.K(trees, None, hyp, GEN, DIM)  # the oracle's SOT Gram over all 75 trees (K_diag == variance on its diagonal)
    assert np.allclose(np.diag(G), hyp.variance, rtol=0, atol=1e-15)
    idx_e = np.arange(len(ev))[:, None].astype(float)
    idx_c = (len(ev) + np.arange(len(ca)))[:, None].astype(float)

    gpr = sklearn_gp.GaussianProcessRegressor(kernel=IndexGram(G), alpha=hyp.noise_variance, optimizer=None, normalize_y=False)
    gpr.fit(idx_e, y - hyp.mean_c)  # constant mean function: object_mean_functions.py:40-48
    mu_sk, sd_sk = gpr.predict(idx_c, return_std=True)
    mu_sk = mu_sk + hyp.mean_c

    K_EE, K_EN = G[: len(ev), : len(ev)], G[: len(ev), len(ev):]
    mu, sigma = rn.gp_posterior_from_grams(K_EE, K_EN, np.full(len(ca), hyp.variance), y, hyp)
    np.testing.assert_allclose(mu, mu_sk, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(sigma, sd_sk, rtol=1e-7, atol=1e-10)  # sqrt of a cancelling difference on both sides

    # the whole acquisition of the oracle on top of scikit-learn's posterior gives the oracle's EI
    ref = rn.acquisition_function(ev, y, ca, hyp, GEN, DIM)
    mu_data_sk = gpr.predict(idx_e) + hyp.mean_c
    ei_sk = rn.expected_improvement(mu_sk, sd_sk, float(np.max(mu_data_sk)))
    np.testing.assert_allclose(ref["acq"], ei_sk, rtol=1e-6, atol=1e-12)

    # log marginal likelihood (GPR.training_loss = -LML, models/gp_model.py:195-199)
    lml_sk = gpr.log_marginal_likelihood_value_
    assert abs(rn.log_marginal_likelihood(K_EE, y, hyp) - lml_sk) <= 1e-9 * max(1.0, abs(lml_sk))
