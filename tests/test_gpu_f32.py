"""fp32 variants of the hot-path entry points (SURVEY.md 8b "fp32 variants"; north_star: "1e-4 in fp32").

Stated tolerance (also in include/bosot_b200.h): per-feature distances, D and the Gram against the fp64 oracle at
1e-4 relative (+ 1e-6 absolute: a distance of exactly 0 must stay 0, which the integer formulation guarantees); mu,
sigma and EI at 1e-4 relative + 1e-4 * sqrt(variance) absolute -- L^-1 amplifies the fp32 rounding of K* (6e-8 relative)
by up to ||L^-1|| ~ 1 / sqrt(noise), and sigma^2 = variance - ||L^-1 k*||^2 cancels.
The reference computes in float64 (kernels/kernel_kernel_grammar_tree.py:35): these entry points are optional."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from bosot_b200 import _lib, sot  # noqa: E402
from bosot_b200.encoding import Grammar, decode_to_tuple, random_stress_trees  # noqa: E402
from oracle import ref_numpy as rn  # noqa: E402

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
GEN, DIM = "CKSHighDimSearchSpace", 5
RTOL = 1e-4


def _setup(n_eval, n_cand, hyper=rn.H1, seed=11):
    gram = Grammar.get(GEN, DIM)
    ev_np = random_stress_trees(n_eval, gram, seed=seed)
    ca_np = random_stress_trees(n_cand, gram, seed=seed + 1)
    ca_np[: min(5, n_eval)] = ev_np[: min(5, n_eval)]  # planted copies: d = 0, K = variance
    hyp = rn.Hyper(**hyper)
    w = hyp.alphas / hyp.alphas.sum()
    y = np.random.default_rng(seed).standard_normal(n_eval)
    return gram, ev_np, ca_np, hyp, w, y


@pytest.mark.parametrize("n_eval,n_cand", [(1, 40), (37, 300), (50, 2000), (200, 1500), (300, 700)])
def test_pairs_f32_against_oracle(n_eval, n_cand):
    gram, ev_np, ca_np, hyp, w, y = _setup(n_eval, n_cand)
    state = sot.EvalState(sot.as_device_trees(ev_np, DEV), gram)
    out = sot.sot_pairs(sot.as_device_trees(ca_np, DEV), state, w, hyp.variance, hyp.lengthscale, want=("K", "D", "Df"),
                        dtype=torch.float32)
    assert all(v.dtype == torch.float32 for v in out.values())
    ev = [decode_to_tuple(r, gram) for r in ev_np]
    ca = [decode_to_tuple(r, gram) for r in ca_np]
    Df_ref = rn.per_feature_distances(ca, ev, GEN, DIM)
    D_ref = rn.wasserstein_distance(ca, ev, hyp, GEN, DIM)
    K_ref = rn.K(ca, ev, hyp, GEN, DIM)
    Df = out["Df"].cpu().numpy().astype(np.float64)
    for f, name in enumerate(rn.DEFAULT_FEATURE_TYPES):  # DIM_WISE, PATHS, SUBTREES
        np.testing.assert_allclose(Df[f], Df_ref[f], rtol=RTOL, atol=1e-6, err_msg=str(name))
    np.testing.assert_allclose(out["D"].cpu().numpy(), D_ref, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(out["K"].cpu().numpy(), K_ref, rtol=RTOL, atol=1e-30)
    # exact zeros and the exact variance on the planted copies, like the fp64 path
    k = min(5, n_eval)
    Dn, Kn = out["D"].cpu().numpy(), out["K"].cpu().numpy()
    assert np.all(Dn[np.arange(k), np.arange(k)] == 0.0)
    assert np.all(Kn[np.arange(k), np.arange(k)] == np.float32(hyp.variance))
    # the fp32 Gram is the fp64 Gram rounded, up to the accuracy of the hardware exp
    K64 = sot.sot_pairs(sot.as_device_trees(ca_np, DEV), state, w, hyp.variance, hyp.lengthscale, want=("K",))["K"].cpu().numpy()
    np.testing.assert_allclose(Kn, K64, rtol=2e-6)


@pytest.mark.parametrize("acq_kind", [_lib.ACQ_EI, _lib.ACQ_UCB])
@pytest.mark.parametrize("n_eval,n_cand", [(24, 200), (50, 10_000), (200, 30_000)])
def test_acquire_device_f32(n_eval, n_cand, acq_kind):
    gram, ev_np, ca_np, hyp, w, y = _setup(n_eval, n_cand, seed=23)
    eng = sot.AcquisitionEngine(sot.as_device_trees(ev_np, DEV), torch.from_numpy(y), gram, w, hyp.variance, hyp.lengthscale,
                                hyp.noise_variance, hyp.mean_c, acq_kind=acq_kind)
    d_ca = sot.as_device_trees(ca_np, DEV)
    mu64, sg64, aq64 = [t.cpu().numpy() for t in eng.acquire_device(d_ca)]
    mu, sg, aq = eng.acquire_device(d_ca, dtype=torch.float32)
    assert mu.dtype == sg.dtype == aq.dtype == torch.float32
    assert eng.bad_tree_count() == 0
    atol = 1e-4 * np.sqrt(hyp.variance)
    np.testing.assert_allclose(mu.cpu().numpy(), mu64, rtol=RTOL, atol=atol)
    np.testing.assert_allclose(sg.cpu().numpy(), sg64, rtol=RTOL, atol=atol)
    np.testing.assert_allclose(aq.cpu().numpy(), aq64, rtol=RTOL, atol=atol)
    # and against the oracle itself on a sample
    idx = np.random.default_rng(3).choice(n_cand, size=min(n_cand, 200), replace=False)
    ref = rn.acquisition_function([decode_to_tuple(r, gram) for r in ev_np], y, [decode_to_tuple(ca_np[i], gram) for i in idx], hyp,
                                  GEN, DIM, kind="EI" if acq_kind == _lib.ACQ_EI else "UCB")
    np.testing.assert_allclose(mu.cpu().numpy()[idx], ref["mu"], rtol=RTOL, atol=atol)
    np.testing.assert_allclose(sg.cpu().numpy()[idx], ref["sigma"], rtol=RTOL, atol=atol)
    np.testing.assert_allclose(aq.cpu().numpy()[idx], ref["acq"], rtol=RTOL, atol=atol)


def test_posterior_f32_entry_point_and_argument_checks():
    gram, ev_np, ca_np, hyp, w, y = _setup(50, 777, seed=31)
    eng = sot.AcquisitionEngine(sot.as_device_trees(ev_np, DEV), torch.from_numpy(y), gram, w, hyp.variance, hyp.lengthscale,
                                hyp.noise_variance, hyp.mean_c)
    d_ca = sot.as_device_trees(ca_np, DEV)
    K32 = sot.sot_pairs(d_ca, eng.state, w, hyp.variance, hyp.lengthscale, dtype=torch.float32)["K"]
    mu, sg, aq = sot.posterior_acq(K32, eng.Linv, eng.beta, hyp.mean_c, hyp.variance, _lib.ACQ_EI, eng.current_max)
    mu2, sg2, aq2 = eng.acquire_device(d_ca, dtype=torch.float32)
    # the composed call runs the same two kernels on the same values (its Gram panel only has a wider row stride)
    assert torch.equal(mu, mu2) and torch.equal(sg, sg2) and torch.equal(aq, aq2)
    with pytest.raises(ValueError):
        eng.acquire_device(d_ca, dtype=torch.float16)
    with pytest.raises(ValueError):
        eng.acquire_device(d_ca, out=tuple(torch.empty(777, dtype=torch.float64, device=DEV) for _ in range(3)), dtype=torch.float32)


def test_f32_rows_of_malformed_records_are_nan():
    gram, ev_np, ca_np, hyp, w, y = _setup(40, 300, seed=41)
    ca_np = ca_np.copy()
    ca_np[17, 0] = 4  # even node count: not a tree
    eng = sot.AcquisitionEngine(sot.as_device_trees(ev_np, DEV), torch.from_numpy(y), gram, w, hyp.variance, hyp.lengthscale,
                                hyp.noise_variance, hyp.mean_c)
    mu, sg, aq = eng.acquire_device(sot.as_device_trees(ca_np, DEV), dtype=torch.float32)
    assert eng.bad_tree_count() == 1
    bad = torch.isnan(aq).cpu().numpy()
    assert bad[17] and bad.sum() == 1
    out = sot.sot_pairs(sot.as_device_trees(ca_np, DEV), eng.state, w, hyp.variance, hyp.lengthscale, want=("K", "D"), check=False,
                        dtype=torch.float32)
    assert torch.isnan(out["K"][17]).all() and torch.isnan(out["D"][17]).all()
    assert not torch.isnan(out["K"][16]).any() and not torch.isnan(out["K"][18]).any()
