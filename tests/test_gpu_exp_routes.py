"""The two exp routes of the pair kernel against the oracle (kernels/kernel_kernel_grammar_tree.py:110-113).

The host picks the route per launch from the hyper-parameters alone (csrc/sot_pairs.cu, launch_pairs): the table-driven
exp when -2 (n_blocks w_dim + w_path + w_sub) / l^2 >= -600 and the variance is an ordinary number, the general
exp_nonpos route (clamp, two-step scaling through the denormal range) otherwise.  Both must meet the same bar: Gram
entries at 1e-12 relative wherever exp(x) carries its full precision, exact `variance` on planted copies, and -- for
arguments beyond the fp64 range -- gradual underflow to 0 without NaN."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from bosot_b200 import sot  # noqa: E402
from bosot_b200.encoding import Grammar, decode_to_tuple, random_stress_trees  # noqa: E402
from oracle import ref_numpy as rn  # noqa: E402

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
GEN, DIM = "CKSHighDimSearchSpace", 5


def _grams(lengthscale, variance, alphas, n_eval=70, n_cand=400, seed=3):
    gram = Grammar.get(GEN, DIM)
    ev_np = random_stress_trees(n_eval, gram, seed=seed)
    ca_np = random_stress_trees(n_cand, gram, seed=seed + 1)
    ca_np[:6] = ev_np[:6]
    hyp = rn.Hyper(alphas=alphas, lengthscale=lengthscale, variance=variance, noise_variance=1e-3, mean_c=0.0)
    w = hyp.alphas / hyp.alphas.sum()
    state = sot.EvalState(sot.as_device_trees(ev_np, DEV), gram)
    out = sot.sot_pairs(sot.as_device_trees(ca_np, DEV), state, w, hyp.variance, hyp.lengthscale, want=("K", "D"))
    ev = [decode_to_tuple(r, gram) for r in ev_np]
    ca = [decode_to_tuple(r, gram) for r in ca_np]
    return out["K"].cpu().numpy(), out["D"].cpu().numpy(), rn.K(ca, ev, hyp, GEN, DIM), rn.wasserstein_distance(ca, ev, hyp, GEN, DIM), hyp


# (lengthscale, variance, alphas, route): x_min = -2 (5 w_dim + w_path + w_sub) / l^2 with w = alphas / sum(alphas)
CASES = [
    (1.0, 1.0, (0.5, 0.5, 0.5), "table"),      # config defaults: x in [-4.7, 0]
    (0.8, 1.3, (0.2, 0.7, 0.4), "table"),      # the bench hyper-parameters
    (0.08, 2.5, (0.2, 0.7, 0.4), "table"),     # x_min = -505: exponents down to 2^-728
    (0.075, 1e-20, (0.3, 0.3, 0.3), "general"),  # x_min = -830 < -600
    (0.05, 1.0, (0.2, 0.7, 0.4), "general"),   # x_min = -1290: results run through the denormal range into 0
    (1.0, 1e-40, (0.5, 0.5, 0.5), "general"),  # variance below 2^-100
    (3.0, 1e35, (0.9, 0.1, 0.6), "general"),   # variance above 2^100
]


@pytest.mark.parametrize("lengthscale,variance,alphas,route", CASES)
def test_gram_matches_oracle_on_both_exp_routes(lengthscale, variance, alphas, route):
    K, D, K_ref, D_ref, hyp = _grams(lengthscale, variance, alphas)
    np.testing.assert_allclose(D, D_ref, rtol=0, atol=1e-12)
    assert not np.isnan(K).any()
    x = -D_ref / lengthscale ** 2
    # rounding of the argument costs |x| ulp of the result whatever the implementation (the oracle's exp included)
    normal = K_ref > variance * 1e-290
    tol = 1e-12 + 4e-16 * np.abs(x)
    assert np.all(np.abs(K - K_ref)[normal] <= (tol * K_ref)[normal])
    # below that: gradual underflow, never more than the reference value's own spacing away, and 0 where it is 0
    assert np.all(np.abs(K - K_ref)[~normal] <= 1e-9 * variance * 1e-290 + 4 * np.spacing(K_ref[~normal]))
    # planted copies: distance exactly 0, Gram entry exactly the variance
    idx = np.arange(6)
    assert np.all(D[idx, idx] == 0.0)
    assert np.all(K[idx, idx] == variance)
    # symmetry of the route: K(a, b) over the evaluated set is bitwise symmetric
    gram = Grammar.get(GEN, DIM)
    ev_np = random_stress_trees(70, gram, seed=3)
    state = sot.EvalState(sot.as_device_trees(ev_np, DEV), gram)
    w = hyp.alphas / hyp.alphas.sum()
    K_EE = sot.sot_pairs(sot.as_device_trees(ev_np, DEV), state, w, variance, lengthscale)["K"]
    assert torch.equal(K_EE, K_EE.t().contiguous())


def test_large_and_small_pools_take_the_same_route():
    """The evaluated-set call (a few dozen rows, fused-scan kernel) and the pool call (two-launch path, block-vector cache)
    must produce the same bits for the same tree: the posterior at planted copies depends on it."""
    gram = Grammar.get(GEN, DIM)
    ev_np = random_stress_trees(64, gram, seed=9)
    big = random_stress_trees(40_000, gram, seed=10)
    big[1000:1064] = ev_np
    state = sot.EvalState(sot.as_device_trees(ev_np, DEV), gram)
    w = np.array([0.2, 0.7, 0.4]) / 1.3
    K_small = sot.sot_pairs(sot.as_device_trees(ev_np, DEV), state, w, 1.3, 0.8)["K"]
    K_big = sot.sot_pairs(sot.as_device_trees(big, DEV), state, w, 1.3, 0.8)["K"]
    assert torch.equal(K_big[1000:1064], K_small)


@pytest.mark.parametrize("gen,dim", [("CKSHighDimSearchSpace", 5), ("CKSHighDimSearchSpace", 3), ("CKSWithRQSearchSpace", 2)])
def test_block_vector_cache_changes_no_bit(gen, dim):
    """Pools beyond the fused-scan limit go through the two-launch path, where (for search spaces whose blocks have at most
    two kinds) the DIM_WISE distances come from the per-CTA block-vector cache and only candidates with more than five
    leaves of one kind build their own table; small pools never use the cache.  Same trees, same bits, for D_f, D and K."""
    gram = Grammar.get(gen, dim)
    ev_np = random_stress_trees(150, gram, seed=21)
    big = random_stress_trees(20_000, gram, seed=22)
    state = sot.EvalState(sot.as_device_trees(ev_np, DEV), gram)
    w = np.array([0.3, 0.5, 0.2])
    d_big = sot.as_device_trees(big, DEV)
    whole = sot.sot_pairs(d_big, state, w, 0.9, 1.1, want=("K", "D", "Df"))
    K_only = sot.sot_pairs(d_big, state, w, 0.9, 1.1)["K"]
    assert torch.equal(K_only, whole["K"])
    for c0 in range(0, 20_000, 5_000):  # 5000 rows: the fused-scan path, no cache
        part = sot.sot_pairs(d_big[c0:c0 + 5_000].contiguous(), state, w, 0.9, 1.1, want=("K", "D", "Df"))
        assert torch.equal(part["K"], whole["K"][c0:c0 + 5_000])
        assert torch.equal(part["D"], whole["D"][c0:c0 + 5_000])
        assert torch.equal(part["Df"], whole["Df"][:, c0:c0 + 5_000])
    # the pool does contain candidates on the fallback route (a kind with more than five leaves)
    n_leaves = (big[:, 0].astype(int) + 1) // 2
    assert (n_leaves > 12).sum() > 100
