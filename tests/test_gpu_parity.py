"""GPU parity tests proper: the CUDA path, called through the C-ABI, against (a) the golden fixtures
generated from the live reference featuriser and (b) the CPU oracle on seeded inputs.

Bars: featurisation / indexing bit-exact (dense [N, F] fp64 matrices equal column for column);
per-feature distances abs 1e-12; Gram rel 1e-12; posterior mean / sigma and EI rel 1e-6 + abs 1e-12
(north_star: 1e-6 relative in fp64).
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from bosot_b200 import _lib, sot  # noqa: E402
from bosot_b200.encoding import Grammar, encode, random_stress_trees  # noqa: E402
from oracle import ref_numpy as rn  # noqa: E402

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
FEATS = (("DIM_WISE", rn.DIM_WISE), ("PATHS", rn.PATHS), ("SUBTREES", rn.SUBTREES))


def _setup(g):
    gram = Grammar.get(str(g["generator_name"]), int(g["input_dimension"]))
    te = sot.as_device_trees(encode([str(s) for s in g["eval_names"]], gram), DEV)
    tc = sot.as_device_trees(encode([str(s) for s in g["cand_names"]], gram), DEV)
    h = rn.Hyper(**getattr(rn, str(g["hyper_name"])))
    w = h.alphas / np.sum(h.alphas)
    return gram, te, tc, h, w


def test_featurisation_bit_exact_vs_reference(golden):
    gram, te, tc, _, _ = _setup(golden)
    fe, fc = sot.featurize(te, gram), sot.featurize(tc, gram)
    joint = sot.dense_feature_matrices(fe, fc)
    alone = sot.dense_feature_matrices(fe)
    for ours, ref in FEATS:
        FE, FC = joint[ours]
        assert FE.shape == golden["F_eval_joint_" + ref].shape and FC.shape == golden["F_cand_joint_" + ref].shape
        assert np.array_equal(FE, golden["F_eval_joint_" + ref]), ref
        assert np.array_equal(FC, golden["F_cand_joint_" + ref]), ref
        assert np.array_equal(alone[ours][0], golden["F_eval_self_" + ref]), ref
    n_leaves = fc["n_leaves"].cpu().numpy()
    assert np.array_equal(n_leaves, [rn.count_leaves(rn.parse_name(str(s))) for s in golden["cand_names"]])


def test_distances_and_gram_vs_reference(golden):
    gram, te, tc, h, w = _setup(golden)
    st = sot.EvalState(te, gram)
    o = sot.sot_pairs(tc, st, w, h.variance, h.lengthscale, want=("K", "D", "Df"))
    Df = o["Df"].cpu().numpy().transpose(0, 2, 1)  # fixtures are [f, eval, cand]
    np.testing.assert_allclose(Df, golden["Df_eval_cand"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(o["D"].cpu().numpy().T, golden["d_eval_cand"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(o["K"].cpu().numpy().T, golden["K_eval_cand"], rtol=1e-12, atol=0)
    # evaluated set against itself: symmetric, exact zeros on the diagonal (d(x, x) = 0 in the reference)
    s = sot.sot_pairs(te, st, w, h.variance, h.lengthscale, want=("K", "D", "Df"))
    np.testing.assert_allclose(s["Df"].cpu().numpy(), golden["Df_eval_eval"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(s["K"].cpu().numpy(), golden["K_eval_eval"], rtol=1e-12, atol=0)
    D = s["D"].cpu().numpy()
    assert np.all(np.diag(D) == 0.0)
    assert np.array_equal(np.diag(s["K"].cpu().numpy()), np.full(st.n_eval, h.variance))  # K_diag == variance
    assert np.array_equal(D, D.T)
    # candidates 5 and 6 are copies of evaluated trees
    Dc = o["D"].cpu().numpy()
    assert Dc[5, st.n_eval - 1] == 0.0 and Dc[6, st.n_eval // 2] == 0.0


def test_posterior_and_acquisition_vs_reference(golden):
    gram, te, tc, h, w = _setup(golden)
    y = torch.from_numpy(golden["y"]).to(DEV)
    for kind, key in ((_lib.ACQ_EI, "ei"), (_lib.ACQ_UCB, "ucb")):
        eng = sot.AcquisitionEngine(te, y, gram, w, h.variance, h.lengthscale, h.noise_variance, h.mean_c, acq_kind=kind)
        np.testing.assert_allclose(eng.mu_data.cpu().numpy(), golden["mu_eval"], rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(eng.sigma_data.cpu().numpy(), golden["sigma_eval"], rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(float(eng.current_max), float(golden["current_max"]), rtol=1e-9)
        mu, sigma, acq = eng.acquire_device(tc)
        np.testing.assert_allclose(mu.cpu().numpy(), golden["mu"], rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(sigma.cpu().numpy(), golden["sigma"], rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(acq.cpu().numpy(), golden[key], rtol=1e-6, atol=1e-12)
        assert int(np.argmax(acq.cpu().numpy())) == int(np.argmax(golden[key]))
        # host-buffer (end-to-end) entry point gives the same bits as the device one
        h_trees = tc.cpu().pin_memory()
        h_acq = torch.empty(tc.shape[0], dtype=torch.float64).pin_memory()
        h_mu = torch.empty_like(h_acq).pin_memory()
        h_sig = torch.empty_like(h_acq).pin_memory()
        eng.acquire_host(h_trees, h_acq, h_mu, h_sig)
        assert torch.equal(h_acq, acq.cpu()) and torch.equal(h_mu, mu.cpu()) and torch.equal(h_sig, sigma.cpu())


def test_appendix_a_known_answers():
    SE0, PER0, LIN1, RQ1 = "RBFWithPrior_on_0", "PeriodicKernelWithPrior_on_0", "LinearWithPrior_on_1", "RQWithPrior_on_1"
    T1 = ("ADD", ("MULTIPLY", SE0, RQ1), PER0)
    T2 = ("ADD", PER0, ("MULTIPLY", RQ1, SE0))
    T3 = ("ADD", ("MULTIPLY", ("MULTIPLY", SE0, SE0), RQ1), PER0)
    T4 = LIN1
    T5 = ("ADD", ("ADD", SE0, SE0), ("MULTIPLY", SE0, LIN1))
    gram = Grammar.get("CKSWithRQSearchSpace", 2)
    X = sot.as_device_trees(encode([T1, T2, T3, T4, T5], gram), DEV)
    st = sot.EvalState(X, gram)
    o = sot.sot_pairs(X, st, (1 / 3, 1 / 3, 1 / 3), 1.0, 1.0, want=("K", "D", "Df"))
    Df, D = o["Df"].cpu().numpy(), o["D"].cpu().numpy()
    assert D[0, 1] == 0.0 and np.all(Df[:, 0, 1] == 0.0)  # T2 = T1 up to child swaps
    np.testing.assert_allclose(Df[:, 0, 2], [1 / 3, 1 / 3, 36 / 35], rtol=1e-15)
    np.testing.assert_allclose(Df[:, 0, 3], [4, 2, 2], rtol=1e-15)
    np.testing.assert_allclose(Df[:, 2, 4], [8 / 3, 3 / 2, 10 / 7], rtol=1e-15)
    np.testing.assert_allclose([D[0, 2], D[0, 3], D[0, 4]], [178 / 315, 8 / 3, 61 / 30], rtol=1e-15)
    ev = sot.as_device_trees(encode([T1, T3, T4], gram), DEV)
    eng = sot.AcquisitionEngine(ev, torch.tensor([0.3, -0.1, 0.5], dtype=torch.float64), gram, (1 / 3, 1 / 3, 1 / 3), 1.0, 1.0, 1e-4, 0.0)
    np.testing.assert_allclose(float(eng.current_max), 0.4999505867642072, rtol=1e-12)
    mu, sigma, ei = eng.acquire_device(sot.as_device_trees(encode([T2, T5], gram), DEV))
    np.testing.assert_allclose(mu.cpu().numpy(), [0.299949493874, 0.074410747987], atol=1e-11)
    np.testing.assert_allclose(sigma.cpu().numpy(), [0.009999260597, 0.977054092246], atol=1e-11)
    np.testing.assert_allclose(ei.cpu().numpy()[1], 2.101167998194e-01, rtol=1e-10)
    np.testing.assert_allclose(ei.cpu().numpy()[0], 1.500882682270e-101, rtol=1e-5)  # far tail


def _chain(names, n_leaves, left_deep, ops):
    t = names[0]
    for i in range(1, n_leaves):
        op = ops[i % len(ops)]
        t = (op, t, names[i % len(names)]) if left_deep else (op, names[i % len(names)], t)
    return t


def test_edge_cases_sizes_and_shapes():
    gen, d = "CKSWithRQSearchSpace", 2
    gram = Grammar.get(gen, d)
    names = gram.symbols
    trees = [names[0], names[7]]  # single leaves (empty path, kernel_grammar.py:311-316)
    trees += [_chain(names, 32, True, ("ADD", "MULTIPLY")), _chain(names, 32, False, ("ADD", "MULTIPLY"))]  # 63 nodes, depth 31
    trees += [_chain(names, 32, True, ("ADD",)), _chain(names[:1], 32, False, ("MULTIPLY",))]  # one collapsed run / one symbol
    trees += [_chain(names, k, k % 2 == 0, ("MULTIPLY", "ADD", "ADD")) for k in range(2, 30)]
    rng = np.random.default_rng(5)
    trees += [rn.random_stress_tree(rng, names) for _ in range(67)]  # ragged total: 101 = 3 warps + 5
    ev = trees[:37]  # E = 37: neither a multiple of 8 nor of 32
    hyp = rn.Hyper(**rn.H1)
    w = hyp.alphas / hyp.alphas.sum()
    X, XE = sot.as_device_trees(encode(trees, gram), DEV), sot.as_device_trees(encode(ev, gram), DEV)
    # features
    dense = sot.dense_feature_matrices(sot.featurize(XE, gram), sot.featurize(X, gram))
    for ours, ref in FEATS:
        FE, FC = rn.dense_features(ev, trees, ref, gen, d)
        assert np.array_equal(dense[ours][0], FE) and np.array_equal(dense[ours][1], FC), ref
    # distances / posterior / EI against the oracle
    y = rng.standard_normal(len(ev))
    ref = rn.acquisition_function(ev, y, trees, hyp, gen, d)
    eng = sot.AcquisitionEngine(XE, torch.from_numpy(y), gram, w, hyp.variance, hyp.lengthscale, hyp.noise_variance, hyp.mean_c)
    o = sot.sot_pairs(X, eng.state, w, hyp.variance, hyp.lengthscale, want=("K", "Df"))
    np.testing.assert_allclose(o["Df"].cpu().numpy(), rn.per_feature_distances(trees, ev, gen, d), rtol=0, atol=1e-12)
    np.testing.assert_allclose(o["K"].cpu().numpy(), rn.K(trees, ev, hyp, gen, d), rtol=1e-12)
    mu, sigma, ei = eng.acquire_device(X)
    np.testing.assert_allclose(mu.cpu().numpy(), ref["mu"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(sigma.cpu().numpy(), ref["sigma"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(ei.cpu().numpy(), ref["acq"], rtol=1e-6, atol=1e-12)
    # E = 1 and N = 1
    e1 = sot.AcquisitionEngine(XE[:1], torch.tensor([0.7], dtype=torch.float64), gram, w, hyp.variance, hyp.lengthscale, hyp.noise_variance, hyp.mean_c)
    r1 = rn.acquisition_function(ev[:1], np.array([0.7]), trees[:1] + trees[40:41], hyp, gen, d)
    m1, s1, a1 = e1.acquire_device(torch.cat([X[:1], X[40:41]]))
    np.testing.assert_allclose(m1.cpu().numpy(), r1["mu"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(s1.cpu().numpy(), r1["sigma"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(a1.cpu().numpy(), r1["acq"], rtol=1e-6, atol=1e-12)
    # empty candidate list
    m0, s0, a0 = eng.acquire_device(X[:0])
    assert m0.numel() == 0 and a0.numel() == 0
    assert sot.sot_pairs(X[:0], eng.state, w, 1.0, 1.0)["K"].shape == (0, len(ev))


def test_malformed_records_are_reported():
    gram = Grammar.get("CKSHighDimSearchSpace", 5)
    good = random_stress_trees(40, gram, seed=3)
    r = int(np.nonzero(good[:8, 0] >= 5)[0][0])  # a record with at least 5 nodes, inside the first 8
    st = sot.EvalState(sot.as_device_trees(good[:8], DEV), gram)

    def even_length(a):
        a[r, 0] = 4

    def zero_length(a):
        a[r, 0] = 0

    def unknown_operator(a):
        a[r, 1] = 0x85

    def symbol_out_of_range(a):
        a[r, int(a[r, 0])] = 60  # last token of a tree is always a leaf

    def too_few_operands(a):
        a[r, 0] = 3
        a[r, 1:4] = [0x80, 0x80, 1]

    def too_many_operands(a):
        a[r, 0] = 3
        a[r, 1:4] = [1, 2, 3]

    for mutate in (even_length, zero_length, unknown_operator, symbol_out_of_range, too_few_operands, too_many_operands):
        bad = good.copy()
        mutate(bad)
        with pytest.raises(ValueError, match="malformed"):
            sot.sot_pairs(sot.as_device_trees(bad, DEV), st, (0.3, 0.3, 0.4), 1.0, 1.0)
        o = sot.sot_pairs(sot.as_device_trees(bad, DEV), st, (0.3, 0.3, 0.4), 1.0, 1.0, check=False)
        K = o["K"].cpu().numpy()
        assert np.all(np.isnan(K[r])) and not np.any(np.isnan(np.delete(K, r, axis=0))), mutate.__name__
        with pytest.raises(ValueError, match="malformed"):
            sot.EvalState(sot.as_device_trees(bad[:8], DEV), gram)
        with pytest.raises(ValueError, match="malformed"):
            sot.dense_feature_matrices(sot.featurize(sot.as_device_trees(bad, DEV), gram))


def test_feature_subsets_and_weights():
    """feature_type_list may hold any subset (kernel_kernel_grammar_tree.py:91-108): a missing feature = weight 0."""
    gen, d = "CKSHighDimSearchSpace", 4
    gram = Grammar.get(gen, d)
    rng = np.random.default_rng(9)
    names = gram.symbols
    ev = [rn.random_stress_tree(rng, names, max_depth=4) for _ in range(16)]
    ca = [rn.random_stress_tree(rng, names, max_depth=4) for _ in range(50)]
    XE, XC = sot.as_device_trees(encode(ev, gram), DEV), sot.as_device_trees(encode(ca, gram), DEV)
    st = sot.EvalState(XE, gram)
    for feats, alphas in (((rn.SUBTREES,), (0.5,)), ((rn.PATHS, rn.SUBTREES), (0.9, 0.3)), ((rn.DIM_WISE,), (0.2,))):
        hyp = rn.Hyper(alphas=alphas, lengthscale=1.7, variance=0.6)
        wmap = dict(zip(feats, hyp.alphas / hyp.alphas.sum()))
        w = [wmap.get(f, 0.0) for f in (rn.DIM_WISE, rn.PATHS, rn.SUBTREES)]
        K = sot.sot_pairs(XC, st, w, hyp.variance, hyp.lengthscale)["K"].cpu().numpy()
        np.testing.assert_allclose(K, rn.K(ca, ev, hyp, gen, d, feature_types=feats), rtol=1e-12)


@pytest.mark.parametrize("n_cand,n_eval", [(1_000_000, 200)])
def test_full_size_properties(n_cand, n_eval):
    """BASELINE.json stress shape (1M trees x 200 evaluated): size-independent properties plus an
    oracle check on a random sample of rows."""
    gen, d = "CKSHighDimSearchSpace", 5
    gram = Grammar.get(gen, d)
    ev_np = random_stress_trees(n_eval, gram, seed=77)
    ca_np = random_stress_trees(n_cand, gram, seed=1234)
    rng = np.random.default_rng(1)
    planted = rng.choice(n_cand, size=n_eval, replace=False)
    ca_np[planted] = ev_np  # plant every evaluated tree once among the candidates
    XE, XC = sot.as_device_trees(ev_np, DEV), sot.as_device_trees(ca_np, DEV)
    hyp = rn.Hyper(**rn.H1)
    w = hyp.alphas / hyp.alphas.sum()
    y = rng.standard_normal(n_eval)
    eng = sot.AcquisitionEngine(XE, torch.from_numpy(y), gram, w, hyp.variance, hyp.lengthscale, hyp.noise_variance, hyp.mean_c)
    o = sot.sot_pairs(XC, eng.state, w, hyp.variance, hyp.lengthscale, want=("K", "D"))
    D, K = o["D"], o["K"]
    # 1. planted copies: distance exactly 0 to their source, Gram == variance
    idx = torch.from_numpy(planted).to(DEV)
    cols = torch.arange(n_eval, device=DEV)
    assert torch.all(D[idx, cols] == 0.0) and torch.all(K[idx, cols] == hyp.variance)
    # 2. ranges: 0 <= d <= sum_f w_f * max_f, 0 < K <= variance
    dmax = w[0] * 2 * d + w[1] * 2 + w[2] * 2
    assert float(D.min()) >= 0.0 and float(D.max()) <= dmax + 1e-12
    assert float(K.min()) > 0.0 and float(K.max()) <= hyp.variance
    # 3. determinism + permutation equivariance: shuffled candidates give shuffled rows, bit for bit
    perm = torch.randperm(n_cand, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
    mu, sigma, ei = eng.acquire_device(XC)
    mu_p, sigma_p, ei_p = eng.acquire_device(XC[perm].contiguous())
    assert torch.equal(mu[perm], mu_p) and torch.equal(sigma[perm], sigma_p) and torch.equal(ei[perm], ei_p)
    # 3b. the host-buffer entry point (chunked three-stream pipeline at this size) returns the same bits
    h_trees = torch.from_numpy(ca_np).pin_memory()
    h_acq, h_mu, h_sig = (torch.empty(n_cand, dtype=torch.float64).pin_memory() for _ in range(3))
    eng.acquire_host(h_trees, h_acq, h_mu, h_sig)
    assert torch.equal(h_acq, ei.cpu()) and torch.equal(h_mu, mu.cpu()) and torch.equal(h_sig, sigma.cpu())
    # 4. posterior at planted copies reproduces the posterior at the evaluated kernels
    assert torch.equal(mu[idx], eng.mu_data) and torch.equal(sigma[idx], eng.sigma_data)
    assert float(sigma.min()) > 0 and float(sigma.max()) <= np.sqrt(hyp.variance) * (1 + 1e-12)
    # 5. oracle on a random sample of rows
    sample = np.sort(rng.choice(n_cand, size=300, replace=False))
    from bosot_b200.encoding import decode_to_tuple

    ev_t = [decode_to_tuple(r, gram) for r in ev_np]
    ca_t = [decode_to_tuple(ca_np[i], gram) for i in sample]
    ref = rn.acquisition_function(ev_t, y, ca_t, hyp, gen, d)
    s_idx = torch.from_numpy(sample).to(DEV)
    np.testing.assert_allclose(K[s_idx].cpu().numpy(), rn.K(ca_t, ev_t, hyp, gen, d), rtol=1e-12)
    np.testing.assert_allclose(mu[s_idx].cpu().numpy(), ref["mu"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(sigma[s_idx].cpu().numpy(), ref["sigma"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(ei[s_idx].cpu().numpy(), ref["acq"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(float(eng.current_max), ref["current_max"], rtol=1e-9)


@pytest.mark.parametrize("n_eval", [300, 333, 600])
def test_large_and_odd_evaluated_sets(n_eval):
    """E beyond one register tile (256), odd E (TMA rows are padded to 16 bytes) and E that needs the smaller
    posterior tiles: the device path against itself through independent routes plus the oracle on a sample."""
    gen, d = "CKSHighDimSearchSpace", 5
    gram = Grammar.get(gen, d)
    ev_np, ca_np = random_stress_trees(n_eval, gram, seed=5), random_stress_trees(3000, gram, seed=6)
    ca_np[7] = ev_np[n_eval - 1]
    XE, XC = sot.as_device_trees(ev_np, DEV), sot.as_device_trees(ca_np, DEV)
    hyp = rn.Hyper(**rn.H1)
    w = hyp.alphas / hyp.alphas.sum()
    y = np.random.default_rng(2).standard_normal(n_eval)
    eng = sot.AcquisitionEngine(XE, torch.from_numpy(y), gram, w, hyp.variance, hyp.lengthscale, hyp.noise_variance, hyp.mean_c)
    mu, sigma, ei = eng.acquire_device(XC)
    # route 2: Gram from the pair kernel alone, posterior in torch fp64 (cuSOLVER / cuBLAS)
    K = sot.sot_pairs(XC, eng.state, w, hyp.variance, hyp.lengthscale)["K"]
    A = eng.K_EE + hyp.noise_variance * torch.eye(n_eval, dtype=torch.float64, device=DEV)
    L = torch.linalg.cholesky(A)
    V = torch.linalg.solve_triangular(L, K.t(), upper=False)
    var = hyp.variance - (V * V).sum(0)
    mean = K @ torch.cholesky_solve((torch.from_numpy(y).to(DEV) - hyp.mean_c).reshape(-1, 1), L)[:, 0] + hyp.mean_c
    np.testing.assert_allclose(mu.cpu().numpy(), mean.cpu().numpy(), rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(sigma.cpu().numpy(), torch.sqrt(var).cpu().numpy(), rtol=1e-7, atol=1e-10)
    assert float(K[7, n_eval - 1]) == hyp.variance
    # oracle on a sample of rows
    from bosot_b200.encoding import decode_to_tuple

    idx = np.arange(0, 3000, 100)
    ev_t = [decode_to_tuple(r, gram) for r in ev_np]
    ca_t = [decode_to_tuple(ca_np[i], gram) for i in idx]
    ref = rn.acquisition_function(ev_t, y, ca_t, hyp, gen, d)
    np.testing.assert_allclose(K[torch.from_numpy(idx).to(DEV)].cpu().numpy(), rn.K(ca_t, ev_t, hyp, gen, d), rtol=1e-12)
    np.testing.assert_allclose(mu.cpu().numpy()[idx], ref["mu"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(sigma.cpu().numpy()[idx], ref["sigma"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(ei.cpu().numpy()[idx], ref["acq"], rtol=1e-6, atol=1e-12)


def test_dim_wise_state_tables_and_dense_columns_agree():
    """DIM_WISE has two evaluation modes in the pair kernel: per-block tables over the distinct block vectors of the
    evaluated set (<= 32 per block) and the dense column loop (more).  Both against the oracle, and against each
    other bit for bit (D(A, B) computed with B as the evaluated set must equal D(B, A) transposed)."""
    gen, d = "CKSWithRQSearchSpace", 1  # one block with four kinds: deep trees give many distinct count vectors
    gram = Grammar.get(gen, d)
    rng = np.random.default_rng(11)
    big = [rn.random_stress_tree(rng, gram.symbols, max_depth=5, p_leaf=0.15) for _ in range(230)]
    small = [rn.random_stress_tree(rng, gram.symbols, max_depth=2) for _ in range(40)]
    XB, XS = sot.as_device_trees(encode(big, gram), DEV), sot.as_device_trees(encode(small, gram), DEV)
    st_big, st_small = sot.EvalState(XB, gram), sot.EvalState(XS, gram)
    assert st_big.info()["dim_wise_states"] == 0 and st_big.info()["max_states_per_block"] == 33
    assert st_small.info()["dim_wise_states"] == 1 and 1 <= st_small.info()["max_states_per_block"] <= 32
    w = np.array([0.2, 0.5, 0.3])
    a = sot.sot_pairs(XS, st_big, w, 1.3, 0.8, want=("K", "D", "Df"))    # dense columns
    b = sot.sot_pairs(XB, st_small, w, 1.3, 0.8, want=("K", "D", "Df"))  # state tables
    for key in ("K", "D"):
        assert np.array_equal(a[key].cpu().numpy(), b[key].cpu().numpy().T), key
    assert np.array_equal(a["Df"].cpu().numpy(), b["Df"].cpu().numpy().transpose(0, 2, 1))
    np.testing.assert_allclose(a["Df"].cpu().numpy(), rn.per_feature_distances(small, big, gen, d), rtol=0, atol=1e-12)
    np.testing.assert_allclose(b["Df"].cpu().numpy(), rn.per_feature_distances(big, small, gen, d), rtol=0, atol=1e-12)
    # self distances in both modes: exact zeros on the diagonal, bitwise symmetric
    for X, st in ((XB, st_big), (XS, st_small)):
        D = sot.sot_pairs(X, st, w, 1.3, 0.8, want=("D",))["D"].cpu().numpy()
        assert np.all(np.diag(D) == 0.0) and np.array_equal(D, D.T)


@pytest.mark.parametrize("gen,d", [("CKSWithRQSearchSpace", 6), ("CompositionalKernelSearchSpace", 8), ("CKSHighDimSearchSpace", 12)])
def test_large_grammars(gen, d):
    """More base kernels / dim-wise columns than the benchmark grammars (B up to 24, F up to 36)."""
    gram = Grammar.get(gen, d)
    rng = np.random.default_rng(3)
    ev = [rn.random_stress_tree(rng, gram.symbols, max_depth=4) for _ in range(40)]
    ca = [rn.random_stress_tree(rng, gram.symbols, max_depth=5) for _ in range(150)]
    ca[3] = ev[7]
    XE, XC = sot.as_device_trees(encode(ev, gram), DEV), sot.as_device_trees(encode(ca, gram), DEV)
    dense = sot.dense_feature_matrices(sot.featurize(XE, gram), sot.featurize(XC, gram))
    for ours, ref in FEATS:
        FE, FC = rn.dense_features(ev, ca, ref, gen, d)
        assert np.array_equal(dense[ours][0], FE) and np.array_equal(dense[ours][1], FC), ref
    hyp = rn.Hyper(**rn.H1)
    w = hyp.alphas / hyp.alphas.sum()
    y = rng.standard_normal(len(ev))
    eng = sot.AcquisitionEngine(XE, torch.from_numpy(y), gram, w, hyp.variance, hyp.lengthscale, hyp.noise_variance, hyp.mean_c)
    o = sot.sot_pairs(XC, eng.state, w, hyp.variance, hyp.lengthscale, want=("K", "Df"))
    np.testing.assert_allclose(o["Df"].cpu().numpy(), rn.per_feature_distances(ca, ev, gen, d), rtol=0, atol=1e-12)
    np.testing.assert_allclose(o["K"].cpu().numpy(), rn.K(ca, ev, hyp, gen, d), rtol=1e-12)
    ref = rn.acquisition_function(ev, y, ca, hyp, gen, d)
    mu, sigma, ei = eng.acquire_device(XC)
    np.testing.assert_allclose(mu.cpu().numpy(), ref["mu"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(sigma.cpu().numpy(), ref["sigma"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(ei.cpu().numpy(), ref["acq"], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("seed", list(range(40 + int(os.environ.get("BOSOT_EXTRA_FUZZ", "0")))))
def test_randomised_configurations(seed):
    """Differential fuzz: random search space, dimension, tree depth, set sizes and hyper-parameters
    (BOSOT_EXTRA_FUZZ=n adds n more seeded configurations)."""
    rng = np.random.default_rng(1000 + seed)
    gen = ["CKSWithRQSearchSpace", "CKSHighDimSearchSpace", "CompositionalKernelSearchSpace", "NDimFullKernelsSearchSpace"][seed % 4]
    d = int(rng.integers(1, 7))
    gram = Grammar.get(gen, d)
    n_eval, n_cand = int(rng.integers(1, 140)), int(rng.integers(1, 260))
    depth_e, depth_c = int(rng.integers(0, 6)), int(rng.integers(0, 6))
    p_leaf = float(rng.uniform(0.05, 0.6))
    ev = [rn.random_stress_tree(rng, gram.symbols, max_depth=depth_e, p_leaf=p_leaf) for _ in range(n_eval)]
    ca = [rn.random_stress_tree(rng, gram.symbols, max_depth=depth_c, p_leaf=p_leaf) for _ in range(n_cand)]
    for k in range(0, n_cand, 17):  # sprinkle exact copies and child-swapped copies of evaluated trees
        t = ev[int(rng.integers(n_eval))]
        ca[k] = t if k % 2 else (t if isinstance(t, str) else (t[0], t[2], t[1]))
    hyp = rn.Hyper(alphas=rng.uniform(0.05, 0.95, 3), lengthscale=float(rng.uniform(0.3, 3.0)), variance=float(rng.uniform(0.2, 3.0)),
                   noise_variance=float(10 ** rng.uniform(-5, -2)), mean_c=float(rng.normal()))
    w = hyp.alphas / hyp.alphas.sum()
    y = rng.standard_normal(n_eval)
    XE, XC = sot.as_device_trees(encode(ev, gram), DEV), sot.as_device_trees(encode(ca, gram), DEV)
    dense = sot.dense_feature_matrices(sot.featurize(XE, gram), sot.featurize(XC, gram))
    for ours, ref in FEATS:
        FE, FC = rn.dense_features(ev, ca, ref, gen, d)
        assert np.array_equal(dense[ours][0], FE) and np.array_equal(dense[ours][1], FC), (ref, gen, d)
    eng = sot.AcquisitionEngine(XE, torch.from_numpy(y), gram, w, hyp.variance, hyp.lengthscale, hyp.noise_variance, hyp.mean_c)
    o = sot.sot_pairs(XC, eng.state, w, hyp.variance, hyp.lengthscale, want=("K", "D", "Df"))
    np.testing.assert_allclose(o["Df"].cpu().numpy(), rn.per_feature_distances(ca, ev, gen, d), rtol=0, atol=1e-12)
    np.testing.assert_allclose(o["K"].cpu().numpy(), rn.K(ca, ev, hyp, gen, d), rtol=1e-12)
    ref = rn.acquisition_function(ev, y, ca, hyp, gen, d)
    mu, sigma, ei = eng.acquire_device(XC)
    # sqrt of a rounding-negative variance is NaN on either side: where the NaN patterns differ, the variance itself
    # (second, Cholesky-free restatement of the oracle) must be zero to rounding -- nothing else may differ
    ok_ref, ok_gpu = np.isfinite(ref["sigma"]), np.isfinite(sigma.cpu().numpy())
    ok = ok_ref & ok_gpu
    if not np.array_equal(ok_ref, ok_gpu):
        Kee_, Ken_ = rn.K(ev, None, hyp, gen, d), rn.K(ev, ca, hyp, gen, d)
        _, var2 = rn.gp_posterior_direct(Kee_, Ken_, np.full(n_cand, hyp.variance), y, hyp)
        assert np.all(np.abs(var2[ok_ref != ok_gpu]) <= 1e-9 * hyp.variance), var2[ok_ref != ok_gpu]
    # mu = K* beta + c cancels when the noise is tiny (beta ~ 1 / noise): the absolute floor follows the size of the
    # terms that are summed, 1e-12 relative to sum_j |K*_j beta_j| (both sides round differently at that level)
    Ks, Kee = rn.K(ca, ev, hyp, gen, d), rn.K(ev, ev, hyp, gen, d)
    beta = np.linalg.solve(Kee + hyp.noise_variance * np.eye(n_eval), y - hyp.mean_c)
    floor = 1e-10 + 1e-12 * (np.abs(Ks) @ np.abs(beta))
    assert np.all(np.abs(mu.cpu().numpy() - ref["mu"]) <= 1e-6 * np.abs(ref["mu"]) + floor)
    np.testing.assert_allclose(sigma.cpu().numpy()[ok], ref["sigma"][ok], rtol=1e-6, atol=1e-10)
    assert np.all(np.abs(ei.cpu().numpy()[ok] - ref["acq"][ok]) <= 1e-6 * np.abs(ref["acq"][ok]) + floor[ok])
