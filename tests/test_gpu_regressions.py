"""GPU regression tests for defects found in review of round 1 (ADVICE.md / VERDICT.md):

* posterior_ws_kernel: a unit of CTA tile i + 2 could pass the stage's mbarrier before tile i had landed when a tile
  has few units (small evaluated sets) and a CTA owns many tiles (large pools) -- the grid E x N below;
* the padding k-steps of the posterior's software pipeline multiplied structural zeros of L^-1 with the NEXT
  candidate's K* row, so a malformed (NaN-poisoned) record turned its neighbour's sigma / EI into NaN;
* malformed candidates were invisible to callers of the composed acquire_* entry points;
* the first call on a non-blocking stream raced with a synchronous table upload (tables are static data now);
* the row-sharded scatter entry points had no test that runs on ONE GPU.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from bosot_b200 import _lib, sot  # noqa: E402
from bosot_b200.encoding import Grammar, decode_to_tuple, random_stress_trees  # noqa: E402
from oracle import ref_numpy as rn  # noqa: E402

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
GEN, DIM = "CKSHighDimSearchSpace", 5


def _engine(n_eval, seed=5, acq_kind=_lib.ACQ_EI):
    gram = Grammar.get(GEN, DIM)
    ev_np = random_stress_trees(n_eval, gram, seed=seed)
    hyp = rn.Hyper(**rn.H1)
    w = hyp.alphas / hyp.alphas.sum()
    y = np.random.default_rng(seed).standard_normal(n_eval)
    eng = sot.AcquisitionEngine(sot.as_device_trees(ev_np, DEV), torch.from_numpy(y), gram, w, hyp.variance, hyp.lengthscale,
                                hyp.noise_variance, hyp.mean_c, acq_kind=acq_kind)
    return gram, ev_np, y, hyp, eng


@pytest.mark.parametrize("n_cand", [40_000, 200_000, 1_000_000])
@pytest.mark.parametrize("n_eval", [1, 2, 8, 16, 17, 32, 33, 48])
def test_posterior_small_evaluated_sets_large_pools(n_eval, n_cand):
    """bosot_posterior_acq on a consistent synthetic problem (Laplace kernel on random points, so K_EE is positive
    definite and the posterior variance is positive) against torch fp64 (cuSOLVER / cuBLAS), run twice: every
    value right, and the same bits both times."""
    g = torch.Generator(device=DEV).manual_seed(1000 * n_eval + n_cand % 977)
    variance, noise, c = 1.3, 1e-3, -0.5
    Xe = torch.rand((n_eval, 6), dtype=torch.float64, device=DEV, generator=g)
    Xc = torch.rand((n_cand, 6), dtype=torch.float64, device=DEV, generator=g)
    y = torch.randn(n_eval, dtype=torch.float64, device=DEV, generator=g)
    K_EE = variance * torch.exp(-torch.cdist(Xe, Xe, p=1.0))
    ld = (n_eval + 7) // 8 * 8
    Kbuf = torch.zeros((n_cand, ld), dtype=torch.float64, device=DEV)
    Kbuf[:, :n_eval] = variance * torch.exp(-torch.cdist(Xc, Xe, p=1.0))
    Kstar = Kbuf[:, :n_eval]  # row stride ld, like the work buffer of the composed calls
    Linv, beta = sot.gp_factorize(K_EE, noise, y - c)
    L = torch.linalg.cholesky(K_EE + noise * torch.eye(n_eval, dtype=torch.float64, device=DEV))
    V = torch.linalg.solve_triangular(L, Kstar.t(), upper=False)
    var = variance - (V * V).sum(0)
    assert float(var.min()) > 0
    sg_ref = torch.sqrt(var)
    mu_ref = Kstar @ torch.cholesky_solve((y - c).reshape(-1, 1), L)[:, 0] + c
    cur = mu_ref[:7].max().reshape(1).contiguous()
    d = mu_ref - cur - 0.01
    z = d / sg_ref
    ei_ref = d * 0.5 * torch.erfc(-z / np.sqrt(2.0)) + sg_ref * torch.exp(-0.5 * z * z) / np.sqrt(2 * np.pi)
    outs = []
    for _ in range(2):
        mu, sigma, ei = sot.posterior_acq(Kstar, Linv, beta, c, variance, _lib.ACQ_EI, cur)
        outs.append((mu, sigma, ei))
    torch.cuda.synchronize()
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    mu, sigma, ei = outs[0]
    scale = float((Kstar.abs() @ beta.abs()).max())
    assert float((mu - mu_ref).abs().max()) <= 1e-12 * scale + 1e-9 * float(mu_ref.abs().max())
    torch.testing.assert_close(sigma, sg_ref, rtol=1e-7, atol=1e-10)
    assert float(((ei - ei_ref).abs() - 1e-6 * ei_ref.abs()).max()) <= 1e-10 * scale


@pytest.mark.parametrize("n_eval,n_cand", [(8, 200_000), (24, 60_000), (48, 40_000)])
def test_full_path_small_evaluated_sets_large_pools(n_eval, n_cand):
    """The whole acquire_device path at the sizes of the defect (few evaluated kernels, many candidates -- every BO run
    starts like this) against the oracle on a sample of rows, twice for determinism."""
    gram, ev_np, y, hyp, eng = _engine(n_eval)
    ca_np = random_stress_trees(n_cand, gram, seed=6)
    ca_np[11] = ev_np[n_eval - 1]
    XC = sot.as_device_trees(ca_np, DEV)
    a = eng.acquire_device(XC)
    b = eng.acquire_device(XC)
    torch.cuda.synchronize()
    for x, z in zip(a, b):
        assert torch.equal(x, z)
    assert eng.bad_tree_count() == 0
    mu, sigma, ei = a
    assert torch.equal(mu[11], eng.mu_data[n_eval - 1]) and torch.equal(sigma[11], eng.sigma_data[n_eval - 1])
    idx = np.sort(np.random.default_rng(0).choice(n_cand, size=200, replace=False))
    ev_t = [decode_to_tuple(r, gram) for r in ev_np]
    ca_t = [decode_to_tuple(ca_np[i], gram) for i in idx]
    ref = rn.acquisition_function(ev_t, y, ca_t, hyp, GEN, DIM)
    np.testing.assert_allclose(mu.cpu().numpy()[idx], ref["mu"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(sigma.cpu().numpy()[idx], ref["sigma"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(ei.cpu().numpy()[idx], ref["acq"], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("n_eval", [200, 37, 6])
def test_malformed_record_poisons_only_its_own_row(n_eval):
    gram, ev_np, y, hyp, eng = _engine(n_eval)
    n = 3000
    good = random_stress_trees(n, gram, seed=9)
    XC = sot.as_device_trees(good, DEV)
    clean = [t.clone() for t in eng.acquire_device(XC)]
    assert eng.bad_tree_count() == 0
    for r in (0, 1, 63, 64, 1500, n - 1):
        bad = good.copy()
        bad[r, 0] = 4  # even node count
        got = eng.acquire_device(sot.as_device_trees(bad, DEV))
        torch.cuda.synchronize()
        assert eng.bad_tree_count() == 1
        with pytest.raises(ValueError, match="malformed"):
            eng.raise_on_bad_trees()
        keep = torch.ones(n, dtype=torch.bool, device=DEV)
        keep[r] = False
        for g_, c_ in zip(got, clean):
            assert bool(torch.isnan(g_[r])), r
            assert torch.equal(g_[keep], c_[keep]), r
        # host entry point: reports the count, raises when it synchronises
        h_trees = torch.from_numpy(bad).pin_memory()
        h_acq = torch.empty(n, dtype=torch.float64).pin_memory()
        with pytest.raises(ValueError, match="malformed"):
            eng.acquire_host(h_trees, h_acq)
        assert torch.equal(h_acq[keep.cpu()], clean[2].cpu()[keep.cpu()]) and bool(torch.isnan(h_acq[r]))
    two = good.copy()
    two[5, 0] = 0
    two[2000, 1] = 0x85
    eng.acquire_device(sot.as_device_trees(two, DEV))
    assert eng.bad_tree_count() == 2
    eng.acquire_device(XC)
    assert eng.bad_tree_count() == 0  # the count is per call


def test_first_call_may_run_on_a_non_blocking_stream():
    """No table upload precedes the first launch any more: a fresh side stream gives the same bits."""
    gram, ev_np, y, hyp, eng = _engine(50)
    XC = sot.as_device_trees(random_stress_trees(5000, gram, seed=10), DEV)
    side = torch.cuda.Stream(device=DEV)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        a = eng.acquire_device(XC)
    side.synchronize()
    b = eng.acquire_device(XC)
    torch.cuda.synchronize()
    for x, z in zip(a, b):
        assert torch.equal(x, z)


@pytest.mark.parametrize("n_cand,n_eval", [(5000, 50), (450_000, 24)])
def test_scatter_entry_points_on_one_gpu(n_cand, n_eval):
    """bosot_acquire_device_scatter / bosot_acquire_host_scatter with n_peers = 2 where both "peer" vectors live on
    this GPU and the shard starts at a non-zero offset: both vectors must hold the very bits that the plain entry
    point returns, at the shard's rows and nowhere else (450k candidates take the chunked host pipeline)."""
    gram, ev_np, y, hyp, eng = _engine(n_eval)
    ca_np = random_stress_trees(n_cand, gram, seed=12)
    XC = sot.as_device_trees(ca_np, DEV)
    mu0, sg0, ei0 = [t.clone() for t in eng.acquire_device(XC)]
    off, tail = 12_345, 77
    peers = [torch.full((off + n_cand + tail,), -7.0, dtype=torch.float64, device=DEV) for _ in range(2)]
    table = torch.tensor([p.data_ptr() for p in peers], dtype=torch.int64, device=DEV)
    mu, sg, ei = eng.acquire_device(XC, scatter=(table.data_ptr(), 2, off))
    torch.cuda.synchronize()
    assert torch.equal(mu, mu0) and torch.equal(sg, sg0) and torch.equal(ei, ei0)
    for p in peers:
        assert torch.equal(p[off:off + n_cand], ei0)
        assert bool((p[:off] == -7.0).all()) and bool((p[off + n_cand:] == -7.0).all())
        p.fill_(-7.0)
    h_trees = torch.from_numpy(ca_np).pin_memory()
    h_acq = torch.empty(n_cand, dtype=torch.float64).pin_memory()
    eng.acquire_host(h_trees, h_acq, scatter=(table.data_ptr(), 2, off))
    assert torch.equal(h_acq, ei0.cpu())
    for p in peers:
        assert torch.equal(p[off:off + n_cand], ei0)
        assert bool((p[:off] == -7.0).all()) and bool((p[off + n_cand:] == -7.0).all())
        p.fill_(-7.0)
    eng.acquire_host(h_trees, None, scatter=(table.data_ptr(), 2, off))  # device-side consumers only: no D2H at all
    for p in peers:
        assert torch.equal(p[off:off + n_cand], ei0)
    # oracle on a sample of the scattered vector
    idx = np.sort(np.random.default_rng(1).choice(n_cand, size=100, replace=False))
    ref = rn.acquisition_function([decode_to_tuple(r, gram) for r in ev_np], y, [decode_to_tuple(ca_np[i], gram) for i in idx], hyp, GEN, DIM)
    np.testing.assert_allclose(peers[1][off:off + n_cand].cpu().numpy()[idx], ref["acq"], rtol=1e-6, atol=1e-12)


def test_shutdown_releases_the_host_pipeline_and_the_library_stays_usable():
    gram, ev_np, y, hyp, eng = _engine(16)
    n = 420_000  # chunked pipeline
    h_trees = torch.from_numpy(random_stress_trees(n, gram, seed=13)).pin_memory()
    a = torch.empty(n, dtype=torch.float64).pin_memory()
    b = torch.empty(n, dtype=torch.float64).pin_memory()
    eng.acquire_host(h_trees, a)
    assert _lib.load().bosot_shutdown() == 0
    assert _lib.load().bosot_shutdown() == 0  # idempotent
    eng.acquire_host(h_trees, b)
    assert torch.equal(a, b)


@pytest.mark.parametrize("n_eval", [50, 200])
def test_fused_scan_and_scan_launch_paths_agree_bit_for_bit(n_eval):
    """Pools of up to 16384 candidates are scanned inside the pair kernel (warp-cooperative closed-form scan), larger
    ones by the thread-per-tree scan launch: the same candidates through both routes -- alone, and as the head of a
    pool large enough for the scan launch -- must give the same bits, defect codes of malformed records included."""
    gram, ev_np, y, hyp, eng = _engine(n_eval)
    big = random_stress_trees(40_000, gram, seed=21)
    names = gram.symbols
    from bosot_b200.encoding import encode

    chains = [names[0]]
    for k in range(2, 33):  # left- and right-deep chains up to 63 nodes: the deepest level structure the scan can meet
        t = names[k % len(names)]
        for i in range(1, k):
            t = ("ADD" if (i * k) % 3 else "MULTIPLY", t, names[i % len(names)]) if k % 2 else ("MULTIPLY" if i % 2 else "ADD", names[i % len(names)], t)
        chains.append(t)
    big[100:100 + len(chains)] = encode(chains, gram)
    big[7, 0] = 4                      # even length
    big[8, 0] = 0                      # zero length
    big[9, :4] = [3, 0x85, 1, 2]       # unknown operator
    big[10, :4] = [3, 0x80, 70, 1]     # symbol out of range
    big[11, 0] = 65                    # too long
    big[12, :4] = [3, 0x80, 0x80, 1]  # too few operands
    big[13, :4] = [3, 1, 2, 3]        # too many operands
    small = big[:6000]
    w = hyp.alphas / hyp.alphas.sum()
    a = sot.sot_pairs(sot.as_device_trees(small, DEV), eng.state, w, hyp.variance, hyp.lengthscale, want=("K", "D", "Df"), check=False)
    b = sot.sot_pairs(sot.as_device_trees(big, DEV), eng.state, w, hyp.variance, hyp.lengthscale, want=("K", "D", "Df"), check=False)
    for key in ("K", "D", "Df"):
        x, z = a[key], b[key][..., :6000, :] if key == "Df" else b[key][:6000]
        assert torch.equal(torch.isnan(x), torch.isnan(z)), key
        assert torch.equal(torch.nan_to_num(x), torch.nan_to_num(z)), key
    assert bool(torch.isnan(a["K"][7:14]).all()) and not bool(torch.isnan(a["K"][14:]).any())
    lib = _lib.load()
    import ctypes as C

    def status_of(trees_np):
        t = sot.as_device_trees(trees_np, DEV)
        n = t.shape[0]
        st = torch.empty(n, dtype=torch.int32, device=DEV)
        K = torch.empty((n, n_eval), dtype=torch.float64, device=DEV)
        scratch = torch.empty(lib.bosot_sot_pairs_scratch_bytes(n), dtype=torch.uint8, device=DEV)
        _lib.check(lib.bosot_sot_pairs(t.data_ptr(), n, C.byref(gram.c_struct), eng.state.blob.data_ptr(), n_eval, (C.c_double * 3)(*w), 1.0, 1.0,
                                       K.data_ptr(), None, None, n_eval, st.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                       torch.cuda.current_stream().cuda_stream))
        return st.cpu().numpy()

    s_small, s_big = status_of(small), status_of(big)
    assert np.array_equal(s_small, s_big[:6000])
    assert list(np.nonzero(s_small)[0]) == [7, 8, 9, 10, 11, 12, 13]
    mu_s, sg_s, ei_s = eng.acquire_device(sot.as_device_trees(small[14:], DEV))
    mu_b, sg_b, ei_b = eng.acquire_device(sot.as_device_trees(big[14:], DEV))
    assert torch.equal(mu_s, mu_b[:5986]) and torch.equal(sg_s, sg_b[:5986]) and torch.equal(ei_s, ei_b[:5986])
