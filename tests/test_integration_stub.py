"""INTEGRATION.md section 2 is executable documentation: the ctypes stub a maintainer of the reference would add is
extracted from the markdown and run.  CPU: it loads the in-tree library and every prototype it declares is checked
against the header (through ``bosot_b200._lib.SIGNATURES``, which ``test_library_exports_every_declared_symbol`` ties to
``include/bosot_b200.h``).  GPU: ``sot_gram`` and ``Acquisition`` of the stub against the package and the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from bosot_b200 import _lib
from bosot_b200.encoding import Grammar, decode_to_tuple, random_stress_trees

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source() -> str:
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    mine = [b for b in blocks if b.startswith("# bosot/kernels/_b200.py")]
    assert len(mine) == 1, "INTEGRATION.md must hold exactly one `# bosot/kernels/_b200.py` block"
    return mine[0]


@pytest.fixture(scope="module")
def stub():
    pytest.importorskip("torch")
    os.environ["BOSOT_B200_LIB"] = _lib.LIB_PATH
    ns = {"__name__": "bosot.kernels._b200"}
    exec(compile(_stub_source(), "INTEGRATION.md::_b200.py", "exec"), ns)
    return ns


def _kind(t):
    """Coarse ABI class of a ctypes type: what matters for the call is width and pointer-ness."""
    if t is None:
        return "void"
    if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or issubclass(t, ctypes._Pointer):
        return "ptr"
    return "%s%d" % ("f" if t in (ctypes.c_double, ctypes.c_float) else "i", ctypes.sizeof(t))


def test_stub_prototypes_match_the_header(stub):
    protos = stub["PROTOTYPES"]
    assert {"bosot_eval_build", "bosot_sot_pairs", "bosot_pack_linv", "bosot_posterior_acq", "bosot_acquire_host"} <= set(protos)
    for name, (res, args) in protos.items():
        assert name in _lib.SIGNATURES, name
        ref_res, ref_args = _lib.SIGNATURES[name]
        assert _kind(res) == _kind(ref_res), name
        assert [_kind(a) for a in args] == [_kind(a) for a in ref_args], name
    # the struct mirrors agree byte for byte
    assert ctypes.sizeof(stub["Grammar"]) == ctypes.sizeof(_lib.BosotGrammar)
    assert [(n, ctypes.sizeof(t)) for n, t in stub["Grammar"]._fields_] == [(n, ctypes.sizeof(t)) for n, t in _lib.BosotGrammar._fields_]
    # every call in the stub passes as many arguments as its prototype declares
    src = _stub_source()
    for name, (_, args) in protos.items():
        for m in re.finditer(r"lib\." + name + r"\(", src):
            depth, i, n_args, any_arg = 1, m.end(), 0, False
            while depth:
                ch = src[i]
                depth += ch in "([{"
                depth -= ch in ")]}"
                if ch == "," and depth == 1:
                    n_args += 1
                elif depth >= 1 and not ch.isspace():
                    any_arg = True
                i += 1
            assert n_args + (1 if any_arg else 0) == len(args), (name, src[m.start():i])


@pytest.mark.gpu
def test_stub_runs_and_matches_the_package(stub):
    import torch

    from bosot_b200 import sot
    from oracle import ref_numpy as rn

    dev = torch.device("cuda:0")
    gen, d = "CKSHighDimSearchSpace", 5
    gram = Grammar.get(gen, d)
    g = stub["Grammar"].from_buffer_copy(bytes(gram.c_struct))
    ev_np, ca_np = random_stress_trees(37, gram, seed=1), random_stress_trees(3000, gram, seed=2)
    hyp = rn.Hyper(**rn.H1)
    y = np.random.default_rng(0).standard_normal(37)
    XE, XC = sot.as_device_trees(ev_np, dev), sot.as_device_trees(ca_np, dev)
    # K(X, X2) of the stub == the package's pair kernel, bit for bit; == the oracle to 1e-12
    K = stub["sot_gram"](XC, XE, g, hyp.alphas, hyp.variance, hyp.lengthscale)
    w = hyp.alphas / hyp.alphas.sum()
    K_pkg = sot.sot_pairs(XC, sot.EvalState(XE, gram), w, hyp.variance, hyp.lengthscale)["K"]
    assert torch.equal(K, K_pkg)
    ev_t = [decode_to_tuple(r, gram) for r in ev_np]
    ca_t = [decode_to_tuple(r, gram) for r in ca_np[:200]]
    np.testing.assert_allclose(K[:200].cpu().numpy(), rn.K(ca_t, ev_t, hyp, gen, d), rtol=1e-12)
    # the acquisition object of the stub == AcquisitionEngine == the oracle
    acq = stub["Acquisition"](XE, torch.from_numpy(y), g, hyp.alphas, hyp.variance, hyp.lengthscale, hyp.noise_variance, hyp.mean_c)
    h_acq = acq(torch.from_numpy(ca_np).pin_memory())
    eng = sot.AcquisitionEngine(XE, torch.from_numpy(y), gram, w, hyp.variance, hyp.lengthscale, hyp.noise_variance, hyp.mean_c)
    assert torch.equal(h_acq, eng.acquire_device(XC)[2].cpu())
    ref = rn.acquisition_function(ev_t, y, ca_t, hyp, gen, d)
    np.testing.assert_allclose(h_acq.numpy()[:200], ref["acq"], rtol=1e-6, atol=1e-12)
    bad = ca_np.copy()
    bad[17, 0] = 2
    with pytest.raises(ValueError, match="malformed"):
        acq(torch.from_numpy(bad).pin_memory())
