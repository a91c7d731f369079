"""CPU: flat encoding round trips, host-side validation, and the C-ABI library (loads, exports every
symbol include/bosot_b200.h declares; no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from bosot_b200 import _lib
from bosot_b200.encoding import Grammar, decode_to_name, decode_to_tuple, encode, random_stress_trees, validate
from oracle import ref_numpy as rn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_grammar_tables_follow_reference_order():
    g = Grammar.get("CKSWithRQSearchSpace", 2)
    assert g.symbols == rn.base_kernel_names("CKSWithRQSearchSpace", 2)
    # dim-wise columns: NULL_i then [RBF, PER, LIN, RQ] on dimension i (optimal_transport_mappings.py:58-69)
    assert g.dim_col_names == list(rn.dimwise_dict("RBFWithPrior_on_0", "CKSWithRQSearchSpace", 2).keys())
    g5 = Grammar.get("CKSHighDimSearchSpace", 5)
    assert g5.n_symbols == 10 and g5.n_dim_cols == 15 and g5.n_blocks == 5
    gn = Grammar.get("NDimFullKernelsSearchSpace", 3)
    assert gn.n_blocks == 1 and gn.dim_col_names == list(rn.dimwise_dict("RBFWithPrior", "NDimFullKernelsSearchSpace", 3).keys())
    with pytest.raises(NotImplementedError):
        Grammar.get("SomethingElse", 2)  # optimal_transport_mappings.py:85-86


def test_encode_roundtrip_names_tuples(golden):
    g = Grammar.get(str(golden["generator_name"]), int(golden["input_dimension"]))
    names = [str(s) for s in golden["cand_names"]]
    enc = encode(names, g)
    validate(enc, g)
    assert enc.shape == (len(names), 64) and enc.dtype == np.uint8
    assert [decode_to_name(r, g) for r in enc] == names
    tuples = [rn.parse_name(s) for s in names]
    assert np.array_equal(encode(tuples, g), enc)
    assert [decode_to_tuple(r, g) for r in enc] == tuples
    assert np.all(enc[:, 0] == [2 * rn.count_leaves(t) - 1 for t in tuples])


def test_encode_roundtrip_random_trees_property():
    """Property test (hypothesis): for random binary trees over every search space, record -> name -> record and
    record -> tuple -> record are identities, the node-count byte is 2 * leaves - 1, and the oracle's own parser
    agrees with the product's printer."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    spaces = [("CKSWithRQSearchSpace", 2), ("CKSHighDimSearchSpace", 5), ("CompositionalKernelSearchSpace", 3), ("NDimFullKernelsSearchSpace", 2)]

    def trees(symbols):
        leaf = st.sampled_from(symbols)
        return st.recursive(leaf, lambda kids: st.tuples(st.sampled_from(["ADD", "MULTIPLY"]), kids, kids), max_leaves=32)

    @settings(max_examples=150, deadline=None)
    @given(st.data())
    def run(data):
        gen, d = data.draw(st.sampled_from(spaces))
        g = Grammar.get(gen, d)
        t = data.draw(trees(g.symbols))
        if rn.count_leaves(t) > 32:
            with pytest.raises(ValueError):
                encode([t], g)
            return
        rec = encode([t], g)
        validate(rec, g)
        assert rec[0, 0] == 2 * rn.count_leaves(t) - 1
        assert decode_to_tuple(rec[0], g) == t
        name = decode_to_name(rec[0], g)
        assert rn.parse_name(name) == t and np.array_equal(encode([name], g), rec)

    run()


def test_encode_rejects_bad_input():
    g = Grammar.get("CKSHighDimSearchSpace", 2)
    with pytest.raises(ValueError):
        encode(["LinearWithPrior_on_0"], g)  # not a base kernel of this space
    big = "RBFWithPrior_on_0"
    for _ in range(32):
        big = "(" + big + " ADD RQWithPrior_on_1)"
    with pytest.raises(ValueError):
        encode([big], g)  # 33 leaves > 32
    ok = encode(["(RBFWithPrior_on_0 ADD RQWithPrior_on_1)"], g)
    for mutate in (lambda a: a.__setitem__((0, 0), 2), lambda a: a.__setitem__((0, 1), 0), lambda a: a.__setitem__((0, 2), 0x82), lambda a: a.__setitem__((0, 3), 77)):
        bad = ok.copy()
        mutate(bad)
        with pytest.raises(ValueError):
            validate(bad, g)


def test_random_stress_trees_are_valid_and_shaped_like_c5():
    g = Grammar.get("CKSHighDimSearchSpace", 5)
    t = random_stress_trees(20000, g, seed=1234)
    validate(t, g)
    leaves = (t[:, 0].astype(int) + 1) // 2
    assert leaves.max() <= 32 and 10.5 < leaves.mean() < 13.5  # mean 11.9
    assert np.array_equal(t, random_stress_trees(20000, g, seed=1234))
    depth_ok = [rn.count_leaves(decode_to_tuple(r, g)) == l for r, l in zip(t[:50], leaves[:50])]
    assert all(depth_ok)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "bosot_b200.h")).read()
    declared = set(re.findall(r"\b(bosot_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    loaded = _lib.load()
    assert loaded.bosot_abi_version() == _lib.ABI_VERSION == 2
    assert loaded.bosot_eval_state_bytes(0) == 0 and loaded.bosot_eval_state_bytes(50) > 0
    assert loaded.bosot_acquire_work_bytes(10, 50) >= 10 * (56 * 8 + 64)


def test_no_cpu_fallback():
    torch = pytest.importorskip("torch")
    from bosot_b200 import sot

    g = Grammar.get("CKSHighDimSearchSpace", 5)
    cpu_trees = torch.from_numpy(random_stress_trees(4, g, seed=0))
    with pytest.raises(RuntimeError, match="no CPU"):
        sot.featurize(cpu_trees, g)
    with pytest.raises(RuntimeError, match="no CPU"):
        sot.EvalState(cpu_trees, g)
