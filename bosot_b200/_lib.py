"""ctypes binding of the C-ABI in include/bosot_b200.h.

There is deliberately no fallback: if ``libbosot_b200.so`` is missing the import of any
compute entry point fails loudly (build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C bosot_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbosot_b200.so")

MAX_SYMBOLS = 64
MAX_BLOCKS = 16
MAX_DIM_COLS = 80
PATH_CODES = 64
TREE_BYTES = 64
MAX_NODES = 63
MAX_LEAVES = 32
MAX_EVAL = 4095
OP_ADD = 0x80
OP_MUL = 0x81
PAD = 0xFF
ACQ_NONE, ACQ_EI, ACQ_UCB = 0, 1, 2
ERAW = -4  # bosot_host_*: block of raw generator outputs too short
ABI_VERSION = 2


class BosotGrammar(C.Structure):
    """Mirror of ``struct bosot_grammar``."""

    _fields_ = [
        ("n_symbols", C.c_int32),
        ("n_blocks", C.c_int32),
        ("n_dim_cols", C.c_int32),
        ("reserved", C.c_int32),
        ("sym_block", C.c_int8 * MAX_SYMBOLS),
        ("sym_col", C.c_int8 * MAX_SYMBOLS),
        ("block_null_col", C.c_int8 * MAX_BLOCKS),
    ]


class BosotHellingerSpace(C.Structure):
    """Mirror of ``struct bosot_hellinger_space``."""

    _fields_ = [
        ("n_symbols", C.c_int32),
        ("input_dimension", C.c_int32),
        ("n_virtual", C.c_int32),
        ("n_samples", C.c_int32),
        ("n_sobol_dims", C.c_int32),
        ("n_priors", C.c_int32),
        ("sym_kind", C.c_uint8 * MAX_SYMBOLS),
        ("sym_dim", C.c_int8 * MAX_SYMBOLS),
        ("prior_of", (C.c_uint8 * 3) * 4),
    ]


HK_RBF, HK_LINEAR, HK_PERIODIC, HK_RQ = 0, 1, 2, 3
HELLINGER_MAX_VIRTUAL = 20
HELLINGER_MAX_DIM = 16

# every symbol include/bosot_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_D3 = C.POINTER(C.c_double)
_G = C.POINTER(BosotGrammar)
SIGNATURES = {
    "bosot_last_error": (C.c_char_p, []),
    "bosot_abi_version": (C.c_int, []),
    "bosot_shutdown": (C.c_int, []),
    "bosot_host_pipeline_config": (C.c_int, [C.c_int64, C.c_int64]),
    "bosot_featurize": (C.c_int, [_P, C.c_int64, _G] + [_P] * 11 + [_P]),
    "bosot_eval_state_bytes": (C.c_size_t, [C.c_int]),
    "bosot_eval_build": (C.c_int, [_P, C.c_int, _G, _P, C.c_size_t, _P, _P]),
    "bosot_eval_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), _P]),
    "bosot_sot_pairs_scratch_bytes": (C.c_size_t, [C.c_int64]),
    "bosot_sot_pairs": (C.c_int, [_P, C.c_int64, _G, _P, C.c_int, _D3, C.c_double, C.c_double, _P, _P, _P, C.c_int64, _P, _P, C.c_size_t, _P]),
    "bosot_posterior_acq": (
        C.c_int,
        [_P, C.c_int64, C.c_int64, C.c_int, _P, C.c_int, _P, C.c_double, C.c_double, C.c_int, _P, C.c_double, C.c_double, _P, _P, _P, _P],
    ),
    "bosot_sot_pairs_f32": (C.c_int, [_P, C.c_int64, _G, _P, C.c_int, _D3, C.c_double, C.c_double, _P, _P, _P, C.c_int64, _P, _P, C.c_size_t, _P]),
    "bosot_posterior_acq_f32": (
        C.c_int,
        [_P, C.c_int64, C.c_int64, C.c_int, _P, C.c_int, _P, C.c_double, C.c_double, C.c_int, _P, C.c_double, C.c_double, _P, _P, _P, _P],
    ),
    "bosot_acquire_device_f32": (
        C.c_int,
        [_P, C.c_int64, _G, _P, C.c_int, _D3, C.c_double, C.c_double, _P, C.c_int, _P, C.c_double, C.c_int, _P, C.c_double, C.c_double,
         _P, C.c_size_t, _P, _P, _P, _P, _P],
    ),
    "bosot_pack_linv": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P]),
    "bosot_gp_factorize": (C.c_int, [_P, C.c_int, C.c_int64, C.c_double, _P, _P, C.c_int, _P, _P, _P]),
    "bosot_max_f64": (C.c_int, [_P, C.c_int64, _P, _P]),
    "bosot_acquire_work_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "bosot_acquire_host": (
        C.c_int,
        [_P, C.c_int64, _G, _P, C.c_int, _D3, C.c_double, C.c_double, _P, C.c_int, _P, C.c_double, C.c_int, _P, C.c_double, C.c_double,
         _P, C.c_size_t, _P, _P, _P, _P, _P],
    ),
    "bosot_acquire_device": (
        C.c_int,
        [_P, C.c_int64, _G, _P, C.c_int, _D3, C.c_double, C.c_double, _P, C.c_int, _P, C.c_double, C.c_int, _P, C.c_double, C.c_double,
         _P, C.c_size_t, _P, _P, _P, _P, _P],
    ),
    "bosot_acquire_host_scatter": (
        C.c_int,
        [_P, C.c_int64, _G, _P, C.c_int, _D3, C.c_double, C.c_double, _P, C.c_int, _P, C.c_double, C.c_int, _P, C.c_double, C.c_double,
         _P, C.c_size_t, _P, _P, _P, _P, _P, C.c_int, C.c_int64, _P],
    ),
    "bosot_enable_peer_access": (C.c_int, [C.c_int]),
    "bosot_posterior_acq_scatter": (
        C.c_int,
        [_P, C.c_int64, C.c_int64, C.c_int, _P, C.c_int, _P, C.c_double, C.c_double, C.c_int, _P, C.c_double, C.c_double, _P, _P, _P,
         _P, C.c_int, C.c_int64, _P],
    ),
    "bosot_acquire_device_scatter": (
        C.c_int,
        [_P, C.c_int64, _G, _P, C.c_int, _D3, C.c_double, C.c_double, _P, C.c_int, _P, C.c_double, C.c_int, _P, C.c_double, C.c_double,
         _P, C.c_size_t, _P, _P, _P, _P, _P, C.c_int, C.c_int64, _P],
    ),
    "bosot_topk_work_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "bosot_topk_f64": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P, C.c_size_t, _P]),
    "bosot_argmax_work_bytes": (C.c_size_t, []),
    "bosot_argmax_f64": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, C.c_size_t, _P]),
    "bosot_neighbours": (C.c_int, [_P, C.c_int64, _P, C.c_int, _P, _P, _P]),
    "bosot_host_choices": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, C.POINTER(C.c_int64)]),
    "bosot_host_recursive_dataset": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int64, C.c_int, _P, C.POINTER(C.c_int64)]),
    "bosot_host_choices_cb": (C.c_int, [_P, _P, _P, C.c_int64, _P, C.POINTER(C.c_int64)]),
    "bosot_host_recursive_dataset_cb": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int, _P, C.POINTER(C.c_int64)]),
    "bosot_gp_fit_max_eval": (C.c_int, []),
    "bosot_gp_nll_grad": (C.c_int, [_P, C.c_int, C.c_int64, _P, _D3, C.c_double, C.c_double, C.c_double, C.c_double, _P, _P]),
    "bosot_hellinger_record_doubles": (C.c_int, [C.c_int]),
    "bosot_hellinger_grams": (C.c_int, [_P, C.c_int64, C.POINTER(BosotHellingerSpace), _P, _P, _P, _P, _P, _P]),
    "bosot_hellinger_pairs": (C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, _P, _P, C.c_int64, _P, _P]),
    "bosot_fp64_fma_probe": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "bosot_fp64_mma_probe": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "bosot_fp64_mma_probe2": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "bosot_kernel_launch_count": (C.c_int64, []),
}

_lib = None


class BosotLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library once and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BosotLibraryError(
            "bosot_b200: CUDA extension %s is missing -- the hot path has no CPU fallback. "
            "Build it with `make -C bosot_b200/csrc` or `__graft_entry__.build()`." % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.bosot_abi_version() != ABI_VERSION:
        raise BosotLibraryError("bosot_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    """Map a negative return code to a Python exception (never silently ignored)."""
    if rc == 0:
        return
    msg = load().bosot_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError("bosot_b200 %s: %s" % (what, msg))
    raise RuntimeError("bosot_b200 %s failed (code %d): %s" % (what, rc, msg))


def launch_count() -> int:
    return int(load().bosot_kernel_launch_count())
