"""The C-ABI library loads on a machine without a GPU, exports every symbol include/ns_b200.h declares,
its struct layouts match the ctypes mirror, and the product path fails loudly without a CUDA device
(no compute calls here)."""
import ctypes as C
import os
import re

import pytest

import needlestack_b200 as ns
from needlestack_b200 import _abi, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from needlestack_b200 import build
    build.build()
    return ns.load_library()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "ns_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decl = set(re.findall(r"\b(ns_[a-z0-9_]+)\s*\(", body))
    assert decl, "no declarations parsed"
    assert decl == set(api.EXPORTED_SYMBOLS)
    for name in decl:
        assert getattr(lib, name) is not None


def test_struct_layouts(lib):
    lib.ns_sizeof.restype = C.c_size_t
    lib.ns_sizeof.argtypes = [C.c_int]
    assert lib.ns_sizeof(0) == C.sizeof(_abi.NsParams)
    assert lib.ns_sizeof(1) == C.sizeof(_abi.NsOut)
    assert lib.ns_sizeof(2) == C.sizeof(_abi.NsTimings)
    assert lib.ns_abi_version() == _abi.NS_ABI_VERSION


def test_defaults_match_reference(lib):
    p = _abi.NsParams()
    lib.ns_params_glm_default(C.byref(p))  # bin/glm_rob_nb.r:16-18
    assert (p.c_tukey_beta, p.c_tukey_sig, p.minsig, p.maxsig, p.maxit, p.maxit_sig) == (5, 3, 1e-3, 10, 30, 50)
    assert (p.min_coverage, p.min_reads, p.size_min, p.extra_rob, p.max_prop_extra_rob) == (1, 1, 10, 1, 0.5)
    lib.ns_params_caller_default(C.byref(p))  # bin/needlestack.r:64-92
    assert (p.min_coverage, p.min_reads, p.min_af, p.extra_rob, p.max_prop_extra_rob) == (30, 3, 0, 0, 0.8)
    assert (p.GQ_threshold, p.SB_threshold_SNV, p.SB_threshold_indel, p.sigma_normal, p.afmin_power) == \
        (50, 100, 100, 0.1, -1)
    q = ns.make_params()
    for k, _ in _abi.NsParams._fields_[:-5]:
        assert getattr(p, k) == getattr(q, k), k


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ns.NeedlestackError, match="no CPU path"):
        ns.Caller(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "needlestack_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".r", ".R")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liborc" not in txt and "import orc" not in txt and "oracle/" not in txt, f


def test_unsupported_options_raise_like_the_reference():
    c = object.__new__(ns.Caller)  # no context needed for argument checks
    with pytest.raises(ns.NeedlestackError, match='Available bounding.func is "T/T"'):
        ns.Caller.glmrob_nb(c, [1, 2], [10, 20], bounding_func="BY/T")
    with pytest.raises(ns.NeedlestackError, match="weights.on.x"):
        ns.Caller.glmrob_nb(c, [1, 2], [10, 20], weights_on_x="soft")
