"""ctypes binding of the serial emulation build of the kernel bodies (tests/emul/ns_emul.cpp).
TEST INFRASTRUCTURE ONLY — the product library never runs on the host."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
sys.path.insert(0, ROOT)
from needlestack_b200 import _abi  # noqa: E402

_LIB = None


def build():
    src = os.path.join(_HERE, "emul", "ns_emul.cpp")
    out = os.path.join(_HERE, "emul", "libnsemul.so")
    deps = [src] + [os.path.join(ROOT, "needlestack_b200", "csrc", f) for f in
                    ("ns_math.cuh", "ns_quad.cuh", "ns_core.cuh", "ns_fit.cuh", "ns_logtab.h")] + [os.path.join(ROOT, "include", "ns_b200.h")]
    if (not os.path.exists(out)) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", out, src], check=True)
    return out


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        d = C.c_double
        for name, nargs in (("emul_dnbinom_mu", 3), ("emul_nb_pvalue", 3), ("emul_nb_pvalue_lb", 3), ("emul_digamma", 1), ("emul_fexp", 1), ("emul_flog", 1), ("emul_flog_tab", 1), ("emul_fexp_bf", 1),
                            ("emul_qnbinom_mu", 3), ("emul_qbinom", 3), ("emul_fisher", 4)):
            f = getattr(L, name)
            f.restype = d
            f.argtypes = [d] * nargs
        L.emul_ai.restype = d
        L.emul_ai.argtypes = [d, d, d, C.c_int]
        L.emul_E12.argtypes = [d, d, d, C.c_int, _abi.c_dp, _abi.c_dp]
        _LIB = L
    return _LIB


def make_params(pairs=None, **kw):
    p = _abi.NsParams()
    vals = dict(_abi.CALLER_DEFAULTS, emit_mode=_abi.NS_EMIT_ALL)  # the checker wants every row
    vals.update(kw)
    for k, v in vals.items():
        setattr(p, k, v)
    keep = []
    if pairs is not None:
        p.is_tn = 1
        for name, cnt in (("tindex", "n_pairs"), ("nindex", "n_pairs"), ("only_t", "n_only_t"),
                          ("only_n", "n_only_n")):
            arr = np.ascontiguousarray(pairs.get(name, []), dtype=np.int32)
            keep.append(arr)
            setattr(p, name, arr.ctypes.data_as(_abi.c_ip))
            setattr(p, cnt, len(arr))
    p._keep = keep
    return p


def roles(n, pairs):
    """per-sample role exactly as the library builds it (assignment order of needlestack.r:343-347)"""
    role = np.zeros(n, dtype=np.uint8)
    if pairs is None:
        return role
    role[np.asarray(pairs.get("nindex", []), dtype=np.int64)] = 1
    role[np.asarray(pairs.get("tindex", []), dtype=np.int64)] = 2
    role[np.asarray(pairs.get("only_n", []), dtype=np.int64)] = 3
    role[np.asarray(pairs.get("only_t", []), dtype=np.int64)] = 4
    return role


def call_unit(dp, vp, vm, rp, rm, prm, pairs=None, is_indel=False, plain=False, fast=True):
    arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (dp, vp, vm, rp, rm)]
    n = len(arrs[0])
    out = {k: np.empty(n) for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs")}
    gt = np.zeros(n, dtype=np.uint8)
    st = np.zeros(n, dtype=np.uint8)
    pooled = np.zeros(8)
    sigma, slope, mg = C.c_double(), C.c_double(), C.c_double()
    flags, iters = C.c_uint32(), C.c_uint32()
    cnt = (C.c_uint64 * 4)()
    role = roles(n, pairs)
    lib().emul_unit(n, *[a.ctypes.data_as(_abi.c_ip) for a in arrs], int(is_indel), int(plain), int(fast), C.byref(prm),
                    role.ctypes.data_as(_abi.c_bp),
                    *[out[k].ctypes.data_as(_abi.c_dp) for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor",
                                                                  "rvsb", "fs")],
                    gt.ctypes.data_as(_abi.c_bp), st.ctypes.data_as(_abi.c_bp), pooled.ctypes.data_as(_abi.c_dp),
                    C.byref(sigma), C.byref(slope), C.byref(mg), C.byref(flags), C.byref(iters), cnt)
    out.update(gt=gt, status=st, pooled=pooled, sigma=sigma.value, slope=slope.value, max_gq=mg.value,
               flags=flags.value, n_outer=iters.value & 0xff, n_irls=(iters.value >> 8) & 0x3ff,
               n_obj=iters.value >> 18, n_seed=cnt[0], n_lattice=cnt[1], n_quad=cnt[2], deferred=cnt[3])
    return out
