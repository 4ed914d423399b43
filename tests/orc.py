"""ctypes bindings for the CPU oracle (oracle/liborc.so) — TEST INFRASTRUCTURE ONLY.

The oracle is the checker for the CUDA path; nothing in needlestack_b200/ imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_bp = C.POINTER(C.c_uint8)


class GlmParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "c_tukey_beta", "c_tukey_sig", "minsig", "maxsig", "minmu", "sig_prec", "tol",
        "min_coverage", "min_reads", "min_af",
        "min_af_extra_rob", "min_prop_extra_rob", "max_prop_extra_rob")] + [
        (k, C.c_int32) for k in ("maxit", "maxit_sig", "n_ai_sig_tukey", "size_min", "extra_rob", "_pad")]


class GlmInfo(C.Structure):
    _fields_ = [("sigma", C.c_double), ("slope", C.c_double)] + [
        (k, C.c_int32) for k in ("gated", "extra_rob", "n_outer", "n_irls", "n_obj_eval", "sig_at_min",
                                 "maxit_hit", "quad_error", "nan_rows", "_pad")] + [
        (k, C.c_int64) for k in ("n_seed", "n_lattice", "n_quad")]


class CallParams(C.Structure):
    _fields_ = [("glm", GlmParams)] + [
        (k, C.c_double) for k in ("GQ_threshold", "SB_threshold_SNV", "SB_threshold_indel", "sigma_normal",
                                  "afmin_power")] + [
        (k, C.c_int32) for k in ("SB_type", "output_all_SNVs", "is_tn", "n_pairs", "n_only_t", "n_only_n")] + [
        (k, c_ip) for k in ("tindex", "nindex", "only_t", "only_n")]


class UnitInfo(C.Structure):
    _fields_ = [("glm", GlmInfo), ("ref_inv", C.c_int32), ("gate1", C.c_int32), ("gate2", C.c_int32),
                ("_pad", C.c_int32), ("max_gq", C.c_double), ("pooled", C.c_double * 8)]


class QuadOut(C.Structure):
    _fields_ = [("result", C.c_double), ("abserr", C.c_double), ("neval", C.c_int), ("ier", C.c_int),
                ("last", C.c_int)]


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    return os.path.join(ORACLE_DIR, "liborc.so")


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ORACLE_DIR, "liborc.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("orc_math.c", "orc_model.c", "orc_math.h", "orc_model.h")]
    if (not os.path.exists(path)) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        build()
    L = C.CDLL(path)
    d = C.c_double
    for name, nargs in (("orc_stirlerr", 1), ("orc_bd0", 2), ("orc_dbinom_raw", 4), ("orc_dnbinom_mu", 3),
                        ("orc_pnbinom_mu_upper", 3), ("orc_pnbinom_mu_lower", 3), ("orc_pbinom_lower", 3),
                        ("orc_digamma", 1), ("orc_qnbinom_mu", 3), ("orc_qbinom", 3), ("orc_dhyper_log", 4),
                        ("orc_fisher_2x2", 4)):
        f = getattr(L, name)
        f.restype = d
        f.argtypes = [d] * nargs
    L.orc_pbeta_raw.restype = d
    L.orc_pbeta_raw.argtypes = [d, d, d, d, C.c_int]
    L.orc_natural_spline.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp]
    L.orc_spline_eval.restype = d
    L.orc_spline_eval.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, d]
    L.orc_integrate_spline.argtypes = [C.c_int, c_dp, c_dp, d, d, C.POINTER(QuadOut)]
    L.orc_dqagse_lists.argtypes = [C.POINTER(C.c_int), c_dp, c_dp, c_dp, c_dp, c_ip]
    L.orc_quantile7.restype = d
    L.orc_quantile7.argtypes = [C.c_int, c_dp, d]
    L.orc_median.restype = d
    L.orc_median.argtypes = [C.c_int, c_dp]
    L.orc_padjust_bh.argtypes = [C.c_int, c_dp, c_dp]
    L.orc_padjust_holm.argtypes = [C.c_int, c_dp, c_dp]
    L.orc_glm_params_default.argtypes = [C.POINTER(GlmParams)]
    L.orc_call_params_default.argtypes = [C.POINTER(CallParams)]
    L.orc_glmrob_nb.argtypes = [C.c_int, c_dp, c_dp, C.POINTER(GlmParams), c_dp, c_dp, c_dp, C.POINTER(GlmInfo)]
    L.orc_call_unit.argtypes = [C.c_int, c_ip, c_ip, c_ip, c_ip, c_ip, C.c_int, C.POINTER(CallParams),
                                c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_bp, c_bp, C.POINTER(UnitInfo)]
    L.orc_call_snv_batch.argtypes = [C.c_int64, C.c_int, c_bp, c_ip, C.POINTER(CallParams), C.c_int,
                                     c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_bp, c_bp, C.POINTER(UnitInfo)]
    L.orc_plot_toqvalue.argtypes = [C.c_int, c_ip, c_ip, d, d, C.c_int64, c_ip, c_ip, c_dp]
    _LIB = L
    return L


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def _bp(a):
    return a.ctypes.data_as(c_bp)


def glm_params(**kw):
    p = GlmParams()
    lib().orc_glm_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def call_params(pairs=None, **kw):
    """pairs = dict(tindex=[..], nindex=[..], only_t=[..], only_n=[..]) with 0-based sample indices."""
    p = CallParams()
    lib().orc_call_params_default(C.byref(p))
    for k, v in kw.items():
        if hasattr(p.glm, k) and not hasattr(p, k):
            setattr(p.glm, k, v)
        else:
            setattr(p, k, v)
    keep = []
    if pairs is not None:
        p.is_tn = 1
        for name, cnt in (("tindex", "n_pairs"), ("nindex", "n_pairs"), ("only_t", "n_only_t"),
                          ("only_n", "n_only_n")):
            arr = np.ascontiguousarray(pairs.get(name, []), dtype=np.int32)
            keep.append(arr)
            setattr(p, name, _ip(arr))
            setattr(p, cnt, len(arr))
    p._keep = keep
    return p


def glmrob_nb(y, x, prm=None, **kw):
    """Oracle glmrob.nb(y, x, ...): returns dict like the reference's result list."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = len(y)
    prm = prm or glm_params(**kw)
    pv, qv, gq = np.empty(n), np.empty(n), np.empty(n)
    info = GlmInfo()
    lib().orc_glmrob_nb(n, _dp(y), _dp(x), C.byref(prm), _dp(pv), _dp(qv), _dp(gq), C.byref(info))
    return dict(coverage=x, ma_count=y, sigma=info.sigma, slope=info.slope, pvalues=pv, qvalues=qv, GQ=gq,
                extra_rob=bool(info.extra_rob), info=info)


def call_unit(dp, vp, vm, rp, rm, prm, is_indel=False):
    arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (dp, vp, vm, rp, rm)]
    n = len(arrs[0])
    out = {k: np.empty(n) for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs")}
    gt = np.zeros(n, dtype=np.uint8)
    st = np.zeros(n, dtype=np.uint8)
    info = UnitInfo()
    lib().orc_call_unit(n, *[_ip(a) for a in arrs], int(is_indel), C.byref(prm),
                        *[_dp(out[k]) for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs")],
                        _bp(gt), _bp(st), C.byref(info))
    out["gt"] = gt
    out["status"] = st
    out["info"] = info
    return out


def call_snv_batch(ref, counts, prm, n_threads=1):
    """counts: int32 [9][P][n] plane-major; ref: uint8 [P] codes (A,T,C,G = 0..3)."""
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    ref = np.ascontiguousarray(ref, dtype=np.uint8)
    _, P, n = counts.shape
    U = 3 * P
    out = {k: np.empty((U, n)) for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs")}
    gt = np.zeros((U, n), dtype=np.uint8)
    st = np.zeros((U, n), dtype=np.uint8)
    info = (UnitInfo * U)()
    lib().orc_call_snv_batch(P, n, _bp(ref), _ip(counts), C.byref(prm), n_threads,
                             *[_dp(out[k]) for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs")],
                             _bp(gt), _bp(st), info)
    out["gt"] = gt
    out["status"] = st
    ia = np.frombuffer(info, dtype=np.dtype(UnitInfo))  # structured view: no per-unit Python loop
    g = ia["glm"]
    out["sigma"] = g["sigma"].copy()
    out["slope"] = g["slope"].copy()
    for k, src in (("gated", "gated"), ("extra_rob", "extra_rob"), ("n_outer", "n_outer"), ("n_irls", "n_irls"),
                   ("n_obj", "n_obj_eval"), ("maxit_hit", "maxit_hit"), ("quad_error", "quad_error"),
                   ("sig_at_min", "sig_at_min"), ("nan_rows", "nan_rows")):
        out[k] = g[src].astype(np.int32)
    for k in ("ref_inv", "gate1", "gate2"):
        out[k] = ia[k].astype(np.int32)
    out["max_gq"] = ia["max_gq"].copy()
    for k in ("n_seed", "n_lattice", "n_quad"):
        out[k] = g[k].astype(np.int64)
    out["pooled"] = ia["pooled"].reshape(U, 8).copy()
    out["info"] = info
    return out


def plot_toqvalue(coverage, ma_count, sigma, slope, gx, gy):
    """toQvalue(x, y) of plot_rob_nb.r:83-87 for every (gx[g], gy[g])"""
    cov = np.ascontiguousarray(coverage, dtype=np.int32)
    ma = np.ascontiguousarray(ma_count, dtype=np.int32)
    gx = np.ascontiguousarray(gx, dtype=np.int32).reshape(-1)
    gy = np.ascontiguousarray(gy, dtype=np.int32).reshape(-1)
    out = np.empty(len(gx))
    lib().orc_plot_toqvalue(len(cov), _ip(cov), _ip(ma), float(sigma), float(slope), len(gx), _ip(gx), _ip(gy), _dp(out))
    return out
