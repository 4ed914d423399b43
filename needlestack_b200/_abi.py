"""ctypes mirror of include/ns_b200.h (struct layouts and constants only)."""
import ctypes as C

NS_ABI_VERSION = 2

NS_OK = 0
NS_ERR_CUDA = -1
NS_ERR_ARG = -2
NS_ERR_EMIT_OVERFLOW = -3
NS_ERR_NOMEM = -4

NS_F_GATED = 0x001
NS_F_EXTRA_ROB = 0x002
NS_F_REF_INV = 0x004
NS_F_SIG_AT_MIN = 0x008
NS_F_MAXIT = 0x010
NS_F_QUAD_ERROR = 0x020
NS_F_NAN_ROWS = 0x040
NS_F_GATE1 = 0x080
NS_F_GATE2 = 0x100
NS_F_SKIPPED = 0x200
NS_F_QUAD_OVERFLOW = 0x400

NS_GT = ("0/0", "0/1", "1/1", "./.")
NS_STATUS = (".", "SOMATIC", "GERMLINE_CONFIRMED", "GERMLINE_UNCONFIRMED", "GERMLINE_UNCONFIRMABLE", "UNKNOWN",
             "POSSIBLE_CONTAMINATION")

NS_EMIT_CALLS = 0
NS_EMIT_FITTED = 1
NS_EMIT_ALL = 2

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_bp = C.POINTER(C.c_uint8)
c_up = C.POINTER(C.c_uint32)
c_lp = C.POINTER(C.c_int64)


class NsParams(C.Structure):
    _fields_ = (
        [(k, C.c_double) for k in ("c_tukey_beta", "c_tukey_sig", "minsig", "maxsig", "minmu", "sig_prec", "tol",
                                   "min_coverage", "min_reads", "min_af", "min_af_extra_rob", "min_prop_extra_rob",
                                   "max_prop_extra_rob")]
        + [(k, C.c_int32) for k in ("maxit", "maxit_sig", "n_ai_sig_tukey", "size_min", "extra_rob", "SB_type",
                                    "output_all_SNVs", "is_tn")]
        + [(k, C.c_double) for k in ("GQ_threshold", "SB_threshold_SNV", "SB_threshold_indel", "sigma_normal",
                                     "afmin_power")]
        + [(k, C.c_int32) for k in ("n_pairs", "n_only_t", "n_only_n", "emit_mode")]
        + [(k, c_ip) for k in ("tindex", "nindex", "only_t", "only_n")]
        + [("role", c_bp)]
    )


class NsOut(C.Structure):
    _fields_ = (
        [("sigma", c_dp), ("slope", c_dp), ("max_gq", c_dp), ("flags", c_up), ("iters", c_up), ("emit_index", c_ip), ("unit_cycles", C.POINTER(C.c_uint64)),
         ("emit_cap", C.c_int64), ("n_emitted", C.c_int64), ("emit_unit", c_lp)]
        + [(k, c_dp) for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs")]
        + [("gt", c_bp), ("status", c_bp), ("pooled", c_dp)]
        + [(k, C.c_int64) for k in ("n_fitted", "n_gated", "n_skipped")]
        + [(k, C.c_uint64) for k in ("n_seed", "n_lattice", "n_quad")]
        + [("row_sigma", c_dp), ("row_slope", c_dp), ("row_max_gq", c_dp), ("row_flags", c_up)]
    )


class NsTimings(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("h2d_ms", "gate_ms", "fit_ms", "heavy_ms", "post_ms", "pack_ms", "d2h_ms",
                                          "total_ms", "heavy_pre_ms", "fit_phase_ms")] + [("kernel_launches", C.c_int64), ("n_heavy", C.c_int64), ("n_deferred", C.c_int64),
                                                            ("fast_seed", C.c_uint64), ("fast_lattice", C.c_uint64),
                                                            ("pool_units", C.c_int64), ("chunks", C.c_int64), ("pool_tail_ms", C.c_double)]


GLM_DEFAULTS = dict(  # bin/glm_rob_nb.r:16-18
    c_tukey_beta=5.0, c_tukey_sig=3.0, minsig=1e-3, maxsig=10.0, minmu=1e-10, sig_prec=1e-8, tol=1e-6,
    min_coverage=1.0, min_reads=1.0, min_af=0.0, min_af_extra_rob=0.2, min_prop_extra_rob=0.1,
    max_prop_extra_rob=0.5, maxit=30, maxit_sig=50, n_ai_sig_tukey=100, size_min=10, extra_rob=1)

CALLER_DEFAULTS = dict(  # bin/needlestack.r:64-92
    GLM_DEFAULTS, min_coverage=30.0, min_reads=3.0, min_af=0.0, extra_rob=0, max_prop_extra_rob=0.8,
    GQ_threshold=50.0, SB_type=0, SB_threshold_SNV=100.0, SB_threshold_indel=100.0, sigma_normal=0.1,
    afmin_power=-1.0, output_all_SNVs=0, is_tn=0, emit_mode=NS_EMIT_CALLS)
