"""Host-side mirror of the reference interface for the error-model path, over the C ABI of
include/ns_b200.h (libns_b200.so, CUDA sm_100a).

    glmrob_nb(y, x, ...)        <- glmrob.nb()           bin/glm_rob_nb.r:16-228
    Caller.call_snv(ref, counts)<- per-allele SNV loop   bin/needlestack.r:321-413
    Caller.call_units(...)      <- DEL / INS blocks      bin/needlestack.r:461-567, :624-730
                                   annotate_vcf DP/AD    bin/annotate_vcf.r:57-69

There is no CPU implementation behind these calls: a missing library or a machine without a
CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import (NS_EMIT_ALL, NS_EMIT_CALLS, NS_EMIT_FITTED, NsOut, NsParams, NsTimings, c_bp, c_dp, c_ip, c_lp,
                   c_up)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libns_b200.so")
_LIB = None


class NeedlestackError(RuntimeError):
    pass


def load_library():
    """dlopen libns_b200.so and declare every symbol of include/ns_b200.h (no compute calls)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise NeedlestackError(f"{LIB_PATH} is missing: run `python -m needlestack_b200.build` (needs nvcc); "
                               "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.ns_abi_version.restype = C.c_int
    L.ns_params_glm_default.argtypes = [C.POINTER(NsParams)]
    L.ns_params_caller_default.argtypes = [C.POINTER(NsParams)]
    L.ns_create.restype = vp
    L.ns_create.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
    L.ns_destroy.argtypes = [vp]
    L.ns_last_error.restype = C.c_char_p
    L.ns_last_error.argtypes = [vp]
    L.ns_set_stream.argtypes = [vp, vp]
    L.ns_alloc_pinned.restype = vp
    L.ns_alloc_pinned.argtypes = [C.c_size_t]
    L.ns_free_pinned.argtypes = [vp]
    L.ns_get_timings.argtypes = [vp, C.POINTER(NsTimings)]
    L.ns_call_snv.argtypes = [vp, C.POINTER(NsParams), C.c_int, C.c_int64, vp, vp, C.POINTER(NsOut)]
    L.ns_call_snv_u16.argtypes = [vp, C.POINTER(NsParams), C.c_int, C.c_int64, vp, vp, C.POINTER(NsOut)]
    L.ns_call_snv_device.argtypes = [vp, C.POINTER(NsParams), C.c_int, C.c_int64, vp, vp, C.POINTER(NsOut)]
    L.ns_call_units.argtypes = [vp, C.POINTER(NsParams), C.c_int, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_int,
                                C.POINTER(NsOut)]
    L.ns_qvalue_grid.argtypes = [vp, C.c_double, C.c_double, C.c_int, vp, vp, C.c_int64, vp, vp, c_dp]
    L.ns_glmrob_nb.argtypes = [vp, C.POINTER(NsParams), C.c_int, vp, vp, c_dp, c_dp, c_dp, c_dp, c_dp, c_ip, c_up]
    L.ns_measure_fp64_peak.argtypes = [vp, c_dp]
    L.ns_create_multi.restype = vp
    L.ns_create_multi.argtypes = [c_ip, C.c_int, C.c_char_p, C.c_size_t]
    L.ns_destroy_multi.argtypes = [vp]
    L.ns_multi_devices.argtypes = [vp]
    L.ns_multi_last_error.restype = C.c_char_p
    L.ns_multi_last_error.argtypes = [vp]
    L.ns_multi_call_snv.argtypes = [vp, C.POINTER(NsParams), C.c_int, C.c_int64, vp, vp, C.POINTER(NsOut)]
    L.ns_multi_call_snv_u16.argtypes = [vp, C.POINTER(NsParams), C.c_int, C.c_int64, vp, vp, C.POINTER(NsOut)]
    L.ns_multi_get_timings.argtypes = [vp, C.c_int, C.POINTER(NsTimings), c_dp]
    L.ns_table_parse.restype = vp
    L.ns_table_parse.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    L.ns_table_free.argtypes = [vp]
    L.ns_table_positions.restype = C.c_int64
    L.ns_table_indel_units.restype = C.c_int64
    for name in ("ns_table_positions", "ns_table_samples", "ns_table_counts_pinned", "ns_table_counts",
                 "ns_table_ref", "ns_table_ref_text", "ns_table_loc", "ns_table_chrom", "ns_table_n_chrom",
                 "ns_table_indel_cov", "ns_table_indel_units", "ns_table_indel_pos", "ns_table_indel_type",
                 "ns_table_indel_vp", "ns_table_indel_vm"):
        getattr(L, name).argtypes = [vp]
    for name in ("ns_table_counts", "ns_table_ref", "ns_table_ref_text", "ns_table_loc", "ns_table_chrom",
                 "ns_table_indel_cov", "ns_table_indel_pos", "ns_table_indel_type", "ns_table_indel_vp",
                 "ns_table_indel_vm"):
        getattr(L, name).restype = vp
    L.ns_table_chrom_name.restype = C.c_char_p
    L.ns_table_chrom_name.argtypes = [vp, C.c_int]
    L.ns_table_indel_seq.restype = C.c_char_p
    L.ns_table_indel_seq.argtypes = [vp, C.c_int64]
    L.ns_fasta_load.restype = vp
    L.ns_fasta_load.argtypes = [C.c_char_p]
    L.ns_fasta_free.argtypes = [vp]
    for name in ("ns_vcf_format_snv", "ns_vcf_format_indel"):
        getattr(L, name).restype = vp
        getattr(L, name).argtypes = [vp, C.POINTER(NsOut), C.POINTER(C.c_char_p), vp, C.c_double, C.c_int]
    L.ns_vcf_text_data.restype = vp
    L.ns_vcf_text_data.argtypes = [vp]
    L.ns_vcf_text_size.restype = C.c_size_t
    L.ns_vcf_text_size.argtypes = [vp]
    L.ns_vcf_text_offsets.restype = vp
    L.ns_vcf_text_offsets.argtypes = [vp]
    L.ns_vcf_text_free.argtypes = [vp]
    L.ns_vcf_format_number.argtypes = [C.c_double, C.c_int, C.c_char_p, C.c_size_t]
    if L.ns_abi_version() != _abi.NS_ABI_VERSION:
        raise NeedlestackError("libns_b200.so ABI version mismatch")
    _LIB = L
    return L


EXPORTED_SYMBOLS = ("ns_abi_version", "ns_sizeof", "ns_params_glm_default", "ns_params_caller_default", "ns_create", "ns_destroy",
                    "ns_last_error", "ns_set_stream", "ns_alloc_pinned", "ns_free_pinned", "ns_get_timings",
                    "ns_call_snv", "ns_call_snv_u16", "ns_call_snv_device", "ns_call_units", "ns_glmrob_nb", "ns_qvalue_grid", "ns_measure_fp64_peak",
                    "ns_create_multi", "ns_destroy_multi", "ns_multi_devices", "ns_multi_last_error", "ns_multi_call_snv",
                    "ns_multi_call_snv_u16", "ns_multi_get_timings",
                    "ns_fasta_load", "ns_fasta_free", "ns_vcf_format_snv", "ns_vcf_format_indel", "ns_vcf_text_data",
                    "ns_vcf_text_size", "ns_vcf_text_offsets", "ns_vcf_text_free", "ns_vcf_format_number",
                    "ns_table_parse", "ns_table_free", "ns_table_positions", "ns_table_samples",
                    "ns_table_counts_pinned", "ns_table_counts", "ns_table_ref", "ns_table_ref_text", "ns_table_loc",
                    "ns_table_chrom", "ns_table_n_chrom", "ns_table_chrom_name", "ns_table_indel_cov",
                    "ns_table_indel_units", "ns_table_indel_pos", "ns_table_indel_type", "ns_table_indel_seq",
                    "ns_table_indel_vp", "ns_table_indel_vm")


def make_params(pairs=None, defaults="caller", **kw):
    """ns_params with needlestack.r's defaults (defaults='caller') or glmrob.nb's own ('glm')."""
    p = NsParams()
    vals = dict(_abi.CALLER_DEFAULTS if defaults == "caller" else dict(_abi.GLM_DEFAULTS, GQ_threshold=50.0,
                SB_threshold_SNV=100.0, SB_threshold_indel=100.0, sigma_normal=0.1, afmin_power=-1.0,
                emit_mode=NS_EMIT_ALL))
    for k in kw:
        if not hasattr(p, k):
            raise TypeError(f"unknown parameter {k!r}")
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
            setattr(p, name, arr.ctypes.data_as(c_ip))
            setattr(p, cnt, len(arr))
        if len(keep[0]) != len(keep[1]):
            raise ValueError("tindex and nindex must have the same length")
    p._keep = keep
    return p


class PinnedArray:
    """numpy view over page-locked host memory from ns_alloc_pinned (the loader's target)."""

    def __init__(self, shape, dtype):
        L = load_library()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = L.ns_alloc_pinned(max(self.nbytes, 1))
        if not self.ptr:
            raise NeedlestackError("ns_alloc_pinned failed")
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            load_library().ns_free_pinned(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Result(dict):
    """per-unit arrays + emitted rows; attribute access for convenience.  n_outer / n_irls / n_obj
    (iteration counts unpacked from `iters`) are computed on first use."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __missing__(self, k):
        it = dict.__getitem__(self, "iters")
        if k == "n_outer":
            v = (it & 0xff).astype(np.int64)
        elif k == "n_irls":
            v = ((it >> 8) & 0x3ff).astype(np.int64)
        elif k == "n_obj":
            v = (it >> 18).astype(np.int64)
        else:
            raise KeyError(k)
        self[k] = v
        return v

    def row_of(self, unit):
        r = int(self["emit_index"][unit])
        return None if r < 0 else r

    def dense(self, key, fill=np.nan):
        """[U][n] array of an emitted plane with `fill` for units that were not emitted"""
        U = len(self["flags"])
        n = self[key].shape[1] if self[key].ndim == 2 else 0
        out = np.full((U, n), fill, dtype=self[key].dtype)
        idx = self["emit_index"]
        m = idx >= 0
        out[m] = self[key][idx[m]]
        return out


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def result_as_out(R):
    """ns_out view of a Result (host arrays) for the VCF writer; keeps the arrays alive through the returned tuple"""
    o = NsOut()
    keep = []

    def setp(field, key, typ):
        a = np.ascontiguousarray(R[key])
        keep.append(a)
        setattr(o, field, a.ctypes.data_as(typ))

    for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs", "pooled"):
        setp(k, k, c_dp)
    for k in ("sigma", "slope", "max_gq", "row_sigma", "row_slope", "row_max_gq"):
        if k in R:
            setp(k, k, c_dp)
    for k in ("flags", "row_flags"):
        if k in R:
            setp(k, k, c_up)
    setp("emit_unit", "emit_unit", c_lp)
    setp("gt", "gt", c_bp)
    setp("status", "status", c_bp)
    o.n_emitted = int(R["n_emitted"])
    o.emit_cap = int(R["n_emitted"])
    return o, keep


def vcf_records(table, R, names, fasta=None, gq_threshold=50.0, indel=False, n_threads=0):
    """VCF record text (bytes) of every emitted row of R, formatted by the C++ writer (ns_vcf_format_snv / _indel), and
    the [records + 1] offsets of the records in it"""
    L = load_library()
    o, keep = result_as_out(R)
    arr = (C.c_char_p * len(names))(*[nm.encode() for nm in names])
    fn = L.ns_vcf_format_indel if indel else L.ns_vcf_format_snv
    h = fn(table._h, C.byref(o), arr, fasta, float(gq_threshold), int(n_threads) or (os.cpu_count() or 1))
    if not h:
        raise NeedlestackError("VCF writer: result planes missing")
    try:
        size = L.ns_vcf_text_size(h)
        text = C.string_at(L.ns_vcf_text_data(h), size)
        off = np.frombuffer(C.string_at(L.ns_vcf_text_offsets(h), 8 * (int(R["n_emitted"]) + 1)), dtype=np.int64).copy()
    finally:
        L.ns_vcf_text_free(h)
    return text, off


class Caller:
    """One GPU context (ns_ctx).  Not thread-safe; one per device and host thread."""

    def __init__(self, device=0, reuse_outputs=False):
        """reuse_outputs=True: results are views into page-locked buffers owned by the Caller and are
        overwritten by its next call (fast D2H, no per-call allocation); False: fresh numpy arrays."""
        self._reuse = bool(reuse_outputs)
        self._pool = {}
        self._L = load_library()
        err = C.create_string_buffer(512)
        self._ctx = self._L.ns_create(int(device), err, 512)
        if not self._ctx:
            raise NeedlestackError(err.value.decode() or "ns_create failed")
        self.device = int(device)

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.ns_destroy(self._ctx)
            self._ctx = None
        for pa in getattr(self, "_pool", {}).values():
            pa.free()
        self._pool = {}

    def _buf(self, name, shape, dtype, fill=None):
        """output array: pinned and pooled when reuse_outputs, else a fresh numpy array"""
        if not self._reuse:
            return np.empty(shape, dtype=dtype) if fill is None else np.full(shape, fill, dtype=dtype)
        need = int(np.prod(shape)) * np.dtype(dtype).itemsize
        pa = self._pool.get(name)
        if pa is None or pa.nbytes < need:
            if pa is not None:
                pa.free()
            pa = PinnedArray((max(need + need // 4, 64),), np.uint8)
            self._pool[name] = pa
        arr = pa.array[:need].view(dtype).reshape(shape)
        if fill is not None:
            arr[...] = fill
        return arr

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            msg = self._L.ns_last_error(self._ctx).decode()
            raise NeedlestackError(f"ns error {rc}: {msg}")

    def set_stream(self, cuda_stream):
        self._check(self._L.ns_set_stream(self._ctx, C.c_void_p(int(cuda_stream) if cuda_stream else 0)))

    def timings(self):
        t = NsTimings()
        self._L.ns_get_timings(self._ctx, C.byref(t))
        return {k: getattr(t, k) for k, _ in NsTimings._fields_}

    def fp64_peak_tflops(self):
        v = C.c_double()
        self._check(self._L.ns_measure_fp64_peak(self._ctx, C.byref(v)))
        return v.value

    # -- output allocation ---------------------------------------------------------------
    def _alloc_out(self, U, n, emit_cap, want_rows=True, per_unit=True):
        o = NsOut()
        B = self._buf
        a = {}
        if per_unit:
            a = dict(sigma=B("sigma", (U,), np.float64), slope=B("slope", (U,), np.float64),
                     max_gq=B("max_gq", (U,), np.float64), flags=B("flags", (U,), np.uint32),
                     iters=B("iters", (U,), np.uint32), emit_index=B("emit_index", (U,), np.int32),
                     unit_cycles=B("unit_cycles", (U,), np.uint64))
            o.sigma, o.slope, o.max_gq = (a[k].ctypes.data_as(c_dp) for k in ("sigma", "slope", "max_gq"))
            o.flags = a["flags"].ctypes.data_as(c_up)
            o.iters = a["iters"].ctypes.data_as(c_up)
            o.emit_index = a["emit_index"].ctypes.data_as(c_ip)
            o.unit_cycles = a["unit_cycles"].ctypes.data_as(C.POINTER(C.c_uint64))
        o.emit_cap = emit_cap if want_rows else 0
        if want_rows:
            for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs"):
                a[k] = B(k, (emit_cap, n), np.float64)
                setattr(o, k, a[k].ctypes.data_as(c_dp))
            for k in ("gt", "status"):
                a[k] = B(k, (emit_cap, n), np.uint8)
                setattr(o, k, a[k].ctypes.data_as(c_bp))
            a["pooled"] = B("pooled", (emit_cap, 8), np.float64)
            o.pooled = a["pooled"].ctypes.data_as(c_dp)
            a["emit_unit"] = B("emit_unit", (emit_cap,), np.int64)
            o.emit_unit = a["emit_unit"].ctypes.data_as(c_lp)
            for k in ("row_sigma", "row_slope", "row_max_gq"):
                a[k] = B(k, (emit_cap,), np.float64)
                setattr(o, k, a[k].ctypes.data_as(c_dp))
            a["row_flags"] = B("row_flags", (emit_cap,), np.uint32)
            o.row_flags = a["row_flags"].ctypes.data_as(c_up)
        return o, a

    @staticmethod
    def _finish(o, a, want_rows):
        E = int(o.n_emitted)
        if want_rows:
            for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs", "gt", "status", "pooled",
                      "emit_unit", "row_sigma", "row_slope", "row_max_gq", "row_flags"):
                a[k] = a[k][:E]
        a.update(n_emitted=E, n_fitted=int(o.n_fitted), n_gated=int(o.n_gated), n_skipped=int(o.n_skipped),
                 n_seed=int(o.n_seed), n_lattice=int(o.n_lattice), n_quad=int(o.n_quad))
        return Result(a)

    def _emit_cap(self, prm, U, emit_cap):
        if emit_cap is not None:
            return int(emit_cap)
        if prm.emit_mode == NS_EMIT_CALLS and not prm.output_all_SNVs:
            return max(1024, U // 16)
        return U

    def _run(self, fn, prm, U, n, emit_cap, want_rows, per_unit=True):
        cap = self._emit_cap(prm, U, emit_cap)
        while True:
            o, a = self._alloc_out(U, n, cap, want_rows, per_unit)
            rc = fn(o)
            if rc == _abi.NS_ERR_EMIT_OVERFLOW and emit_cap is None:
                cap = int(o.n_emitted)
                continue
            self._check(rc)
            return self._finish(o, a, want_rows)

    # -- the reference's per-allele SNV loop (needlestack.r:321-413) -----------------------
    def call_snv(self, ref, counts, prm=None, emit_cap=None, want_rows=True, per_unit=True, **kw):
        """ref uint8 [P] (0..3 = A,T,C,G), counts int32 [P][9][n] (depth,A,T,C,G,a,t,c,g), host arrays.  A uint16
        `counts` array takes the narrow wire format (ns_call_snv_u16: half the host-to-device bytes).
        per_unit=False: only the emitted rows come back (with row_sigma / row_slope / row_max_gq / row_flags per row):
        what a caller that writes the VCF needs, without 44 B of device-to-host traffic for every gated unit."""
        prm = prm or make_params(**kw)
        u16 = getattr(counts, "dtype", None) == np.uint16
        counts = np.ascontiguousarray(counts, dtype=np.uint16 if u16 else np.int32)
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        P, nine, n = counts.shape
        if nine != 9 or len(ref) != P:
            raise ValueError("counts must be [P][9][n] and ref [P]")
        fn = self._L.ns_call_snv_u16 if u16 else self._L.ns_call_snv
        return self._run(lambda o: fn(self._ctx, C.byref(prm), n, P, _ptr(ref), _ptr(counts), C.byref(o)),
                         prm, 3 * P, n, emit_cap, want_rows, per_unit)

    def call_snv_device(self, d_ref_ptr, d_counts_ptr, P, n, prm, emit_cap=None, want_rows=True, per_unit=True):
        """same with device-resident inputs (raw device pointers, e.g. torch tensor .data_ptr())"""
        return self._run(lambda o: self._L.ns_call_snv_device(self._ctx, C.byref(prm), n, P, C.c_void_p(d_ref_ptr),
                                                              C.c_void_p(d_counts_ptr), C.byref(o)),
                         prm, 3 * P, n, emit_cap, want_rows, per_unit)

    # -- explicit units: indel alleles, annotate_vcf, batched glmrob.nb ---------------------
    def call_units(self, dp, vp, vm=None, rp=None, rm=None, is_indel=None, plain=False, prm=None, emit_cap=None,
                   want_rows=True, **kw):
        prm = prm or make_params(**kw)
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.int32) for a in (dp, vp, vm, rp, rm)]
        U, n = arrs[0].shape
        for a in arrs[1:]:
            if a is not None and a.shape != (U, n):
                raise ValueError("all count planes must be [U][n]")
        ind = None if is_indel is None else np.ascontiguousarray(is_indel, dtype=np.uint8)
        return self._run(lambda o: self._L.ns_call_units(self._ctx, C.byref(prm), n, U, *[_ptr(a) for a in arrs],
                                                         _ptr(ind), int(bool(plain)), C.byref(o)),
                         prm, U, n, emit_cap, want_rows)

    # -- glmrob.nb(y, x, ...) (glm_rob_nb.r:16) -------------------------------------------
    def glmrob_nb(self, y, x, bounding_func="T/T", weights_on_x="none", **kw):
        """Mirrors glmrob.nb(): returns dict(coverage, ma_count, coef=dict(sigma, slope), pvalues, qvalues, GQ,
        extra_rob).  Unsupported options raise with the reference's messages (glm_rob_nb.r:99,226)."""
        if weights_on_x != "none":
            raise NeedlestackError('Only "hard" and "none" are implemented for weights.on.x.'
                                   if weights_on_x != "hard" else
                                   'weights.on.x="hard" (MASS::cov.rob, random) is not available on the GPU path')
        if bounding_func != "T/T":
            raise NeedlestackError('Available bounding.func is "T/T"')
        y = np.asarray(y)
        x = np.asarray(x)
        if y.shape != x.shape or y.ndim != 1:
            raise ValueError("y and x must be vectors of the same length")
        if np.any(y != np.round(y)) or np.any(x != np.round(x)) or np.any(y < 0) or np.any(x < 0):
            raise NeedlestackError("counts must be non-negative integers")
        yi = np.ascontiguousarray(y, dtype=np.int32)
        xi = np.ascontiguousarray(x, dtype=np.int32)
        n = len(yi)
        prm = make_params(defaults="glm", **kw)
        sigma, slope = C.c_double(), C.c_double()
        pv, qv, gq = np.empty(n), np.empty(n), np.empty(n)
        er, fl = C.c_int32(), C.c_uint32()
        self._check(self._L.ns_glmrob_nb(self._ctx, C.byref(prm), n, _ptr(yi), _ptr(xi), C.byref(sigma),
                                         C.byref(slope), pv.ctypes.data_as(c_dp), qv.ctypes.data_as(c_dp),
                                         gq.ctypes.data_as(c_dp), C.byref(er), C.byref(fl)))
        if fl.value & _abi.NS_F_QUAD_ERROR:
            raise NeedlestackError("integrate() failed inside the IRLS loop (the reference stops here)")
        return dict(coverage=np.asarray(x, dtype=float), ma_count=np.asarray(y, dtype=float),
                    coef=dict(sigma=sigma.value, slope=slope.value), pvalues=pv, qvalues=qv, GQ=gq,
                    extra_rob=bool(er.value), flags=fl.value)


class MultiCaller(Caller):
    """One workload over several GPUs from this ONE host process (ns_multi_*): a queue of position chunks, one host
    thread per device inside the library, results gathered in genomic order - the role of the reference's --nsplit
    region split and final merge (needlestack.nf:421-503, :546-553).  call_snv returns exactly what Caller.call_snv
    returns on one device.  devices: CUDA ordinals (an ordinal listed twice = two contexts on that GPU)."""

    def __init__(self, devices, reuse_outputs=False):
        self._reuse = bool(reuse_outputs)
        self._pool = {}
        self._L = load_library()
        self._ctx = None
        self.devices = [int(d) for d in devices]
        err = C.create_string_buffer(512)
        arr = (C.c_int32 * len(self.devices))(*self.devices)
        self._m = self._L.ns_create_multi(arr, len(self.devices), err, 512)
        if not self._m:
            raise NeedlestackError(err.value.decode() or "ns_create_multi failed")

    def close(self):
        if getattr(self, "_m", None):
            self._L.ns_destroy_multi(self._m)
            self._m = None
        for pa in getattr(self, "_pool", {}).values():
            pa.free()
        self._pool = {}

    def _check(self, rc):
        if rc != 0:
            raise NeedlestackError(f"ns error {rc}: {self._L.ns_multi_last_error(self._m).decode()}")

    def call_snv(self, ref, counts, prm=None, emit_cap=None, want_rows=True, per_unit=True, **kw):
        prm = prm or make_params(**kw)
        u16 = getattr(counts, "dtype", None) == np.uint16
        counts = np.ascontiguousarray(counts, dtype=np.uint16 if u16 else np.int32)
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        P, nine, n = counts.shape
        if nine != 9 or len(ref) != P:
            raise ValueError("counts must be [P][9][n] and ref [P]")
        fn = self._L.ns_multi_call_snv_u16 if u16 else self._L.ns_multi_call_snv
        return self._run(lambda o: fn(self._m, C.byref(prm), n, P, _ptr(ref), _ptr(counts), C.byref(o)),
                         prm, 3 * P, n, emit_cap, want_rows, per_unit)

    def timings(self):
        """per device slot: the ns_timings of its context in the last call plus chunks pulled and host-thread wall ms"""
        out = []
        for s in range(len(self.devices)):
            t, w = NsTimings(), C.c_double()
            self._L.ns_multi_get_timings(self._m, s, C.byref(t), C.byref(w))
            d = {k: getattr(t, k) for k, _ in NsTimings._fields_}
            d["wall_ms"] = w.value
            d["device"] = self.devices[s]
            out.append(d)
        return out

    def _unsupported(self, *a, **k):
        raise NeedlestackError("MultiCaller shards ns_call_snv only; use a Caller for this entry point")

    call_snv_device = call_units = glmrob_nb = set_stream = fp64_peak_tflops = qvalue_grid = _unsupported


def _qvalue_grid(self, sigma, slope, coverage, ma_count, gx, gy):
    """toQvalue(x, y) of plot_rob_nb.r:83-87 for the points (gx[g], gy[g]) against one fitted unit: Phred-scaled
    Holm-adjusted tail probability of the new point among the n + 1; 0 where gy > gx."""
    cov = np.ascontiguousarray(coverage, dtype=np.int32)
    ma = np.ascontiguousarray(ma_count, dtype=np.int32)
    gx = np.ascontiguousarray(gx, dtype=np.int32)
    gy = np.ascontiguousarray(gy, dtype=np.int32)
    if cov.shape != ma.shape or cov.ndim != 1 or gx.shape != gy.shape:
        raise ValueError("coverage/ma_count and gx/gy must be pairs of equal-length vectors")
    out = np.empty(gx.shape, dtype=np.float64)
    self._check(self._L.ns_qvalue_grid(self._ctx, float(sigma), float(slope), len(cov), _ptr(cov), _ptr(ma), gx.size,
                                       _ptr(gx.reshape(-1)), _ptr(gy.reshape(-1)), out.ctypes.data_as(c_dp)))
    return out


Caller.qvalue_grid = _qvalue_grid


def glmrob_nb(y, x, device=0, **kw):
    """Convenience one-shot glmrob.nb() on `device`."""
    with Caller(device) as c:
        return c.glmrob_nb(y, x, **kw)
