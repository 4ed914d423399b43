"""Builds needlestack_b200/r/ns_rshim.c against the stub R API of tests/rstub/ (TEST INFRASTRUCTURE: R is not
installed in this image) and wraps the harness entry points, so that the .Call routines a maintainer would link
are compiled and driven exactly as R would drive them: registered-routine lookup, argument-count check, R vectors
with dim / names attributes, PROTECT accounting, error() as a non-local exit."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STUB = os.path.join(HERE, "rstub")
SO = os.path.join(STUB, "librstub_shim.so")
_L = None


def build():
    srcs = [os.path.join(STUB, "rstub.c"), os.path.join(ROOT, "needlestack_b200", "r", "ns_rshim.c")]
    deps = srcs + [os.path.join(STUB, "Rinternals.h"), os.path.join(ROOT, "include", "ns_b200.h"),
                   os.path.join(ROOT, "needlestack_b200", "lib", "libns_b200.so")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        libdir = os.path.join(ROOT, "needlestack_b200", "lib")
        subprocess.run(["gcc", "-O1", "-g", "-fPIC", "-shared", "-Wall", "-Werror", "-I" + STUB,
                        "-I" + os.path.join(ROOT, "include"), "-o", SO] + srcs +
                       ["-L" + libdir, "-lns_b200", "-Wl,-rpath," + libdir], check=True)
    return SO


def lib():
    global _L
    if _L is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        for name, res, args in (("rstub_init", C.c_int, []), ("rstub_dynamic_symbols", C.c_int, []),
                                ("rstub_routine_name", C.c_char_p, [C.c_int]), ("rstub_routine_nargs", C.c_int, [C.c_int]),
                                ("rstub_last_error", C.c_char_p, []), ("rstub_depth", C.c_int, []), ("rstub_nil", vp, []),
                                ("rstub_ints", vp, [vp, C.c_ssize_t]), ("rstub_reals", vp, [vp, C.c_ssize_t]),
                                ("rstub_raws", vp, [vp, C.c_ssize_t]), ("rstub_set_dim", None, [vp, vp, C.c_int]),
                                ("rstub_list", vp, [C.c_ssize_t]), ("rstub_list_set", None, [vp, C.c_ssize_t, C.c_char_p, vp]),
                                ("rstub_list_get", vp, [vp, C.c_char_p]), ("rstub_type", C.c_int, [vp]),
                                ("rstub_len", C.c_ssize_t, [vp]), ("rstub_data", vp, [vp]), ("rstub_dim", C.c_int, [vp, C.c_int]),
                                ("rstub_finalize", None, [vp]), ("rstub_call", vp, [C.c_char_p, C.c_int, C.POINTER(vp)])):
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        L.n_routines = L.rstub_init()
        _L = L
    return _L


class RError(RuntimeError):
    pass


def r_int(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    s = lib().rstub_ints(a.ctypes.data_as(C.c_void_p), a.size)
    if a.ndim > 1:  # numpy [P][9][n] in C order == R array dim c(n, 9, P) in column-major order
        d = np.array(a.shape[::-1], dtype=np.int32)
        lib().rstub_set_dim(s, d.ctypes.data_as(C.c_void_p), len(d))
    return s


def r_real(a):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    return lib().rstub_reals(a.ctypes.data_as(C.c_void_p), a.size)


def r_raw(a):
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)
    return lib().rstub_raws(a.ctypes.data_as(C.c_void_p), a.size)


def r_list(items):
    """named list of numbers / SEXPs (an R `list(a = 1, b = 2L)`); names may be None for positional lists"""
    L = lib()
    s = L.rstub_list(len(items))
    for i, (k, v) in enumerate(items):
        if isinstance(v, Sexp):
            v = v.h
        else:
            v = r_real([float(v)])
        L.rstub_list_set(s, i, None if k is None else k.encode(), v)
    return s


class Sexp:
    """marks a list element that is already a stub SEXP (a vector built by r_int / r_real / r_raw)"""

    def __init__(self, h):
        self.h = h


def call(name, *args):
    """.Call(name, ...): NULL result = error() was raised inside the routine"""
    L = lib()
    arr = (C.c_void_p * len(args))(*args)
    d0 = L.rstub_depth()
    r = L.rstub_call(name.encode(), len(args), arr)
    if not r:
        raise RError(L.rstub_last_error().decode())
    assert L.rstub_depth() == d0, f"{name}: PROTECT / UNPROTECT unbalanced ({L.rstub_depth() - d0:+d})"
    return r


def get(lst, name, dtype=None):
    """element `name` of a result list as a numpy array (matrices come back as [cols][rows] = the C layout)"""
    L = lib()
    v = L.rstub_list_get(lst, name.encode())
    assert v, name
    t, n = L.rstub_type(v), L.rstub_len(v)
    np_t = {13: np.int32, 10: np.int32, 14: np.float64, 24: np.uint8}[t]
    a = np.frombuffer((C.c_char * (n * np.dtype(np_t).itemsize)).from_address(L.rstub_data(v)), dtype=np_t).copy()
    if L.rstub_dim(v, 0) >= 0:
        a = a.reshape(L.rstub_dim(v, 1), L.rstub_dim(v, 0))
    return a if dtype is None else a.astype(dtype)
