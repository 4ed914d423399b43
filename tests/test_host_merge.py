"""Host-side bookkeeping of the emitted rows (needlestack_b200/csrc/ns_hostmerge.hpp, pure C++): the late rows of the
call-wide cluster-path pool are merged into the rows already copied out, and the private rows of the devices of a
multi-GPU call are gathered in chunk order — both must leave the rows in increasing unit order with a consistent
per-unit row index (the role of the reference's final sort, needlestack.nf:546-553).  No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def hm():
    src = os.path.join(HERE, "emul", "ns_hostmerge_test.cpp")
    out = os.path.join(HERE, "emul", "libnshostmerge.so")
    dep = os.path.join(ROOT, "needlestack_b200", "csrc", "ns_hostmerge.hpp")
    if not os.path.exists(out) or max(os.path.getmtime(src), os.path.getmtime(dep)) > os.path.getmtime(out):
        subprocess.run(["g++", "-O1", "-g", "-fPIC", "-shared", "-std=c++17", "-o", out, src], check=True)
    L = C.CDLL(out)
    L.hm_merge.restype = C.c_int64
    L.hm_concat.restype = C.c_int64
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("use_emit_unit", [1, 0])
def test_pool_rows_merge_in_unit_order(hm, seed, use_emit_unit):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 7))
    U = int(rng.integers(5, 400))
    emitted = np.sort(rng.choice(U, size=int(rng.integers(0, U // 2 + 1)), replace=False))
    rest = np.setdiff1d(np.arange(U), emitted)
    late = np.sort(rng.choice(rest, size=int(rng.integers(0, min(len(rest), 30) + 1)), replace=False))
    E1, E2 = len(emitted), len(late)
    cap = E1 + E2 + 3
    gq = np.full((cap, n), -1.0)
    gt = np.full((cap, n), 255, dtype=np.uint8)
    pooled = np.full((cap, 8), -1.0)
    eu = np.full(cap, -1, dtype=np.int64)
    gq[:E1] = emitted[:, None] * 10.0 + np.arange(n)
    gt[:E1] = (emitted[:, None] + np.arange(n)) % 4
    pooled[:E1] = emitted[:, None] + 0.5
    eu[:E1] = emitted
    idx = np.full(U, -1, dtype=np.int32)
    idx[emitted] = np.arange(E1)
    # the pool's staging holds its rows in arbitrary (slot) order
    perm = rng.permutation(E2)
    l_gq = np.zeros((max(E2, 1), n))
    l_gt = np.zeros((max(E2, 1), n), dtype=np.uint8)
    l_pooled = np.zeros((max(E2, 1), 8))
    l_gq[perm] = late[:, None] * 10.0 + np.arange(n)
    l_gt[perm] = (late[:, None] + np.arange(n)) % 4
    l_pooled[perm] = late[:, None] + 0.5
    before = np.searchsorted(emitted, late).astype(np.int64)
    src = perm.astype(np.int64)
    E = hm.hm_merge(n, C.c_int64(E1), _p(gq), _p(gt), _p(pooled), _p(eu), use_emit_unit, C.c_int64(E2),
                    _p(late.astype(np.int64)), _p(before), _p(src), _p(l_gq), _p(l_gt), _p(l_pooled), _p(idx),
                    C.c_int64(U))
    assert E == E1 + E2
    allu = np.sort(np.concatenate([emitted, late]))
    assert np.array_equal(gq[:E], allu[:, None] * 10.0 + np.arange(n))
    assert np.array_equal(gt[:E], (allu[:, None] + np.arange(n)) % 4)
    assert np.array_equal(pooled[:E], np.repeat(allu[:, None] + 0.5, 8, axis=1))
    if use_emit_unit:
        assert np.array_equal(eu[:E], allu)
    want = np.full(U, -1, dtype=np.int32)
    want[allu] = np.arange(E)
    assert np.array_equal(idx, want)
    assert np.all(gq[E:] == -1.0)  # nothing written past the merged rows


@pytest.mark.parametrize("seed", range(6))
def test_multi_device_rows_gathered_in_chunk_order(hm, seed):
    rng = np.random.default_rng(100 + seed)
    n = 3
    n_chunks = int(rng.integers(1, 12))
    units_per = int(rng.integers(3, 40))
    dev = rng.integers(0, 2, size=n_chunks).astype(np.int32)
    priv_gq = [[], []]
    priv_eu = [[], []]
    row0 = np.zeros(n_chunks, dtype=np.int64)
    rows = np.zeros(n_chunks, dtype=np.int64)
    unit0 = np.arange(n_chunks, dtype=np.int64) * units_per
    units = np.full(n_chunks, units_per, dtype=np.int64)
    U = n_chunks * units_per
    idx = np.full(U, -1, dtype=np.int32)
    all_units = []
    for c in range(n_chunks):
        em = np.sort(rng.choice(units_per, size=int(rng.integers(0, units_per)), replace=False)) + unit0[c]
        d = int(dev[c])
        row0[c] = len(priv_eu[d])
        rows[c] = len(em)
        idx[em] = row0[c] + np.arange(len(em))  # device-private row numbers
        priv_eu[d] += list(em)
        priv_gq[d] += [u * 10.0 + np.arange(n) for u in em]
        all_units += list(em)
    g = [np.array(x, dtype=np.float64).reshape(-1, n) if x else np.zeros((1, n)) for x in priv_gq]
    e = [np.array(x, dtype=np.int64) if x else np.zeros(1, dtype=np.int64) for x in priv_eu]
    E = len(all_units)
    gq = np.full((E + 2, n), -1.0)
    eu = np.full(E + 2, -1, dtype=np.int64)
    got = hm.hm_concat(n, C.c_int64(E + 2), _p(gq), _p(eu), _p(g[0]), _p(e[0]), _p(g[1]), _p(e[1]), C.c_int64(n_chunks),
                       _p(dev), _p(row0), _p(rows), _p(unit0), _p(units), _p(idx))
    assert got == E
    au = np.array(all_units, dtype=np.int64)
    assert np.array_equal(eu[:E], au) and np.all(np.diff(au) > 0)
    assert np.array_equal(gq[:E], au[:, None] * 10.0 + np.arange(n))
    want = np.full(U, -1, dtype=np.int32)
    want[au] = np.arange(E)
    assert np.array_equal(idx, want)
    # too small a destination: the needed row count comes back negated, nothing is written
    if E > 0:
        assert hm.hm_concat(n, C.c_int64(E - 1), _p(gq), _p(eu), _p(g[0]), _p(e[0]), _p(g[1]), _p(e[1]),
                            C.c_int64(n_chunks), _p(dev), _p(row0), _p(rows), _p(unit0), _p(units), _p(idx)) == -E
