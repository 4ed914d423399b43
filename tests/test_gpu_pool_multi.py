"""The call-wide pool of cluster-path units and the multi-GPU call: both only change WHEN and WHERE units are
fitted, never the results.  Pool on / off and one / several contexts must agree bit for bit (every floating-point sum
inside a unit is ordered by the unit's own geometry, not by its neighbours), rows must stay in unit order, and the
pooled batch must still meet the oracle."""
import numpy as np
import pytest

import parity
import needlestack_b200 as ns
from needlestack_b200 import NS_EMIT_ALL, NS_EMIT_CALLS, NS_EMIT_FITTED, make_params, synth

pytestmark = pytest.mark.gpu

KEYS_UNIT = ("sigma", "slope", "max_gq", "flags", "iters", "emit_index")
KEYS_ROWS = ("emit_unit", "pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs", "gt", "status", "pooled")


def _same(A, B, label):
    assert A.n_emitted == B.n_emitted and A.n_fitted == B.n_fitted and A.n_gated == B.n_gated, label
    assert A.n_skipped == B.n_skipped, label
    for k in KEYS_UNIT + KEYS_ROWS:
        assert np.array_equal(A[k], B[k], equal_nan=A[k].dtype.kind == "f"), (label, k)


def _batch(P=2400, seed=31, cfgname="C2", pg=0.01, prs=0.004):
    cfg = synth.CONFIGS[cfgname]
    ref, counts = synth.generate(cfg, P=P, seed=seed, p_variant=0.01, p_germline=pg, p_refswap=prs)
    ref[5::211] = 4  # a few skipped positions
    return ref, counts


@pytest.mark.parametrize("emit_mode", [NS_EMIT_CALLS, NS_EMIT_FITTED, NS_EMIT_ALL])
def test_pool_on_off_bit_identical_over_several_chunks(monkeypatch, emit_mode):
    ref, counts = _batch()
    monkeypatch.setenv("NS_CHUNK_UNITS", str(3 * 500))  # 5 chunks
    res = {}
    for pool in ("1", "0"):
        monkeypatch.setenv("NS_POOL", pool)
        with ns.Caller(0) as c:
            res[pool] = c.call_snv(ref, counts, prm=make_params(emit_mode=emit_mode))
            tm = c.timings()
            assert tm["chunks"] == 5
            if pool == "1":
                assert tm["pool_units"] > 0 and tm["pool_units"] == tm["n_heavy"]
            else:
                assert tm["pool_units"] == 0 and tm["n_heavy"] > 0
    _same(res["1"], res["0"], f"pool on/off emit_mode {emit_mode}")
    R = res["1"]
    assert np.all(np.diff(R.emit_unit) > 0)  # rows in unit order, pooled ones merged in
    idx = R.emit_index
    assert np.array_equal(idx[R.emit_unit], np.arange(R.n_emitted)) and (idx >= 0).sum() == R.n_emitted


def test_pooled_batch_meets_the_oracle(oracle, monkeypatch):
    ref, counts = _batch(P=1500, seed=32)
    monkeypatch.setenv("NS_CHUNK_UNITS", str(3 * 400))
    O = oracle.call_snv_batch(ref, np.ascontiguousarray(counts.transpose(1, 0, 2)), oracle.call_params(), n_threads=8)
    with ns.Caller(0) as c:
        R = c.call_snv(ref, counts, prm=make_params(emit_mode=NS_EMIT_ALL))
        assert c.timings()["pool_units"] > 0
    d = {k: R.dense(k) for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs", "pooled")}
    d["gt"] = R.dense("gt", fill=0)
    d["status"] = R.dense("status", fill=0)
    for k in ("sigma", "slope", "max_gq", "flags", "n_outer", "n_irls", "n_obj"):
        d[k] = R[k]
    rep = parity.compare(O, d, label="pooled")
    assert rep["n_fitted"] == R.n_fitted and rep["nonconverged_compared"] == rep["nonconverged"] > 0
    # ... and the same through the explicit-unit entry point, where the pool's rows may be absent (vm/rp/rm = NULL)
    units = np.where((O["gated"] == 0) & (O["n_quad"] > 0) & (O["ref_inv"] == 0))[0][:6]
    assert len(units) > 0
    rows = [np.stack([parity.snv_unit_rows(ref, counts, int(u))[k] for u in units]) for k in range(5)]
    with ns.Caller(0) as c:
        A = c.call_units(rows[0], rows[1] + rows[2], plain=True, prm=make_params(emit_mode=NS_EMIT_ALL))
        assert c.timings()["pool_units"] == len(units)
    assert np.array_equal(A.sigma, R.sigma[units]) and np.array_equal(A.slope, R.slope[units])
    assert np.array_equal(A.gq, R.dense("gq")[units])


def test_pool_overflow_falls_back_to_the_in_chunk_kernel(monkeypatch):
    ref, counts = _batch(P=3000, seed=34)
    monkeypatch.setenv("NS_POOL_CAP", "16")
    monkeypatch.setenv("NS_CHUNK_UNITS", str(3 * 300))
    with ns.Caller(0) as c:
        A = c.call_snv(ref, counts, prm=make_params(emit_mode=NS_EMIT_ALL))
        tm = c.timings()
        assert 0 < tm["pool_units"] <= 16 and tm["n_heavy"] > tm["pool_units"]
    monkeypatch.delenv("NS_POOL_CAP")
    with ns.Caller(0) as c:
        B = c.call_snv(ref, counts, prm=make_params(emit_mode=NS_EMIT_ALL))
    _same(A, B, "pool overflow")


def test_glmrob_nb_of_a_cluster_path_unit_goes_through_the_pool(caller, oracle):
    rng = np.random.default_rng(5)
    x = rng.integers(80, 400, size=40)
    y = rng.binomial(x, 0.45)
    R = caller.glmrob_nb(y, x, min_coverage=30, min_reads=3, extra_rob=0)
    assert caller.timings()["pool_units"] == 1
    O = oracle.glmrob_nb(y.astype(float), x.astype(float), min_coverage=30, min_reads=3, extra_rob=0)
    assert O["info"].n_quad > 0
    assert abs(R["coef"]["sigma"] - O["sigma"]) <= max(1e-6 * O["sigma"], parity.SIGMA_ABS)
    assert abs(R["coef"]["slope"] - O["slope"]) <= 1e-6 * O["slope"]
    assert np.max(np.abs(R["GQ"] - O["GQ"]) / np.maximum(1, O["GQ"])) <= 1e-6


@pytest.mark.parametrize("emit_mode", [NS_EMIT_CALLS, NS_EMIT_ALL])
def test_multi_context_call_equals_single_context_call(monkeypatch, emit_mode):
    """two contexts on device 0 (the multi-GPU code path on a one-GPU box): one queue of chunks, two host threads,
    private rows gathered in chunk order"""
    ref, counts = _batch(P=3000, seed=34)
    with ns.Caller(0) as c:
        A = c.call_snv(ref, counts, prm=make_params(emit_mode=emit_mode))
    monkeypatch.setenv("NS_MULTI_CHUNK_POSITIONS", "250")  # 12 chunks
    with ns.MultiCaller([0, 0]) as m:
        B = m.call_snv(ref, counts, prm=make_params(emit_mode=emit_mode))
        tm = m.timings()
        assert sum(t["chunks"] for t in tm) == 12 and all(t["chunks"] > 0 for t in tm)
        # a too small emit_cap is reported with the needed row count, then succeeds
        if emit_mode == NS_EMIT_CALLS and A.n_emitted > 1:
            with pytest.raises(ns.NeedlestackError):
                m.call_snv(ref, counts, prm=make_params(emit_mode=emit_mode), emit_cap=A.n_emitted - 1)
        B2 = m.call_snv(ref, counts[..., :].astype(np.uint16), prm=make_params(emit_mode=emit_mode))
    _same(A, B, "multi vs single")
    _same(A, B2, "multi u16 vs single")


def test_multi_gpu_call_over_all_visible_devices(monkeypatch):
    import torch
    nd = torch.cuda.device_count()
    if nd < 2:
        pytest.skip("one GPU visible: the multi-device path is covered by the two-context test")
    ref, counts = _batch(P=4000, seed=35)
    with ns.Caller(0) as c:
        A = c.call_snv(ref, counts, prm=make_params())
    monkeypatch.setenv("NS_MULTI_CHUNK_POSITIONS", "200")
    with ns.MultiCaller(list(range(nd))) as m:
        B = m.call_snv(ref, counts, prm=make_params())
        assert sum(t["chunks"] for t in m.timings()) == 20
    _same(A, B, f"{nd} GPUs vs one")


@pytest.mark.parametrize("multi", [False, True])
def test_rows_only_call_returns_the_same_rows(monkeypatch, multi):
    """per_unit=False (all per-unit output pointers NULL): the emitted rows are unchanged and carry their unit's sigma,
    slope, max(GQ) and flags per row - including the rows that come from the pool and from other contexts"""
    ref, counts = _batch(P=2400, seed=36)
    monkeypatch.setenv("NS_CHUNK_UNITS", str(3 * 700))
    monkeypatch.setenv("NS_MULTI_CHUNK_POSITIONS", "300")
    mk = (lambda: ns.MultiCaller([0, 0])) if multi else (lambda: ns.Caller(0))
    with mk() as c:
        A = c.call_snv(ref, counts, prm=make_params())
        B = c.call_snv(ref, counts, prm=make_params(), per_unit=False)
    assert "sigma" not in B and "emit_index" not in B
    assert A.n_emitted == B.n_emitted > 10 and A.n_fitted == B.n_fitted and A.n_skipped == B.n_skipped
    for k in KEYS_ROWS:
        assert np.array_equal(A[k], B[k], equal_nan=A[k].dtype.kind == "f"), k
    for R in (A, B):
        assert np.array_equal(R.row_sigma, A.sigma[A.emit_unit]) and np.array_equal(R.row_slope, A.slope[A.emit_unit])
        assert np.array_equal(R.row_max_gq, A.max_gq[A.emit_unit]) and np.array_equal(R.row_flags, A.flags[A.emit_unit])
    assert (A.flags[A.emit_unit] & ns.NS_F_GATE2).all()
