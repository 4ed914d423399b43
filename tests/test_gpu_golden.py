"""CUDA path (C ABI) vs the committed golden fixtures, the explicit-unit and glmrob.nb entry points,
and size-independent properties at larger sizes."""
import os

import numpy as np
import pytest

import golden_util
import parity
import needlestack_b200 as ns
from needlestack_b200 import NS_EMIT_ALL, NS_EMIT_CALLS, NS_EMIT_FITTED, _abi, make_params, synth

pytestmark = pytest.mark.gpu


def _dense(R):
    out = {k: R.dense(k) for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs", "pooled")}
    out["gt"] = R.dense("gt", fill=0)
    out["status"] = R.dense("status", fill=0)
    for k in ("sigma", "slope", "max_gq", "flags", "n_outer"):
        out[k] = R[k]
    return out


@pytest.mark.parametrize("name", sorted(golden_util.CASES))
def test_golden(caller, name):
    G = golden_util.load(name)
    kw, pairs = golden_util.case_params(name)
    R = caller.call_snv(G["ref"], G["counts"], prm=make_params(pairs=pairs, emit_mode=NS_EMIT_ALL, **kw))
    rep = parity.compare(G, _dense(R), label=name)
    assert rep["n_fitted"] == int((G["gated"] == 0).sum()) == R.n_fitted
    # the default emission = exactly the units the reference writes to the VCF (needlestack.r:387)
    R2 = caller.call_snv(G["ref"], G["counts"], prm=make_params(pairs=pairs, emit_mode=NS_EMIT_CALLS, **kw))
    assert np.array_equal(R2.emit_unit, np.where(G["gate2"] != 0)[0])
    assert np.array_equal(R2.gt, G["gt"][R2.emit_unit]) and np.array_equal(R2.status, G["status"][R2.emit_unit])


def test_explicit_units_match_snv_path_and_indel_semantics(caller, oracle):
    """ns_call_units with the five rows of each SNV unit == ns_call_snv; is_indel switches the SB threshold,
    all_DP (needlestack.r:536,700) and disables output_all (needlestack.r:534,698)"""
    G = golden_util.load("c1_default")
    ref, counts = G["ref"], G["counts"]
    ok = np.repeat(ref < 4, 3)
    units = np.where(ok)[0]
    rows = [np.stack([parity.snv_unit_rows(ref, counts, int(u))[k] for u in units]) for k in range(5)]
    prm = make_params(emit_mode=NS_EMIT_ALL)
    A = caller.call_units(*rows, prm=prm)
    B = caller.call_snv(ref, counts, prm=prm)
    assert np.array_equal(A.flags, B.flags[units])
    assert np.array_equal(A.sigma, B.sigma[units], equal_nan=True)
    assert np.array_equal(A.dense("gq"), B.dense("gq")[units], equal_nan=True)
    # indel units against the oracle's is_indel path
    sel = units[(B.flags[units] & _abi.NS_F_GATE1) != 0][:6]
    prm_i = make_params(emit_mode=NS_EMIT_ALL, SB_threshold_indel=0.5, SB_threshold_SNV=100.0)
    rows_i = [np.stack([parity.snv_unit_rows(ref, counts, int(u))[k] for u in sel]) for k in range(5)]
    I = caller.call_units(*rows_i, is_indel=np.ones(len(sel), dtype=np.uint8), prm=prm_i)
    op = oracle.call_params(SB_threshold_indel=0.5, SB_threshold_SNV=100.0)
    for k, u in enumerate(sel):
        O = oracle.call_unit(*[r[k] for r in rows_i], op, is_indel=True)
        assert bool(I.flags[k] & _abi.NS_F_GATE2) == bool(O["info"].gate2)
        assert np.array_equal(I.dense("gt", fill=0)[k], O["gt"])
        assert abs(I.dense("pooled")[k][7] - O["info"].pooled[7]) < 1e-9


def test_glmrob_nb_entry_point(caller, oracle):
    """the mirror of glmrob.nb(y, x, ...) with its own defaults (extra_rob=TRUE, min_coverage=1, ...)"""
    rng = np.random.default_rng(21)
    for trial in range(12):
        n = int(rng.integers(12, 150))
        x = rng.integers(20, 400, size=n)
        y = rng.poisson(10 ** rng.uniform(-3, -1.5) * x)
        if trial % 3 == 0:
            y[rng.integers(0, n, size=2)] += rng.integers(10, 60, size=2)
        y = np.minimum(y, x)
        R = caller.glmrob_nb(y, x)
        O = oracle.glmrob_nb(y.astype(float), x.astype(float))
        assert R["extra_rob"] == O["extra_rob"]
        if O["info"].gated:
            assert np.isnan(R["coef"]["sigma"]) and np.all(R["pvalues"] == 1) and np.all(R["GQ"] == 0)
            continue
        if O["info"].maxit_hit:
            continue
        assert abs(R["coef"]["sigma"] - O["sigma"]) <= max(1e-6 * O["sigma"], parity.SIGMA_ABS)
        assert abs(R["coef"]["slope"] - O["slope"]) <= 1e-6 * O["slope"]
        assert np.all(np.abs(R["GQ"] - O["GQ"]) <= 1e-6 * np.maximum(1, O["GQ"]))
        assert np.all(np.abs(R["pvalues"] - O["pvalues"]) <= 1e-6 * np.maximum(1e-300, O["pvalues"]) + 1e-300)
    with pytest.raises(ns.NeedlestackError):
        caller.glmrob_nb([1.5, 2], [10, 20])


def test_edge_cases(caller, oracle):
    prm = make_params(emit_mode=NS_EMIT_ALL)
    # empty batch
    R = caller.call_snv(np.zeros(0, dtype=np.uint8), np.zeros((0, 9, 20), dtype=np.int32), prm=prm)
    assert R.n_emitted == 0 and R.n_fitted == 0
    # fewer than size_min = 10 covered samples can never be fitted (glm_rob_nb.r:38)
    cfg = synth.Config("tiny", 9, 50, 100, -3, -2, -2, 0, seed=5)
    ref, counts = synth.generate(cfg, P=50, seed=5)
    R = caller.call_snv(ref, counts, prm=prm)
    assert R.n_fitted == 0 and np.all(R.flags & _abi.NS_F_GATED)
    # all-zero depth, all-REF and all-ALT (ref-inversion) positions
    n = 24
    counts = np.zeros((3, 9, n), dtype=np.int32)
    counts[1, 0] = 100
    counts[1, 1] = 50
    counts[1, 5] = 50  # A fwd/rev = reference
    counts[2, 0] = 100
    counts[2, 2] = 48
    counts[2, 6] = 50  # all T although REF = A -> inverted regression on the REF counts
    counts[2, 1] = 1
    counts[2, 5] = 1
    ref = np.zeros(3, dtype=np.uint8)
    R = caller.call_snv(ref, counts, prm=prm)
    O = oracle.call_snv_batch(ref, np.ascontiguousarray(counts.transpose(1, 0, 2)), oracle.call_params())
    parity.compare(O, _dense(R), label="edge")
    assert R.flags[6] & _abi.NS_F_REF_INV  # unit 2*3+0 = alt T at position 2


def test_properties_at_scale(caller):
    """determinism, chunking independence, host == device-resident path, emission modes nest"""
    cfg = synth.CONFIGS["C2"]
    ref, counts = synth.generate(cfg, P=6000, seed=77)
    prm = make_params(emit_mode=NS_EMIT_FITTED)
    A = caller.call_snv(ref, counts, prm=prm)
    B = caller.call_snv(ref, counts, prm=prm)
    for k in ("sigma", "slope", "flags", "max_gq"):
        assert np.array_equal(A[k], B[k], equal_nan=True), k
    assert np.array_equal(A.gq, B.gq, equal_nan=True)
    os.environ["NS_CHUNK_UNITS"] = "3000"
    try:
        with ns.Caller(0) as c2:
            C2 = c2.call_snv(ref, counts, prm=prm)
    finally:
        del os.environ["NS_CHUNK_UNITS"]
    assert np.array_equal(A.flags, C2.flags) and np.array_equal(A.emit_unit, C2.emit_unit)
    assert np.array_equal(A.sigma, C2.sigma, equal_nan=True) and np.array_equal(A.gq, C2.gq, equal_nan=True)
    import torch
    d_ref = torch.from_numpy(ref).cuda()
    d_counts = torch.from_numpy(counts).cuda()
    torch.cuda.synchronize()
    D = caller.call_snv_device(d_ref.data_ptr(), d_counts.data_ptr(), len(ref), cfg.n, prm)
    assert np.array_equal(A.sigma, D.sigma, equal_nan=True) and np.array_equal(A.gq, D.gq, equal_nan=True)
    calls = caller.call_snv(ref, counts, prm=make_params(emit_mode=NS_EMIT_CALLS))
    assert set(calls.emit_unit) <= set(A.emit_unit)
    assert A.n_fitted + A.n_gated + A.n_skipped == 3 * len(ref)
    # gate flags are what the thresholds say
    gq = A.dense("gq")
    g1 = (A.flags & _abi.NS_F_GATE1) != 0
    fitted = (A.flags & _abi.NS_F_GATED) == 0
    assert np.array_equal(g1[fitted], (np.nanmax(gq[fitted], axis=1) >= 50))
    # sample order does not matter to the fit (a permutation of the samples)
    perm = np.random.default_rng(1).permutation(cfg.n)
    Pm = caller.call_snv(ref[:600], counts[:600][:, :, perm], prm=prm)
    A6 = caller.call_snv(ref[:600], counts[:600], prm=prm)
    conv = ((A6.flags & (_abi.NS_F_GATED | _abi.NS_F_MAXIT)) == 0)
    assert np.array_equal(Pm.flags & _abi.NS_F_GATED, A6.flags & _abi.NS_F_GATED)
    assert np.allclose(Pm.sigma[conv], A6.sigma[conv], rtol=1e-6, atol=2e-8)
    assert np.allclose(Pm.slope[conv], A6.slope[conv], rtol=1e-6)


def test_tn_index_validation(caller):
    G = golden_util.load("c1_default")
    with pytest.raises(ns.NeedlestackError, match="out of range"):
        caller.call_snv(G["ref"], G["counts"], prm=make_params(pairs=dict(tindex=[0], nindex=[99])))


def test_u16_wire_format_is_identical(caller):
    """ns_call_snv_u16: the 16-bit planes are widened on the device; results are those of the int32 call, bit for bit"""
    import numpy as np
    from needlestack_b200 import NS_EMIT_ALL, make_params, synth
    ref, counts = synth.generate(synth.CONFIGS["C2"], P=1501, seed=77, p_variant=0.02, p_germline=0.002)
    assert counts.max() < 65536
    prm = make_params(emit_mode=NS_EMIT_ALL)
    A = caller.call_snv(ref, counts, prm=prm)
    B = caller.call_snv(ref, counts.astype(np.uint16), prm=prm)
    assert A.n_fitted == B.n_fitted > 50 and A.n_emitted == B.n_emitted
    for k in ("sigma", "slope", "flags", "max_gq"):
        assert np.array_equal(A[k], B[k], equal_nan=True), k
    for k in ("pvalue", "gq", "qval_minaf", "sor"):
        assert np.array_equal(A.dense(k), B.dense(k), equal_nan=True), k
    for k in ("gt", "status"):
        assert np.array_equal(A.dense(k, fill=0), B.dense(k, fill=0)), k
