"""The R side, checked without R: needlestack_b200/r/ns_rshim.c - the only code a maintainer links into R - compiled
(-Wall -Werror) against a stub of the R API (tests/rstub/, test infrastructure) and driven through the registered
.Call table the way R drives it.  CPU part: it builds, registers its routines with the right arity, refuses wrong
argument counts, and turns a failing ns_create() into an R error.  GPU part (-m gpu): every routine against the
ctypes binding on the same inputs - catches PROTECT, index, layout (column-major matrices = [U][n] rows) and
ownership mistakes.  Reference interface: bin/glm_rob_nb.r:16-18, bin/needlestack.r:312-413."""
import numpy as np
import pytest

import rshim_util as R

ROUTINES = {"C_ns_create": 1, "C_ns_glmrob_nb": 4, "C_ns_call_snv": 6, "C_ns_call_units": 8, "C_ns_call_units_rows": 9,
            "C_ns_qvalue_grid": 7,
            "C_ns_create_multi": 1, "C_ns_multi_call_snv": 6}


def test_shim_builds_and_registers_its_routines():
    L = R.lib()
    got = {L.rstub_routine_name(i).decode(): L.rstub_routine_nargs(i) for i in range(L.n_routines)}
    assert got == ROUTINES
    assert L.rstub_dynamic_symbols() == 0  # R_useDynamicSymbols(dll, FALSE): only the registered names resolve
    with pytest.raises(R.RError, match="Incorrect number of arguments"):
        R.call("C_ns_glmrob_nb", L.rstub_nil())
    with pytest.raises(R.RError, match="not in load table"):
        R.call("C_ns_nonexistent", L.rstub_nil())


def test_create_failure_is_an_r_error_not_a_crash():
    """an impossible device ordinal: ns_create() fails, the shim must raise error() with the library's message (and on
    a machine without a GPU this is every ordinal: the library has no CPU path)"""
    with pytest.raises(R.RError, match="no usable CUDA device|CUDA resource creation failed"):
        R.call("C_ns_create", R.r_int([4096]))
    with pytest.raises(R.RError, match="no usable CUDA device|CUDA resource creation failed"):
        R.call("C_ns_create_multi", R.r_int([0, 4096]))
    assert R.lib().rstub_depth() == 0


# ---------------------------------------------------------------------------------------------- GPU part
GLM = dict(min_coverage=30, min_reads=3, extra_rob=0)


@pytest.fixture(scope="module")
def rctx():
    ctx = R.call("C_ns_create", R.r_int([0]))
    yield ctx
    R.lib().rstub_finalize(ctx)   # what R's garbage collector does with the external pointer
    with pytest.raises(R.RError, match="context was destroyed"):
        R.call("C_ns_glmrob_nb", ctx, R.r_int([1]), R.r_int([1]), R.r_list([]))


@pytest.mark.gpu
def test_glmrob_nb_through_the_shim(rctx, caller):
    rng = np.random.default_rng(3)
    for trial in range(4):
        n = int(rng.integers(12, 90))
        x = rng.integers(20, 300, size=n)
        y = rng.poisson(0.004 * x) + (rng.random(n) < 0.05) * rng.binomial(x, 0.3)
        if trial == 3:
            y[:] = 0  # gated: the reference returns NA coefficients, p = q = 1, GQ = 0 (glm_rob_nb.r:48)
        res = R.call("C_ns_glmrob_nb", rctx, R.r_int(y), R.r_int(x), R.r_list(list(GLM.items())))
        want = caller.glmrob_nb(y, x, **GLM)
        sig, slope = R.get(res, "sigma")[0], R.get(res, "slope")[0]
        if np.isnan(want["coef"]["sigma"]):
            assert np.isnan(sig) and np.isnan(slope)
        else:
            assert sig == want["coef"]["sigma"] and slope == want["coef"]["slope"]
        for k_r, k_w in (("pvalues", "pvalues"), ("qvalues", "qvalues"), ("GQ", "GQ")):
            assert np.array_equal(R.get(res, k_r), want[k_w])
        assert bool(R.get(res, "extra_rob")[0]) == want["extra_rob"] and int(R.get(res, "flags")[0]) == want["flags"]
    with pytest.raises(R.RError, match="unknown parameter"):
        R.call("C_ns_glmrob_nb", rctx, R.r_int([1] * 12), R.r_int([50] * 12), R.r_list([("min_covrage", 30)]))
    with pytest.raises(R.RError, match="length"):
        R.call("C_ns_glmrob_nb", rctx, R.r_int([1] * 12), R.r_int([50] * 11), R.r_list([]))


def _snv_case(tn):
    from needlestack_b200 import synth
    cfg = synth.Config(**{**synth.CONFIGS["C4" if tn else "C1"].__dict__, "n": 24, "tn_pairs": 12 if tn else 0})
    ref, counts = synth.generate(cfg, P=700, seed=55, p_variant=0.03, p_germline=0.01, p_refswap=0.01)
    ref[9::97] = 4
    return cfg, ref, counts, (synth.tn_pairs(cfg) if tn else None)


@pytest.mark.gpu
@pytest.mark.parametrize("tn", [False, True])
@pytest.mark.parametrize("routine", ["C_ns_call_snv", "C_ns_multi_call_snv"])
def test_call_snv_through_the_shim(rctx, caller, tn, routine):
    from needlestack_b200 import make_params
    cfg, ref, counts, pairs = _snv_case(tn)
    want = caller.call_snv(ref, counts, prm=make_params(pairs=pairs, min_af=0.0))
    tn_arg = R.lib().rstub_nil() if not tn else R.r_list([(None, R.Sexp(R.r_int(pairs[k]))) for k in
                                                         ("tindex", "nindex", "only_t", "only_n")])
    ctx = rctx if routine == "C_ns_call_snv" else R.call("C_ns_create_multi", R.r_int([0, 0]))
    params = R.r_list([("min_coverage", 30), ("min_reads", 3), ("min_af", 0), ("GQ_threshold", 50), ("SB_type", 0),
                       ("sigma_normal", 0.1), ("afmin_power", -1)])
    # too small a cap first: n_emitted reports what is needed, the caller asks again
    res = R.call(routine, ctx, R.r_int(counts), R.r_raw(ref), params, tn_arg, R.r_real([2]))
    need = int(R.get(res, "n_emitted")[0])
    assert need == want.n_emitted > 2
    res = R.call(routine, ctx, R.r_int(counts), R.r_raw(ref), params, tn_arg, R.r_real([need + 5]))
    E = need
    assert np.array_equal(R.get(res, "emit_unit")[:E].astype(np.int64), want.emit_unit)
    for k in ("sigma", "slope", "max_gq"):
        assert np.array_equal(R.get(res, k), want[k], equal_nan=True), k
    assert np.array_equal(R.get(res, "flags").view(np.uint32), want.flags)
    assert np.array_equal(R.get(res, "emit_index"), want.emit_index)
    for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs", "gt", "status", "pooled"):
        m = R.get(res, k)                      # R matrix [n x cap] column-major == [cap][n] rows
        assert m.shape == (need + 5, 8 if k == "pooled" else cfg.n)
        assert np.array_equal(m[:E], want[k], equal_nan=m.dtype.kind == "f"), k
    if tn:
        assert (want.status != 0).any()
    if routine != "C_ns_call_snv":
        R.lib().rstub_finalize(ctx)


@pytest.mark.gpu
def test_call_units_and_qvalue_grid_through_the_shim(rctx, caller):
    from needlestack_b200 import NS_EMIT_ALL, make_params
    import parity
    cfg, ref, counts, _ = _snv_case(False)
    units = [u for u in range(0, 3 * len(ref), 7) if ref[u // 3] < 4][:120]
    rows = [np.stack([parity.snv_unit_rows(ref, counts, u)[k] for u in units]) for k in range(5)]
    ind = (np.arange(len(units)) % 2).astype(np.uint8)
    want = caller.call_units(*rows, is_indel=ind, prm=make_params(emit_mode=NS_EMIT_ALL))
    params = R.r_list([("min_coverage", 30), ("min_reads", 3)])
    res = R.call("C_ns_call_units", rctx, *[R.r_int(r) for r in rows], R.r_raw(ind), params)
    assert np.array_equal(R.get(res, "sigma"), want.sigma, equal_nan=True)
    assert np.array_equal(R.get(res, "flags").view(np.uint32), want.flags)
    for k in ("pvalue", "qvalue", "gq"):
        assert np.array_equal(R.get(res, k), want.dense(k), equal_nan=True), k
    # optional planes absent (annotate_vcf.r gives DP and AO only): NULL arguments
    nil = R.lib().rstub_nil()
    res2 = R.call("C_ns_call_units", rctx, R.r_int(rows[0]), R.r_int(rows[1] + rows[2]), nil, nil, nil, nil, params)
    want2 = caller.call_units(rows[0], rows[1] + rows[2], prm=make_params(emit_mode=NS_EMIT_ALL))
    assert np.array_equal(R.get(res2, "gq"), want2.dense("gq"), equal_nan=True)
    # caller-level rows of explicit units (the indel blocks of the drop-in driver): the library's default emission
    want3 = caller.call_units(*rows, is_indel=ind, prm=make_params(min_af=0.0))
    res3 = R.call("C_ns_call_units_rows", rctx, *[R.r_int(r) for r in rows], R.r_raw(ind), params, nil)
    E3 = int(R.get(res3, "n_emitted")[0])
    assert E3 == want3.n_emitted > 0
    assert np.array_equal(R.get(res3, "emit_unit")[:E3].astype(np.int64), want3.emit_unit)
    for k in ("gq", "qval_minaf", "sor", "gt", "status", "pooled"):
        assert np.array_equal(R.get(res3, k)[:E3], want3[k], equal_nan=k not in ("gt", "status")), k
    # the contour grid of plot_rob_nb (bin/plot_rob_nb.r:83-87)
    fitted = np.where(~np.isnan(want.sigma))[0]
    u = int(fitted[0])
    gx, gy = np.meshgrid(np.arange(10, 400, 13), np.arange(0, 40, 3))
    q = R.call("C_ns_qvalue_grid", rctx, R.r_real([want.sigma[u]]), R.r_real([want.slope[u]]), R.r_int(rows[0][u]),
               R.r_int(rows[1][u] + rows[2][u]), R.r_int(gx.reshape(-1)), R.r_int(gy.reshape(-1)))
    L = R.lib()
    import ctypes as C
    got = np.frombuffer((C.c_char * (8 * gx.size)).from_address(L.rstub_data(q)), dtype=np.float64)
    wantq = caller.qvalue_grid(want.sigma[u], want.slope[u], rows[0][u], rows[1][u] + rows[2][u], gx.reshape(-1), gy.reshape(-1))
    assert np.array_equal(got, wantq)
