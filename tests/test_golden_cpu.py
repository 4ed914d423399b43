"""CPU suite: the oracle reproduces the committed golden fixtures, and the serial emulation of the
CUDA kernel bodies (tests/emul) agrees with them within the parity tolerances."""
import numpy as np
import pytest

import emul
import golden_util
import orc
import parity


@pytest.mark.parametrize("name", sorted(golden_util.CASES))
def test_oracle_reproduces_golden(name):
    G = golden_util.load(name)
    kw, pairs = golden_util.case_params(name)
    O = orc.call_snv_batch(G["ref"], np.ascontiguousarray(G["counts"].transpose(1, 0, 2)),
                           orc.call_params(pairs=pairs, **kw), n_threads=4)
    for k in ("gated", "extra_rob", "ref_inv", "gate1", "gate2", "n_outer", "gt", "status"):
        assert np.array_equal(O[k], G[k]), k
    for k in ("sigma", "slope", "pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs", "pooled", "max_gq"):
        a, b = O[k], G[k]
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        m = ~np.isnan(a) & np.isfinite(b)
        assert np.allclose(a[m], b[m], rtol=1e-11, atol=1e-13), k


@pytest.mark.parametrize("name,fast", [(n, f) for n in sorted(golden_util.CASES) for f in (True, False)
                                       if not (n == "c4_tn_pairs" and not f)])
def test_kernel_bodies_emulated_vs_golden(name, fast):
    """fast=True: the warp kernel's table/recurrence path with hand-over of spline units;
    fast=False: the generic path."""
    G = golden_util.load(name)
    kw, pairs = golden_util.case_params(name)
    E = parity.emul_snv_batch(emul, G["ref"], G["counts"], emul.make_params(pairs=pairs, **kw), pairs=pairs, fast=fast)
    rep = parity.compare(G, E, label=name)
    assert rep["n_fitted"] == int((G["gated"] == 0).sum())
    assert rep["gt_mismatch"] == 0 and rep["status_mismatch"] == 0 and rep["gate_mismatch"] == 0


def test_emulated_numeric_layer_matches_oracle():
    L, O = emul.lib(), orc.lib()
    rng = np.random.default_rng(4)
    for _ in range(500):
        size = 10 ** rng.uniform(-1, 3)
        mu = 10 ** rng.uniform(-4, 3)
        y = float(rng.integers(0, 80))
        a, b = L.emul_dnbinom_mu(y, size, mu), O.orc_dnbinom_mu(y, size, mu)
        assert abs(a - b) <= 2e-12 * b + 1e-300  # |log pmf| up to ~700: eps*|log| relative
        a = L.emul_nb_pvalue(y, size, mu)
        b = O.orc_dnbinom_mu(y, size, mu) + O.orc_pnbinom_mu_upper(y, size, mu)
        assert abs(a - b) <= 5e-12 * b + 1e-300
    for x in (0, 1, 30, 137, 1000, 52341):
        assert L.emul_qnbinom_mu(0.01, 10.0, 0.5 * x) == O.orc_qnbinom_mu(0.01, 10.0, 0.5 * x)
        assert L.emul_qbinom(0.01, float(x), 0.01) == O.orc_qbinom(0.01, float(x), 0.01)
    for t in ((3, 1, 50, 48), (0, 0, 10, 12), (250, 3, 5000, 4800), (7, 7, 7, 7), (0, 0, 0, 0)):
        a, b = L.emul_fisher(*map(float, t)), O.orc_fisher_2x2(*map(float, t))
        assert abs(a - b) <= 1e-11 * b


def test_window_expectations_sum_and_spline_branches():
    """E.tukeypsi.1/.2 and ai.sig.tukey of the kernel code vs direct lattice sums / the oracle's spline path"""
    L = emul.lib()
    import ctypes as C
    for mu, sig, c in ((0.3, 0.05, 5.0), (2.0, 1.0, 5.0), (40.0, 0.01, 3.0), (100.0, 0.05, 5.0), (500.0, 0.01, 3.0),
                       (22.0, 10.0, 3.0)):
        e1, e2 = C.c_double(), C.c_double()
        L.emul_E12(mu, sig, c, 100, C.byref(e1), C.byref(e2))
        ai = L.emul_ai(mu, sig, 3.0, 100)
        r = orc.glm_params()
        # the oracle exposes the same quantities only through the fit; compare with plain sums when K <= 200
        sd = np.sqrt(mu + sig * mu * mu)
        j1, j2 = max(np.ceil(mu - c * sd), 0), np.floor(mu + c * sd)
        if j2 - j1 + 1 <= 200:
            j = np.arange(j1, j2 + 1)
            pm = np.array([orc.lib().orc_dnbinom_mu(float(v), 1 / sig, mu) for v in j])
            w = (((j - mu) / (c * sd)) ** 2 - 1) ** 2
            assert abs(e1.value - np.sum(w * (j - mu) * pm) / sd) <= 1e-12 * max(1, abs(e1.value))
            assert abs(e2.value - np.sum(w * (j - mu) ** 2 * pm) / sd) <= 1e-12 * max(1, abs(e2.value))
        assert np.isfinite(ai)


def test_calls_mode_early_out_keeps_the_calls():
    """NS_EMIT_CALLS skips the multiple-testing step of units whose smallest p-value rules out a call:
    the gate flags and every row of the units that ARE calls must be unchanged"""
    from needlestack_b200 import _abi
    G = golden_util.load("c1_default")
    full = parity.emul_snv_batch(emul, G["ref"], G["counts"], emul.make_params(emit_mode=_abi.NS_EMIT_ALL))
    calls = parity.emul_snv_batch(emul, G["ref"], G["counts"], emul.make_params(emit_mode=_abi.NS_EMIT_CALLS))
    g = _abi.NS_F_GATE1 | _abi.NS_F_GATE2 | _abi.NS_F_GATED | _abi.NS_F_REF_INV
    assert np.array_equal(full["flags"] & g, calls["flags"] & g)
    sel = (full["flags"] & _abi.NS_F_GATE1) != 0
    assert sel.sum() > 3
    for k in ("gq", "qval_minaf", "sor", "gt", "status", "max_gq"):
        assert np.array_equal(full[k][sel], calls[k][sel], equal_nan=True), k
    skipped = ((full["flags"] & (_abi.NS_F_GATED | _abi.NS_F_GATE1)) == 0)
    assert np.all(np.isnan(calls["max_gq"][skipped]) | (calls["max_gq"][skipped] < 50))
