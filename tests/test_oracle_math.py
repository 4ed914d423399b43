"""Known-answer tests that pin the oracle's numeric layer (oracle/orc_math.c) — R is not
available here and the reference ships no golden vectors, so every base-R routine the hot path
uses is checked against an independent high-precision or library implementation:
mpmath (exact NB pmf / tails / digamma / hypergeometric), scipy.stats, and scipy.integrate.quad,
which wraps the same QUADPACK dqagse that R's integrate() calls."""
import ctypes as C
import math

import mpmath as mp
import numpy as np
import pytest
import scipy.integrate
import scipy.interpolate
import scipy.stats

import orc

mp.mp.dps = 50


@pytest.fixture(scope="module")
def L():
    return orc.lib()


def nb_pmf_mp(x, size, mu):
    size, mu = mp.mpf(size), mp.mpf(mu)
    p = size / (size + mu)
    return mp.exp(mp.loggamma(x + size) - mp.loggamma(size) - mp.loggamma(x + 1) + size * mp.log(p)
                  + x * mp.log1p(-p))


def nb_upper_mp(q, size, mu):
    """P(Y > q) = I_{mu/(size+mu)}(q+1, size)"""
    size, mu = mp.mpf(size), mp.mpf(mu)
    return mp.betainc(q + 1, size, 0, mu / (size + mu), regularized=True)


GRID = [(x, size, mu) for size in (0.1, 0.37, 1.0, 7.5, 10.0, 123.4, 1000.0)
        for mu in (1e-10, 1e-3, 0.1, 1.0, 9.7, 150.0, 5e3, 1e5)
        for x in (0, 1, 2, 5, 17, 100, 1234, 50000)]


def test_dnbinom_mu_vs_mpmath(L):
    worst = 0
    for x, size, mu in GRID:
        ref = nb_pmf_mp(x, size, mu)
        got = L.orc_dnbinom_mu(float(x), size, mu)
        if ref < mp.mpf("1e-300"):
            assert got < 1e-299
            continue
        rel = abs(mp.mpf(got) - ref) / ref
        worst = max(worst, float(rel))
        assert rel < 5e-13, (x, size, mu, got, ref)
    assert worst < 5e-13


def test_dnbinom_mu_edge_cases(L):
    assert L.orc_dnbinom_mu(0.0, 10.0, 0.0) == 1.0  # x_i = 0 samples: mu = slope*0
    assert L.orc_dnbinom_mu(3.0, 10.0, 0.0) == 0.0
    assert L.orc_dnbinom_mu(2.5, 10.0, 1.0) == 0.0  # R: warning + 0 for non-integer x


def test_pnbinom_upper_vs_mpmath_including_far_tails(L):
    n_mp = 0
    for x, size, mu in GRID:
        got = L.orc_pnbinom_mu_upper(float(x), size, mu)
        if x <= 1234:
            try:
                ref = nb_upper_mp(x, size, mu)
            except ValueError:  # mpmath's hypergeometric series gives up on some corners
                continue
            n_mp += 1
            tol = 2e-12
        else:  # boost's ibeta through scipy for the very large arguments
            ref = mp.mpf(scipy.stats.nbinom.sf(x, size, size / (size + mu)))
            tol = 1e-9
        if ref < mp.mpf("1e-290"):
            assert got < 1e-289
            continue
        rel = abs(mp.mpf(got) - ref) / ref
        assert rel < tol, (x, size, mu, got, float(ref))
    assert n_mp > 250
    # p-values down to 1e-100 and below decide GQ (capped only at 1000)
    for y, size, mu in ((60, 1000.0, 1.0), (200, 10.0, 2.0), (40, 2.5, 0.05)):
        ref = nb_upper_mp(y, size, mu) + nb_pmf_mp(y, size, mu)
        got = L.orc_dnbinom_mu(float(y), size, mu) + L.orc_pnbinom_mu_upper(float(y), size, mu)
        assert ref < 1e-60
        assert abs(mp.mpf(got) - ref) / ref < 2e-12


def test_pnbinom_lower_and_pbinom_vs_scipy(L):
    for size, mu in ((10.0, 50.0), (0.5, 3.0), (10.0, 0.5 * 137)):
        for k in (0, 3, 17, 60):
            ref = scipy.stats.nbinom.cdf(k, size, size / (size + mu))
            assert abs(L.orc_pnbinom_mu_lower(float(k), size, mu) - ref) <= 1e-12 * max(ref, 1e-3)
    for n, p in ((100.0, 0.01), (57.0, 0.3), (20000.0, 0.001)):
        for k in (0, 1, 5, 30):
            ref = scipy.stats.binom.cdf(k, n, p)
            assert abs(L.orc_pbinom_lower(float(k), n, p) - ref) <= 1e-12 * max(ref, 1e-3)


def test_digamma_vs_mpmath(L):
    for x in (0.1, 0.25, 1.0, 1.4616321449683623, 3.7, 11.99, 12.0, 57.3, 1000.0, 1e5 + 0.1):
        ref = mp.digamma(x)
        got = L.orc_digamma(x)
        assert abs(mp.mpf(got) - ref) <= 4e-16 * max(1, abs(ref)) + 1e-16


def _q_def(cdf, p, hi):
    """smallest k with cdf(k) >= p*(1-64 eps) — the definition R's search implements"""
    p = p * (1 - 64 * np.finfo(float).eps)
    k = 0
    while cdf(k) < p and k < hi:
        k += 1
    return k


def test_discrete_quantiles_by_definition(L):
    for x in (0, 1, 7, 30, 100, 137, 1000, 52341):
        mu = 0.5 * x
        size = 10.0
        got = L.orc_qnbinom_mu(0.01, size, mu)
        if mu == 0:
            assert got == 0
            continue
        want = _q_def(lambda k: float(1 - nb_upper_mp(k, size, mu)), 0.01, 10 ** 7) if x <= 1000 else \
            int(scipy.stats.nbinom.ppf(0.01, size, size / (size + mu)))
        assert got == want, (x, got, want)
    for n, pr in ((0, 0.01), (1, 0.01), (50, 0.01), (100, 0.05), (997, 0.01), (50000, 0.001)):
        got = L.orc_qbinom(0.01, float(n), pr)
        want = int(scipy.stats.binom.ppf(0.01, n, pr)) if n else 0
        assert got == want, (n, pr, got, want)


def test_fisher_2x2_vs_scipy(L):
    rng = np.random.default_rng(3)
    tabs = [(3, 1, 50, 48), (0, 0, 10, 12), (5, 0, 0, 7), (12, 15, 40, 38), (1, 9, 60, 2), (0, 4, 30, 30),
            (250, 3, 5000, 4800)]
    tabs += [tuple(int(v) for v in rng.integers(0, 80, size=4)) for _ in range(40)]
    for vp, vm, rp, rm in tabs:
        got = L.orc_fisher_2x2(float(vp), float(vm), float(rp), float(rm))
        # R: fisher.test(matrix(c(Vp,Vm,Rp,Rm), nrow=2)) -> rows (Vp,Rp),(Vm,Rm)
        want = scipy.stats.fisher_exact([[vp, rp], [vm, rm]], alternative="two-sided")[1]
        assert abs(got - want) <= 1e-9 * max(want, 1e-12) + 1e-15, ((vp, vm, rp, rm), got, want)


def _bh(p):
    p = np.asarray(p, float)
    n = len(p)
    o = np.argsort(-p, kind="stable")
    i = np.arange(n, 0, -1)
    q = np.minimum(1, np.minimum.accumulate(n / i * p[o]))
    out = np.empty(n)
    out[o] = q
    return out


def _holm(p):
    p = np.asarray(p, float)
    n = len(p)
    o = np.argsort(p, kind="stable")
    i = np.arange(1, n + 1)
    q = np.minimum(1, np.maximum.accumulate((n - i + 1) * p[o]))
    out = np.empty(n)
    out[o] = q
    return out


def test_padjust_bh_and_holm(L):
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 20, 101):
        p = rng.random(n) ** 3
        if n > 3:
            p[1] = p[2]  # ties
            p[0] = 1.0
        q = np.empty(n)
        L.orc_padjust_bh(n, p.ctypes.data_as(orc.c_dp), q.ctypes.data_as(orc.c_dp))
        np.testing.assert_allclose(q, _bh(p) if n > 1 else p, rtol=0, atol=1e-16)
        L.orc_padjust_holm(n, p.ctypes.data_as(orc.c_dp), q.ctypes.data_as(orc.c_dp))
        np.testing.assert_allclose(q, _holm(p) if n > 1 else p, rtol=0, atol=1e-16)
    # R documentation example values
    p = np.array([0.01, 0.02, 0.03, 0.04, 0.05])
    q = np.empty(5)
    L.orc_padjust_bh(5, p.ctypes.data_as(orc.c_dp), q.ctypes.data_as(orc.c_dp))
    np.testing.assert_allclose(q, [0.05] * 5, atol=1e-15)
    L.orc_padjust_holm(5, p.ctypes.data_as(orc.c_dp), q.ctypes.data_as(orc.c_dp))
    np.testing.assert_allclose(q, [0.05, 0.08, 0.09, 0.09, 0.09], atol=1e-15)


def test_quantile_type7_and_median(L):
    rng = np.random.default_rng(7)
    for n in (1, 2, 5, 10, 11, 100):
        x = np.sort(rng.random(n))
        for pr in (0.25, 0.75, 0.5):
            got = L.orc_quantile7(n, x.ctypes.data_as(orc.c_dp), pr)
            assert abs(got - np.quantile(x, pr, method="linear")) < 1e-15
        y = rng.permutation(x)
        assert abs(L.orc_median(n, y.ctypes.data_as(orc.c_dp)) - np.median(x)) < 1e-16


def test_natural_spline_vs_scipy_and_linear_extrapolation(L):
    rng = np.random.default_rng(11)
    x = np.round(np.linspace(3, 900, 100))
    y = np.exp(-((x - 400) / 150.0) ** 2) * (x - 380)
    b, c, d = np.empty(100), np.empty(100), np.empty(100)
    L.orc_natural_spline(100, *[a.ctypes.data_as(orc.c_dp) for a in (x, y, b, c, d)])
    cs = scipy.interpolate.CubicSpline(x, y, bc_type="natural")
    for u in np.concatenate([rng.uniform(3, 900, 200), x[:5], [900.0]]):
        got = L.orc_spline_eval(100, *[a.ctypes.data_as(orc.c_dp) for a in (x, y, b, c, d)], float(u))
        assert abs(got - cs(u)) <= 1e-11 * max(1, abs(cs(u)))
    # right of the last knot R's natural spline is LINEAR (scipy extrapolates cubically)
    slope = cs(900.0, 1)
    got = L.orc_spline_eval(100, *[a.ctypes.data_as(orc.c_dp) for a in (x, y, b, c, d)], 900.75)
    assert abs(got - (y[-1] + 0.75 * slope)) <= 1e-10 * max(1, abs(y[-1]))


def _spline_fn(x, y):
    cs = scipy.interpolate.CubicSpline(x, y, bc_type="natural")
    s_last = cs(x[-1], 1)

    def f(u):
        return float(cs(u)) if u <= x[-1] else float(y[-1] + (u - x[-1]) * s_last)
    return f


@pytest.mark.parametrize("mu,sig,c,term", [(100.0, 0.05, 5, 1), (100.0, 0.05, 5, 2), (500.0, 0.01, 3, 2),
                                           (5000.0, 0.002, 5, 1), (50000.0, 0.001, 5, 2), (22.0, 10.0, 3, 0)])
def test_dqagse_matches_scipy_quadpack_on_the_reference_integrands(L, mu, sig, c, term):
    """the integrands of glm_rob_nb.r:118,134,154: natural spline through 100 lattice points"""
    sd = math.sqrt(mu + sig * mu * mu)
    j1, j2 = max(math.ceil(mu - c * sd), 0), math.floor(mu + c * sd)
    assert j2 - j1 + 1 > 200
    by = (j2 - j1) / 99.0
    x = np.array([j1] + [round(j1 + k * by) for k in range(1, 99)] + [j2], dtype=float)
    size = 1 / sig
    pm = np.array([L.orc_dnbinom_mu(v, size, mu) for v in x])
    w = (((x - mu) / (c * sd)) ** 2 - 1) ** 2
    if term == 1:
        y = w * (x - mu) * pm
    elif term == 2:
        y = w * (x - mu) ** 2 * pm
    else:
        psi = np.array([L.orc_digamma(v + size) for v in x]) - L.orc_digamma(size) - math.log(mu / size + 1) \
            - (x - mu) / (mu + size)
        y = w * psi * pm
    q = orc.QuadOut()
    ier = L.orc_integrate_spline(100, x.ctypes.data_as(orc.c_dp), y.ctypes.data_as(orc.c_dp), float(x[0]),
                                 float(x[-1] + 1), C.byref(q))
    tol = 2.0 ** -13
    val, err, info = scipy.integrate.quad(_spline_fn(x, y), x[0], x[-1] + 1, epsabs=tol, epsrel=tol, limit=100,
                                          full_output=1)[:3]
    assert ier == 0
    assert q.neval == info["neval"] and q.last == info["last"]
    assert abs(q.result - val) <= 1e-10 * max(abs(val), 1e-300)
    assert abs(q.abserr - err) <= 1e-6 * max(abs(err), 1e-300)


def test_dqagse_extrapolation_path_vs_scipy(L):
    """an integrand hard enough to reach bisection + epsilon extrapolation (log singularity)"""
    x = np.linspace(0, 1, 100)
    x[0] = 1e-9
    y = np.log(x) * np.sqrt(x)
    q = orc.QuadOut()
    L.orc_integrate_spline(100, x.ctypes.data_as(orc.c_dp), y.ctypes.data_as(orc.c_dp), 1e-9, 1.0, C.byref(q))
    tol = 2.0 ** -13
    val, err, info = scipy.integrate.quad(_spline_fn(x, y), 1e-9, 1.0, epsabs=tol, epsrel=tol, limit=100,
                                          full_output=1)[:3]
    assert q.last == info["last"] and q.neval == info["neval"]
    assert abs(q.result - val) <= 1e-9 * abs(val)


def test_zeroin_brent(L):
    CB = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
    L.orc_uniroot.argtypes = [CB, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, orc.c_dp,
                              C.POINTER(C.c_int)]
    root, ne = C.c_double(), C.c_int()
    f = CB(lambda x, _: math.cos(x) - x)
    assert L.orc_uniroot(f, None, 0.0, 1.0, 1e-8, 50, C.byref(root), C.byref(ne)) == 0
    assert abs(root.value - 0.7390851332151607) < 1e-8
    # R's uniroot: same sign at both ends is an error (-> NA -> minsig in glm_rob_nb.r:175-178)
    g = CB(lambda x, _: x * x + 1)
    assert L.orc_uniroot(g, None, 1e-3, 10.0, 1e-8, 50, C.byref(root), C.byref(ne)) != 0
    h = CB(lambda x, _: float("nan"))
    assert L.orc_uniroot(h, None, 1e-3, 10.0, 1e-8, 50, C.byref(root), C.byref(ne)) != 0


def test_pvalue_lower_bound_and_fast_math_helpers():
    """ns_nb_pvalue_lb (the post kernel's call test) never exceeds the exact tail probability and equals it where the
    direct sum applies; ns_flog_tab / ns_fexp_bf (the window seed of k_fit) stay within 1 ulp of mpmath."""
    import mpmath as mp
    import emul
    L = emul.lib()
    rng = np.random.default_rng(12)
    n_eq = 0
    for _ in range(4000):
        size = 1.0 / 10 ** rng.uniform(-3, 1)
        mu = 10 ** rng.uniform(-3, 2.5)
        y = float(rng.integers(0, 60))
        exact, lb = L.emul_nb_pvalue(y, size, mu), L.emul_nb_pvalue_lb(y, size, mu)
        assert lb <= exact * (1 + 1e-12) + 1e-300, (y, size, mu, lb, exact)
        n_eq += lb == exact
    assert n_eq > 1000
    mp.mp.dps = 40
    for x in np.concatenate([1 + 10 ** rng.uniform(-15, 0, 300), 10 ** rng.uniform(0, 7, 300)]):
        got, want = L.emul_flog_tab(float(x)), mp.log(mp.mpf(float(x)))
        assert abs((mp.mpf(got) - want) / want) < 2 ** -52 * 1.01
    for x in -10 ** rng.uniform(-12, 2.84, 600):
        got, want = L.emul_fexp_bf(float(x)), mp.exp(mp.mpf(float(x)))
        assert abs((mp.mpf(got) - want) / want) < 2 ** -52
    assert L.emul_fexp_bf(-5000.0) == 0.0
