"""Second, independent CPU restatement of glmrob.nb() (bin/glm_rob_nb.r:16-228) — TEST INFRASTRUCTURE ONLY.

Written straight from the R source with scipy / numpy primitives so that it shares no numerics with
oracle/orc_math.c: scipy.stats.nbinom for dnbinom / pnbinom (boost), scipy.special.digamma,
scipy.integrate.quad (QUADPACK dqagse, the routine behind R's integrate()) on a natural cubic
spline with R's linear extrapolation, numpy for quantile / median / BH.  Slow; used on a handful
of units to cross-check the C oracle (SURVEY.md section 8c, item v)."""
import math

import numpy as np
import scipy.integrate
import scipy.interpolate
import scipy.special
import scipy.stats

EPS = np.finfo(float).eps


def dnbinom(x, size, mu):
    return scipy.stats.nbinom.pmf(x, size, size / (size + mu))


def pnbinom_upper(q, size, mu):
    return scipy.stats.nbinom.sf(q, size, size / (size + mu))


def r_round(x):
    return np.rint(x)  # R's round(x) for digits = 0: IEEE round-half-even


def splinefun_natural(x, y):
    cs = scipy.interpolate.CubicSpline(x, y, bc_type="natural")
    slope_last = cs(x[-1], 1)

    def f(u):  # linear beyond the last knot, as R's natural splinefun
        return float(cs(u)) if u <= x[-1] else float(y[-1] + (u - x[-1]) * slope_last)
    return f


def integrate(f, lo, hi):
    tol = EPS ** 0.25
    val, err, info = scipy.integrate.quad(f, lo, hi, epsabs=tol, epsrel=tol, limit=100, full_output=1)[:3]
    return val


def zeroin(f, ax, bx, fa, fb, tol, maxit):
    """R_zeroin2 (Brent), transcribed from the algorithm description in SURVEY.md App. B.4"""
    a, b, c, fc = ax, bx, ax, fa
    if fa == 0:
        return a
    if fb == 0:
        return b
    for _ in range(maxit + 1):
        prev_step = b - a
        if abs(fc) < abs(fb):
            a, b, c = b, c, b
            fa, fb, fc = fb, fc, fb
        tol_act = 2 * EPS * abs(b) + tol / 2
        new_step = (c - b) / 2
        if abs(new_step) <= tol_act or fb == 0:
            return b
        if abs(prev_step) >= tol_act and abs(fa) > abs(fb):
            cb = c - b
            if a == c:
                t1 = fb / fa
                p = cb * t1
                q = 1.0 - t1
            else:
                q = fa / fc
                t1 = fb / fc
                t2 = fb / fa
                p = t2 * (cb * q * (q - t1) - (b - a) * (t1 - 1.0))
                q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0)
            if p > 0:
                q = -q
            else:
                p = -p
            if p < (0.75 * cb * q - abs(tol_act * q) / 2) and p < abs(prev_step * q / 2):
                new_step = p / q
        if abs(new_step) < tol_act:
            new_step = tol_act if new_step > 0 else -tol_act
        a, fa = b, fb
        b += new_step
        fb = f(b)
        if math.isnan(fb) or fb == math.inf:
            raise ArithmeticError("invalid function value in 'zeroin'")
        if (fb > 0 and fc > 0) or (fb < 0 and fc < 0):
            c, fc = a, fa
    return b


def uniroot(f, lo, hi, tol, maxiter):
    flo, fhi = f(lo), f(hi)
    if math.isnan(flo) or math.isnan(fhi):
        raise ArithmeticError("f.lower/f.upper is NA")
    if flo * fhi > 0:
        raise ArithmeticError("f() values at end points not of opposite sign")
    return zeroin(f, lo, hi, flo, fhi, tol, maxiter)


def p_adjust_bh(p):
    p = np.asarray(p, float)
    n = len(p)
    if n <= 1:
        return p.copy()
    o = np.argsort(-p, kind="stable")
    q = np.minimum(1, np.minimum.accumulate(n / np.arange(n, 0, -1) * p[o]))
    out = np.empty(n)
    out[o] = q
    return out


def glmrob_nb(y, x, c_tukey_beta=5, c_tukey_sig=3, minsig=1e-3, maxsig=10, minmu=1e-10, maxit=30, maxit_sig=50,
              sig_prec=1e-8, tol=1e-6, n_ai=100, min_coverage=1, min_reads=1, min_af=0, size_min=10, extra_rob=True,
              min_af_extra_rob=0.2, min_prop_extra_rob=0.1, max_prop_extra_rob=0.5):
    y = np.asarray(y, float)
    x = np.asarray(x, float)
    N = len(x)
    maxmu = x.max()
    with np.errstate(divide="ignore", invalid="ignore"):
        af_all = y / x
    S = np.where(af_all > min_af_extra_rob)[0]
    extra_out = False
    if extra_rob and len(S) > min_prop_extra_rob * N and len(S) < max_prop_extra_rob * N and N - len(S) > size_min:
        extra_out = True
    keep = np.ones(N, bool)
    if extra_out:
        keep[S] = False
        min_reads = 0
        min_af = 0
    xx, yy = x[keep], y[keep]
    with np.errstate(divide="ignore", invalid="ignore"):
        r = yy / xx
    maxr = np.nanmax(r) if np.any(~np.isnan(r)) else -math.inf
    if (np.median(xx) < min_coverage or np.sum(xx > min_coverage) < size_min or yy.max() < min_reads
            or (np.sum(yy > 0) == 0 and not extra_out) or maxr < min_af):
        return dict(sigma=math.nan, slope=math.nan, pvalues=np.ones(N), qvalues=np.ones(N), GQ=np.zeros(N),
                    extra_rob=extra_out)
    m = xx > 0
    yf, xf = yy[m], xx[m]
    n = len(yf)
    X = np.log(xf)
    af = yf / np.exp(X)
    af[yf == 0] = 0
    q25, q75 = np.quantile(af, [0.25, 0.75], method="linear")
    inislope = af[af <= q75 + 1.5 * (q75 - q25)].mean()
    if inislope <= 1 / (10 * np.exp(X).mean()):
        inislope = 1 / (10 * np.exp(X).mean())
    mu = inislope * np.exp(X)
    mu[mu > maxmu] = maxmu
    mu[mu < minmu] = minmu
    eta = np.log(mu)
    sig = float(np.sum((yf / mu - 1) ** 2) / n)
    sig = min(max(sig, minsig), maxsig)

    def varfunc(mu_, s):
        return mu_ + s * mu_ ** 2

    def tukeypsi(r_, c):
        r_ = np.asarray(r_, float)
        return np.where(np.abs(r_) > c, 0.0, ((r_ / c) ** 2 - 1) ** 2 * r_)

    def window(mui, sd, c):
        j1 = max(math.ceil(mui - c * sd), 0)
        j2 = math.floor(mui + c * sd)
        return j1, j2

    def expect(mui, s, c, kind):
        if kind == "ai":
            sd = math.sqrt(mui * (s * mui + 1))
        else:
            sd = math.sqrt(varfunc(mui, s))
        j1, j2 = window(mui, sd, c)
        if j1 > j2:
            return 0.0
        invsig = 1 / s

        def g(j):
            w = (((j - mui) / (c * sd)) ** 2 - 1) ** 2
            pm = dnbinom(j, invsig, mui)
            if kind == "e1":
                return w * (j - mui) * pm
            if kind == "e2":
                return w * (j - mui) ** 2 * pm
            psi = scipy.special.digamma(j + invsig) - scipy.special.digamma(invsig) - np.log(mui / invsig + 1) \
                - (j - mui) / (mui + invsig)
            return w * psi * pm
        if j2 - j1 + 1 > 2 * n_ai:
            j12 = r_round(np.concatenate([[j1], j1 + np.arange(1, n_ai - 1) * ((j2 - j1) / (n_ai - 1)), [j2]]))
            val = integrate(splinefun_natural(j12, g(j12)), j12.min(), j12.max() + 1)
        else:
            j12 = np.arange(j1, j2 + 1, dtype=float)
            val = float(np.sum(g(j12)))
        return val if kind == "ai" else val / sd

    def sig_rob(s):
        r_ = (yf - mu) / np.sqrt(varfunc(mu, s))
        with np.errstate(divide="ignore", invalid="ignore"):
            wi = tukeypsi(r_, c_tukey_sig) / r_
        psiml = (scipy.special.digamma(r_ * np.sqrt(mu * (s * mu + 1)) + mu + 1 / s) - s * r_ * np.sqrt(mu / (s * mu + 1))
                 - scipy.special.digamma(1 / s) - np.log(s * mu + 1))
        ai = np.array([expect(mi, s, c_tukey_sig, "ai") for mi in mu])
        return float(np.sum(wi * psiml - ai))

    def D(u, v):  # abs(max(u - v)) with R recycling
        u, v = np.atleast_1d(u).astype(float), np.atleast_1d(v).astype(float)
        L = max(len(u), len(v))
        d = np.resize(u, L) - np.resize(v, L)
        return abs(np.max(d))

    sig0 = sig + 1 + tol
    beta11 = np.array([math.log(inislope)])
    beta00 = beta11 + tol + 1
    beta1 = beta11
    it = 0
    while (abs(sig - sig0) > tol or D(beta11, beta00) > tol) and it < maxit:
        sig0 = sig
        beta00 = beta11
        try:
            sig = uniroot(sig_rob, minsig, maxsig, sig_prec, maxit_sig)
            sig = min(max(sig, minsig), maxsig)
        except ArithmeticError:
            sig = minsig
        beta1 = np.array([math.log(inislope)])
        beta0 = beta1 + tol + 1
        it_mu = 0
        while D(beta1, beta0) > tol and it_mu < maxit:
            beta0 = beta1
            sap = np.array([expect(mi, sig, c_tukey_beta, "e2") for mi in mu])
            e1 = np.array([expect(mi, sig, c_tukey_beta, "e1") for mi in mu])
            V = varfunc(mu, sig)
            bi = sap * V ** (-1.5) * np.exp(eta) ** 2
            ei = (tukeypsi((yf - mu) / np.sqrt(V), c_tukey_beta) - e1) / sap * V * (1 / mu)
            zi = eta + ei
            ok = (bi != 0) & ~np.isnan(bi) & ~np.isnan(zi)
            beta = float(np.sum(bi[ok] * (zi[ok] - X[ok])) / np.sum(bi[ok]))
            beta1 = np.array([beta, 1.0])
            eta = X + beta
            mu = np.exp(eta)
            mu[mu > maxmu] = maxmu
            mu[mu == 0] = minmu
            eta = np.log(mu)
            it_mu += 1
        beta11 = beta1
        it += 1
    slope = min(math.exp(beta1[0]), 1.0)
    p = dnbinom(y, 1 / sig, slope * x) + pnbinom_upper(y, 1 / sig, slope * x)
    p = np.where(x == 0, 1.0, p)
    q = p_adjust_bh(p)
    with np.errstate(divide="ignore"):
        gq = np.minimum(-10 * np.log10(q), 1000)
    return dict(sigma=sig, slope=slope, pvalues=p, qvalues=q, GQ=gq, extra_rob=extra_out, n_outer=it)
