/*
 * oracle/orc_math.c — TEST INFRASTRUCTURE ONLY (CPU oracle, numeric layer).
 * See orc_math.h for the scope statement.  PARITY UNPINNED (no R available).
 *
 * Each routine names the R builtin it restates and the reference call site
 * (file:line under /root/reference) that makes it part of the hot path.
 */
#include "orc_math.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_LN_SQRT_2PI 0.918938533204672741780329736406
#define ORC_LN_2PI 1.837877066409345483560659472811

/* ------------------------------------------------------------------------
 * Loader's saddle-point pieces (R nmath stirlerr.c / bd0.c / dbinom.c).
 * Used by dnbinom(): glm_rob_nb.r:117,121,133,137,153,157,221.
 * ---------------------------------------------------------------------- */
static const double sferr_halves[31] = {
    0.0,
    0.1534264097200273452913839, 0.08106146679532725821967026,
    0.0548141210519176538961387, 0.04134069595540929409382208,
    0.03316287351993628748511051, 0.02767792568499833914878929,
    0.02374616365629749597133028, 0.02079067210376509311152277,
    0.01848845053267318523077936, 0.01664469118982119216319487,
    0.01513497322191737887351384, 0.01387612882307074799874573,
    0.01281046524292022692425066, 0.01189670994589177009505572,
    0.01110455975820691732663076, 0.01041126526197209649747857,
    0.009799416126158803298390373, 0.009255462182712732917728637,
    0.008768700134139385462955047, 0.008330563433362871256469319,
    0.007934114564314020547249562, 0.007573675487951840794972024,
    0.007244554301320383179546197, 0.006942840107209529865664153,
    0.006665247032707682442356181, 0.006408994188004207068439631,
    0.006171712263039457647534605, 0.005951370112758847735624416,
    0.005746216513010115682026102, 0.00555473355196280137103869};

double orc_stirlerr(double n)
{
    const double S0 = 1.0 / 12.0, S1 = 1.0 / 360.0, S2 = 1.0 / 1260.0, S3 = 1.0 / 1680.0,
                 S4 = 1.0 / 1188.0;
    double nn;
    if (n <= 15.0) {
        nn = n + n;
        if (nn == (double)(int)nn)
            return sferr_halves[(int)nn];
        return lgamma(n + 1.0) - (n + 0.5) * log(n) + n - ORC_LN_SQRT_2PI;
    }
    nn = n * n;
    if (n > 500)
        return (S0 - S1 / nn) / n;
    if (n > 80)
        return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35)
        return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

double orc_bd0(double x, double np)
{
    if (!isfinite(x) || !isfinite(np) || np == 0.0)
        return NAN;
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        if (fabs(s) < DBL_MIN)
            return s;
        double ej = 2 * x * v;
        v = v * v;
        for (int j = 1; j < 1000; j++) {
            ej *= v;
            double s1 = s + ej / ((j << 1) + 1);
            if (s1 == s)
                return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

/* log of dbinom_raw(x, n, p, q) */
static double dbinom_raw_log(double x, double n, double p, double q)
{
    double lc, lf;
    if (p == 0)
        return (x == 0) ? 0.0 : -INFINITY;
    if (q == 0)
        return (x == n) ? 0.0 : -INFINITY;
    if (x == 0) {
        if (n == 0)
            return 0.0;
        lc = (p < 0.1) ? -orc_bd0(n, n * q) - n * p : n * log(q);
        return lc;
    }
    if (x == n) {
        lc = (q < 0.1) ? -orc_bd0(n, n * p) - n * q : n * log(p);
        return lc;
    }
    if (x < 0 || x > n)
        return -INFINITY;
    lc = orc_stirlerr(n) - orc_stirlerr(x) - orc_stirlerr(n - x) - orc_bd0(x, n * p) -
         orc_bd0(n - x, n * q);
    lf = ORC_LN_2PI + log(x) + log1p(-x / n);
    return lc - 0.5 * lf;
}

double orc_dbinom_raw(double x, double n, double p, double q)
{
    return exp(dbinom_raw_log(x, n, p, q));
}

/* dnbinom_mu (R nmath dnbinom.c) */
double orc_dnbinom_mu(double x, double size, double mu)
{
    if (isnan(x) || isnan(size) || isnan(mu))
        return x + size + mu;
    if (mu < 0 || size < 0)
        return NAN;
    if (fabs(x - nearbyint(x)) > 1e-7 * fmax(1.0, fabs(x)))
        return 0.0; /* R: warning + 0 */
    if (x < 0 || !isfinite(x))
        return 0.0;
    if (x == 0 && size == 0)
        return 1.0;
    x = nearbyint(x);
    if (!isfinite(size)) /* Poisson limit: never reached (size <= 1000) */
        return exp(-orc_stirlerr(x) - orc_bd0(x, mu)) / sqrt(2 * M_PI * x);
    if (x == 0)
        return exp(size * (size < mu ? log(size / (size + mu)) : log1p(-mu / (size + mu))));
    if (x < 1e-10 * size) {
        double p = (size < mu ? log(size / (1 + size / mu)) : log(mu / (1 + mu / size)));
        return exp(x * p - mu - lgamma(x + 1) + log1p(x * (x - 1) / (2 * size)));
    }
    double p = size / (size + x);
    double ans = orc_dbinom_raw(size, x + size, size / (size + mu), mu / (size + mu));
    return p * ans;
}

/* ------------------------------------------------------------------------
 * Regularised incomplete beta.  R uses TOMS 708 bratio(); its source is not
 * available here, so this restates the *function* (I_x(a,b)) with a
 * continued fraction (modified Lentz) and a saddle-point prefactor that
 * keeps full relative accuracy in the far tails:
 *     x^a y^b / (a B(a,b)) = dbinom_raw(a, a+b, x, y) * b/(a+b).
 * Known answers: mpmath.betainc (tests/test_oracle_math.py).
 * ---------------------------------------------------------------------- */
static double dbinom_ab(double a, double b, double p, double q)
{
    /* dbinom_raw(a, a+b, p, q) with n-x passed explicitly (no cancellation) */
    double n = a + b;
    double lc = orc_stirlerr(n) - orc_stirlerr(a) - orc_stirlerr(b) - orc_bd0(a, n * p) -
                orc_bd0(b, n * q);
    double lf = ORC_LN_2PI + log(a) + log(b / n);
    return exp(lc - 0.5 * lf);
}

static double betacf(double a, double b, double x)
{
    const double FPMIN = 1e-300, EPS = 1e-16;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < FPMIN)
        d = FPMIN;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 200000; m++) {
        double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < FPMIN)
            d = FPMIN;
        c = 1.0 + aa / c;
        if (fabs(c) < FPMIN)
            c = FPMIN;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < FPMIN)
            d = FPMIN;
        c = 1.0 + aa / c;
        if (fabs(c) < FPMIN)
            c = FPMIN;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < EPS)
            break;
    }
    return h;
}

double orc_pbeta_raw(double x, double y, double a, double b, int lower)
{
    if (x <= 0)
        return lower ? 0.0 : 1.0;
    if (y <= 0)
        return lower ? 1.0 : 0.0;
    if (x < (a + 1.0) / (a + b + 2.0)) {
        double I = dbinom_ab(a, b, x, y) * (b / (a + b)) * betacf(a, b, x);
        return lower ? I : 1.0 - I;
    } else {
        double J = dbinom_ab(b, a, y, x) * (a / (a + b)) * betacf(b, a, y);
        return lower ? 1.0 - J : J;
    }
}

double orc_pnbinom_mu_upper(double q, double size, double mu)
{
    if (isnan(q) || isnan(size) || isnan(mu))
        return q + size + mu;
    if (q < 0)
        return 1.0;
    if (!isfinite(q))
        return 0.0;
    q = floor(q + 1e-7);
    /* 1 - I_pr(size, q+1) = I_{1-pr}(q+1, size) */
    return orc_pbeta_raw(mu / (size + mu), size / (size + mu), q + 1.0, size, 1);
}

double orc_pnbinom_mu_lower(double q, double size, double mu)
{
    if (q < 0)
        return 0.0;
    if (!isfinite(q))
        return 1.0;
    q = floor(q + 1e-7);
    return orc_pbeta_raw(mu / (size + mu), size / (size + mu), q + 1.0, size, 0);
}

double orc_pbinom_lower(double q, double n, double p)
{
    if (q < 0)
        return 0.0;
    q = floor(q + 1e-7);
    if (n <= q)
        return 1.0;
    /* pbeta(p, q+1, n-q, lower=FALSE) */
    return orc_pbeta_raw(p, 1.0 - p, q + 1.0, n - q, 0);
}

/* digamma (R: psigamma(x,0) via Amos dpsifn).  Restated as upward recurrence
 * to x>=12 followed by the asymptotic series; |err| < 2e-16*max(1,|psi|). */
double orc_digamma(double x)
{
    if (isnan(x))
        return x;
    if (x <= 0)
        return NAN; /* never reached: arguments are >= 1/maxsig = 0.1 */
    double acc = 0.0;
    while (x < 12.0) {
        acc -= 1.0 / x;
        x += 1.0;
    }
    double r = 1.0 / x, r2 = r * r;
    double s = r2 * (1.0 / 12 -
                     r2 * (1.0 / 120 -
                           r2 * (1.0 / 252 -
                                 r2 * (1.0 / 240 -
                                       r2 * (1.0 / 132 -
                                             r2 * (691.0 / 32760 - r2 * (1.0 / 12)))))));
    return acc + log(x) - 0.5 * r - s;
}

/* ------------------------------------------------------------------------
 * Discrete quantiles (R nmath qnbinom.c / qbinom.c, pre-4.2 search):
 * Cornish-Fisher start, then unit search for the smallest k with
 * CDF(k) >= p*(1-64eps).  needlestack.r:243 (qnbinom(0.01,...)), :249 (qbinom).
 * ---------------------------------------------------------------------- */
static double qnorm_lower(double p)
{
    /* Acklam's rational approximation refined by one Halley step: only the
     * START of the search depends on it, never the returned quantile. */
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02,
                               -2.759285104469687e+02, 1.383577518672690e+02,
                               -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02,
                               -1.556989798598866e+02, 6.680131188771972e+01,
                               -1.328068155288572e+01};
    static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01,
                               -2.400758277161838e+00, -2.549732539343734e+00,
                               4.374664141464968e+00,  2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01,
                               2.445134137142996e+00, 3.754408661907416e+00};
    double q, r, x;
    if (p < 0.02425) {
        q = sqrt(-2 * log(p));
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    } else if (p <= 1 - 0.02425) {
        q = p - 0.5;
        r = q * q;
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
    } else {
        q = sqrt(-2 * log(1 - p));
        x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
    double e = 0.5 * erfc(-x / M_SQRT2) - p;
    double u = e * sqrt(2 * M_PI) * exp(x * x / 2);
    x = x - u / (1 + x * u / 2);
    return x;
}

typedef double (*cdf_fn)(double k, double p1, double p2);
static double cdf_nb(double k, double size, double mu) { return orc_pnbinom_mu_lower(k, size, mu); }
static double cdf_bin(double k, double n, double pr) { return orc_pbinom_lower(k, n, pr); }

static double discrete_search(double y, double p, cdf_fn cdf, double p1, double p2, double ymax)
{
    double z = cdf(y, p1, p2);
    p *= 1 - 64 * DBL_EPSILON;
    if (z >= p) { /* search to the left */
        for (;;) {
            if (y == 0 || cdf(y - 1, p1, p2) < p)
                return y;
            y = fmax(0, y - 1);
        }
    } else { /* search to the right */
        for (;;) {
            y = y + 1;
            if (y >= ymax || cdf(y, p1, p2) >= p)
                return y;
        }
    }
}

double orc_qnbinom_mu(double p, double size, double mu)
{
    if (isnan(p) || isnan(size) || isnan(mu))
        return p + size + mu;
    double prob = size / (size + mu);
    if (prob == 1 || size == 0)
        return 0;
    if (p <= 0)
        return 0;
    if (p >= 1)
        return INFINITY;
    double Q = 1.0 / prob, P = (1.0 - prob) * Q;
    double m = size * P, sigma = sqrt(size * P * Q), gamma = (Q + P) / sigma;
    double z = qnorm_lower(p);
    double y = nearbyint(m + sigma * (z + gamma * (z * z - 1) / 6));
    if (y < 0)
        y = 0;
    return discrete_search(y, p, cdf_nb, size, mu, INFINITY);
}

double orc_qbinom(double p, double n, double pr)
{
    if (isnan(p) || isnan(n) || isnan(pr))
        return p + n + pr;
    if (pr == 0 || n == 0)
        return 0;
    if (p <= 0)
        return 0;
    if (p >= 1)
        return n;
    double q = 1 - pr;
    if (q == 0)
        return n;
    double m = n * pr, sigma = sqrt(n * pr * q), gamma = (q - pr) / sigma;
    double z = qnorm_lower(p);
    double y = nearbyint(m + sigma * (z + gamma * (z * z - 1) / 6));
    if (y < 0)
        y = 0;
    if (y > n)
        y = n;
    return discrete_search(y, p, cdf_bin, n, pr, n);
}

/* ------------------------------------------------------------------------
 * Hypergeometric / Fisher exact 2x2 (R: dhyper.c, fisher.test).
 * needlestack.r:234,236 — matrix(c(Vp,Vm,Rp,Rm),nrow=2): column 1 = alt.
 * ---------------------------------------------------------------------- */
double orc_dhyper_log(double x, double r, double b, double n)
{
    if (x < 0)
        return -INFINITY;
    if (n < x || r < x || n - x > b)
        return -INFINITY;
    if (n == 0)
        return (x == 0) ? 0.0 : -INFINITY;
    double p = n / (r + b), q = (r + b - n) / (r + b);
    double p1 = dbinom_raw_log(x, r, p, q);
    double p2 = dbinom_raw_log(n - x, b, p, q);
    double p3 = dbinom_raw_log(n, r + b, p, q);
    return p1 + p2 - p3;
}

double orc_fisher_2x2(double a11, double a21, double a12, double a22)
{
    double m = a11 + a21, n = a12 + a22, k = a11 + a12, x = a11;
    double lo = fmax(0, k - n), hi = fmin(k, m);
    int ns = (int)(hi - lo) + 1;
    double *d = (double *)malloc(sizeof(double) * (size_t)ns);
    double mx = -INFINITY;
    for (int i = 0; i < ns; i++) {
        d[i] = orc_dhyper_log(lo + i, m, n, k);
        if (d[i] > mx)
            mx = d[i];
    }
    double sum = 0;
    for (int i = 0; i < ns; i++) {
        d[i] = exp(d[i] - mx);
        sum += d[i];
    }
    for (int i = 0; i < ns; i++)
        d[i] /= sum;
    const double relErr = 1 + 1e-7;
    double dobs = d[(int)(x - lo)] * relErr, pv = 0;
    for (int i = 0; i < ns; i++)
        if (d[i] <= dobs)
            pv += d[i];
    free(d);
    return pv;
}

/* ------------------------------------------------------------------------
 * Brent's zeroin as R_zeroin2 + the R-level uniroot() wrapper.
 * glm_rob_nb.r:175-178.
 * ---------------------------------------------------------------------- */
double orc_zeroin2(double ax, double bx, double fa, double fb, orc_fn f, void *info, double *Tol,
                   int *Maxit)
{
    double a = ax, b = bx, c = a, fc = fa;
    int maxit = *Maxit + 1;
    double tol = *Tol;
    if (fa == 0.0) {
        *Tol = 0.0;
        *Maxit = 0;
        return a;
    }
    if (fb == 0.0) {
        *Tol = 0.0;
        *Maxit = 0;
        return b;
    }
    while (maxit--) {
        double prev_step = b - a;
        double tol_act, p, q, new_step;
        if (fabs(fc) < fabs(fb)) {
            a = b;
            b = c;
            c = a;
            fa = fb;
            fb = fc;
            fc = fa;
        }
        tol_act = 2 * DBL_EPSILON * fabs(b) + tol / 2;
        new_step = (c - b) / 2;
        if (fabs(new_step) <= tol_act || fb == 0.0) {
            *Maxit -= maxit;
            *Tol = fabs(c - b);
            return b;
        }
        if (fabs(prev_step) >= tol_act && fabs(fa) > fabs(fb)) {
            double t1, cb, t2;
            cb = c - b;
            if (a == c) {
                t1 = fb / fa;
                p = cb * t1;
                q = 1.0 - t1;
            } else {
                q = fa / fc;
                t1 = fb / fc;
                t2 = fb / fa;
                p = t2 * (cb * q * (q - t1) - (b - a) * (t1 - 1.0));
                q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0);
            }
            if (p > 0.0)
                q = -q;
            else
                p = -p;
            if (p < (0.75 * cb * q - fabs(tol_act * q) / 2) && p < fabs(prev_step * q / 2))
                new_step = p / q;
        }
        if (fabs(new_step) < tol_act) {
            if (new_step > 0.0)
                new_step = tol_act;
            else
                new_step = -tol_act;
        }
        a = b;
        fa = fb;
        b += new_step;
        fb = f(b, info);
        if (isnan(fb) || fb == INFINITY) { /* R: error("invalid function value in 'zeroin'") */
            *Tol = -2.0;
            *Maxit = -2;
            return NAN;
        }
        if (fb == -INFINITY)
            fb = -DBL_MAX; /* R: warning "-Inf replaced by maximally negative value" */
        if ((fb > 0 && fc > 0) || (fb < 0 && fc < 0)) {
            c = a;
            fc = fa;
        }
    }
    *Tol = -1.0;
    *Maxit = -1;
    return b; /* R: warning "_NOT_ converged", root still returned */
}

int orc_uniroot(orc_fn f, void *info, double lo, double hi, double tol, int maxiter, double *root,
                int *n_eval)
{
    double flo = f(lo, info), fhi = f(hi, info);
    if (n_eval)
        *n_eval = 2;
    if (isnan(flo))
        return 1; /* "f.lower = f(lower) is NA" */
    if (isnan(fhi))
        return 2;
    if (flo * fhi > 0)
        return 3; /* "f() values at end points not of opposite sign" */
    /* zeroin2's R callback is not applied to the end points; +-Inf there goes through */
    double t = tol;
    int mi = maxiter;
    double r = orc_zeroin2(lo, hi, flo, fhi, f, info, &t, &mi);
    if (mi == -2)
        return 4; /* invalid function value */
    if (n_eval)
        *n_eval += (mi < 0 ? maxiter + 1 : mi);
    *root = r;
    return 0;
}

/* ------------------------------------------------------------------------
 * Natural cubic spline (R splines.c natural_spline / spline_eval), as used by
 * splinefun(method="natural"): glm_rob_nb.r:118,134,154.
 * ---------------------------------------------------------------------- */
void orc_natural_spline(int n, const double *x0, const double *y0, double *b0, double *c0,
                        double *d0)
{
    if (n < 2)
        return;
    const double *x = x0 - 1, *y = y0 - 1;
    double *b = b0 - 1, *c = c0 - 1, *d = d0 - 1, t;
    int i;
    if (n < 3) {
        t = (y[2] - y[1]);
        b[1] = t / (x[2] - x[1]);
        b[2] = b[1];
        c[1] = c[2] = d[1] = d[2] = 0.0;
        return;
    }
    const int nm1 = n - 1;
    d[1] = x[2] - x[1];
    c[2] = (y[2] - y[1]) / d[1];
    for (i = 2; i < n; i++) {
        d[i] = x[i + 1] - x[i];
        b[i] = 2 * (d[i - 1] + d[i]);
        c[i + 1] = (y[i + 1] - y[i]) / d[i];
        c[i] = c[i + 1] - c[i];
    }
    for (i = 3; i < n; i++) {
        t = d[i - 1] / b[i - 1];
        b[i] = b[i] - t * d[i - 1];
        c[i] = c[i] - t * c[i - 1];
    }
    c[nm1] = c[nm1] / b[nm1];
    for (i = n - 2; i > 1; i--)
        c[i] = (c[i] - d[i] * c[i + 1]) / b[i];
    c[1] = c[n] = 0.0;
    b[1] = (y[2] - y[1]) / d[1] - d[i] * c[2]; /* i == 1 here */
    c[1] = 0.0;
    d[1] = c[2] / d[1];
    b[n] = (y[n] - y[nm1]) / d[nm1] + d[nm1] * c[nm1];
    for (i = 2; i < n; i++) {
        b[i] = (y[i + 1] - y[i]) / d[i] - d[i] * (c[i + 1] + 2.0 * c[i]);
        d[i] = (c[i + 1] - c[i]) / d[i];
        c[i] = 3.0 * c[i];
    }
    c[n] = 0.0;
    d[n] = 0.0;
}

double orc_spline_eval(int n, const double *x, const double *y, const double *b, const double *c,
                       const double *d, double u)
{
    int i = 0, j = n;
    do { /* x[i] <= u <= x[i+1] (bisection exactly as spline_eval) */
        int k = (i + j) / 2;
        if (u < x[k])
            j = k;
        else
            i = k;
    } while (j > i + 1);
    double dx = u - x[i];
    double tmp = (u < x[0]) ? 0.0 : d[i]; /* natural: linear to the left */
    return y[i] + dx * (b[i] + dx * (c[i] + dx * tmp));
}

/* ------------------------------------------------------------------------
 * QUADPACK dqagse + dqk21 + dqpsrt + dqelg (netlib, public domain), i.e. what
 * R's integrate() runs: glm_rob_nb.r:118,134,154.
 * ---------------------------------------------------------------------- */
static const double xgk21[11] = {
    0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
    0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
    0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
    0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
    0.294392862701460198131126603103866, 0.148874338981631210884826001129720,
    0.0};
static const double wgk21[11] = {
    0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
    0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
    0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
    0.123491976262065851077958109831074, 0.134709217311473325928054001771707,
    0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
    0.149445554002916905664936468389821};
static const double wg10[5] = {
    0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
    0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
    0.295524224714752870173892994651338};

static void dqk21(orc_fn f, void *info, double a, double b, double *result, double *abserr,
                  double *resabs, double *resasc)
{
    double fv1[10], fv2[10];
    const double epmach = DBL_EPSILON, uflow = DBL_MIN;
    double centr = (a + b) * .5, hlgth = (b - a) * .5, dhlgth = fabs(hlgth);
    double resg = 0.0;
    double fc = f(centr, info);
    double resk = wgk21[10] * fc;
    *resabs = fabs(resk);
    for (int j = 0; j < 5; j++) {
        int jtw = 2 * j + 1;
        double absc = hlgth * xgk21[jtw];
        double fval1 = f(centr - absc, info), fval2 = f(centr + absc, info);
        fv1[jtw] = fval1;
        fv2[jtw] = fval2;
        double fsum = fval1 + fval2;
        resg += wg10[j] * fsum;
        resk += wgk21[jtw] * fsum;
        *resabs += wgk21[jtw] * (fabs(fval1) + fabs(fval2));
    }
    for (int j = 0; j < 5; j++) {
        int jtwm1 = 2 * j;
        double absc = hlgth * xgk21[jtwm1];
        double fval1 = f(centr - absc, info), fval2 = f(centr + absc, info);
        fv1[jtwm1] = fval1;
        fv2[jtwm1] = fval2;
        double fsum = fval1 + fval2;
        resk += wgk21[jtwm1] * fsum;
        *resabs += wgk21[jtwm1] * (fabs(fval1) + fabs(fval2));
    }
    double reskh = resk * .5;
    *resasc = wgk21[10] * fabs(fc - reskh);
    for (int j = 0; j < 10; j++)
        *resasc += wgk21[j] * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
    *result = resk * hlgth;
    *resabs *= dhlgth;
    *resasc *= dhlgth;
    *abserr = fabs((resk - resg) * hlgth);
    if (*resasc != 0.0 && *abserr != 0.0)
        *abserr = *resasc * fmin(1.0, pow(*abserr * 200.0 / *resasc, 1.5));
    if (*resabs > uflow / (epmach * 50.0))
        *abserr = fmax(epmach * 50.0 * *resabs, *abserr);
}

/* 1-based arrays inside, as in the Fortran */
static void dqpsrt(int limit, int last, int *maxerr, double *ermax, double *elist, int *iord,
                   int *nrmax)
{
    int i, j, k, ido, jbnd, isucc, jupbn, ibeg;
    double errmin, errmax;
    if (last <= 2) {
        iord[1] = 1;
        iord[2] = 2;
        goto Last;
    }
    errmax = elist[*maxerr];
    if (*nrmax > 1) {
        ido = *nrmax - 1;
        for (i = 1; i <= ido; ++i) {
            isucc = iord[*nrmax - 1];
            if (errmax <= elist[isucc])
                break;
            iord[*nrmax] = isucc;
            --(*nrmax);
        }
    }
    if (last > limit / 2 + 2)
        jupbn = limit + 3 - last;
    else
        jupbn = last;
    errmin = elist[last];
    jbnd = jupbn - 1;
    ibeg = *nrmax + 1;
    if (ibeg <= jbnd) {
        for (i = ibeg; i <= jbnd; ++i) {
            isucc = iord[i];
            if (errmax >= elist[isucc]) {
                iord[i - 1] = *maxerr;
                for (j = i, k = jbnd; j <= jbnd; j++, k--) {
                    isucc = iord[k];
                    if (errmin < elist[isucc]) {
                        iord[k + 1] = last;
                        goto Last;
                    }
                    iord[k + 1] = isucc;
                }
                iord[i] = last;
                goto Last;
            }
            iord[i - 1] = isucc;
        }
    }
    iord[jbnd] = *maxerr;
    iord[jupbn] = last;
Last:
    *maxerr = iord[*nrmax];
    *ermax = elist[*maxerr];
}

/* epstab, res3la 1-based */
static void dqelg(int *n, double *epstab, double *result, double *abserr, double *res3la, int *nres)
{
    int i, indx, ib, ib2, ie, k1, k2, k3, num, newelm, limexp;
    double delta1, delta2, delta3, e0, e1, e1abs, e2, e3, epsinf, ss = 0, res;
    double errA, err1, err2, err3, tol1, tol2, tol3;
    const double epmach = DBL_EPSILON, oflow = DBL_MAX;
    ++(*nres);
    *abserr = oflow;
    *result = epstab[*n];
    if (*n < 3)
        goto L100;
    limexp = 50;
    epstab[*n + 2] = epstab[*n];
    newelm = (*n - 1) / 2;
    epstab[*n] = oflow;
    num = *n;
    k1 = *n;
    for (i = 1; i <= newelm; ++i) {
        k2 = k1 - 1;
        k3 = k1 - 2;
        res = epstab[k1 + 2];
        e0 = epstab[k3];
        e1 = epstab[k2];
        e2 = res;
        e1abs = fabs(e1);
        delta2 = e2 - e1;
        err2 = fabs(delta2);
        tol2 = fmax(fabs(e2), e1abs) * epmach;
        delta3 = e1 - e0;
        err3 = fabs(delta3);
        tol3 = fmax(e1abs, fabs(e0)) * epmach;
        if (err2 <= tol2 && err3 <= tol3) {
            *result = res;
            *abserr = err2 + err3;
            goto L100;
        }
        e3 = epstab[k1];
        epstab[k1] = e1;
        delta1 = e1 - e3;
        err1 = fabs(delta1);
        tol1 = fmax(e1abs, fabs(e3)) * epmach;
        if (err1 > tol1 && err2 > tol2 && err3 > tol3) {
            ss = 1. / delta1 + 1. / delta2 - 1. / delta3;
            epsinf = fabs(ss * e1);
            if (epsinf > 1e-4)
                goto L30;
        }
        *n = i + i - 1;
        goto L50;
    L30:
        res = e1 + 1. / ss;
        epstab[k1] = res;
        k1 += -2;
        errA = err2 + fabs(res - e2) + err3;
        if (errA <= *abserr) {
            *abserr = errA;
            *result = res;
        }
    }
L50:
    if (*n == limexp)
        *n = (limexp / 2 << 1) - 1;
    if (num / 2 << 1 == num)
        ib = 2;
    else
        ib = 1;
    ie = newelm + 1;
    for (i = 1; i <= ie; ++i) {
        ib2 = ib + 2;
        epstab[ib] = epstab[ib2];
        ib = ib2;
    }
    if (num != *n) {
        indx = num - *n + 1;
        for (i = 1; i <= *n; ++i) {
            epstab[i] = epstab[indx];
            ++indx;
        }
    }
    if (*nres >= 4) {
        *abserr = fabs(*result - res3la[3]) + fabs(*result - res3la[2]) + fabs(*result - res3la[1]);
        res3la[1] = res3la[2];
        res3la[2] = res3la[3];
        res3la[3] = *result;
    } else {
        res3la[*nres] = *result;
        *abserr = oflow;
    }
L100:
    *abserr = fmax(*abserr, epmach * 5. * fabs(*result));
}

static __thread double g_alist[ORC_QLIMIT + 2], g_blist[ORC_QLIMIT + 2], g_rlist[ORC_QLIMIT + 2],
    g_elist[ORC_QLIMIT + 2];
static __thread int g_iord[ORC_QLIMIT + 2], g_last;

void orc_dqagse(orc_fn f, void *info, double a, double b, double epsabs, double epsrel, int limit,
                orc_quad_out *out)
{
    double *alist = g_alist, *blist = g_blist, *rlist = g_rlist, *elist = g_elist;
    int *iord = g_iord;
    double rlist2[53], res3la[4];
    const double epmach = DBL_EPSILON, uflow = DBL_MIN, oflow = DBL_MAX;
    double result = 0, abserr = 0, defabs, resabs, dres, errbnd;
    double area = 0, area1, area2, area12, a1, a2, b1, b2, error1, error2, erro12, errmax = 0,
           errsum = 0, erlast, erlarg = 0, ertest = 0, defab1, defab2, correc = 0, small = 0,
           reseps, abseps;
    int ier = 0, ierro = 0, last = 0, maxerr = 1, nrmax = 1, nres = 0, numrl2 = 2, ktmin = 0,
        iroff1 = 0, iroff2 = 0, iroff3 = 0, ksgn, extrap = 0, noext = 0, k, id, jupbnd;
    if (limit > ORC_QLIMIT)
        limit = ORC_QLIMIT;
    alist[1] = a;
    blist[1] = b;
    rlist[1] = 0;
    elist[1] = 0;
    if (epsabs <= 0 && epsrel < fmax(epmach * 50., 5e-29)) {
        ier = 6;
        goto L999;
    }
    dqk21(f, info, a, b, &result, &abserr, &defabs, &resabs);
    dres = fabs(result);
    errbnd = fmax(epsabs, epsrel * dres);
    last = 1;
    rlist[1] = result;
    elist[1] = abserr;
    iord[1] = 1;
    if (abserr <= epmach * 100. * defabs && abserr > errbnd)
        ier = 2;
    if (limit == 1)
        ier = 1;
    if (ier != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.)
        goto L140;

    rlist2[1] = result;
    errmax = abserr;
    maxerr = 1;
    area = result;
    errsum = abserr;
    abserr = oflow;
    ksgn = -1;
    if (dres >= (1. - epmach * 50.) * defabs)
        ksgn = 1;

    for (last = 2; last <= limit; ++last) {
        a1 = alist[maxerr];
        b1 = (alist[maxerr] + blist[maxerr]) * .5;
        a2 = b1;
        b2 = blist[maxerr];
        erlast = errmax;
        dqk21(f, info, a1, b1, &area1, &error1, &resabs, &defab1);
        dqk21(f, info, a2, b2, &area2, &error2, &resabs, &defab2);
        area12 = area1 + area2;
        erro12 = error1 + error2;
        errsum = errsum + erro12 - errmax;
        area = area + area12 - rlist[maxerr];
        if (!(defab1 == error1 || defab2 == error2)) {
            if (fabs(rlist[maxerr] - area12) <= fabs(area12) * 1e-5 && erro12 >= errmax * .99) {
                if (extrap)
                    ++iroff2;
                else
                    ++iroff1;
            }
            if (last > 10 && erro12 > errmax)
                ++iroff3;
        }
        rlist[maxerr] = area1;
        rlist[last] = area2;
        errbnd = fmax(epsabs, epsrel * fabs(area));
        if (iroff1 + iroff2 >= 10 || iroff3 >= 20)
            ier = 2;
        if (iroff2 >= 5)
            ierro = 3;
        if (last == limit)
            ier = 1;
        if (fmax(fabs(a1), fabs(b2)) <= (epmach * 100. + 1.) * (fabs(a2) + uflow * 1e3))
            ier = 4;
        if (error2 > error1) {
            alist[maxerr] = a2;
            alist[last] = a1;
            blist[last] = b1;
            rlist[maxerr] = area2;
            rlist[last] = area1;
            elist[maxerr] = error2;
            elist[last] = error1;
        } else {
            alist[last] = a2;
            blist[maxerr] = b1;
            blist[last] = b2;
            elist[maxerr] = error1;
            elist[last] = error2;
        }
        dqpsrt(limit, last, &maxerr, &errmax, elist, iord, &nrmax);
        if (errsum <= errbnd)
            goto L115;
        if (ier != 0)
            break;
        if (last == 2) {
            small = fabs(b - a) * .375;
            erlarg = errsum;
            ertest = errbnd;
            rlist2[2] = area;
            continue;
        }
        if (noext)
            continue;
        erlarg -= erlast;
        if (fabs(b1 - a1) > small)
            erlarg += erro12;
        if (!extrap) {
            if (fabs(blist[maxerr] - alist[maxerr]) > small)
                continue;
            extrap = 1;
            nrmax = 2;
        }
        if (ierro != 3 && erlarg > ertest) {
            id = nrmax;
            jupbnd = last;
            if (last > limit / 2 + 2)
                jupbnd = limit + 3 - last;
            int found = 0;
            for (k = id; k <= jupbnd; ++k) {
                maxerr = iord[nrmax];
                errmax = elist[maxerr];
                if (fabs(blist[maxerr] - alist[maxerr]) > small) {
                    found = 1;
                    break;
                }
                ++nrmax;
            }
            if (found)
                continue; /* L90 */
        }
        ++numrl2;
        rlist2[numrl2] = area;
        dqelg(&numrl2, rlist2, &reseps, &abseps, res3la, &nres);
        ++ktmin;
        if (ktmin > 5 && abserr < errsum * .001)
            ier = 5;
        if (abseps < abserr) {
            ktmin = 0;
            abserr = abseps;
            result = reseps;
            correc = erlarg;
            ertest = fmax(epsabs, epsrel * fabs(reseps));
            if (abserr <= ertest)
                break;
        }
        if (numrl2 == 1)
            noext = 1;
        if (ier == 5)
            break;
        maxerr = iord[1];
        errmax = elist[maxerr];
        nrmax = 1;
        extrap = 0;
        small *= .5;
        erlarg = errsum;
    }
    if (last > limit)
        last = limit;
    /* L100 */
    if (abserr == oflow)
        goto L115;
    if (ier + ierro == 0)
        goto L110;
    if (ierro == 3)
        abserr += correc;
    if (ier == 0)
        ier = 3;
    if (result == 0. || area == 0.) {
        if (abserr > errsum)
            goto L115;
        if (area == 0.)
            goto L130;
        goto L110;
    }
    if (abserr / fabs(result) > errsum / fabs(area))
        goto L115;
L110:
    if (ksgn == -1 && fmax(fabs(result), fabs(area)) <= defabs * .01)
        goto L130;
    if (.01 > result / area || result / area > 100. || errsum > fabs(area))
        ier = 6;
    goto L130;
L115:
    result = 0.;
    for (k = 1; k <= last; ++k)
        result += rlist[k];
    abserr = errsum;
L130:
    if (ier > 2)
        --ier;
L140:
L999:
    g_last = last;
    out->result = result;
    out->abserr = abserr;
    out->neval = last * 42 - 21;
    out->ier = ier;
    out->last = last;
}

void orc_dqagse_lists(int *last, double *alist, double *blist, double *rlist, double *elist,
                      int *iord)
{
    *last = g_last;
    for (int k = 1; k <= g_last; k++) {
        alist[k - 1] = g_alist[k];
        blist[k - 1] = g_blist[k];
        rlist[k - 1] = g_rlist[k];
        elist[k - 1] = g_elist[k];
        iord[k - 1] = g_iord[k];
    }
}

typedef struct {
    int n;
    const double *x, *y, *b, *c, *d;
} spl_info;
static double spl_fn(double u, void *info)
{
    spl_info *s = (spl_info *)info;
    return orc_spline_eval(s->n, s->x, s->y, s->b, s->c, s->d, u);
}

int orc_integrate_spline(int n, const double *x, const double *y, double lo, double hi,
                         orc_quad_out *out)
{
    double *b = (double *)malloc(sizeof(double) * 3 * (size_t)n), *c = b + n, *d = c + n;
    orc_natural_spline(n, x, y, b, c, d);
    spl_info s = {n, x, y, b, c, d};
    const double tol = 1.220703125e-4; /* .Machine$double.eps^0.25 */
    orc_dqagse(spl_fn, &s, lo, hi, tol, tol, 100, out);
    free(b);
    return out->ier;
}

/* ------------------------------------------------------------------------
 * quantile(type=7), median, p.adjust — glm_rob_nb.r:38,77,222; needlestack.r:244,250
 * ---------------------------------------------------------------------- */
double orc_quantile7(int n, const double *s, double p)
{
    double index = 1 + (n > 1 ? (n - 1) : 0) * p;
    double lo = floor(index), hi = ceil(index);
    double qs = s[(int)lo - 1];
    double xh = s[(int)hi - 1];
    if (index > lo && xh != qs) {
        double h = index - lo;
        qs = (1 - h) * qs + h * xh;
    }
    return qs;
}

static int cmp_dbl(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

double orc_median(int n, const double *x)
{
    if (n == 0)
        return NAN;
    double *t = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(t, x, sizeof(double) * (size_t)n);
    qsort(t, (size_t)n, sizeof(double), cmp_dbl);
    double m = (n & 1) ? t[n / 2] : (t[n / 2 - 1] + t[n / 2]) / 2;
    free(t);
    return m;
}

typedef struct {
    double p;
    int i;
} pidx;
static int cmp_pidx_asc(const void *a, const void *b)
{
    const pidx *x = (const pidx *)a, *y = (const pidx *)b;
    if (x->p < y->p)
        return -1;
    if (x->p > y->p)
        return 1;
    return (x->i > y->i) - (x->i < y->i);
}

void orc_padjust_bh(int n, const double *p, double *q)
{
    if (n <= 1) {
        for (int i = 0; i < n; i++)
            q[i] = p[i];
        return;
    }
    pidx *o = (pidx *)malloc(sizeof(pidx) * (size_t)n);
    for (int i = 0; i < n; i++) {
        o[i].p = p[i];
        o[i].i = i;
    }
    qsort(o, (size_t)n, sizeof(pidx), cmp_pidx_asc);
    double cm = INFINITY; /* cummin over decreasing p of n/i*p */
    for (int r = n; r >= 1; r--) {
        double v = ((double)n / (double)r) * o[r - 1].p;
        if (v < cm)
            cm = v;
        q[o[r - 1].i] = cm < 1 ? cm : 1;
    }
    free(o);
}

void orc_padjust_holm(int n, const double *p, double *q)
{
    if (n <= 1) {
        for (int i = 0; i < n; i++)
            q[i] = p[i];
        return;
    }
    pidx *o = (pidx *)malloc(sizeof(pidx) * (size_t)n);
    for (int i = 0; i < n; i++) {
        o[i].p = p[i];
        o[i].i = i;
    }
    qsort(o, (size_t)n, sizeof(pidx), cmp_pidx_asc);
    double cm = -INFINITY;
    for (int r = 1; r <= n; r++) {
        double v = (double)(n + 1 - r) * o[r - 1].p;
        if (v > cm)
            cm = v;
        q[o[r - 1].i] = cm < 1 ? cm : 1;
    }
    free(o);
}
