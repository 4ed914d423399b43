/*
 * oracle/orc_model.c — TEST INFRASTRUCTURE ONLY (CPU oracle, model layer).
 * Literal restatement of /root/reference/bin/glm_rob_nb.r:16-228 and of the
 * per-allele body of /root/reference/bin/needlestack.r (see orc_model.h).
 * PARITY UNPINNED (no R in this image; no golden vectors in the reference).
 */
#include "orc_model.h"
#include "orc_math.h"
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

void orc_glm_params_default(orc_glm_params *p)
{ /* glm_rob_nb.r:16-18 */
    memset(p, 0, sizeof(*p));
    p->c_tukey_beta = 5;
    p->c_tukey_sig = 3;
    p->minsig = 1e-3;
    p->maxsig = 10;
    p->minmu = 1e-10;
    p->sig_prec = 1e-8;
    p->tol = 1e-6;
    p->maxit = 30;
    p->maxit_sig = 50;
    p->n_ai_sig_tukey = 100;
    p->min_coverage = 1;
    p->min_reads = 1;
    p->min_af = 0;
    p->size_min = 10;
    p->extra_rob = 1;
    p->min_af_extra_rob = 0.2;
    p->min_prop_extra_rob = 0.1;
    p->max_prop_extra_rob = 0.5;
}

void orc_call_params_default(orc_call_params *p)
{ /* needlestack.r:64-92 */
    memset(p, 0, sizeof(*p));
    orc_glm_params_default(&p->glm);
    p->glm.min_coverage = 30;
    p->glm.min_reads = 3;
    p->glm.min_af = 0;
    p->glm.extra_rob = 0;
    p->glm.max_prop_extra_rob = 0.8;
    p->GQ_threshold = 50;
    p->SB_type = 0;
    p->SB_threshold_SNV = 100;
    p->SB_threshold_indel = 100;
    p->sigma_normal = 0.1;
    p->afmin_power = -1;
}

/* ---- fit state --------------------------------------------------------- */
typedef struct {
    const orc_glm_params *prm;
    int n;
    const double *y;
    const double *mu;
    orc_glm_info *info;
    int quad_err; /* set when integrate() would have raised an R error */
} fit_state;

static double varfunc(double mu, double sig) { return mu + sig * (mu * mu); } /* :73 */

static double tukeypsi(double r, double c)
{ /* :106-108 */
    if (isnan(r))
        return r;
    if (fabs(r) > c)
        return 0.0;
    double t = (r / c) * (r / c) - 1;
    return t * t * r;
}

typedef struct {
    int n;
    double x[ORC_QLIMIT + 1], yv[ORC_QLIMIT + 1], b[ORC_QLIMIT + 1], c[ORC_QLIMIT + 1],
        d[ORC_QLIMIT + 1];
} spline100;

static double spl_eval_cb(double u, void *info)
{
    spline100 *s = (spline100 *)info;
    return orc_spline_eval(s->n, s->x, s->yv, s->b, s->c, s->d, u);
}

/* integrate(splinefun(x=j12, y=yj12, method="natural"), lower=min(j12), upper=max(j12)+1)[[1]] */
static double integrate_spline(spline100 *s, fit_state *st)
{
    orc_natural_spline(s->n, s->x, s->yv, s->b, s->c, s->d);
    orc_quad_out q;
    const double tol = 1.220703125e-4; /* .Machine$double.eps^0.25 */
    orc_dqagse(spl_eval_cb, s, s->x[0], s->x[s->n - 1] + 1, tol, tol, 100, &q);
    st->info->n_quad += q.neval / 21;
    if (q.ier != 0 || !isfinite(q.result))
        st->quad_err = 1; /* R: stop(res$message) */
    return q.result;
}

/* round(seq(j1, j2, length.out = l)) — seq.default: c(from, from + (1:(l-2))*by, to) */
static void knots(double j1, double j2, int l, double *out)
{
    double by = (j2 - j1) / (double)(l - 1);
    out[0] = nearbyint(j1);
    for (int k = 1; k < l - 1; k++)
        out[k] = nearbyint(j1 + (double)k * by);
    out[l - 1] = nearbyint(j2);
}

/* E.tukeypsi.1 (:109-124) and E.tukeypsi.2 (:125-140), evaluated together */
static void E_tukeypsi_12(double mui, double sig, double c, fit_state *st, double *e1, double *e2)
{
    const int n_ai = st->prm->n_ai_sig_tukey;
    double sqrtV = sqrt(varfunc(mui, sig));
    double j1 = fmax(ceil(mui - c * sqrtV), 0);
    double j2 = floor(mui + c * sqrtV);
    st->info->n_seed += 1;
    if (j1 > j2) {
        *e1 = 0;
        *e2 = 0;
        return;
    }
    double size = 1 / sig;
    if (j2 - j1 + 1 > 2 * n_ai) {
        spline100 s1, s2;
        s1.n = s2.n = n_ai;
        knots(j1, j2, n_ai, s1.x);
        memcpy(s2.x, s1.x, sizeof(double) * (size_t)n_ai);
        for (int k = 0; k < n_ai; k++) {
            double j = s1.x[k];
            double t = (j - mui) / (c * sqrtV);
            double w = (t * t - 1) * (t * t - 1);
            double pm = orc_dnbinom_mu(j, size, mui);
            s1.yv[k] = w * (j - mui) * pm;
            s2.yv[k] = w * ((j - mui) * (j - mui)) * pm;
        }
        st->info->n_lattice += n_ai;
        *e1 = integrate_spline(&s1, st) / sqrtV;
        *e2 = integrate_spline(&s2, st) / sqrtV;
    } else {
        long double a1 = 0, a2 = 0; /* R's sum() accumulates in long double */
        for (double j = j1; j <= j2; j += 1) {
            double t = (j - mui) / (c * sqrtV);
            double w = (t * t - 1) * (t * t - 1);
            double pm = orc_dnbinom_mu(j, size, mui);
            a1 += w * (j - mui) * pm;
            a2 += w * ((j - mui) * (j - mui)) * pm;
        }
        st->info->n_lattice += (int64_t)(j2 - j1 + 1);
        *e1 = (double)a1 / sqrtV;
        *e2 = (double)a2 / sqrtV;
    }
}

/* ai.sig.tukey (:141-160) */
static double ai_sig_tukey(double mui, double sig, double c, fit_state *st)
{
    const int n_ai = st->prm->n_ai_sig_tukey;
    double sqrtV = sqrt(mui * (sig * mui + 1));
    double invsig = 1 / sig;
    double j1 = fmax(ceil(mui - c * sqrtV), 0);
    double j2 = floor(mui + c * sqrtV);
    st->info->n_seed += 1;
    if (j1 > j2)
        return 0;
    double dg0 = orc_digamma(invsig), lg = log(mui / invsig + 1);
    if (j2 - j1 + 1 > 2 * n_ai) {
        spline100 s;
        s.n = n_ai;
        knots(j1, j2, n_ai, s.x);
        for (int k = 0; k < n_ai; k++) {
            double j = s.x[k];
            double t = (j - mui) / (c * sqrtV);
            double w = (t * t - 1) * (t * t - 1);
            double psi = orc_digamma(j + invsig) - dg0 - lg - (j - mui) / (mui + invsig);
            s.yv[k] = w * psi * orc_dnbinom_mu(j, invsig, mui);
        }
        st->info->n_lattice += n_ai;
        return integrate_spline(&s, st);
    } else {
        long double a = 0;
        for (double j = j1; j <= j2; j += 1) {
            double t = (j - mui) / (c * sqrtV);
            double w = (t * t - 1) * (t * t - 1);
            double psi = orc_digamma(j + invsig) - dg0 - lg - (j - mui) / (mui + invsig);
            a += w * psi * orc_dnbinom_mu(j, invsig, mui);
        }
        st->info->n_lattice += (int64_t)(j2 - j1 + 1);
        return (double)a;
    }
}

/* psi.sig.ML (:101-103) */
static double psi_sig_ML(double r, double mu, double sig)
{
    return orc_digamma(r * sqrt(mu * (sig * mu + 1)) + mu + 1 / sig) -
           sig * r * sqrt(mu / (sig * mu + 1)) - orc_digamma(1 / sig) - log(sig * mu + 1);
}

/* sig.rob.tukey (:161-165) */
static double sig_rob_tukey(double sig, void *vst)
{
    fit_state *st = (fit_state *)vst;
    const double c = st->prm->c_tukey_sig;
    long double acc = 0;
    st->info->n_obj_eval += 1;
    for (int i = 0; i < st->n; i++) {
        double mu = st->mu[i];
        double r = (st->y[i] - mu) / sqrt(varfunc(mu, sig));
        double wi = tukeypsi(r, c) / r;
        double term = wi * psi_sig_ML(r, mu, sig) - ai_sig_tukey(mu, sig, c, st);
        acc += term;
    }
    if (st->quad_err)
        return NAN; /* integrate() error inside uniroot -> tryCatch -> NA (:175-177) */
    return (double)acc;
}

static int cmp_dbl(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* the c(coef,1) recycling quirk of :171,187,195 — see SURVEY.md App. A6 */
typedef struct {
    double v[2];
    int len;
} rvec;
static double D_absmax(rvec u, rvec v)
{ /* abs(max(u - v)) with R recycling */
    int L = u.len > v.len ? u.len : v.len;
    double m = -INFINITY;
    int has_nan = 0;
    for (int k = 0; k < L; k++) {
        double d = u.v[k % u.len] - v.v[k % v.len];
        if (isnan(d))
            has_nan = 1;
        else if (d > m)
            m = d;
    }
    if (has_nan)
        return NAN;
    return fabs(m);
}

/* optional trace of the outer alternation (sigma and beta after every pass), for the divergence diagnostics of
 * tests/test_chaotic_units.py; thread-local, set by orc_glmrob_nb_trace() */
static __thread double *g_trace_sig = NULL, *g_trace_beta = NULL;
static __thread int g_trace_cap = 0;

int orc_glmrob_nb(int N, const double *y_in, const double *x_in, const orc_glm_params *prm,
                  double *pvalues, double *qvalues, double *gq, orc_glm_info *info)
{
    memset(info, 0, sizeof(*info));
    info->sigma = NAN;
    info->slope = NAN;
    double *x = (double *)malloc(sizeof(double) * (size_t)(6 * N + 6));
    double *y = x + N, *X = y + N, *mu = X + N, *eta = mu + N, *af = eta + N;
    unsigned char *removed = (unsigned char *)calloc((size_t)N + 1, 1);
    double min_reads = prm->min_reads, min_af = prm->min_af;

    /* :20 */
    double maxmu = -INFINITY;
    for (int i = 0; i < N; i++)
        if (x_in[i] > maxmu)
            maxmu = x_in[i];

    /* :22-36 extra-robust pre-filter */
    int nS = 0;
    for (int i = 0; i < N; i++)
        if (y_in[i] / x_in[i] > prm->min_af_extra_rob)
            nS++;
    int extra_rob_out = 0;
    if (prm->extra_rob && nS > prm->min_prop_extra_rob * N && nS < prm->max_prop_extra_rob * N) {
        if (N - nS > prm->size_min)
            extra_rob_out = 1;
    }
    int n0 = 0;
    for (int i = 0; i < N; i++) {
        int rem = extra_rob_out && (y_in[i] / x_in[i] > prm->min_af_extra_rob);
        removed[i] = (unsigned char)rem;
        if (!rem) {
            x[n0] = x_in[i];
            y[n0] = y_in[i];
            n0++;
        }
    }
    if (extra_rob_out) {
        min_reads = 0;
        min_af = 0;
    }
    info->extra_rob = extra_rob_out;

    /* :38 gate (on the possibly reduced vectors) */
    {
        double med = orc_median(n0, x);
        int n_cov = 0, n_pos = 0;
        double maxy = -INFINITY, maxaf = -INFINITY;
        for (int i = 0; i < n0; i++) {
            if (x[i] > prm->min_coverage)
                n_cov++;
            if (y[i] > 0)
                n_pos++;
            if (y[i] > maxy)
                maxy = y[i];
            double a = y[i] / x[i];
            if (!isnan(a) && a > maxaf)
                maxaf = a;
        }
        int gate = (med < prm->min_coverage) || (n_cov < prm->size_min) || (maxy < min_reads) ||
                   (n_pos == 0 && !extra_rob_out) || (maxaf < min_af);
        if (n0 == 0 || isnan(med))
            gate = 1;
        if (gate) {
            info->gated = 1;
            for (int i = 0; i < N; i++) {
                pvalues[i] = 1;
                qvalues[i] = 1;
                gq[i] = 0;
            }
            free(x);
            free(removed);
            return 0;
        }
    }

    /* :58-63 keep x>0 */
    int n = 0;
    for (int i = 0; i < n0; i++)
        if (x[i] > 0) {
            x[n] = x[i];
            y[n] = y[i];
            n++;
        }
    for (int i = 0; i < n; i++)
        X[i] = log(x[i]);

    /* :75-78 initial slope */
    long double sum_ex = 0;
    for (int i = 0; i < n; i++) {
        double ex = exp(X[i]);
        af[i] = (y[i] == 0) ? 0.0 : y[i] / ex;
        sum_ex += ex;
    }
    double *srt = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(srt, af, sizeof(double) * (size_t)n);
    qsort(srt, (size_t)n, sizeof(double), cmp_dbl);
    double q25 = orc_quantile7(n, srt, 0.25), q75 = orc_quantile7(n, srt, 0.75);
    free(srt);
    double fence = q75 + 1.5 * (q75 - q25);
    long double s_af = 0;
    int c_af = 0;
    for (int i = 0; i < n; i++)
        if (af[i] <= fence) {
            s_af += af[i];
            c_af++;
        }
    double inislope = (double)(s_af / c_af);
    {
        double mean_ex = (double)(sum_ex / n);
        if (inislope <= 1 / (10 * mean_ex))
            inislope = 1 / (10 * mean_ex);
    }
    /* :79-86 */
    long double s_sig = 0;
    for (int i = 0; i < n; i++) {
        mu[i] = inislope * exp(X[i]);
        if (mu[i] > maxmu)
            mu[i] = maxmu;
        if (mu[i] < prm->minmu)
            mu[i] = prm->minmu;
        eta[i] = log(mu[i]);
        double t = y[i] / mu[i] - 1;
        s_sig += t * t;
    }
    double sig = (double)(s_sig / n);
    if (sig > prm->maxsig)
        sig = prm->maxsig;
    if (sig < prm->minsig)
        sig = prm->minsig;

    /* :166-205 robust alternation */
    fit_state st = {prm, n, y, mu, info, 0};
    const double tol = prm->tol;
    double sig0 = sig + 1 + tol;
    rvec beta11 = {{log(inislope), 0}, 1};
    rvec beta00 = {{beta11.v[0] + tol + 1, 0}, 1};
    rvec beta1 = beta11;
    int it = 0;
    while ((fabs(sig - sig0) > tol || D_absmax(beta11, beta00) > tol) && it < prm->maxit) {
        sig0 = sig;
        beta00 = beta11;
        /* :175-182 sigma given mu */
        double root = NAN;
        st.quad_err = 0;
        int rc = orc_uniroot(sig_rob_tukey, &st, prm->minsig, prm->maxsig, prm->sig_prec,
                             prm->maxit_sig, &root, NULL);
        if (rc != 0 || isnan(root)) {
            sig = prm->minsig;
            info->sig_at_min += 1;
        } else {
            sig = root;
            if (sig > prm->maxsig)
                sig = prm->maxsig;
            if (sig < prm->minsig)
                sig = prm->minsig;
        }
        st.quad_err = 0;
        /* :184-202 mu given sigma */
        beta1.v[0] = log(inislope);
        beta1.len = 1;
        rvec beta0 = {{beta1.v[0] + tol + 1, 0}, 1};
        int it_mu = 0;
        while (D_absmax(beta1, beta0) > tol && it_mu < prm->maxit) {
            beta0 = beta1;
            long double sw = 0, swz = 0;
            for (int i = 0; i < n; i++) {
                double e1, sap;
                E_tukeypsi_12(mu[i], sig, prm->c_tukey_beta, &st, &e1, &sap);
                double V = varfunc(mu[i], sig);
                double bi = sap * pow(V, -1.5) * (exp(eta[i]) * exp(eta[i]));
                double ei = (tukeypsi((y[i] - mu[i]) / sqrt(V), prm->c_tukey_beta) - e1) / sap * V *
                            (1 / mu[i]);
                double zi = eta[i] + ei;
                if (isnan(bi) || isnan(zi)) { /* lm(): na.omit drops the row */
                    info->nan_rows += 1;
                    continue;
                }
                if (bi == 0)
                    continue; /* lm(): zero-weight rows excluded */
                sw += bi;
                swz += bi * (zi - X[i]);
            }
            double beta = (double)(swz / sw);
            beta1.v[0] = beta;
            beta1.v[1] = 1;
            beta1.len = 2;
            for (int i = 0; i < n; i++) {
                eta[i] = X[i] + beta;
                mu[i] = exp(eta[i]);
                if (mu[i] > maxmu)
                    mu[i] = maxmu;
                if (mu[i] == 0)
                    mu[i] = prm->minmu;
                eta[i] = log(mu[i]);
            }
            it_mu++;
            info->n_irls += 1;
        }
        if (st.quad_err)
            info->quad_error = 1; /* R would have stopped the script here */
        beta11 = beta1;
        if (g_trace_sig && it < g_trace_cap) {
            g_trace_sig[it] = sig;
            g_trace_beta[it] = beta1.v[0];
        }
        it++;
    }
    info->n_outer = it;
    info->maxit_hit = (it >= prm->maxit);

    /* :207-224 result on the full vectors */
    double slope = exp(beta1.v[0]);
    if (slope > 1)
        slope = 1;
    info->sigma = sig;
    info->slope = slope;
    for (int i = 0; i < N; i++) {
        double m = slope * x_in[i];
        pvalues[i] = orc_dnbinom_mu(y_in[i], 1 / sig, m) + orc_pnbinom_mu_upper(y_in[i], 1 / sig, m);
    }
    orc_padjust_bh(N, pvalues, qvalues);
    for (int i = 0; i < N; i++) {
        double g = -log10(qvalues[i]) * 10;
        if (g > 1000)
            g = 1000;
        gq[i] = g;
    }
    free(x);
    free(removed);
    return 0;
}

int orc_glmrob_nb_trace(int N, const double *y, const double *x, const orc_glm_params *prm, int cap, double *sig_trace,
                        double *beta_trace, orc_glm_info *info)
{
    double *p = (double *)malloc(sizeof(double) * 3 * (size_t)(N + 1));
    g_trace_sig = sig_trace;
    g_trace_beta = beta_trace;
    g_trace_cap = cap;
    int rc = orc_glmrob_nb(N, y, x, prm, p, p + N, p + 2 * N, info);
    g_trace_sig = g_trace_beta = NULL;
    g_trace_cap = 0;
    free(p);
    return rc;
}

/* ---- needlestack.r per-allele body -------------------------------------- */

/* SOR(), needlestack.r:206-215 (IEEE semantics reproduce R's NaN/Inf) */
static double sor_fn(double RO_f, double AO_f, double RO_r, double AO_r)
{
    double X00 = RO_f, X10 = AO_f, X01 = RO_r, X11 = AO_r;
    double refRatio = fmax(X00, X01) / fmin(X00, X01);
    double altRatio = fmax(X10, X11) / fmin(X10, X11);
    double R = (X00 / X01) * (X11 / X10);
    return log((R + 1 / R) / (refRatio / altRatio));
}

/* toQvalueN / toQvalueT, needlestack.r:242-252: Holm-adjusted (p.adjust default) tail
 * probability of the 1% quantile count at depth x, among the n fitted p-values */
static double to_qvalue(double x, int use_binom, int n, const double *pv, double sigma_fit,
                        double slope_fit, double sigma_normal, double afmin, double *ptmp,
                        double *qtmp)
{
    if (isnan(sigma_fit) || isnan(slope_fit))
        return NAN; /* dnbinom(size=1/NA) = NA */
    double ystar = use_binom ? orc_qbinom(0.01, x, afmin) : orc_qnbinom_mu(0.01, 1 / sigma_normal, 0.5 * x);
    double m = slope_fit * x;
    double pstar = orc_dnbinom_mu(ystar, 1 / sigma_fit, m) + orc_pnbinom_mu_upper(ystar, 1 / sigma_fit, m);
    memcpy(ptmp, pv, sizeof(double) * (size_t)n);
    ptmp[n] = pstar;
    orc_padjust_holm(n + 1, ptmp, qtmp);
    return -10 * log10(qtmp[n]);
}

int orc_call_unit(int n, const int32_t *dp, const int32_t *vp, const int32_t *vm,
                  const int32_t *rp, const int32_t *rm, int is_indel, const orc_call_params *prm,
                  double *pvalue, double *qvalue, double *gq, double *qval_minaf, double *sor,
                  double *rvsb, double *fs, uint8_t *gt, uint8_t *status, orc_unit_info *info)
{
    memset(info, 0, sizeof(*info));
    double *DP = (double *)malloc(sizeof(double) * (size_t)(4 * n + 4));
    double *ma = DP + n, *yreg = ma + n, *ptmp = yreg + n; /* ptmp has n+1.. use separate */
    double *hp = (double *)malloc(sizeof(double) * (size_t)(2 * n + 2));
    double *hq = hp + n + 1;
    (void)ptmp;
    const double thr = prm->GQ_threshold;

    /* :326-337 */
    int cnt_inv = 0;
    for (int i = 0; i < n; i++) {
        DP[i] = dp[i];
        ma[i] = (double)vp[i] + (double)vm[i];
        if (ma[i] / DP[i] > 0.8)
            cnt_inv++; /* NaN compares FALSE = na.rm */
    }
    int ref_inv = cnt_inv > 0.5 * n;
    info->ref_inv = ref_inv;
    for (int i = 0; i < n; i++)
        yreg[i] = ref_inv ? (double)rp[i] + (double)rm[i] : ma[i];

    orc_glmrob_nb(n, yreg, DP, &prm->glm, pvalue, qvalue, gq, &info->glm);
    double sigma_fit = info->glm.sigma, slope_fit = info->glm.slope;

    /* :340-380 qval_minAF + somatic status */
    for (int i = 0; i < n; i++) {
        qval_minaf[i] = 0;
        status[i] = ORC_ST_NONE;
    }
#define TOQ_N(i) to_qvalue(DP[i], 0, n, pvalue, sigma_fit, slope_fit, prm->sigma_normal, 0, hp, hq)
#define TOQ_T(i) to_qvalue(DP[i], 1, n, pvalue, sigma_fit, slope_fit, 0, afmin, hp, hq)
    if (prm->is_tn) {
        double afmin = prm->afmin_power == -1 ? 0.01 : prm->afmin_power; /* :166 */
        const int32_t *T = prm->tindex, *Nn = prm->nindex;
        int np = prm->n_pairs;
        for (int k = 0; k < np; k++)
            qval_minaf[Nn[k]] = TOQ_N(Nn[k]);
        int germ = 0;
        for (int k = 0; k < np; k++)
            if (gq[Nn[k]] > thr)
                germ++;
        for (int k = 0; k < np; k++)
            qval_minaf[T[k]] = germ > 0 ? TOQ_N(T[k]) : TOQ_T(T[k]);
        for (int k = 0; k < prm->n_only_n; k++)
            qval_minaf[prm->only_n[k]] = TOQ_N(prm->only_n[k]);
        for (int k = 0; k < prm->n_only_t; k++)
            qval_minaf[prm->only_t[k]] = TOQ_T(prm->only_t[k]);
        /* :349-369, in the reference's assignment order */
        for (int k = 0; k < prm->n_only_t; k++)
            if (gq[prm->only_t[k]] > thr)
                status[prm->only_t[k]] = ORC_ST_UNKNOWN;
        for (int k = 0; k < prm->n_only_n; k++)
            if (gq[prm->only_n[k]] > thr)
                status[prm->only_n[k]] = ORC_ST_GERMLINE_UNCONFIRMABLE;
#define RULE(cond, val, both)                                                                      \
    for (int k = 0; k < np; k++)                                                                   \
        if (cond)                                                                                  \
            status[T[k]] = (val);                                                                  \
    if (both)                                                                                      \
        for (int k = 0; k < np; k++)                                                               \
            if (cond)                                                                              \
                status[Nn[k]] = (val);
        RULE((gq[T[k]] < thr) && (qval_minaf[T[k]] > thr) && (gq[Nn[k]] > thr), ORC_ST_GERMLINE_UNCONFIRMED, 1)
        RULE((gq[T[k]] < thr) && (qval_minaf[T[k]] < thr) && (gq[Nn[k]] > thr), ORC_ST_GERMLINE_UNCONFIRMABLE, 1)
        RULE((gq[T[k]] > thr) && (qval_minaf[Nn[k]] < thr) && (gq[Nn[k]] < thr), ORC_ST_UNKNOWN, 1)
        RULE((gq[T[k]] > thr) && (gq[Nn[k]] > thr), ORC_ST_GERMLINE_CONFIRMED, 1)
        RULE((gq[T[k]] > thr) && (qval_minaf[Nn[k]] > thr) && (gq[Nn[k]] < thr), ORC_ST_SOMATIC, 0)
#undef RULE
        /* :372-373 possible contamination */
        int n_germ = 0;
        for (int i = 0; i < n; i++)
            if (status[i] == ORC_ST_GERMLINE_CONFIRMED || status[i] == ORC_ST_GERMLINE_UNCONFIRMED ||
                status[i] == ORC_ST_GERMLINE_UNCONFIRMABLE || status[i] == ORC_ST_UNKNOWN)
                n_germ++;
        if (n_germ > 0)
            for (int k = 0; k < np; k++)
                if (status[T[k]] == ORC_ST_SOMATIC)
                    status[T[k]] = ORC_ST_POSSIBLE_CONTAMINATION;
    } else {
        double afmin = prm->afmin_power;
        for (int i = 0; i < n; i++)
            qval_minaf[i] = (prm->afmin_power == -1) ? TOQ_N(i) : TOQ_T(i);
    }
#undef TOQ_N
#undef TOQ_T

    /* :381 gate 1 (output_all_SNVs only exists in the SNV block; :534,:698 for indels) */
    double maxgq = -INFINITY;
    int any_gq = 0;
    for (int i = 0; i < n; i++) {
        if (gq[i] > maxgq)
            maxgq = gq[i];
        if (gq[i] >= thr)
            any_gq = 1;
    }
    info->max_gq = maxgq;
    int out_all = (!is_indel) && prm->output_all_SNVs;
    info->gate1 = out_all || (!isnan(slope_fit) && any_gq);
    for (int i = 0; i < n; i++) {
        sor[i] = rvsb[i] = fs[i] = NAN;
        gt[i] = ref_inv ? ORC_GT_11 : ORC_GT_00;
    }
    if (info->gate1) {
        /* common_annot(), :217-239 */
        double sVp = 0, sVm = 0, sRp = 0, sRm = 0, sCp = 0, sCm = 0, sDP = 0, sAO = 0;
        for (int i = 0; i < n; i++) {
            double Vp = vp[i], Vm = vm[i], Rp = rp[i], Rm = rm[i];
            double Cp = Vp + Rp, Cm = Vm + Rm;
            double t = fmax(Vp * Cm, Vm * Cp) / (Vp * Cm + Vm * Cp);
            rvsb[i] = isnan(t) ? -1 : t;
            double s = sor_fn(Rp, Vp, Rm, Vm);
            if (isnan(s))
                s = -1;
            else if (isinf(s))
                s = 99;
            sor[i] = s;
            fs[i] = -10 * log10(orc_fisher_2x2(Vp, Vm, Rp, Rm)); /* :234; the cap at :235 is a no-op */
            sVp += Vp;
            sVm += Vm;
            sRp += Rp;
            sRm += Rm;
            sCp += Cp;
            sCm += Cm;
            sDP += DP[i];
            sAO += ma[i];
        }
        double ar = fmax(sVp * sCm, sVm * sCp) / (sVp * sCm + sVm * sCp);
        if (isnan(ar))
            ar = -1;
        double as = sor_fn(sRp, sVp, sRm, sVm);
        if (isnan(as))
            as = -1;
        else if (isinf(as))
            as = 99;
        info->pooled[0] = as;
        info->pooled[1] = ar;
        info->pooled[2] = -10 * log10(orc_fisher_2x2(sVp, sVm, sRp, sRm));
        info->pooled[3] = sRp;
        info->pooled[4] = sRm;
        info->pooled[5] = sVp;
        info->pooled[6] = sVm;
        info->pooled[7] = is_indel ? sDP + sAO : sDP; /* :383 vs :536,:700 */
        const double *sbs = prm->SB_type == 0 ? sor : (prm->SB_type == 1 ? rvsb : fs);
        double sb_thr = is_indel ? prm->SB_threshold_indel : prm->SB_threshold_SNV;
        int any2 = 0;
        for (int i = 0; i < n; i++)
            if (sbs[i] <= sb_thr && gq[i] >= thr)
                any2 = 1;
        info->gate2 = out_all || any2;
        /* genotype, :403-409 (:555-561, :718-724 also use SB_threshold_SNV) */
        for (int i = 0; i < n; i++) {
            double af = yreg[i] / DP[i]; /* reg_res$ma_count/reg_res$coverage */
            if (gq[i] >= thr && sbs[i] <= prm->SB_threshold_SNV && af < 0.75)
                gt[i] = ORC_GT_01;
            if (gq[i] >= thr && sbs[i] <= prm->SB_threshold_SNV && af >= 0.75)
                gt[i] = ref_inv ? ORC_GT_00 : ORC_GT_11;
            if (gq[i] < thr && qval_minaf[i] < thr)
                gt[i] = ORC_GT_MISSING;
        }
    }
    free(DP);
    free(hp);
    return 0;
}

/* ---- SNV batch ----------------------------------------------------------- */
typedef struct {
    int64_t P, p0, p1;
    int n;
    const uint8_t *ref;
    const int32_t *counts;
    const orc_call_params *prm;
    double *pvalue, *qvalue, *gq, *qval_minaf, *sor, *rvsb, *fs;
    uint8_t *gt, *status;
    orc_unit_info *info;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    const int n = j->n;
    const int64_t plane = j->P * (int64_t)n;
    for (int64_t p = j->p0; p < j->p1; p++) {
        int ref = j->ref[p];
        int k = 0;
        for (int alt = 0; alt < 4; alt++) { /* setdiff(c("A","T","C","G"), ref) */
            if (alt == ref)
                continue;
            int64_t u = p * 3 + k;
            k++;
            size_t off = (size_t)u * (size_t)n;
            if (ref > 3) { /* needlestack.r:319: position skipped */
                memset(&j->info[u], 0, sizeof(orc_unit_info));
                j->info[u].glm.gated = 1;
                j->info[u].glm.sigma = j->info[u].glm.slope = NAN;
                for (int i = 0; i < n; i++) {
                    j->pvalue[off + i] = 1;
                    j->qvalue[off + i] = 1;
                    j->gq[off + i] = 0;
                    j->qval_minaf[off + i] = NAN;
                    j->sor[off + i] = j->rvsb[off + i] = j->fs[off + i] = NAN;
                    j->gt[off + i] = 0;
                    j->status[off + i] = 0;
                }
                if (k == 3)
                    break;
                continue;
            }
            const int32_t *row = j->counts + p * (int64_t)n;
            orc_call_unit(n, row, row + (1 + alt) * plane, row + (5 + alt) * plane,
                          row + (1 + ref) * plane, row + (5 + ref) * plane, 0, j->prm,
                          j->pvalue + off, j->qvalue + off, j->gq + off, j->qval_minaf + off,
                          j->sor + off, j->rvsb + off, j->fs + off, j->gt + off, j->status + off,
                          &j->info[u]);
        }
    }
    return NULL;
}

int orc_call_snv_batch(int64_t P, int n, const uint8_t *ref, const int32_t *counts,
                       const orc_call_params *prm, int n_threads, double *pvalue, double *qvalue,
                       double *gq, double *qval_minaf, double *sor, double *rvsb, double *fs,
                       uint8_t *gt, uint8_t *status, orc_unit_info *info)
{
    if (n_threads < 1)
        n_threads = 1;
    if (n_threads > 256)
        n_threads = 256;
    pthread_t th[256];
    batch_job jobs[256];
    /* interleaved blocks of positions keep the load balanced enough for a baseline */
    int64_t per = (P + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; t++) {
        batch_job jb = {P, t * per, (t + 1) * per < P ? (t + 1) * per : P, n, ref, counts, prm,
                        pvalue, qvalue, gq, qval_minaf, sor, rvsb, fs, gt, status, info};
        jobs[t] = jb;
        if (jobs[t].p0 > P)
            jobs[t].p0 = P;
    }
    if (n_threads == 1) {
        batch_worker(&jobs[0]);
        return 0;
    }
    for (int t = 0; t < n_threads; t++)
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    for (int t = 0; t < n_threads; t++)
        pthread_join(th[t], NULL);
    return 0;
}

/* toQvalue(x, y) of /root/reference/bin/plot_rob_nb.r:83-87 (the contour / power grid of
 * plot_rob_nb, :80-121): the new point (x = DP, y = AO) is appended to the n observations, all
 * n + 1 tail probabilities dnbinom + pnbinom(lower.tail = F) are Holm-adjusted (p.adjust's
 * default method) and the new point's value is returned on the Phred scale; 0 when y > x. */
int orc_plot_toqvalue(int n, const int32_t *coverage, const int32_t *ma_count, double sigma, double slope,
                      int64_t m, const int32_t *gx, const int32_t *gy, double *qval)
{
    double *p = (double *)malloc(sizeof(double) * (size_t)(2 * n + 2));
    double *q = p + n + 1;
    if (!p)
        return -1;
    for (int i = 0; i < n; i++) {
        double mu = slope * (double)coverage[i];
        p[i] = orc_dnbinom_mu((double)ma_count[i], 1 / sigma, mu) + orc_pnbinom_mu_upper((double)ma_count[i], 1 / sigma, mu);
    }
    for (int64_t g = 0; g < m; g++) {
        double x = gx[g], y = gy[g];
        if (y > x) {
            qval[g] = 0;
            continue;
        }
        p[n] = orc_dnbinom_mu(y, 1 / sigma, slope * x) + orc_pnbinom_mu_upper(y, 1 / sigma, slope * x);
        orc_padjust_holm(n + 1, p, q);
        qval[g] = -10 * log10(q[n]);
    }
    free(p);
    return 0;
}
