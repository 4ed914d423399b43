/*
 * oracle/orc_math.h — TEST INFRASTRUCTURE ONLY (CPU oracle, numeric layer).
 *
 * Restates the base-R numeric routines that the reference's hot path calls
 * (bin/glm_rob_nb.r, bin/needlestack.r).  R itself is NOT vendored under
 * /root/reference and is not installed in this image, so these are
 * restatements of R's *published* algorithms (nmath: Loader's saddle-point
 * dbinom/dnbinom, Brent's zeroin, QUADPACK dqagse/dqk21/dqelg, the natural
 * cubic spline of splines.c, p.adjust, type-7 quantile, fisher.test 2x2).
 * PARITY UNPINNED: no golden vector from a real R run exists in the tree
 * (SURVEY.md §4, §8c).  Known answers come from mpmath / scipy (tests/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use anything under oracle/.
 */
#ifndef ORC_MATH_H
#define ORC_MATH_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- densities / tails (R nmath semantics) ---------------------------- */
double orc_stirlerr(double n);
double orc_bd0(double x, double np);
double orc_dbinom_raw(double x, double n, double p, double q);
/* dnbinom(x, size=size, mu=mu)        — glm_rob_nb.r:117,121,133,137,153,157,221 */
double orc_dnbinom_mu(double x, double size, double mu);
/* pnbinom(q, size, mu, lower.tail=FALSE) — glm_rob_nb.r:221, needlestack.r:245,251 */
double orc_pnbinom_mu_upper(double q, double size, double mu);
/* pnbinom(q, size, mu) lower tail (used by qnbinom) */
double orc_pnbinom_mu_lower(double q, double size, double mu);
/* pbinom(q, n, p) lower tail (used by qbinom) */
double orc_pbinom_lower(double q, double n, double p);
/* regularised incomplete beta I_x(a,b) (lower) given x and y=1-x */
double orc_pbeta_raw(double x, double y, double a, double b, int lower);
/* digamma(x), x>0 — glm_rob_nb.r:102,143 */
double orc_digamma(double x);
/* qnbinom(p, size, mu) / qbinom(p, n, prob), lower tail — needlestack.r:243,249 */
double orc_qnbinom_mu(double p, double size, double mu);
double orc_qbinom(double p, double n, double prob);
/* dhyper(x, r, b, n, log=TRUE) and fisher.test(2x2)$p.value two-sided — needlestack.r:234,236 */
double orc_dhyper_log(double x, double r, double b, double n);
double orc_fisher_2x2(double a11, double a21, double a12, double a22);

/* ---- root finding / quadrature / spline ------------------------------- */
typedef double (*orc_fn)(double x, void *info);
/* R_zeroin2 (Brent).  *maxit in: max iterations, out: used (or -1).  *tol out: final precision or -1 */
double orc_zeroin2(double ax, double bx, double fa, double fb, orc_fn f, void *info,
                   double *tol, int *maxit);
/* uniroot(f, c(lo,hi), tol=, maxiter=)$root; returns 0 on success, nonzero = R error */
int orc_uniroot(orc_fn f, void *info, double lo, double hi, double tol, int maxiter,
                double *root, int *n_eval);

/* natural cubic spline coefficients (splinefun(method="natural")) */
void orc_natural_spline(int n, const double *x, const double *y, double *b, double *c, double *d);
double orc_spline_eval(int n, const double *x, const double *y, const double *b,
                       const double *c, const double *d, double u);

/* QUADPACK dqagse as called by R's integrate(): limit<=ORC_QLIMIT */
#define ORC_QLIMIT 100
typedef struct {
    double result, abserr;
    int neval, ier, last;
} orc_quad_out;
void orc_dqagse(orc_fn f, void *info, double a, double b, double epsabs, double epsrel,
                int limit, orc_quad_out *out);
/* exposes the subdivision lists of the last orc_dqagse call (for tests vs scipy) */
void orc_dqagse_lists(int *last, double *alist, double *blist, double *rlist, double *elist, int *iord);
/* test hook: integrate a natural spline through (x,y) over [lo,hi] with R's integrate() defaults */
int orc_integrate_spline(int n, const double *x, const double *y, double lo, double hi,
                         orc_quad_out *out);

/* ---- small statistics -------------------------------------------------- */
double orc_quantile7(int n, const double *sorted, double p);
double orc_median(int n, const double *x);          /* copies + sorts */
void orc_padjust_bh(int n, const double *p, double *q);
void orc_padjust_holm(int n, const double *p, double *q);

#ifdef __cplusplus
}
#endif
#endif
