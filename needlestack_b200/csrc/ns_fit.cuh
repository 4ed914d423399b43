/*
 * ns_fit.cuh — the robust negative-binomial regression of one unit
 * (/root/reference/bin/glm_rob_nb.r:58-205), for a group of lanes G.
 *
 * Two code paths compute the same estimator:
 *   FAST = true   what the persistent warp kernel k_fit runs.  A window that starts at 0 and
 *                 is shorter than NS_JTAB lattice points (every sample at exome depth) is
 *                 summed with division-free recurrences: the negative-binomial pmf from
 *                 pmf(0) = exp(-size*log(1+sig*mu)), the digamma difference
 *                 digamma(j+1/sig) - digamma(1/sig) = sum_{k<j} 1/(1/sig+k) from a per-sigma
 *                 table the warp builds once per objective evaluation, 1/(j+1) from a table.
 *                 Longer windows use the generic lattice sum; a unit that needs the spline +
 *                 quadrature branch (window > 2*n_ai.sig.tukey) is handed to the heavy kernel.
 *   FAST = false  the generic path: Loader pmf seeds, spline + QUADPACK dqags
 *                 (glm_rob_nb.r:115-118,131-134,151-154); run by k_fit_heavy with one CTA per
 *                 unit (BlockG), and by the serial emulation in tests/.
 */
#pragma once
#include "ns_math.cuh"
#include "ns_quad.cuh"

#define NS_JTAB 64  /* digamma-difference table entries per sigma */
#define NS_RTAB 256 /* reciprocals 1/j, j = 1..255 */
#ifndef NS_PT_FACTOR
#define NS_PT_FACTOR 2 /* two-phase evaluation of the cluster kernel from this many samples per warp */
#endif
#define NS_STEP_BY 32.0 /* knot spacing up to which the spline knots are reached by lattice recurrence (a direct pmf + digamma per knot costs ~25 recurrence steps) */

/* The hot functions are out of line, so the compiler no longer sees that the working set lives in
 * shared memory and would emit generic loads (LD + address translation).  In the fast path the
 * tables and per-sample arrays are read through explicit shared-space loads. */
#if defined(__CUDA_ARCH__)
struct ns_sptr {
    unsigned a;
    __device__ __forceinline__ explicit ns_sptr(const double *p) : a((unsigned)__cvta_generic_to_shared(p)) {}
    __device__ __forceinline__ double operator[](int i) const
    {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a + 8u * (unsigned)i));
        return v;
    }
};
#else
struct ns_sptr {
    const double *p;
    explicit ns_sptr(const double *q) : p(q) {}
    double operator[](int i) const { return p[i]; }
};
#endif

struct ns_fit_counters {
    unsigned long long n_seed, n_lattice, n_quad;
    int quad_err;      /* an integrate() that R would have raised as an error */
    int quad_overflow; /* on-chip subdivision list exhausted */
};

NS_HD double ns_varfunc(double mu, double sig) { return mu + sig * (mu * mu); } /* :73 */

NS_HD double ns_tukeypsi(double r, double c)
{ /* glm_rob_nb.r:106-108 */
    if (r != r)
        return r;
    if (fabs(r) > c)
        return 0.0;
    double t = (r / c) * (r / c) - 1;
    return t * t * r;
}

/* One copy each of the large scalar functions for the whole cluster kernel (k_fit keeps them inlined).  Force-inlined
 * they made ns_ai_coop and ns_E12_coop 237 KB and 296 KB of code (three copies of dnbinom_mu with its lgamma /
 * stirlerr / bd0 each, three of dqags) against 32 KB of instruction cache per SM: with the 16-20 warps of a CTA in
 * different phases of different samples, 44 % of the stall samples of the throughput regime (C3) were instruction
 * fetches.  OOL = true selects the shared copy. */
#if defined(__CUDACC__)
static __device__ __noinline__ double ns_dnbinom_mu_c(double x, double size, double mu) { return ns_dnbinom_mu(x, size, mu); }
static __device__ __noinline__ double ns_digamma_c(double x) { return ns_digamma(x); }
#endif
template <bool OOL> NS_GF inline double ns_pmf_seed(double x, double size, double mu)
{
#if defined(__CUDA_ARCH__)
    if (OOL)
        return ns_dnbinom_mu_c(x, size, mu);
#endif
    return ns_dnbinom_mu(x, size, mu);
}
template <bool OOL> NS_GF inline double ns_digamma_sel(double x)
{
#if defined(__CUDA_ARCH__)
    if (OOL)
        return ns_digamma_c(x);
#endif
    return ns_digamma(x);
}

/* ---- generic window expectations ----------------------------------------------------- */
/* lattice-sum branch of ai.sig.tukey (glm_rob_nb.r:155-158) */
template <bool OOL = false>
NS_GF inline double ns_ai_sum(double mui, double sig, double c, double j1, double j2, double dg0)
{
    const double sqrtV = sqrt(mui * (sig * mui + 1));
    const double invsig = 1 / sig;
    const double lg = log(mui / invsig + 1);
    const double csd = c * sqrtV, mpi = mui + invsig;
    double pm = ns_pmf_seed<OOL>(j1, invsig, mui);
    double D = (j1 == 0) ? 0.0 : ns_digamma_sel<OOL>(j1 + invsig) - dg0; /* digamma(j+1/sig)-digamma(1/sig) */
    const double q = mui / (mui + invsig);
    const double inv_csd = 1.0 / csd, inv_mpi = 1.0 / mpi; /* the loop itself is division-free */
    double acc = 0;
    for (double j = j1; j <= j2; j += 1) {
        double d = j - mui;
        double t = d * inv_csd;
        double w = (t * t - 1) * (t * t - 1);
        double psi = D - lg - d * inv_mpi;
        acc += w * psi * pm;
        double ja = j + invsig, jb = j + 1;
        double rr = ns_frcp(ja * jb); /* ja, jb >= 1e-1: normal and positive */
        pm *= (ja * ja) * rr * q; /* pmf(j+1)/pmf(j) = (j+size)/(j+1) * mu/(mu+size) */
        D += jb * rr;             /* digamma(x+1) = digamma(x) + 1/x */
    }
    return acc;
}

/* spline + quadrature branch of ai.sig.tukey (glm_rob_nb.r:151-154) */
NS_GF inline double ns_ai_spline(double mui, double sig, double c, int n_ai, double j1, double j2, double dg0,
                                 ns_fit_counters &cnt)
{
    const double sqrtV = sqrt(mui * (sig * mui + 1));
    const double invsig = 1 / sig;
    const double lg = log(mui / invsig + 1);
    const double csd = c * sqrtV, mpi = mui + invsig;
    ns_spline s;
    ns_spline_init(s, n_ai, j1, j2);
    if (s.by <= NS_STEP_BY) {
        /* walk the lattice once by recurrence and sample it at the knots */
        double pm = ns_dnbinom_mu(j1, invsig, mui);
        double D = (j1 == 0) ? 0.0 : ns_digamma(j1 + invsig) - dg0;
        const double q = mui / (mui + invsig);
        int k = 0;
        for (double j = j1;; j += 1) {
            if (j == s.x[k]) {
                double t = (j - mui) / csd;
                double w = (t * t - 1) * (t * t - 1);
                s.y[k] = w * (D - lg - (j - mui) / mpi) * pm;
                if (++k == n_ai)
                    break;
            }
            double ja = j + invsig, jb = j + 1;
            double rr = 1.0 / (ja * jb);
            pm *= (ja * ja) * rr * q;
            D += jb * rr;
        }
    } else {
        for (int k = 0; k < n_ai; k++) {
            double j = s.x[k];
            double t = (j - mui) / csd;
            double w = (t * t - 1) * (t * t - 1);
            double psi = ns_digamma(j + invsig) - dg0 - lg - (j - mui) / mpi;
            s.y[k] = w * psi * ns_dnbinom_mu(j, invsig, mui);
        }
    }
    ns_spline_diag(s);
    ns_spline_solve(s);
    int ier = 0, np = 0;
    double v = ns_integrate_spline(s, ier, np);
    cnt.n_quad += (unsigned long long)np;
    if (ier == 7)
        cnt.quad_overflow = 1;
    if (ier != 0 || !(fabs(v) <= NS_DBL_MAX))
        cnt.quad_err = 1;
    return v;
}

/* ai.sig.tukey(mui, sig, c) (glm_rob_nb.r:141-160) */
NS_GF inline double ns_ai_sig_tukey(double mui, double sig, double c, int n_ai, double dg0, ns_fit_counters &cnt)
{
    const double sqrtV = sqrt(mui * (sig * mui + 1));
    const double j1 = fmax(ceil(mui - c * sqrtV), 0.0);
    const double j2 = floor(mui + c * sqrtV);
    cnt.n_seed += 1;
    if (j1 > j2)
        return 0.0;
    if (j2 - j1 + 1 > 2 * n_ai) {
        cnt.n_lattice += (unsigned long long)n_ai;
        return ns_ai_spline(mui, sig, c, n_ai, j1, j2, dg0, cnt);
    }
    cnt.n_lattice += (unsigned long long)(j2 - j1 + 1);
    return ns_ai_sum(mui, sig, c, j1, j2, dg0);
}

/* lattice-sum branch of E.tukeypsi.1 / .2 (glm_rob_nb.r:119-122, 135-138), both in one pass */
template <bool OOL = false>
NS_GF inline void ns_E12_sum(double mui, double sig, double c, double j1, double j2, double &e1, double &e2)
{
    const double sqrtV = sqrt(ns_varfunc(mui, sig));
    const double size = 1 / sig, csd = c * sqrtV;
    double pm = ns_pmf_seed<OOL>(j1, size, mui);
    const double q = mui / (mui + size);
    const double inv_csd = 1.0 / csd;
    double a1 = 0, a2 = 0;
    for (double j = j1; j <= j2; j += 1) {
        double d = j - mui;
        double t = d * inv_csd;
        double w = (t * t - 1) * (t * t - 1);
        double wp = w * pm;
        a1 += wp * d;
        a2 += wp * (d * d);
        pm *= ((j + size) * ns_frcp(j + 1)) * q;
    }
    e1 = a1 / sqrtV;
    e2 = a2 / sqrtV;
}

NS_GF inline void ns_E12_spline(double mui, double sig, double c, int n_ai, double j1, double j2,
                                ns_fit_counters &cnt, double &e1, double &e2)
{
    const double sqrtV = sqrt(ns_varfunc(mui, sig));
    const double size = 1 / sig, csd = c * sqrtV;
    ns_spline s;
    ns_spline_init(s, n_ai, j1, j2);
    if (s.by <= NS_STEP_BY) {
        double pm = ns_dnbinom_mu(j1, size, mui);
        const double q = mui / (mui + size);
        int k = 0;
        for (double j = j1;; j += 1) {
            if (j == s.x[k]) {
                double t = (j - mui) / csd;
                double w = (t * t - 1) * (t * t - 1);
                s.y[k] = w * (j - mui) * pm;
                if (++k == n_ai)
                    break;
            }
            pm *= ((j + size) / (j + 1)) * q;
        }
    } else {
        for (int k = 0; k < n_ai; k++) {
            double j = s.x[k];
            double t = (j - mui) / csd;
            double w = (t * t - 1) * (t * t - 1);
            s.y[k] = w * (j - mui) * ns_dnbinom_mu(j, size, mui);
        }
    }
    ns_spline_diag(s);
    ns_spline_solve(s);
    int ier1 = 0, ier2 = 0, np = 0;
    double v1 = ns_integrate_spline(s, ier1, np);
    for (int k = 0; k < n_ai; k++)
        s.y[k] *= (ns_knot(s, k) - mui); /* w*(j-mu)^2*pmf */
    ns_spline_solve(s);
    double v2 = ns_integrate_spline(s, ier2, np);
    cnt.n_quad += (unsigned long long)np;
    if (ier1 == 7 || ier2 == 7)
        cnt.quad_overflow = 1;
    if (ier1 != 0 || ier2 != 0 || !(fabs(v1) <= NS_DBL_MAX) || !(fabs(v2) <= NS_DBL_MAX))
        cnt.quad_err = 1;
    e1 = v1 / sqrtV;
    e2 = v2 / sqrtV;
}

/* E.tukeypsi.1 and E.tukeypsi.2 (glm_rob_nb.r:109-140) */
NS_GF inline void ns_E_tukeypsi_12(double mui, double sig, double c, int n_ai, ns_fit_counters &cnt, double &e1,
                                   double &e2)
{
    const double sqrtV = sqrt(ns_varfunc(mui, sig));
    const double j1 = fmax(ceil(mui - c * sqrtV), 0.0);
    const double j2 = floor(mui + c * sqrtV);
    cnt.n_seed += 1;
    if (j1 > j2) {
        e1 = 0;
        e2 = 0;
        return;
    }
    if (j2 - j1 + 1 > 2 * n_ai) {
        cnt.n_lattice += (unsigned long long)n_ai;
        ns_E12_spline(mui, sig, c, n_ai, j1, j2, cnt, e1, e2);
        return;
    }
    cnt.n_lattice += (unsigned long long)(j2 - j1 + 1);
    ns_E12_sum(mui, sig, c, j1, j2, e1, e2);
}

/* warp-cooperative versions for the heavy kernel (ns_coop.cuh, device only) */
struct ns_coop_scratch {
    double *y, *c, *ib, *m; /* NS_MAX_KNOTS each, per warp, shared memory */
    const double *rt;       /* 1/j, j < rtn (CTA-shared) */
    int rtn;
};
#if defined(__CUDACC__)
template <bool ROLLED>
static __device__ __noinline__ double ns_ai_coop(double mui, double sig, double c, int n_ai, double dg0,
                                                 const ns_coop_scratch &sc, ns_fit_counters &cnt);
template <bool ROLLED>
static __device__ __noinline__ void ns_E12_coop(double mui, double sig, double c, int n_ai, const ns_coop_scratch &sc,
                                                ns_fit_counters &cnt, double *e1, double *e2);
#endif

/* ---- per-group working set (slices of shared memory) ----------------------------- */
struct ns_work {
    double *y, *x, *X, *mu, *eta; /* n each; x <= 0 marks a sample that is not in the fit */
    double *scratch;              /* n + 8 */
    double *dtab;                 /* NS_JTAB: 1/(j+1/sig), the increments of digamma(j+1/sig) for the current sigma */
    double *r2tab;                /* NS_JTAB: (j+1/sig)/(j+1) for the current sigma (pmf ratio without mu/(mu+size)) */
    const double *rtab;           /* NS_RTAB: 1/j (shared by the CTA) */
    const double *ltab;           /* MODE 1: the table of ns_flog_tab (shared by the CTA) */
    ns_coop_scratch coop;         /* MODE 2 only: this warp's spline arrays */
};
NS_HD size_t ns_work_doubles(int n) { return (size_t)6 * (size_t)n + 8 + 2 * NS_JTAB; }
NS_HD void ns_work_bind(ns_work &w, double *base, int n, const double *rtab, const double *ltab = nullptr)
{
    w.ltab = ltab;
    w.y = base;
    w.x = base + n;
    w.X = base + 2 * (size_t)n;
    w.mu = base + 3 * (size_t)n;
    w.eta = base + 4 * (size_t)n;
    w.scratch = base + 5 * (size_t)n;
    w.dtab = base + 6 * (size_t)n + 8;
    w.r2tab = w.dtab + NS_JTAB;
    w.rtab = rtab;
}

struct ns_fit_result {
    double sigma, slope;
    uint32_t flags; /* NS_F_SIG_AT_MIN | NS_F_MAXIT | NS_F_QUAD_ERROR | NS_F_NAN_ROWS | NS_F_QUAD_OVERFLOW */
    int n_outer, n_irls, n_obj;
    int deferred; /* FAST only: the unit needs the spline branch, nothing else is valid */
    ns_fit_counters cnt;
};

/* sig.rob.tukey(sig) (glm_rob_nb.r:161-165), summed over the group's samples */
#if defined(__CUDACC__)
#define NS_GFN __host__ __device__ __noinline__
#else
#define NS_GFN
#endif

/* ---- fast path (MODE 1): two samples per lane in flight ------------------------------------
 * The per-sample chains (rsqrt/log/exp seed, pmf recurrence) are long and dependent; ncu showed the
 * warp kernel waiting on fixed-latency dependencies (stall_wait 3 of 7 cycles per issue).  Each lane
 * therefore carries samples i and i+nl through the seed and one joint window loop. */
NS_HD double ns_tukeypsi_f(double r, double c, double inv_c)
{ /* tukeypsi with the division by c as a multiplication; NaN propagates */
    double t = r * inv_c;
    t = fma(t, t, -1.0);
    const double v = t * t * r;
    return (fabs(r) > c) ? 0.0 : v;
}

/* Tukey's biweight (t^2-1)^2 inside |t| <= 1 and 0 outside, from u = t^2-1: a positive u loses its
 * high word (u*u then underflows to 0) with two integer instructions, so the joint window loop of two
 * samples needs no end-of-window test.  At the window's ends (|t| = 1 up to rounding) the weight is 0
 * either way. */
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double ns_biweight(double u)
{
    int hi = __double2hiint(u);
    hi &= hi >> 31;
    u = __hiloint2double(hi, __double2loint(u));
    return u * u;
}
#define NS_UNROLL1 _Pragma("unroll 1")
#else
static inline double ns_biweight(double u) { return u <= 0.0 ? u * u : 0.0; }
#define NS_UNROLL1
#endif

struct ns_fseed {
    double mu, y, s1, sd, isd, pm, ics, t0;
    int J2;    /* last lattice point of the window {0..J2}; -1: nothing to sum here */
    bool slow; /* valid sample whose window is not of the {0..J2 < NS_JTAB} shape */
};
/* quantities shared by ai.sig.tukey and E.tukeypsi.1/.2 on a window that starts at 0:
 * sd = sqrt(V), isd = 1/sd, pm = dnbinom(0) = (1+sig mu)^(-1/sig), t_j = t0 + j*ics */
NS_HD void ns_fast_seed(ns_fseed &S, bool valid, double sig, double invsig, double c, double inv_c, const double *ltab,
                        double &lg)
{
    S.s1 = fma(sig, S.mu, 1.0);
    ns_fsqrt2(S.mu * S.s1, S.sd, S.isd);
    const double csd = c * S.sd;
    const double hi = S.mu + csd, lo = S.mu - csd;
    const bool shape = (lo <= 0.0) && (hi < (double)NS_JTAB); /* j1 == 0 and j2 < NS_JTAB */
    S.slow = valid && !shape;
    S.J2 = (valid && shape) ? ns_d2i_floor(hi) : -1;
    lg = ns_flog_tab(S.s1, ltab); /* s1 = 1 + sig mu >= 1 */
    S.pm = ns_fexp_bf(-invsig * lg);
    S.ics = S.isd * inv_c;
    S.t0 = -S.mu * S.ics;
}

/* a sample of the warp kernel whose window is not {0..J2<NS_JTAB}: generic lattice sum, or hand-over */
static NS_HDN double ns_obj_slow(double mu, double y, double sig, double c, double twon, double *dg0, int *have_dg0,
                                 int *need_heavy, unsigned *n_lat)
{
    const double invsig = 1 / sig;
    const double sd = sqrt(mu * (sig * mu + 1));
    const double j1 = fmax(ceil(mu - c * sd), 0.0), j2 = floor(mu + c * sd);
    double term;
    if (j1 > j2) {
        term = 0; /* cannot happen: kept for fidelity */
        double r = (y - mu) / sd;
        double wi = ns_tukeypsi(r, c) / r;
        if (wi != 0.0)
            term = NAN;
    } else if (j2 - j1 + 1 > twon) {
        *need_heavy = 1;
        term = 0;
    } else {
        if (!*have_dg0) {
            *dg0 = ns_digamma(invsig);
            *have_dg0 = 1;
        }
        double r = (y - mu) / sqrt(ns_varfunc(mu, sig));
        double wi = ns_tukeypsi(r, c) / r;
        term = 0.0;
        if (wi != 0.0) {
            double psiML = ns_digamma(r * sqrt(mu * (sig * mu + 1)) + mu + invsig) - sig * r * sqrt(mu / (sig * mu + 1)) -
                           *dg0 - log(sig * mu + 1);
            term = wi * psiML;
        }
        term -= ns_ai_sum(mu, sig, c, j1, j2, *dg0);
        *n_lat += (unsigned)(j2 - j1 + 1);
    }
    return term;
}

/* same for the IRLS pass: returns 0 when the sample needs the heavy kernel */
static NS_HDN int ns_irls_slow(double mu, double sig, double c, double twon, double *e1, double *sap, unsigned *n_lat)
{
    const double sV = sqrt(ns_varfunc(mu, sig));
    const double j1 = fmax(ceil(mu - c * sV), 0.0), j2 = floor(mu + c * sV);
    if (j1 > j2) {
        *e1 = 0;
        *sap = 0;
    } else if (j2 - j1 + 1 > twon) {
        return 0;
    } else {
        ns_E12_sum(mu, sig, c, j1, j2, *e1, *sap);
        *n_lat += (unsigned)(j2 - j1 + 1);
    }
    return 1;
}

#if defined(__CUDACC__)
/* Deferred-sample list of the cluster kernel's two-phase evaluation (per CTA, in w.scratch): threads append the
 * samples they leave to the warp-cooperative path in any order, ns_deflist_sort() orders the list by sample index so
 * that the assignment of samples to warps - and with it every floating-point sum - is the same in every run. */

struct ns_deflist {
    int *cnt, *raw, *srt;
    __device__ __forceinline__ ns_deflist(double *scratch, int n)
    {
        cnt = reinterpret_cast<int *>(scratch);
        raw = cnt + 2;
        srt = raw + n;
    }
    __device__ __forceinline__ void reset()
    {
        __syncthreads();
        if (threadIdx.x == 0)
            cnt[0] = 0;
        __syncthreads();
    }
    __device__ __forceinline__ void push(int i) { raw[atomicAdd(cnt, 1)] = i; }
    __device__ __forceinline__ int sort()
    {
        __syncthreads();
        const int L = cnt[0];
        for (int t = threadIdx.x; t < L; t += blockDim.x) {
            const int v = raw[t];
            int rank = 0;
            for (int u = 0; u < L; u++)
                rank += raw[u] < v;
            srt[rank] = v;
        }
        __syncthreads();
        return L;
    }
};
#endif

/* out of line on purpose: three call sites in the Brent driver; inlining them tripled the kernel's
 * code and the warps stalled on instruction fetch (ncu: 32 % of issue stalls) */
template <class G, int MODE>
NS_GFN double ns_sig_objective(double sig, const ns_work &w, int n, const ns_params &prm, ns_fit_counters &cnt,
                                     int &need_heavy)
{
    const double c = prm.c_tukey_sig;
    const double invsig = 1 / sig;
    double acc = 0;
    int qerr = 0;
    if (MODE == 1) {
        G::fill_dtab(w.dtab, w.r2tab, w.rtab, invsig, NS_JTAB);
        double dg0 = 0;
        int have_dg0 = 0;
        const double twon = (double)(2 * prm.n_ai_sig_tukey);
        const double inv_c = ns_frcp(c);
        const ns_sptr s_x(w.x), s_mu(w.mu), s_y(w.y), s_dtab(w.dtab), s_r2(w.r2tab);
        unsigned n_seed = 0, n_lat = 0; /* kept in registers, added to cnt once */
        const int nl = G::nl();
        double accB = 0;
        for (int iA = G::lane(); iA < n; iA += 2 * nl) {
            const int iB = (iA + nl < n) ? iA + nl : iA;
            const bool vA = s_x[iA] > 0, vB = (iB != iA) && (s_x[iB] > 0);
            ns_fseed A, B;
            A.mu = vA ? s_mu[iA] : 1.0;
            A.y = vA ? s_y[iA] : 0.0;
            B.mu = vB ? s_mu[iB] : 1.0;
            B.y = vB ? s_y[iB] : 0.0;
            double lgA, lgB;
            ns_fast_seed(A, vA, sig, invsig, c, inv_c, w.ltab, lgA);
            ns_fast_seed(B, vB, sig, invsig, c, inv_c, w.ltab, lgB);
            /* window {0..J2}: everything by recurrence, no division inside the loop.
             * t_j = (j-mu)/(c sd) = t0 + j/(c sd);  psi_j = D_j - lg - (j-mu)/(mu+1/sig) = D_j + C0 - j/(mu+1/sig),
             * 1/(mu+1/sig) = sig/(1+sig mu) = sig mu / V */
            const double impA = (sig * A.mu) * (A.isd * A.isd), impB = (sig * B.mu) * (B.isd * B.isd);
            const double qA = A.mu * impA, qB = B.mu * impB;
            /* a y beyond the window meets weight 0 (or no lattice point at all): tukeypsi(r) = 0 there */
            const int yiA = (A.y < (double)NS_JTAB) ? (int)A.y : NS_JTAB, yiB = (B.y < (double)NS_JTAB) ? (int)B.y : NS_JTAB;
            double DA = qA - lgA, DB = qB - lgB; /* D = digamma(j+1/sig) - digamma(1/sig) + C0 */
            double pmA = A.pm, pmB = B.pm, aA = 0, aB = 0, tyA = 0, tyB = 0, jd = 0.0;
            const int Jm = A.J2 > B.J2 ? A.J2 : B.J2;
            NS_UNROLL1
            for (int j = 0; j <= Jm; j++) {
                const double r2 = s_r2[j], dt = s_dtab[j];
                const double tA = fma(jd, A.ics, A.t0), tB = fma(jd, B.ics, B.t0);
                const double wA = ns_biweight(fma(tA, tA, -1.0)), wB = ns_biweight(fma(tB, tB, -1.0));
                const double wpsiA = wA * fma(-jd, impA, DA), wpsiB = wB * fma(-jd, impB, DB);
                aA = fma(wpsiA, pmA, aA);
                aB = fma(wpsiB, pmB, aB);
                tyA = (j == yiA) ? wpsiA : tyA; /* (psi_c(r)/r) * psi.sig.ML(r) at r = (y-mu)/sd */
                tyB = (j == yiB) ? wpsiB : tyB;
                pmA *= r2 * qA;
                pmB *= r2 * qB;
                DA += dt;
                DB += dt;
                jd += 1.0;
            }
            double termA = tyA - aA, termB = tyB - aB;
            if (A.y == A.mu)
                termA = NAN; /* r == 0: tukeypsi(r)/r = 0/0 in R */
            if (B.y == B.mu)
                termB = NAN;
            if (A.J2 < 0)
                termA = 0; /* no sample here, or one for ns_obj_slow */
            if (B.J2 < 0)
                termB = 0;
            n_seed += (unsigned)vA + (unsigned)vB;
            n_lat += (unsigned)(A.J2 + 1) + (unsigned)(B.J2 + 1);
            if (A.slow)
                termA = ns_obj_slow(A.mu, A.y, sig, c, twon, &dg0, &have_dg0, &need_heavy, &n_lat);
            if (B.slow)
                termB = ns_obj_slow(B.mu, B.y, sig, c, twon, &dg0, &have_dg0, &need_heavy, &n_lat);
            acc += termA;
            accB += termB;
        }
        acc += accB;
        cnt.n_seed += n_seed;
        cnt.n_lattice += n_lat;
    } else {
        const double dg0 = ns_digamma_sel<MODE == 2>(invsig);
        const double twon = (double)(2 * prm.n_ai_sig_tukey);
#if defined(__CUDA_ARCH__)
        /* throughput regime only: with few samples per warp (a unit on a full cluster) the lanes sharing one
         * window is the shorter critical path */
        const bool per_thread = (MODE == 2) && n >= NS_PT_FACTOR * G::nl();
        if (per_thread) {
            /* Windows that stay on the lattice-sum branch (<= 2 n_ai points; 94 % of the sample evaluations of a
             * deep panel, 25 points on average) are taken one sample per THREAD, serially: with the warp's lanes
             * sharing such a window every lane paid a pmf seed and a digamma for one or two lattice points.  The
             * spline + quadrature windows keep the warp-per-sample form below. */
            const int lid = threadIdx.x & 31;
            double accA = 0;
            int nsA = 0, nlA = 0;
            ns_deflist dl(w.scratch, n);
            dl.reset();
            /* consecutive samples go to consecutive CTAs of the cluster, so that every CTA gets its share of both
             * the per-thread sums and the deferred list */
            const int csz = G::nl() / G::lnl(), crk = G::lane() / G::lnl();
            for (int i = (G::llane() * 32 + lid) * csz + crk; i < n; i += G::nl() * 32) {
                if (!(w.x[i] > 0))
                    continue;
                const double mu = w.mu[i], y = w.y[i];
                const double sqrtV = sqrt(mu * (sig * mu + 1)); /* as in ai.sig.tukey */
                const double j1 = fmax(ceil(mu - c * sqrtV), 0.0), j2 = floor(mu + c * sqrtV);
                if (j2 - j1 + 1 > twon) {
                    dl.push(i); /* spline windows: the lanes of a warp share them */
                    continue;
                }
                const double sd = sqrt(ns_varfunc(mu, sig));
                const double r = (y - mu) / sd;
                const double wi = ns_tukeypsi(r, c) / r;
                double term = 0.0;
                if (wi != 0.0) {
                    double psiML = ns_digamma_sel<true>(r * sqrt(mu * (sig * mu + 1)) + mu + 1 / sig) -
                                   sig * r * sqrt(mu / (sig * mu + 1)) - dg0 - log(sig * mu + 1);
                    term = wi * psiML;
                }
                nsA++;
                if (!(j1 > j2)) {
                    nlA += (int)(j2 - j1 + 1);
                    term -= ns_ai_sum<true>(mu, sig, c, j1, j2, dg0);
                }
                accA += term;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                accA += __shfl_xor_sync(0xffffffffu, accA, o);
            acc = accA; /* uniform in the warp, like the terms added below */
            cnt.n_seed += (unsigned long long)__reduce_add_sync(0xffffffffu, nsA);
            cnt.n_lattice += (unsigned long long)__reduce_add_sync(0xffffffffu, nlA);
        }
#endif
        int n_iter = n, i_first = G::lane(), i_step = G::nl();
#if defined(__CUDA_ARCH__)
        ns_deflist dl(w.scratch, n);
        if (per_thread) { /* the deferred samples of this CTA, dealt to its warps in index order */
            n_iter = dl.sort();
            i_first = G::llane();
            i_step = G::lnl();
        }
#endif
        for (int it = i_first; it < n_iter; it += i_step) {
            int i = it;
#if defined(__CUDA_ARCH__)
            if (per_thread)
                i = dl.srt[it];
#endif
            if (!(w.x[i] > 0))
                continue;
            double mu = w.mu[i], y = w.y[i];
            double sd = sqrt(ns_varfunc(mu, sig));
            double r = (y - mu) / sd;
            double wi = ns_tukeypsi(r, c) / r; /* r == 0 -> NaN, as in R */
            double term;
            if (wi == 0.0) {
                term = 0.0; /* 0 * finite psi.sig.ML */
            } else {
                double psiML = ns_digamma_sel<MODE == 2>(r * sqrt(mu * (sig * mu + 1)) + mu + 1 / sig) -
                               sig * r * sqrt(mu / (sig * mu + 1)) - dg0 - log(sig * mu + 1);
                term = wi * psiML;
            }
            int qe0 = cnt.quad_err;
            cnt.quad_err = 0;
#if defined(__CUDA_ARCH__)
            if (MODE == 2)
                term -= ns_ai_coop<G::rolled>(mu, sig, c, prm.n_ai_sig_tukey, dg0, w.coop, cnt);
            else
#endif
                term -= ns_ai_sig_tukey(mu, sig, c, prm.n_ai_sig_tukey, dg0, cnt);
            qerr |= cnt.quad_err;
            cnt.quad_err = qe0;
            acc += term;
        }
    }
    acc = G::sum(acc);
    if (G::any(qerr))
        return NAN; /* integrate() error inside uniroot -> tryCatch -> NA (:175-177) */
    return acc;
}

/* uniroot(sig.rob.tukey, c(minsig,maxsig), tol=sig.prec, maxiter=maxit.sig)$root: R's
 * end-point checks followed by R_zeroin2 (Brent).  Returns false on any R error. */
template <class G, int MODE>
NS_GF inline bool ns_sigma_root(const ns_work &w, int n, const ns_params &prm, ns_fit_result &res, double &root,
                                int &need_heavy)
{
    double a = prm.minsig, b = prm.maxsig;
    double fa = ns_sig_objective<G, MODE>(a, w, n, prm, res.cnt, need_heavy);
    double fb = ns_sig_objective<G, MODE>(b, w, n, prm, res.cnt, need_heavy);
    res.n_obj += 2;
    if (MODE == 1 && G::any(need_heavy))
        return false;
    if (fa != fa || fb != fb)
        return false;
    if (fa * fb > 0)
        return false;
    double c = a, fc = fa;
    if (fa == 0.0) {
        root = a;
        return true;
    }
    if (fb == 0.0) {
        root = b;
        return true;
    }
    int maxit = prm.maxit_sig + 1;
    const double tol = prm.sig_prec;
    while (maxit--) {
        double prev_step = b - a;
        if (fabs(fc) < fabs(fb)) {
            a = b;
            b = c;
            c = a;
            fa = fb;
            fb = fc;
            fc = fa;
        }
        double tol_act = 2 * NS_DBL_EPS * fabs(b) + tol / 2;
        double new_step = (c - b) / 2;
        if (fabs(new_step) <= tol_act || fb == 0.0) {
            root = b;
            return true;
        }
        if (fabs(prev_step) >= tol_act && fabs(fa) > fabs(fb)) {
            double p, q, cb = c - b;
            if (a == c) {
                double t1 = fb / fa;
                p = cb * t1;
                q = 1.0 - t1;
            } else {
                double t1, t2;
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
        if (fabs(new_step) < tol_act)
            new_step = (new_step > 0.0) ? tol_act : -tol_act;
        a = b;
        fa = fb;
        b += new_step;
        fb = ns_sig_objective<G, MODE>(b, w, n, prm, res.cnt, need_heavy);
        res.n_obj += 1;
        if (MODE == 1 && G::any(need_heavy))
            return false;
        if (fb != fb || fb == INFINITY)
            return false; /* "invalid function value in 'zeroin'" */
        if (fb == -INFINITY)
            fb = -NS_DBL_MAX;
        if ((fb > 0 && fc > 0) || (fb < 0 && fc < 0)) {
            c = a;
            fc = fa;
        }
    }
    root = b; /* maxiter reached: a warning in R, the current b is returned */
    return true;
}

/* abs(max(u - v)) with R's recycling, for u, v of length 1 or 2 (second entry of a
 * length-2 vector is the constant 1 of c(coef(wls), 1); glm_rob_nb.r:171,187,195) */
NS_HD double ns_D_absmax(double u0, int ulen, double v0, int vlen)
{
    double d0 = u0 - v0;
    double d1 = (ulen == 2 ? 1.0 : u0) - (vlen == 2 ? 1.0 : v0);
    if (ulen == 1 && vlen == 1)
        return fabs(d0);
    if (d0 != d0 || d1 != d1)
        return NAN;
    return fabs(fmax(d0, d1));
}

/* one IRLS pass (glm_rob_nb.r:188-200): returns the new beta, updates mu/eta */
template <class G, int MODE>
NS_GFN double ns_irls_step(double sig, const ns_work &w, int n, double maxmu, const ns_params &prm,
                                 ns_fit_result &res, int &need_heavy)
{
    const int lane = G::lane();
    const double c = prm.c_tukey_beta;
    const double invsig = 1 / sig;
    const double twon = (double)(2 * prm.n_ai_sig_tukey);
    double sw = 0, swz = 0;
    int nanrows = 0;
    const ns_sptr s_x(w.x), s_mu(w.mu), s_y(w.y), s_eta(w.eta), s_X(w.X), s_r2(w.r2tab);
    unsigned n_seed = 0, n_lat = 0;
    if (MODE == 1) {
        /* two samples per lane in flight (see ns_fast_seed).  With y1 = 1/sqrt(V): 1/(V sqrt V) = y1^3,
         * mu/(mu+1/sig) = sig mu^2 y1^2, V/mu = 1 + sig mu */
        const double inv_c = ns_frcp(c);
        const int nl = G::nl();
        for (int iA = lane; iA < n; iA += 2 * nl) {
            const int iB = (iA + nl < n) ? iA + nl : iA;
            const bool vA = s_x[iA] > 0, vB = (iB != iA) && (s_x[iB] > 0);
            ns_fseed A, B;
            A.mu = vA ? s_mu[iA] : 1.0;
            A.y = vA ? s_y[iA] : 0.0;
            B.mu = vB ? s_mu[iB] : 1.0;
            B.y = vB ? s_y[iB] : 0.0;
            double lgA, lgB;
            ns_fast_seed(A, vA, sig, invsig, c, inv_c, w.ltab, lgA);
            ns_fast_seed(B, vB, sig, invsig, c, inv_c, w.ltab, lgB);
            const double qA = (sig * A.mu * A.mu) * (A.isd * A.isd), qB = (sig * B.mu * B.mu) * (B.isd * B.isd);
            double pmA = A.pm, pmB = B.pm, a1A = 0, a2A = 0, a1B = 0, a2B = 0, jd = 0.0;
            const int Jm = A.J2 > B.J2 ? A.J2 : B.J2;
            NS_UNROLL1
            for (int j = 0; j <= Jm; j++) {
                const double r2 = s_r2[j];
                const double tA = fma(jd, A.ics, A.t0), tB = fma(jd, B.ics, B.t0);
                const double wA = ns_biweight(fma(tA, tA, -1.0)), wB = ns_biweight(fma(tB, tB, -1.0));
                const double dA = jd - A.mu, dB = jd - B.mu;
                const double wpdA = (wA * pmA) * dA, wpdB = (wB * pmB) * dB;
                a1A += wpdA;
                a2A = fma(wpdA, dA, a2A);
                a1B += wpdB;
                a2B = fma(wpdB, dB, a2B);
                pmA *= r2 * qA;
                pmB *= r2 * qB;
                jd += 1.0;
            }
            double e1A = a1A * A.isd, sapA = a2A * A.isd, e1B = a1B * B.isd, sapB = a2B * B.isd;
            n_seed += (unsigned)vA + (unsigned)vB;
            n_lat += (unsigned)(A.J2 + 1) + (unsigned)(B.J2 + 1);
            bool useA = vA, useB = vB;
            if (A.slow && !ns_irls_slow(A.mu, sig, c, twon, &e1A, &sapA, &n_lat)) {
                need_heavy = 1;
                useA = false;
            }
            if (B.slow && !ns_irls_slow(B.mu, sig, c, twon, &e1B, &sapB, &n_lat)) {
                need_heavy = 1;
                useB = false;
            }
            /* derivinvlink(eta) = exp(log(mu)) = mu */
            const double biA = sapA * (A.isd * A.isd * A.isd) * (A.mu * A.mu);
            const double biB = sapB * (B.isd * B.isd * B.isd) * (B.mu * B.mu);
            const double eiA = (ns_tukeypsi_f((A.y - A.mu) * A.isd, c, inv_c) - e1A) * ns_frcp(sapA) * A.s1;
            const double eiB = (ns_tukeypsi_f((B.y - B.mu) * B.isd, c, inv_c) - e1B) * ns_frcp(sapB) * B.s1;
            if (useA) {
                const double zi = s_eta[iA] + eiA;
                if (biA != biA || zi != zi) {
                    nanrows++; /* lm(): na.omit drops the row */
                } else if (biA != 0) { /* lm.wfit(): zero-weight rows excluded */
                    sw += biA;
                    swz += biA * (zi - s_X[iA]);
                }
            }
            if (useB) {
                const double zi = s_eta[iB] + eiB;
                if (biB != biB || zi != zi) {
                    nanrows++;
                } else if (biB != 0) {
                    sw += biB;
                    swz += biB * (zi - s_X[iB]);
                }
            }
        }
    } else {
#if defined(__CUDA_ARCH__)
        const bool per_thread = (MODE == 2) && n >= NS_PT_FACTOR * G::nl();
        if (per_thread) {
            /* lattice-sum windows: one sample per thread (see ns_sig_objective) */
            const int lid = threadIdx.x & 31;
            double swA = 0, swzA = 0;
            int nsA = 0, nlA = 0;
            ns_deflist dl(w.scratch, n);
            dl.reset();
            const int csz = G::nl() / G::lnl(), crk = G::lane() / G::lnl();
            for (int i = (G::llane() * 32 + lid) * csz + crk; i < n; i += G::nl() * 32) {
                if (!(s_x[i] > 0))
                    continue;
                const double mu = s_mu[i], eta = s_eta[i], y = s_y[i];
                const double V = ns_varfunc(mu, sig);
                const double sV = sqrt(V);
                const double j1 = fmax(ceil(mu - c * sV), 0.0), j2 = floor(mu + c * sV);
                if (j2 - j1 + 1 > twon) {
                    dl.push(i);
                    continue;
                }
                double e1 = 0, sap = 0;
                nsA++;
                if (!(j1 > j2)) {
                    nlA += (int)(j2 - j1 + 1);
                    ns_E12_sum<true>(mu, sig, c, j1, j2, e1, sap);
                }
                const double ee = exp(eta);
                const double bi = sap * (1.0 / (V * sV)) * (ee * ee);
                const double ei = (ns_tukeypsi((y - mu) / sV, c) - e1) / sap * V * (1 / mu);
                const double zi = eta + ei;
                if (bi != bi || zi != zi) {
                    nanrows++;
                    continue;
                }
                if (bi == 0)
                    continue;
                swA += bi;
                swzA += bi * (zi - s_X[i]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                swA += __shfl_xor_sync(0xffffffffu, swA, o);
                swzA += __shfl_xor_sync(0xffffffffu, swzA, o);
            }
            sw = swA;
            swz = swzA;
            nanrows = __any_sync(0xffffffffu, nanrows);
            n_seed += (unsigned)__reduce_add_sync(0xffffffffu, nsA);
            n_lat += (unsigned)__reduce_add_sync(0xffffffffu, nlA);
        }
#endif
        int n_iter = n, i_first = lane, i_step = G::nl();
#if defined(__CUDA_ARCH__)
        ns_deflist dl(w.scratch, n);
        if (per_thread) {
            n_iter = dl.sort();
            i_first = G::llane();
            i_step = G::lnl();
        }
#endif
        for (int it = i_first; it < n_iter; it += i_step) {
            int i = it;
#if defined(__CUDA_ARCH__)
            if (per_thread)
                i = dl.srt[it];
#endif
            if (!(s_x[i] > 0))
                continue;
            const double mu = s_mu[i], eta = s_eta[i], y = s_y[i];
            double e1, sap;
            const double V = ns_varfunc(mu, sig);
            const double sV = sqrt(V);
#if defined(__CUDA_ARCH__)
            if (MODE == 2) {
                ns_E12_coop<G::rolled>(mu, sig, c, prm.n_ai_sig_tukey, w.coop, res.cnt, &e1, &sap);
            } else
#endif
                ns_E_tukeypsi_12(mu, sig, c, prm.n_ai_sig_tukey, res.cnt, e1, sap);
            double ee = exp(eta); /* derivinvlink(eta) = exp(log(mu)) */
            double bi = sap * (1.0 / (V * sV)) * (ee * ee);
            double ei = (ns_tukeypsi((y - mu) / sV, c) - e1) / sap * V * (1 / mu);
            double zi = eta + ei;
            if (bi != bi || zi != zi) {
                nanrows++; /* lm(): na.omit drops the row */
                continue;
            }
            if (bi == 0)
                continue; /* lm.wfit(): zero-weight rows excluded */
            sw += bi;
            swz += bi * (zi - s_X[i]);
        }
    }
    res.cnt.n_seed += n_seed;
    res.cnt.n_lattice += n_lat;
    sw = G::sum(sw);
    swz = G::sum(swz);
    if (G::any(nanrows))
        res.flags |= NS_F_NAN_ROWS;
    const double beta = swz / sw;
    G::lsync();
    for (int i = G::llane(); i < n; i += G::lnl()) { /* every replica updates all samples */
        if (!(w.x[i] > 0))
            continue;
        double mu = (MODE == 1) ? ns_fexp(w.X[i] + beta) : exp(w.X[i] + beta);
        if (mu > maxmu)
            mu = maxmu;
        if (mu == 0)
            mu = prm.minmu;
        w.mu[i] = mu;
        w.eta[i] = (MODE == 1) ? ns_flog(mu) : log(mu);
    }
    G::lsync();
    return beta;
}

#if defined(NS_FIT_TRACE) && !defined(__CUDA_ARCH__)
static thread_local double *ns_fit_trace_sig = nullptr, *ns_fit_trace_beta = nullptr;
static thread_local int ns_fit_trace_cap = 0;
#endif

/* The fit proper: glm_rob_nb.r:58-205 on the samples with w.x[i] > 0.
 * Expects w.y / w.x filled (x <= 0 for samples outside the fit); maxmu = max(x) over ALL
 * samples of the unit (:20). */
template <class G, int MODE>
NS_GF inline void ns_fit_unit(const ns_work &w, int n, double maxmu, const ns_params &prm, ns_fit_result &res)
{
    const int lane = G::llane(); /* the initial estimates are computed by every replica */
    res.flags = 0;
    res.n_outer = res.n_irls = res.n_obj = 0;
    res.deferred = 0;
    res.sigma = res.slope = NAN;
    res.cnt.n_seed = res.cnt.n_lattice = res.cnt.n_quad = 0;
    res.cnt.quad_err = res.cnt.quad_overflow = 0;
    /* :58-63, :75-76 */
    int nfit = 0;
    double sum_ex = 0;
    for (int i = lane; i < n; i += G::lnl()) {
        double x = w.x[i];
        if (x > 0) {
            double X = log(x);
            w.X[i] = X;
            double ex = exp(X);
            double y = w.y[i];
            w.eta[i] = (y == 0) ? 0.0 : y / ex; /* af, parked in eta until mu is known */
            sum_ex += ex;
            nfit++;
        }
    }
    nfit = G::lisum(nfit);
    sum_ex = G::lsum(sum_ex);
    G::lsync();
    /* :77 type-7 quartiles of af by rank counting */
    const double h25 = (double)(nfit - 1) * 0.25, h75 = (double)(nfit - 1) * 0.75;
    const int lo25 = (int)floor(h25), hi25 = (int)ceil(h25), lo75 = (int)floor(h75), hi75 = (int)ceil(h75);
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    int r_first = lane, r_step = G::lnl();
#if defined(__CUDA_ARCH__)
    if (MODE == 2) { /* the replicas of a warp would all count the same ranks: one sample per thread instead */
        r_first = threadIdx.x;
        r_step = blockDim.x;
    }
#endif
    for (int i = r_first; i < n; i += r_step) {
        if (!(w.x[i] > 0))
            continue;
        double ai = w.eta[i];
        int rank = 0;
        for (int j = 0; j < n; j++) {
            double aj = w.eta[j];
            rank += (w.x[j] > 0) && (aj < ai || (aj == ai && j < i));
        }
        if (rank == lo25)
            v0 = ai;
        if (rank == hi25)
            v1 = ai;
        if (rank == lo75)
            v2 = ai;
        if (rank == hi75)
            v3 = ai;
    }
#if defined(__CUDA_ARCH__)
    if (MODE == 2) { /* the collectives below take one value per warp */
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
            v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            v3 += __shfl_xor_sync(0xffffffffu, v3, o);
        }
    }
#endif
    v0 = G::lsum(v0);
    v1 = G::lsum(v1);
    v2 = G::lsum(v2);
    v3 = G::lsum(v3);
    double q25 = v0, q75 = v2;
    if (h25 > (double)lo25 && v1 != v0) {
        double h = h25 - (double)lo25;
        q25 = (1 - h) * v0 + h * v1;
    }
    if (h75 > (double)lo75 && v3 != v2) {
        double h = h75 - (double)lo75;
        q75 = (1 - h) * v2 + h * v3;
    }
    const double fence = q75 + 1.5 * (q75 - q25);
    double s_af = 0;
    int c_af = 0;
    for (int i = lane; i < n; i += G::lnl())
        if (w.x[i] > 0 && w.eta[i] <= fence) {
            s_af += w.eta[i];
            c_af++;
        }
    s_af = G::lsum(s_af);
    c_af = G::lisum(c_af);
    double inislope = s_af / (double)c_af;
    {
        double mean_ex = sum_ex / (double)nfit;
        if (inislope <= 1 / (10 * mean_ex))
            inislope = 1 / (10 * mean_ex); /* :78 */
    }
    G::lsync();
    /* :79-86 */
    double s_sig = 0;
    for (int i = lane; i < n; i += G::lnl()) {
        if (!(w.x[i] > 0))
            continue;
        double mu = inislope * exp(w.X[i]);
        if (mu > maxmu)
            mu = maxmu;
        if (mu < prm.minmu)
            mu = prm.minmu;
        w.mu[i] = mu;
        w.eta[i] = log(mu);
        double t = w.y[i] / mu - 1;
        s_sig += t * t;
    }
    s_sig = G::lsum(s_sig);
    double sig = s_sig / (double)nfit;
    if (sig > prm.maxsig)
        sig = prm.maxsig;
    if (sig < prm.minsig)
        sig = prm.minsig;
    G::lsync();

    /* :166-205 alternation */
    const double tol = prm.tol;
    double sig0 = sig + 1 + tol;
    double beta11 = log(inislope);
    int len11 = 1;
    double beta00 = beta11 + tol + 1;
    int len00 = 1;
    double beta1 = beta11;
    int len1 = 1;
    int it = 0;
    int last_sig_failed = 0;
    int need_heavy = 0;
    while ((fabs(sig - sig0) > tol || ns_D_absmax(beta11, len11, beta00, len00) > tol) && it < prm.maxit) {
        sig0 = sig;
        beta00 = beta11;
        len00 = len11;
        /* :175-182 */
        double root = NAN;
        bool ok = ns_sigma_root<G, MODE>(w, n, prm, res, root, need_heavy);
        if (MODE == 1 && G::any(need_heavy)) {
            res.deferred = 1;
            return;
        }
        if (!ok || root != root) {
            sig = prm.minsig;
            last_sig_failed = 1;
        } else {
            sig = root;
            last_sig_failed = 0;
            if (sig > prm.maxsig)
                sig = prm.maxsig;
            if (sig < prm.minsig)
                sig = prm.minsig;
        }
        /* :184-202 */
        beta1 = log(inislope);
        len1 = 1;
        double beta0 = beta1 + tol + 1;
        int len0 = 1;
        int it_mu = 0;
        if (MODE == 1)
            G::fill_r2tab(w.r2tab, w.rtab, 1 / sig, NS_JTAB); /* sigma is fixed over the IRLS passes */
        while (ns_D_absmax(beta1, len1, beta0, len0) > tol && it_mu < prm.maxit) {
            beta0 = beta1;
            len0 = len1;
            beta1 = ns_irls_step<G, MODE>(sig, w, n, maxmu, prm, res, need_heavy);
            len1 = 2;
            if (MODE == 1 && G::any(need_heavy)) {
                res.deferred = 1;
                return;
            }
            it_mu++;
            res.n_irls++;
        }
        if (G::any(res.cnt.quad_err))
            res.flags |= NS_F_QUAD_ERROR; /* R would have stopped the script */
        res.cnt.quad_err = 0;
        beta11 = beta1;
        len11 = len1;
#if defined(NS_FIT_TRACE) && !defined(__CUDA_ARCH__)
        if (ns_fit_trace_sig && it < ns_fit_trace_cap) { /* tests/emul only: the iterates of the alternation */
            ns_fit_trace_sig[it] = sig;
            ns_fit_trace_beta[it] = beta1;
        }
#endif
        it++;
    }
    if (G::any(res.cnt.quad_overflow))
        res.flags |= NS_F_QUAD_OVERFLOW;
    res.n_outer = it;
    if (it >= prm.maxit)
        res.flags |= NS_F_MAXIT;
    if (last_sig_failed)
        res.flags |= NS_F_SIG_AT_MIN;
    double slope = exp(beta1);
    if (slope > 1)
        slope = 1;
    res.sigma = sig;
    res.slope = slope;
}
