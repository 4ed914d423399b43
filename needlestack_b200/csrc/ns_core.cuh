/*
 * ns_core.cuh — the per-(position, allele) error-model unit, written once for a
 * "group" of cooperating lanes: a 32-lane warp on the B200 (WarpG, shuffles) and a
 * single lane in the CPU emulation that tests/ builds without a GPU (SerialG).
 *
 * Mapping: one group = one unit (one glmrob.nb() call of the reference);
 * lanes stride over the samples, per-sample state lives in the group's slice of
 * shared memory, every cross-sample sum is a warp-shuffle all-reduce, and the
 * sequential control flow (Brent, IRLS, the alternating outer loop) is executed
 * redundantly and identically by all lanes.
 *
 * Reference:  /root/reference/bin/glm_rob_nb.r:16-228  (glmrob.nb)
 *             /root/reference/bin/needlestack.r:321-409 (per-allele body)
 */
#pragma once
#include "ns_math.cuh"
#include "ns_quad.cuh"
#include "../../include/ns_b200.h"

/* ---- lane groups ----------------------------------------------------------------- */
#if defined(__CUDACC__)
struct WarpG {
    static constexpr int W = 32;
    __device__ static __forceinline__ int lane() { return threadIdx.x & 31; }
    __device__ static __forceinline__ double sum(double v)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ static __forceinline__ double max(double v)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    __device__ static __forceinline__ int isum(int v) { return __reduce_add_sync(0xffffffffu, v); }
    __device__ static __forceinline__ int imax(int v) { return __reduce_max_sync(0xffffffffu, v); }
    __device__ static __forceinline__ int any(int p) { return __any_sync(0xffffffffu, p); }
    __device__ static __forceinline__ void sync() { __syncwarp(); }
    /* inclusive scans across lanes */
    __device__ static __forceinline__ double scan_min(double v)
    {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane() >= o)
                v = fmin(v, t);
        }
        return v;
    }
    __device__ static __forceinline__ double scan_max(double v)
    {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane() >= o)
                v = fmax(v, t);
        }
        return v;
    }
    __device__ static __forceinline__ double bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
    /* min of v over the lanes above this one (+Inf for the top lane) */
    __device__ static __forceinline__ double excl_suffix_min(double v)
    {
        double rv = __shfl_sync(0xffffffffu, v, 31 - lane());
        rv = scan_min(rv);
        double ex = __shfl_up_sync(0xffffffffu, rv, 1);
        if (lane() == 0)
            ex = INFINITY;
        return __shfl_sync(0xffffffffu, ex, 31 - lane());
    }
    /* max of v over the lanes below this one (-Inf for lane 0) */
    __device__ static __forceinline__ double excl_prefix_max(double v)
    {
        double inc = scan_max(v);
        double ex = __shfl_up_sync(0xffffffffu, inc, 1);
        return lane() == 0 ? -INFINITY : ex;
    }
};
#define NS_GD __device__
#else
#define NS_GD
#endif

struct SerialG {
    static constexpr int W = 1;
    static NS_HD int lane() { return 0; }
    static NS_HD double sum(double v) { return v; }
    static NS_HD double max(double v) { return v; }
    static NS_HD int isum(int v) { return v; }
    static NS_HD int imax(int v) { return v; }
    static NS_HD int any(int p) { return p; }
    static NS_HD void sync() {}
    static NS_HD double scan_min(double v) { return v; }
    static NS_HD double scan_max(double v) { return v; }
    static NS_HD double bcast(double v, int) { return v; }
    static NS_HD double excl_suffix_min(double) { return INFINITY; }
    static NS_HD double excl_prefix_max(double) { return -INFINITY; }
};

#if defined(__CUDACC__)
#define NS_GF __host__ __device__
#else
#define NS_GF
#endif

/* ---- where a unit's five count rows live ----------------------------------------- */
struct ns_unit_src {
    int mode;               /* 0: SNV planes [P][9][n] + ref[P];  1: explicit rows [U][n] */
    int n;                  /* samples */
    const int32_t *counts;  /* mode 0 */
    const uint8_t *ref;     /* mode 0: 0..3 = A,T,C,G (needlestack.r:188-198 column order) */
    const int32_t *dp, *vp, *vm, *rp, *rm; /* mode 1; vm/rp/rm may be null (zeros) */
    const uint8_t *is_indel; /* mode 1, may be null */
    int plain;              /* mode 1: glmrob.nb(y = vp, x = dp) without the ref-inversion test */
};

struct ns_rows {
    const int32_t *dp, *vp, *vm, *rp, *rm;
    int is_indel;
    int valid; /* 0: position skipped (REF not in A,T,C,G; needlestack.r:319) */
};

/* alt allele of SNV unit u = 3*p + k: k-th element of setdiff(c("A","T","C","G"), ref) */
NS_HD ns_rows ns_unit_rows(const ns_unit_src &s, int64_t u)
{
    ns_rows r;
    if (s.mode == 0) {
        int64_t p = u / 3;
        int k = (int)(u - 3 * p);
        int ref = s.ref[p];
        r.valid = ref < 4;
        if (!r.valid)
            ref = 0;
        int alt = k + (k >= ref ? 1 : 0);
        const int32_t *base = s.counts + p * 9 * (int64_t)s.n;
        r.dp = base;
        r.vp = base + (1 + alt) * (int64_t)s.n;
        r.vm = base + (5 + alt) * (int64_t)s.n;
        r.rp = base + (1 + ref) * (int64_t)s.n;
        r.rm = base + (5 + ref) * (int64_t)s.n;
        r.is_indel = 0;
    } else {
        int64_t off = u * (int64_t)s.n;
        r.valid = 1;
        r.dp = s.dp + off;
        r.vp = s.vp + off;
        r.vm = s.vm ? s.vm + off : nullptr;
        r.rp = s.rp ? s.rp + off : nullptr;
        r.rm = s.rm ? s.rm + off : nullptr;
        r.is_indel = s.is_indel ? s.is_indel[u] : 0;
    }
    return r;
}
NS_HD double ns_ld(const int32_t *p, int i) { return p ? (double)p[i] : 0.0; }

/* ---- gate: ref-inversion, extra-robust pre-filter, the NA gate ------------------- */
/* needlestack.r:330-337 + glm_rob_nb.r:20-49.  Returns unit flags
 * (NS_F_REF_INV | NS_F_EXTRA_ROB | NS_F_GATED | NS_F_SKIPPED).
 * scratch: >= n doubles of group-shared memory. */
template <class G>
NS_GF inline uint32_t ns_gate_unit(const ns_rows &r, int n, const ns_params &prm, int plain, double *scratch)
{
    const int lane = G::lane();
    if (!r.valid)
        return NS_F_SKIPPED | NS_F_GATED;
    uint32_t flags = 0;
    /* ref inversion: sum(ma/DP > 0.8, na.rm=T) > 0.5*n  (NaN compares false) */
    int cnt = 0;
    if (!plain) {
        for (int i = lane; i < n; i += G::W) {
            double dp = ns_ld(r.dp, i), ma = ns_ld(r.vp, i) + ns_ld(r.vm, i);
            if (ma / dp > 0.8)
                cnt++;
        }
        cnt = G::isum(cnt);
        if ((double)cnt > 0.5 * (double)n)
            flags |= NS_F_REF_INV;
    }
    const bool inv = flags & NS_F_REF_INV;
    /* extra-robust candidate set S = { y/x > min_af_extra_rob } */
    int nS = 0;
    if (prm.extra_rob) {
        for (int i = lane; i < n; i += G::W) {
            double x = ns_ld(r.dp, i);
            double y = inv ? ns_ld(r.rp, i) + ns_ld(r.rm, i) : ns_ld(r.vp, i) + ns_ld(r.vm, i);
            if (y / x > prm.min_af_extra_rob)
                nS++;
        }
        nS = G::isum(nS);
    }
    bool extra = prm.extra_rob && (double)nS > prm.min_prop_extra_rob * (double)n &&
                 (double)nS < prm.max_prop_extra_rob * (double)n && (n - nS) > prm.size_min;
    if (extra)
        flags |= NS_F_EXTRA_ROB;
    const double min_reads = extra ? 0.0 : prm.min_reads, min_af = extra ? 0.0 : prm.min_af;
    /* cheap gate terms over the kept samples */
    int n0 = 0, n_cov = 0, n_pos = 0;
    double maxy = -INFINITY, maxaf = -INFINITY;
    for (int i = lane; i < n; i += G::W) {
        double x = ns_ld(r.dp, i);
        double y = inv ? ns_ld(r.rp, i) + ns_ld(r.rm, i) : ns_ld(r.vp, i) + ns_ld(r.vm, i);
        double a = y / x;
        bool keep = !(extra && a > prm.min_af_extra_rob);
        if (keep) {
            n0++;
            n_cov += x > prm.min_coverage;
            n_pos += y > 0;
            maxy = fmax(maxy, y);
            if (a == a)
                maxaf = fmax(maxaf, a);
        }
    }
    n0 = G::isum(n0);
    n_cov = G::isum(n_cov);
    n_pos = G::isum(n_pos);
    maxy = G::max(maxy);
    maxaf = G::max(maxaf);
    bool gate = (n0 == 0) || (n_cov < prm.size_min) || (maxy < min_reads) || (n_pos == 0 && !extra) ||
                (maxaf < min_af);
    if (!gate) {
        /* median(x) < min_coverage over the kept samples: the two middle order statistics
         * by rank counting (integers; ties broken by index) */
        G::sync();
        for (int i = lane; i < n; i += G::W) {
            double x = ns_ld(r.dp, i);
            double y = inv ? ns_ld(r.rp, i) + ns_ld(r.rm, i) : ns_ld(r.vp, i) + ns_ld(r.vm, i);
            bool keep = !(extra && y / x > prm.min_af_extra_rob);
            scratch[i] = keep ? x : -1.0; /* depths are >= 0 */
        }
        G::sync();
        const int k_lo = (n0 - 1) / 2, k_hi = n0 / 2;
        double vlo = 0, vhi = 0;
        for (int i = lane; i < n; i += G::W) {
            double xi = scratch[i];
            if (xi < 0)
                continue;
            int rank = 0;
            for (int j = 0; j < n; j++) {
                double xj = scratch[j];
                rank += (xj >= 0) && (xj < xi || (xj == xi && j < i));
            }
            if (rank == k_lo)
                vlo = xi;
            if (rank == k_hi)
                vhi = xi;
        }
        vlo = G::sum(vlo); /* exactly one lane contributes */
        vhi = G::sum(vhi);
        double med = (k_lo == k_hi) ? vlo : (vlo + vhi) / 2;
        G::sync();
        gate = med < prm.min_coverage;
    }
    if (gate)
        flags |= NS_F_GATED;
    return flags;
}

/* ---- window expectations ---------------------------------------------------------- */
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

/* ai.sig.tukey(mui, sig, c) (glm_rob_nb.r:141-160) and, when yi lies in the window,
 * psi.sig.ML at yi through the same digamma recurrence (:101-103).
 * dg0 = digamma(1/sig) is per-sigma and supplied by the caller. */
NS_GF inline double ns_ai_sig_tukey(double mui, double sig, double c, int n_ai, double dg0,
                                    ns_fit_counters &cnt)
{
    const double sqrtV = sqrt(mui * (sig * mui + 1));
    const double invsig = 1 / sig;
    const double j1 = fmax(ceil(mui - c * sqrtV), 0.0);
    const double j2 = floor(mui + c * sqrtV);
    cnt.n_seed += 1;
    if (j1 > j2)
        return 0.0;
    const double lg = log(mui / invsig + 1);
    const double csd = c * sqrtV, mpi = mui + invsig;
    if (j2 - j1 + 1 > 2 * n_ai) {
        ns_spline s;
        ns_spline_init(s, n_ai, j1, j2);
        for (int k = 0; k < n_ai; k++) {
            double j = ns_knot(s, k);
            double t = (j - mui) / csd;
            double w = (t * t - 1) * (t * t - 1);
            double psi = ns_digamma(j + invsig) - dg0 - lg - (j - mui) / mpi;
            s.y[k] = w * psi * ns_dnbinom_mu(j, invsig, mui);
        }
        cnt.n_lattice += (unsigned long long)n_ai;
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
    /* lattice sum, pmf and digamma difference by recurrence from j1 */
    double pm = ns_dnbinom_mu(j1, invsig, mui);
    double D = (j1 == 0) ? 0.0 : ns_digamma(j1 + invsig) - dg0; /* digamma(j+1/sig)-digamma(1/sig) */
    const double q = mui / (mui + invsig);
    double acc = 0;
    for (double j = j1; j <= j2; j += 1) {
        double d = j - mui;
        double t = d / csd;
        double w = (t * t - 1) * (t * t - 1);
        double psi = D - lg - d / mpi;
        acc += w * psi * pm;
        double ja = j + invsig, jb = j + 1;
        double rr = 1.0 / (ja * jb);
        pm *= (ja * ja) * rr * q; /* pmf(j+1)/pmf(j) = (j+size)/(j+1) * mu/(mu+size) */
        D += jb * rr;             /* digamma(x+1) = digamma(x) + 1/x */
    }
    cnt.n_lattice += (unsigned long long)(j2 - j1 + 1);
    return acc;
}

/* E.tukeypsi.1 and E.tukeypsi.2 (glm_rob_nb.r:109-140) in one pass over the window */
NS_GF inline void ns_E_tukeypsi_12(double mui, double sig, double c, int n_ai, ns_fit_counters &cnt,
                                   double &e1, double &e2)
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
    const double size = 1 / sig, csd = c * sqrtV;
    if (j2 - j1 + 1 > 2 * n_ai) {
        ns_spline s;
        ns_spline_init(s, n_ai, j1, j2);
        for (int k = 0; k < n_ai; k++) {
            double j = ns_knot(s, k);
            double t = (j - mui) / csd;
            double w = (t * t - 1) * (t * t - 1);
            s.y[k] = w * (j - mui) * ns_dnbinom_mu(j, size, mui);
        }
        cnt.n_lattice += (unsigned long long)n_ai;
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
        return;
    }
    double pm = ns_dnbinom_mu(j1, size, mui);
    const double q = mui / (mui + size);
    double a1 = 0, a2 = 0;
    for (double j = j1; j <= j2; j += 1) {
        double d = j - mui;
        double t = d / csd;
        double w = (t * t - 1) * (t * t - 1);
        double wp = w * pm;
        a1 += wp * d;
        a2 += wp * (d * d);
        pm *= ((j + size) / (j + 1)) * q;
    }
    cnt.n_lattice += (unsigned long long)(j2 - j1 + 1);
    e1 = a1 / sqrtV;
    e2 = a2 / sqrtV;
}

/* ---- per-group working set (slices of shared memory) ----------------------------- */
struct ns_work {
    double *y, *x, *X, *mu, *eta; /* n each; x <= 0 marks a sample that is not in the fit */
    double *scratch;              /* n + 8 */
};
NS_HD size_t ns_work_doubles(int n) { return (size_t)6 * (size_t)n + 8; }
NS_HD void ns_work_bind(ns_work &w, double *base, int n)
{
    w.y = base;
    w.x = base + n;
    w.X = base + 2 * (size_t)n;
    w.mu = base + 3 * (size_t)n;
    w.eta = base + 4 * (size_t)n;
    w.scratch = base + 5 * (size_t)n;
}

struct ns_fit_result {
    double sigma, slope;
    uint32_t flags; /* NS_F_SIG_AT_MIN | NS_F_MAXIT | NS_F_QUAD_ERROR | NS_F_NAN_ROWS | NS_F_QUAD_OVERFLOW */
    int n_outer, n_irls, n_obj;
    ns_fit_counters cnt;
};

/* sig.rob.tukey(sig) (glm_rob_nb.r:161-165), summed over the group's samples */
template <class G>
NS_GF inline double ns_sig_objective(double sig, const ns_work &w, int n, const ns_params &prm,
                                     ns_fit_counters &cnt)
{
    const double c = prm.c_tukey_sig;
    const double dg0 = ns_digamma(1 / sig);
    double acc = 0;
    int qerr = 0;
    for (int i = G::lane(); i < n; i += G::W) {
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
            double psiML = ns_digamma(r * sqrt(mu * (sig * mu + 1)) + mu + 1 / sig) -
                           sig * r * sqrt(mu / (sig * mu + 1)) - dg0 - log(sig * mu + 1);
            term = wi * psiML;
        }
        int qe0 = cnt.quad_err;
        cnt.quad_err = 0;
        term -= ns_ai_sig_tukey(mu, sig, c, prm.n_ai_sig_tukey, dg0, cnt);
        qerr |= cnt.quad_err;
        cnt.quad_err = qe0;
        acc += term;
    }
    acc = G::sum(acc);
    if (G::any(qerr))
        return NAN; /* integrate() error inside uniroot -> tryCatch -> NA (:175-177) */
    return acc;
}

/* uniroot(sig.rob.tukey, c(minsig,maxsig), tol=sig.prec, maxiter=maxit.sig)$root: R's
 * end-point checks followed by R_zeroin2 (Brent).  Returns false on any R error. */
template <class G>
NS_GF inline bool ns_sigma_root(const ns_work &w, int n, const ns_params &prm, ns_fit_result &res,
                                double &root)
{
    double a = prm.minsig, b = prm.maxsig;
    double fa = ns_sig_objective<G>(a, w, n, prm, res.cnt);
    double fb = ns_sig_objective<G>(b, w, n, prm, res.cnt);
    res.n_obj += 2;
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
        fb = ns_sig_objective<G>(b, w, n, prm, res.cnt);
        res.n_obj += 1;
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

/* The fit proper: glm_rob_nb.r:58-205 on the samples with w.x[i] > 0.
 * Expects w.y / w.x filled (x <= 0 for samples outside the fit); maxmu = max(x) over ALL
 * samples of the unit (:20). */
template <class G>
NS_GF inline void ns_fit_unit(const ns_work &w, int n, double maxmu, const ns_params &prm,
                              ns_fit_result &res)
{
    const int lane = G::lane();
    res.flags = 0;
    res.n_outer = res.n_irls = res.n_obj = 0;
    res.cnt.n_seed = res.cnt.n_lattice = res.cnt.n_quad = 0;
    res.cnt.quad_err = res.cnt.quad_overflow = 0;
    /* :58-63, :75-76 */
    int nfit = 0;
    double sum_ex = 0;
    for (int i = lane; i < n; i += G::W) {
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
    nfit = G::isum(nfit);
    sum_ex = G::sum(sum_ex);
    G::sync();
    /* :77 type-7 quartiles of af by rank counting */
    const double h25 = (double)(nfit - 1) * 0.25, h75 = (double)(nfit - 1) * 0.75;
    const int lo25 = (int)floor(h25), hi25 = (int)ceil(h25), lo75 = (int)floor(h75), hi75 = (int)ceil(h75);
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    for (int i = lane; i < n; i += G::W) {
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
    v0 = G::sum(v0);
    v1 = G::sum(v1);
    v2 = G::sum(v2);
    v3 = G::sum(v3);
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
    for (int i = lane; i < n; i += G::W)
        if (w.x[i] > 0 && w.eta[i] <= fence) {
            s_af += w.eta[i];
            c_af++;
        }
    s_af = G::sum(s_af);
    c_af = G::isum(c_af);
    double inislope = s_af / (double)c_af;
    {
        double mean_ex = sum_ex / (double)nfit;
        if (inislope <= 1 / (10 * mean_ex))
            inislope = 1 / (10 * mean_ex); /* :78 */
    }
    G::sync();
    /* :79-86 */
    double s_sig = 0;
    for (int i = lane; i < n; i += G::W) {
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
    s_sig = G::sum(s_sig);
    double sig = s_sig / (double)nfit;
    if (sig > prm.maxsig)
        sig = prm.maxsig;
    if (sig < prm.minsig)
        sig = prm.minsig;
    G::sync();

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
    while ((fabs(sig - sig0) > tol || ns_D_absmax(beta11, len11, beta00, len00) > tol) && it < prm.maxit) {
        sig0 = sig;
        beta00 = beta11;
        len00 = len11;
        /* :175-182 */
        double root = NAN;
        bool ok = ns_sigma_root<G>(w, n, prm, res, root);
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
        while (ns_D_absmax(beta1, len1, beta0, len0) > tol && it_mu < prm.maxit) {
            beta0 = beta1;
            len0 = len1;
            double sw = 0, swz = 0;
            int nanrows = 0;
            for (int i = lane; i < n; i += G::W) {
                if (!(w.x[i] > 0))
                    continue;
                double mu = w.mu[i], eta = w.eta[i], y = w.y[i];
                double e1, sap;
                ns_E_tukeypsi_12(mu, sig, prm.c_tukey_beta, prm.n_ai_sig_tukey, res.cnt, e1, sap);
                double V = ns_varfunc(mu, sig);
                double sV = sqrt(V);
                double ee = exp(eta);
                double bi = sap * (1.0 / (V * sV)) * (ee * ee);
                double ei = (ns_tukeypsi((y - mu) / sV, prm.c_tukey_beta) - e1) / sap * V * (1 / mu);
                double zi = eta + ei;
                if (bi != bi || zi != zi) {
                    nanrows++; /* lm(): na.omit drops the row */
                    continue;
                }
                if (bi == 0)
                    continue; /* lm.wfit(): zero-weight rows excluded */
                sw += bi;
                swz += bi * (zi - w.X[i]);
            }
            sw = G::sum(sw);
            swz = G::sum(swz);
            if (G::any(nanrows))
                res.flags |= NS_F_NAN_ROWS;
            double beta = swz / sw;
            beta1 = beta;
            len1 = 2;
            G::sync();
            for (int i = lane; i < n; i += G::W) {
                if (!(w.x[i] > 0))
                    continue;
                double mu = exp(w.X[i] + beta);
                if (mu > maxmu)
                    mu = maxmu;
                if (mu == 0)
                    mu = prm.minmu;
                w.mu[i] = mu;
                w.eta[i] = log(mu);
            }
            G::sync();
            it_mu++;
            res.n_irls++;
        }
        if (G::any(res.cnt.quad_err))
            res.flags |= NS_F_QUAD_ERROR; /* R would have stopped the script */
        res.cnt.quad_err = 0;
        beta11 = beta1;
        len11 = len1;
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

/* ---- everything after the fit, per unit ------------------------------------------ */
/* Output rows (length n each) of one unit in the compact planes. */
struct ns_unit_out {
    double *pvalue, *qvalue, *gq, *qval_minaf, *sor, *rvsb, *fs;
    uint8_t *gt, *status;
    double *pooled; /* 8 */
};

/* sorted p-values in srt[0..n-1] (ascending) from p[0..n-1]; rank counting */
template <class G> NS_GF inline void ns_sort_ranks(const double *p, double *srt, int *rank_out, int n)
{
    for (int i = G::lane(); i < n; i += G::W) {
        double pi = p[i];
        int rank = 0;
        for (int j = 0; j < n; j++) {
            double pj = p[j];
            rank += (pj < pi) || (pj == pi && j < i);
        }
        srt[rank] = pi;
        if (rank_out)
            rank_out[i] = rank;
    }
    G::sync();
}

/* toQvalueN / toQvalueT (needlestack.r:242-252) given the ascending fitted p-values and the
 * Holm prefix maxima hmax[k] = max_{j<=k} (n+1-j)*srt[j] (0-based j; m = n+1 hypotheses). */
NS_HD double ns_holm_last(double pstar, const double *srt, const double *hmax, int n)
{
    /* r = #{ p_j <= p* } (stable order puts p*, the last element, after its ties) */
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (srt[mid] <= pstar)
            lo = mid + 1;
        else
            hi = mid;
    }
    int r = lo;
    double v = (double)(n + 1 - r) * pstar; /* (m - (r+1) + 1) with m = n+1 */
    if (r > 0)
        v = fmax(v, hmax[r - 1]);
    return v < 1 ? v : 1;
}

/* Per-unit tail: p-values, BH q-values, GQ (glm_rob_nb.r:219-224); power q-values and
 * tumour/normal status (needlestack.r:340-380); gate 1, strand bias, gate 2, genotype
 * (needlestack.r:381-409).  sm: group-shared scratch of >= 3n+8 doubles.
 * Returns the unit's flag word (input flags | NS_F_GATE1 | NS_F_GATE2). */
template <class G>
NS_GF inline uint32_t ns_post_unit(const ns_rows &r, int n, const ns_params &prm, uint32_t flags,
                                   double sigma, double slope, const double *qtab, int qtab_len,
                                   const ns_unit_out &o, double *sm, double &max_gq_out)
{
    const int lane = G::lane();
    const bool gated = flags & NS_F_GATED;
    const bool inv = flags & NS_F_REF_INV;
    const double thr = prm.GQ_threshold;
    double *srt = sm, *hm = sm + n, *tmp = sm + 2 * (size_t)n;
    const double size = 1 / sigma;
    const int per = (n + G::W - 1) / G::W; /* contiguous block of sorted ranks per lane */
    const int b0 = (lane * per < n) ? lane * per : n, b1 = (b0 + per < n) ? b0 + per : n;
    /* p-values on the full vectors (:221) */
    if (gated) {
        for (int i = lane; i < n; i += G::W) {
            o.pvalue[i] = 1;
            o.qvalue[i] = 1;
            o.gq[i] = 0;
        }
    } else {
        for (int i = lane; i < n; i += G::W) {
            double x = ns_ld(r.dp, i);
            double y = inv ? ns_ld(r.rp, i) + ns_ld(r.rm, i) : ns_ld(r.vp, i) + ns_ld(r.vm, i);
            double p = ns_nb_pvalue(y, size, slope * x);
            o.pvalue[i] = p;
            tmp[i] = p;
        }
        G::sync();
        ns_sort_ranks<G>(tmp, srt, nullptr, n);
        /* p.adjust(p, "BH") (:222): q_(k) = min(1, min_{j>=k} n/(j+1)*p_(j)), ascending ranks.
         * Tied p-values share the value at their first rank, so ties need no special care. */
        if (n > 1) {
            double loc = INFINITY;
            for (int k = b1 - 1; k >= b0; k--)
                loc = fmin(loc, ((double)n / (double)(k + 1)) * srt[k]);
            double run = G::excl_suffix_min(loc);
            for (int k = b1 - 1; k >= b0; k--) {
                run = fmin(run, ((double)n / (double)(k + 1)) * srt[k]);
                hm[k] = run < 1 ? run : 1;
            }
        } else {
            for (int k = b0; k < b1; k++)
                hm[k] = srt[k];
        }
        G::sync();
        for (int i = lane; i < n; i += G::W) {
            double pi = tmp[i];
            int lo = 0, hi = n;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (srt[mid] < pi)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            double q = hm[lo];
            o.qvalue[i] = q;
            double g = -log10(q) * 10; /* :223-224 */
            if (g > 1000)
                g = 1000;
            o.gq[i] = g;
        }
    }
    G::sync();
    /* max(GQ), any GQ >= thr */
    double mg = -INFINITY;
    int anyq = 0;
    for (int i = lane; i < n; i += G::W) {
        mg = fmax(mg, o.gq[i]);
        anyq |= o.gq[i] >= thr;
    }
    mg = G::max(mg);
    anyq = G::any(anyq);
    max_gq_out = mg;

    /* qval_minAF (needlestack.r:340-380) */
    for (int i = lane; i < n; i += G::W) {
        o.qval_minaf[i] = 0;
        o.status[i] = NS_ST_NONE;
    }
    G::sync();
    if (!gated) {
        /* Holm prefix maxima over the ascending p-values, m = n + 1 hypotheses */
        double loc = -INFINITY;
        for (int k = b0; k < b1; k++)
            loc = fmax(loc, (double)(n + 1 - k) * srt[k]);
        double run = G::excl_prefix_max(loc);
        for (int k = b0; k < b1; k++) {
            run = fmax(run, (double)(n + 1 - k) * srt[k]);
            hm[k] = run;
        }
        G::sync();
    }
    /* which power test each sample gets: 0 none, 1 toQvalueN, 2 toQvalueT */
    const bool tn = prm.is_tn != 0;
    double afmin = prm.afmin_power;
    if (tn && afmin == -1)
        afmin = 0.01; /* needlestack.r:166 */
    int germ = 0;
    if (tn) {
        for (int k = lane; k < prm.n_pairs; k += G::W)
            germ |= o.gq[prm.nindex[k]] > thr;
        germ = G::any(germ);
    }
    for (int i = lane; i < n; i += G::W) {
        int kind;
        if (tn) {
            kind = prm.role ? prm.role[i] : 0; /* 0 none, 1 N of a pair, 2 T of a pair, 3 only N, 4 only T */
            if (kind == 1 || kind == 3)
                kind = 1;
            else if (kind == 2)
                kind = germ ? 1 : 2;
            else if (kind == 4)
                kind = 2;
        } else {
            kind = (prm.afmin_power == -1) ? 1 : 2;
        }
        if (kind == 0)
            continue;
        if (gated) {
            o.qval_minaf[i] = NAN; /* dnbinom(size = 1/NA) = NA */
            continue;
        }
        double x = ns_ld(r.dp, i);
        double ystar;
        if (kind == 1) {
            int xi = (int)x;
            ystar = (qtab && xi < qtab_len) ? qtab[xi] : ns_qnbinom_mu(0.01, 1 / prm.sigma_normal, 0.5 * x);
        } else {
            ystar = ns_qbinom(0.01, x, afmin);
        }
        double pstar = ns_nb_pvalue(ystar, size, slope * x);
        o.qval_minaf[i] = -10 * log10(ns_holm_last(pstar, srt, hm, n));
    }
    G::sync();
    /* tumour/normal status, needlestack.r:349-373, in the reference's assignment order */
    if (tn) {
        const int32_t *T = prm.tindex, *N = prm.nindex;
        for (int k = lane; k < prm.n_only_t; k += G::W)
            if (o.gq[prm.only_t[k]] > thr)
                o.status[prm.only_t[k]] = NS_ST_UNKNOWN;
        for (int k = lane; k < prm.n_only_n; k += G::W)
            if (o.gq[prm.only_n[k]] > thr)
                o.status[prm.only_n[k]] = NS_ST_GERMLINE_UNCONFIRMABLE;
        G::sync();
        /* one pair may appear once; a sample belonging to several pairs is resolved by
         * running the rules sequentially per rule over pairs, as R's vector assignment does */
        for (int rule = 0; rule < 5; rule++) {
            for (int pass = 0; pass < 2; pass++) { /* 0: T side, 1: N side */
                if (rule == 4 && pass == 1)
                    break;
                for (int k = lane; k < prm.n_pairs; k += G::W) {
                    double gT = o.gq[T[k]], gN = o.gq[N[k]], qT = o.qval_minaf[T[k]], qN = o.qval_minaf[N[k]];
                    bool cond;
                    uint8_t val;
                    switch (rule) {
                    case 0:
                        cond = (gT < thr) && (qT > thr) && (gN > thr);
                        val = NS_ST_GERMLINE_UNCONFIRMED;
                        break;
                    case 1:
                        cond = (gT < thr) && (qT < thr) && (gN > thr);
                        val = NS_ST_GERMLINE_UNCONFIRMABLE;
                        break;
                    case 2:
                        cond = (gT > thr) && (qN < thr) && (gN < thr);
                        val = NS_ST_UNKNOWN;
                        break;
                    case 3:
                        cond = (gT > thr) && (gN > thr);
                        val = NS_ST_GERMLINE_CONFIRMED;
                        break;
                    default:
                        cond = (gT > thr) && (qN > thr) && (gN < thr);
                        val = NS_ST_SOMATIC;
                        break;
                    }
                    if (cond)
                        o.status[pass == 0 ? T[k] : N[k]] = val;
                }
                G::sync();
            }
        }
        int ngerm = 0;
        for (int i = lane; i < n; i += G::W) {
            uint8_t s = o.status[i];
            ngerm |= (s == NS_ST_GERMLINE_CONFIRMED || s == NS_ST_GERMLINE_UNCONFIRMED ||
                      s == NS_ST_GERMLINE_UNCONFIRMABLE || s == NS_ST_UNKNOWN);
        }
        ngerm = G::any(ngerm);
        if (ngerm)
            for (int k = lane; k < prm.n_pairs; k += G::W)
                if (o.status[T[k]] == NS_ST_SOMATIC)
                    o.status[T[k]] = NS_ST_POSSIBLE_CONTAMINATION;
        G::sync();
    }

    /* gate 1 (needlestack.r:381; :534, :698 for indels, which have no output_all) */
    const bool out_all = (!r.is_indel) && prm.output_all_SNVs;
    const bool gate1 = out_all || (!gated && anyq);
    for (int i = lane; i < n; i += G::W) {
        o.sor[i] = NAN;
        o.rvsb[i] = NAN;
        o.fs[i] = NAN;
        o.gt[i] = inv ? NS_GT_11 : NS_GT_00;
    }
    if (!gate1)
        return flags;
    flags |= NS_F_GATE1;
    /* common_annot(), needlestack.r:217-239 */
    double sVp = 0, sVm = 0, sRp = 0, sRm = 0, sDP = 0;
    for (int i = lane; i < n; i += G::W) {
        double Vp = ns_ld(r.vp, i), Vm = ns_ld(r.vm, i), Rp = ns_ld(r.rp, i), Rm = ns_ld(r.rm, i);
        double Cp = Vp + Rp, Cm = Vm + Rm;
        double t = fmax(Vp * Cm, Vm * Cp) / (Vp * Cm + Vm * Cp);
        o.rvsb[i] = (t != t) ? -1 : t;
        double s = ns_sor(Rp, Vp, Rm, Vm);
        if (s != s)
            s = -1;
        else if (fabs(s) > NS_DBL_MAX)
            s = 99;
        o.sor[i] = s;
        o.fs[i] = -10 * log10(ns_fisher_2x2(Vp, Vm, Rp, Rm)); /* the cap at :235 is a no-op */
        sVp += Vp;
        sVm += Vm;
        sRp += Rp;
        sRm += Rm;
        sDP += ns_ld(r.dp, i);
    }
    sVp = G::sum(sVp);
    sVm = G::sum(sVm);
    sRp = G::sum(sRp);
    sRm = G::sum(sRm);
    sDP = G::sum(sDP);
    {
        double sCp = sVp + sRp, sCm = sVm + sRm;
        double ar = fmax(sVp * sCm, sVm * sCp) / (sVp * sCm + sVm * sCp);
        if (ar != ar)
            ar = -1;
        double as = ns_sor(sRp, sVp, sRm, sVm);
        if (as != as)
            as = -1;
        else if (fabs(as) > NS_DBL_MAX)
            as = 99;
        if (lane == 0) {
            o.pooled[0] = as;
            o.pooled[1] = ar;
            o.pooled[2] = -10 * log10(ns_fisher_2x2(sVp, sVm, sRp, sRm));
            o.pooled[3] = sRp;
            o.pooled[4] = sRm;
            o.pooled[5] = sVp;
            o.pooled[6] = sVm;
            o.pooled[7] = r.is_indel ? sDP + (sVp + sVm) : sDP; /* :383 vs :536,:700 */
        }
    }
    G::sync();
    const double *sbs = prm.SB_type == 0 ? o.sor : (prm.SB_type == 1 ? o.rvsb : o.fs);
    const double sb_thr = r.is_indel ? prm.SB_threshold_indel : prm.SB_threshold_SNV;
    int any2 = 0;
    for (int i = lane; i < n; i += G::W)
        any2 |= (sbs[i] <= sb_thr && o.gq[i] >= thr);
    any2 = G::any(any2);
    if (out_all || any2)
        flags |= NS_F_GATE2;
    /* genotype, needlestack.r:403-409 (:555-561 and :718-724 also use SB_threshold_SNV) */
    for (int i = lane; i < n; i += G::W) {
        double x = ns_ld(r.dp, i);
        double y = inv ? ns_ld(r.rp, i) + ns_ld(r.rm, i) : ns_ld(r.vp, i) + ns_ld(r.vm, i);
        double af = y / x;
        uint8_t g = inv ? NS_GT_11 : NS_GT_00;
        if (o.gq[i] >= thr && sbs[i] <= prm.SB_threshold_SNV && af < 0.75)
            g = NS_GT_01;
        if (o.gq[i] >= thr && sbs[i] <= prm.SB_threshold_SNV && af >= 0.75)
            g = inv ? NS_GT_00 : NS_GT_11;
        if (o.gq[i] < thr && o.qval_minaf[i] < thr)
            g = NS_GT_MISSING;
        o.gt[i] = g;
    }
    return flags;
}
