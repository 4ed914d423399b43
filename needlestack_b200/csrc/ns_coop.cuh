/*
 * ns_coop.cuh — warp-cooperative window expectations for the heavy kernel (device only).
 *
 * k_fit_heavy gives one CTA to one unit; each WARP takes one sample at a time and its 32
 * lanes share that sample's window (glm_rob_nb.r:109-160):
 *   lattice sum (window <= 2*n_ai.sig.tukey points)   lanes own contiguous chunks, each seeds
 *                                                      the pmf at its chunk start and walks it
 *   spline + integrate (wider windows)                 lanes evaluate the 100 knot values,
 *                                                      the spline lives in the warp's slice of
 *                                                      shared memory, the 21 Gauss-Kronrod
 *                                                      nodes of every dqags panel are one per
 *                                                      lane, the interval bookkeeping of dqags
 *                                                      is replicated
 * Every function returns values that are uniform across the warp.
 */
#pragma once
#if defined(__CUDACC__)
#include "ns_fit.cuh"

#ifndef NS_COOP_DOUBLES
#define NS_COOP_DOUBLES (4 * NS_MAX_KNOTS)
#endif

struct ns_cspline {
    int n;
    double j1, j2, by, inv_by;
    const double *y, *c;
    const double *rt;
    int rtn;
    const double *b, *d; /* per-interval first / third order coefficients (ns_cspline_coefs) */
};

/* 1/d for an integer-valued d > 0: table when it reaches, division otherwise */
__device__ __forceinline__ double ns_rcp_int(const double *rt, int rtn, double d)
{
    return (d < (double)rtn) ? rt[(int)d] : 1.0 / d;
}

__device__ __forceinline__ double ns_cknot(double j1, double j2, double by, int n, int k)
{
    if (k <= 0)
        return j1;
    if (k >= n - 1)
        return j2;
    return nearbyint(j1 + (double)k * by);
}

/* Natural-spline second derivatives for y[] in shared memory, all 32 lanes at work.
 * R's natural_spline() is the Thomas algorithm on the tridiagonal system of the interior
 * unknowns c[1..n-2]; here the same elimination is evaluated in parallel:
 *   pivots   b'_i = b_i - a_i^2/b'_{i-1} forget their start value by a factor (a/b')^2 < 0.15
 *            per step, so every lane restarts the chain 20 rows below its own 4 rows and
 *            reaches them with the sequential values to below one ulp;
 *   sweeps   r'_i = r_i - t_i r'_{i-1} and x_i = r'_i/b'_i - (a_{i+1}/b'_i) x_{i+1} are
 *            first-order affine recurrences: 4 rows per lane composed locally, a warp scan
 *            of the affine maps, then the 4 rows again from the scanned inflow.
 * Only rounding differs from the sequential sweep.  n <= 130. */
struct ns_cpiv {
    double piv[4], t[4], aup[4], dlo[4]; /* pivot, a_i/b'_{i-1}, a_{i+1} = d_i, d_{i-1} */
};

__device__ __forceinline__ void ns_cspline_pivots(ns_cpiv &P, int n, double j1, double j2, double by)
{
    const int lane = threadIdx.x & 31;
    const int i0 = 4 * lane + 1, nu = n - 2;
    int iw = i0 - 20;
    if (iw < 1)
        iw = 1;
    double xc = ns_cknot(j1, j2, by, n, iw);
    double dm = xc - ns_cknot(j1, j2, by, n, iw - 1); /* d_{iw-1} */
    double bp = 0;
    /* warm-up rows iw .. i0-1 */
    for (int i = iw; i < i0 && i <= nu; i++) {
        double xp = ns_cknot(j1, j2, by, n, i + 1);
        double di = xp - xc;
        double bi = 2 * (dm + di);
        if (i > iw) {
            /* the first half of the warm-up only has to be right to ~1e-7: its error is
             * damped by more than 1e-10 before the lane's own rows */
            double t = (i < i0 - 10) ? dm * (double)__frcp_rn((float)bp) : dm * ns_frcp(bp); /* pivots are >= 2 */
            bi = bi - t * dm;
        }
        bp = bi;
        dm = di;
        xc = xp;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = i0 + k;
        P.piv[k] = 1.0;
        P.t[k] = 0.0;
        P.aup[k] = 0.0;
        P.dlo[k] = 1.0;
        if (i <= nu) {
            double xp = ns_cknot(j1, j2, by, n, i + 1);
            double di = xp - xc;
            double bi = 2 * (dm + di);
            double t = 0;
            if (i > 1) { /* i == iw can only happen for i == 1 here (iw < i0 unless i0 == 1) */
                t = dm * ns_frcp(bp);
                bi = bi - t * dm;
            }
            P.piv[k] = bi;
            P.t[k] = t;
            P.aup[k] = di;
            P.dlo[k] = dm;
            bp = bi;
            dm = di;
            xc = xp;
        }
    }
}

__device__ __forceinline__ void ns_cspline_solve(const ns_cpiv &P, const ns_coop_scratch &s, int n)
{
    const int lane = threadIdx.x & 31;
    const int i0 = 4 * lane + 1, nu = n - 2;
    double rhs[4], rp[4];
    /* forward: r'_i = rhs_i - t_i r'_{i-1} */
    double A = 1.0, B = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = i0 + k;
        rhs[k] = 0.0;
        if (i <= nu) {
            rhs[k] = (s.y[i + 1] - s.y[i]) * ns_rcp_int(s.rt, s.rtn, P.aup[k]) -
                     (s.y[i] - s.y[i - 1]) * ns_rcp_int(s.rt, s.rtn, P.dlo[k]);
            B = rhs[k] - P.t[k] * B;
            A = -P.t[k] * A;
        }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double Ae = __shfl_up_sync(0xffffffffu, A, o), Be = __shfl_up_sync(0xffffffffu, B, o);
        if (lane >= o) {
            B = B + A * Be;
            A = A * Ae;
        }
    }
    double xin = __shfl_up_sync(0xffffffffu, B, 1);
    if (lane == 0)
        xin = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        rp[k] = 0.0;
        if (i0 + k <= nu) {
            xin = rhs[k] - P.t[k] * xin;
            rp[k] = xin;
        }
    }
    /* backward: x_i = r'_i/b'_i - (a_{i+1}/b'_i) x_{i+1}, x_{n-1} = 0 */
    double Bk[4], Ak[4];
    A = 1.0;
    B = 0.0;
#pragma unroll
    for (int k = 3; k >= 0; k--) {
        Bk[k] = 0.0;
        Ak[k] = 1.0;
        if (i0 + k <= nu) {
            double ib = ns_frcp(P.piv[k]);
            Bk[k] = rp[k] * ib;
            Ak[k] = -P.aup[k] * ib;
            B = Bk[k] + Ak[k] * B;
            A = Ak[k] * A;
        }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double Ae = __shfl_down_sync(0xffffffffu, A, o), Be = __shfl_down_sync(0xffffffffu, B, o);
        if (lane + o < 32) {
            B = B + A * Be;
            A = A * Ae;
        }
    }
    xin = __shfl_down_sync(0xffffffffu, B, 1);
    if (lane == 31)
        xin = 0.0;
#pragma unroll
    for (int k = 3; k >= 0; k--) {
        if (i0 + k <= nu) {
            xin = Bk[k] + Ak[k] * xin;
            s.c[i0 + k] = xin;
        }
    }
    if (lane == 0) {
        s.c[0] = 0.0;
        s.c[n - 1] = 0.0;
    }
}

/* b_i and d_i of R's spline coefficients (c_i as natural_spline() leaves it) for every interval, once per solve: the ~6 x 21 evaluations of a dqags call then cost two knot
 * positions, four shared loads and a Horner form instead of re-deriving the coefficients each time.  Row n-1 is
 * the linear continuation R uses to the right of the last knot (the integral runs to j2 + 1). */
__device__ __forceinline__ void ns_cspline_coefs(const ns_coop_scratch &sc, int n, double j1, double j2, double by)
{
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < n; i += 32) {
        const double xi = ns_cknot(j1, j2, by, n, i);
        if (i < n - 1) {
            const double di = ns_cknot(j1, j2, by, n, i + 1) - xi;
            const double idi = ns_rcp_int(sc.rt, sc.rtn, di);
            const double c0 = sc.c[i], c1 = sc.c[i + 1];
            sc.ib[i] = (sc.y[i + 1] - sc.y[i]) * idi - di * (c1 + 2.0 * c0);
            sc.m[i] = (c1 - c0) * idi;
        } else {
            const double dl = xi - ns_cknot(j1, j2, by, n, n - 2);
            sc.ib[i] = (sc.y[n - 1] - sc.y[n - 2]) * ns_rcp_int(sc.rt, sc.rtn, dl) + dl * sc.c[n - 2];
            sc.m[i] = 0.0;
        }
    }
    __syncwarp();
}

__device__ __forceinline__ double ns_cspline_eval(const ns_cspline &s, double u)
{
    const int n = s.n;
    int i = (int)((u - s.j1) * s.inv_by);
    if (i < 0)
        i = 0;
    if (i > n - 1)
        i = n - 1;
    double xi = ns_cknot(s.j1, s.j2, s.by, n, i);
    while (i > 0 && u < xi) {
        i--;
        xi = ns_cknot(s.j1, s.j2, s.by, n, i);
    }
    while (i < n - 1) {
        const double xn = ns_cknot(s.j1, s.j2, s.by, n, i + 1);
        if (u < xn)
            break;
        i++;
        xi = xn;
    }
    const double dx = u - xi;
    const double c3 = (i == n - 1) ? 0.0 : 3.0 * s.c[i];
    return s.y[i] + dx * (s.b[i] + dx * (c3 + dx * s.d[i]));
}

__device__ __forceinline__ double ns_wsum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/* Gauss-Kronrod 21 laid out one node per lane: signed node, Kronrod weight, Gauss weight */
static __device__ const double ns_lane_sx[32] = {
    0.0, -0.995657163025808080735527280689003, -0.973906528517171720077964012084452,
    -0.930157491355708226001207180059508, -0.865063366688984510732096688423493, -0.780817726586416897063717578345042,
    -0.679409568299024406234327365114874, -0.562757134668604683339000099272694, -0.433395394129247190799265943165784,
    -0.294392862701460198131126603103866, -0.148874338981631210884826001129720, 0.995657163025808080735527280689003,
    0.973906528517171720077964012084452, 0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
    0.780817726586416897063717578345042, 0.679409568299024406234327365114874, 0.562757134668604683339000099272694,
    0.433395394129247190799265943165784, 0.294392862701460198131126603103866, 0.148874338981631210884826001129720,
    0.0, 0.0, 0.0,
    0.0, 0.0, 0.0,
    0.0, 0.0, 0.0,
    0.0, 0.0};
static __device__ const double ns_lane_wk[32] = {
    0.149445554002916905664936468389821, 0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
    0.054755896574351996031381300244580, 0.075039674810919952767043140916190, 0.093125454583697605535065465083366,
    0.109387158802297641899210590325805, 0.123491976262065851077958109831074, 0.134709217311473325928054001771707,
    0.142775938577060080797094273138717, 0.147739104901338491374841515972068, 0.011694638867371874278064396062192,
    0.032558162307964727478818972459390, 0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
    0.093125454583697605535065465083366, 0.109387158802297641899210590325805, 0.123491976262065851077958109831074,
    0.134709217311473325928054001771707, 0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
    0.0, 0.0, 0.0,
    0.0, 0.0, 0.0,
    0.0, 0.0, 0.0,
    0.0, 0.0};
static __device__ const double ns_lane_wg[32] = {
    0.0, 0.0, 0.066671344308688137593568809893332,
    0.0, 0.149451349150580593145776339657697, 0.0,
    0.219086362515982043995534934228163, 0.0, 0.269266719309996355091226921569469,
    0.0, 0.295524224714752870173892994651338, 0.0,
    0.066671344308688137593568809893332, 0.0, 0.149451349150580593145776339657697,
    0.0, 0.219086362515982043995534934228163, 0.0,
    0.269266719309996355091226921569469, 0.0, 0.295524224714752870173892994651338,
    0.0, 0.0, 0.0,
    0.0, 0.0, 0.0,
    0.0, 0.0, 0.0,
    0.0, 0.0};

/* dqk21 with one Gauss-Kronrod node per lane (lanes 0..20); results uniform */
struct ns_k21_coop {
    template <class F>
    static __device__ __forceinline__ void eval(const F &f, double a, double b, double &result, double &abserr,
                                                double &resabs, double &resasc)
    {
        const double epmach = NS_DBL_EPS, uflow = NS_DBL_MIN;
        const int lane = threadIdx.x & 31;
        const double centr = (a + b) * .5, hlgth = (b - a) * .5, dhlgth = fabs(hlgth);
        /* lane 0: centre; lanes 1..10: centr - hlgth*xgk[l-1]; lanes 11..20: centr + hlgth*xgk[l-11] */
        double fv = 0;
        const double wk = ns_lane_wk[lane], wgw = ns_lane_wg[lane];
        if (lane < 21)
            fv = f(centr + hlgth * ns_lane_sx[lane]);
        /* three independent butterflies interleaved: one shuffle latency chain instead of three */
        double resk = wk * fv, resg = wgw * fv, rabs = wk * fabs(fv);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            resk += __shfl_xor_sync(0xffffffffu, resk, o);
            resg += __shfl_xor_sync(0xffffffffu, resg, o);
            rabs += __shfl_xor_sync(0xffffffffu, rabs, o);
        }
        double reskh = resk * .5;
        double rasc = ns_wsum(lane < 21 ? wk * fabs(fv - reskh) : 0.0);
        result = resk * hlgth;
        resabs = rabs * dhlgth;
        resasc = rasc * dhlgth;
        abserr = fabs((resk - resg) * hlgth);
        if (resasc != 0.0 && abserr != 0.0) {
            double r = abserr * 200.0 * ns_frcp(resasc); /* both positive and normal here */
            abserr = resasc * fmin(1.0, r * ns_fsqrt(r));
        }
        if (resabs > NS_DBL_MIN / (NS_DBL_EPS * 50.0))
            abserr = fmax(epmach * 50.0 * resabs, abserr);
    }
};

struct ns_cspline_fn {
    const ns_cspline *s;
    __device__ __forceinline__ double operator()(double u) const { return ns_cspline_eval(*s, u); }
};

template <bool ROLLED>
__device__ __forceinline__ double ns_cintegrate(const ns_cspline &s, int &ier, int &npanels)
{
    const double tol = 1.220703125e-4; /* .Machine$double.eps^0.25 */
    ns_cspline_fn f = {&s};
    ns_quad_out q;
    ns_dqagse_t<ns_k21_coop, ns_cspline_fn, ROLLED>(f, s.j1, s.j2 + 1.0, tol, tol, q);
    ier = q.ier;
    npanels += q.neval / 21;
    return q.result;
}

/* knots handled by this lane: a contiguous block */
__device__ __forceinline__ void ns_lane_block(int n, int &k0, int &k1)
{
    const int lane = threadIdx.x & 31;
    const int per = (n + 31) >> 5;
    k0 = lane * per;
    k1 = k0 + per;
    if (k0 > n)
        k0 = n;
    if (k1 > n)
        k1 = n;
}

/* ai.sig.tukey(mui, sig, c), glm_rob_nb.r:141-160, one sample per warp */
template <bool ROLLED>
static __device__ __noinline__ double ns_ai_coop(double mui, double sig, double c, int n_ai, double dg0,
                                          const ns_coop_scratch &sc, ns_fit_counters &cnt)
{
    const int lane = threadIdx.x & 31;
    const double sqrtV = sqrt(mui * (sig * mui + 1));
    const double invsig = 1 / sig;
    const double j1 = fmax(ceil(mui - c * sqrtV), 0.0);
    const double j2 = floor(mui + c * sqrtV);
    cnt.n_seed += 1;
    if (j1 > j2)
        return 0.0;
    const double lg = log(mui / invsig + 1);
    const double csd = c * sqrtV, mpi = mui + invsig;
    const double inv_csd = 1.0 / csd, inv_mpi = 1.0 / mpi;
    const double q = mui * inv_mpi;
    const double K = j2 - j1 + 1;
    if (K <= 2 * n_ai) {
        cnt.n_lattice += (unsigned long long)K;
        const double per = ceil(K / 32.0);
        const double a = j1 + (double)lane * per;
        const double b = fmin(j2, a + per - 1);
        double acc = 0;
        if (a <= j2) {
            double pm = ns_dnbinom_mu_c(a, invsig, mui);
            double D = (a == 0) ? 0.0 : ns_digamma_c(a + invsig) - dg0;
            for (double j = a; j <= b; j += 1) {
                double d = j - mui;
                double t = d * inv_csd;
                double w = (t * t - 1) * (t * t - 1);
                acc += w * (D - lg - d * inv_mpi) * pm;
                double ja = j + invsig;
                pm *= (ja * ns_rcp_int(sc.rt, sc.rtn, j + 1)) * q;
                D += ns_frcp(ja);
            }
        }
        return ns_wsum(acc);
    }
    cnt.n_lattice += (unsigned long long)n_ai;
    const double by = (j2 - j1) / (double)(n_ai - 1);
    int k0, k1;
    ns_lane_block(n_ai, k0, k1);
    __syncwarp();
    if (k0 < k1) {
        if (by <= NS_STEP_BY) {
            /* pmf by recurrence along the lattice (1/(j+1) from the table), the digamma
             * difference directly at each knot */
            double j = ns_cknot(j1, j2, by, n_ai, k0);
            double pm = ns_dnbinom_mu_c(j, invsig, mui);
            /* digamma(j + 1/sig) - digamma(1/sig): one evaluation at the lane's first knot, then the recurrence
             * digamma(x + 1) = digamma(x) + 1/x along the walk (a digamma per knot cost more than the <= 16 steps) */
            double D = (j == 0) ? 0.0 : ns_digamma_c(j + invsig) - dg0;
            int k = k0;
            double xk = j;
            for (;; j += 1) {
                if (j == xk) {
                    double t = (j - mui) * inv_csd;
                    double w = (t * t - 1) * (t * t - 1);
                    sc.y[k] = w * (D - lg - (j - mui) * inv_mpi) * pm;
                    if (++k == k1)
                        break;
                    xk = ns_cknot(j1, j2, by, n_ai, k);
                }
                const double ja = j + invsig;
                pm *= (ja * ns_rcp_int(sc.rt, sc.rtn, j + 1)) * q;
                D += ns_frcp(ja);
            }
        } else {
            for (int k = k0; k < k1; k++) {
                double j = ns_cknot(j1, j2, by, n_ai, k);
                double t = (j - mui) * inv_csd;
                double w = (t * t - 1) * (t * t - 1);
                double psi = ns_digamma_c(j + invsig) - dg0 - lg - (j - mui) * inv_mpi;
                sc.y[k] = w * psi * ns_dnbinom_mu_c(j, invsig, mui);
            }
        }
    }
    __syncwarp();
    {
        ns_cpiv P;
        ns_cspline_pivots(P, n_ai, j1, j2, by);
        ns_cspline_solve(P, sc, n_ai);
    }
    __syncwarp();
    ns_cspline_coefs(sc, n_ai, j1, j2, by);
    ns_cspline s = {n_ai, j1, j2, by, 1.0 / by, sc.y, sc.c, sc.rt, sc.rtn, sc.ib, sc.m};
    int ier = 0, np = 0;
    double v = ns_cintegrate<ROLLED>(s, ier, np);
    cnt.n_quad += (unsigned long long)np;
    if (ier == 7)
        cnt.quad_overflow = 1;
    if (ier != 0 || !(fabs(v) <= NS_DBL_MAX))
        cnt.quad_err = 1;
    __syncwarp();
    return v;
}

/* E.tukeypsi.1 / E.tukeypsi.2 (glm_rob_nb.r:109-140), one sample per warp */
template <bool ROLLED>
static __device__ __noinline__ void ns_E12_coop(double mui, double sig, double c, int n_ai, const ns_coop_scratch &sc,
                                         ns_fit_counters &cnt, double *e1, double *e2)
{
    const int lane = threadIdx.x & 31;
    const double sqrtV = sqrt(ns_varfunc(mui, sig));
    const double j1 = fmax(ceil(mui - c * sqrtV), 0.0);
    const double j2 = floor(mui + c * sqrtV);
    cnt.n_seed += 1;
    if (j1 > j2) {
        *e1 = 0;
        *e2 = 0;
        return;
    }
    const double size = 1 / sig, csd = c * sqrtV;
    const double inv_csd = 1.0 / csd;
    const double q = mui / (mui + size);
    const double K = j2 - j1 + 1;
    if (K <= 2 * n_ai) {
        cnt.n_lattice += (unsigned long long)K;
        const double per = ceil(K / 32.0);
        const double a = j1 + (double)lane * per;
        const double b = fmin(j2, a + per - 1);
        double a1 = 0, a2 = 0;
        if (a <= j2) {
            double pm = ns_dnbinom_mu_c(a, size, mui);
            for (double j = a; j <= b; j += 1) {
                double d = j - mui;
                double t = d * inv_csd;
                double w = (t * t - 1) * (t * t - 1);
                double wp = w * pm;
                a1 += wp * d;
                a2 += wp * (d * d);
                pm *= ((j + size) * ns_rcp_int(sc.rt, sc.rtn, j + 1)) * q;
            }
        }
        *e1 = ns_wsum(a1) / sqrtV;
        *e2 = ns_wsum(a2) / sqrtV;
        return;
    }
    cnt.n_lattice += (unsigned long long)n_ai;
    const double by = (j2 - j1) / (double)(n_ai - 1);
    int k0, k1;
    ns_lane_block(n_ai, k0, k1);
    __syncwarp();
    if (k0 < k1) {
        if (by <= NS_STEP_BY) {
            double j = ns_cknot(j1, j2, by, n_ai, k0);
            double pm = ns_dnbinom_mu_c(j, size, mui);
            int k = k0;
            double xk = j;
            for (;; j += 1) {
                if (j == xk) {
                    double t = (j - mui) * inv_csd;
                    double w = (t * t - 1) * (t * t - 1);
                    sc.y[k] = w * (j - mui) * pm;
                    if (++k == k1)
                        break;
                    xk = ns_cknot(j1, j2, by, n_ai, k);
                }
                pm *= ((j + size) * ns_rcp_int(sc.rt, sc.rtn, j + 1)) * q;
            }
        } else {
            for (int k = k0; k < k1; k++) {
                double j = ns_cknot(j1, j2, by, n_ai, k);
                double t = (j - mui) * inv_csd;
                double w = (t * t - 1) * (t * t - 1);
                sc.y[k] = w * (j - mui) * ns_dnbinom_mu_c(j, size, mui);
            }
        }
    }
    __syncwarp();
    ns_cpiv P;
    ns_cspline_pivots(P, n_ai, j1, j2, by);
    ns_cspline s = {n_ai, j1, j2, by, 1.0 / by, sc.y, sc.c, sc.rt, sc.rtn, sc.ib, sc.m};
    int ier1 = 0, ier2 = 0, np = 0;
    double v1 = 0, v2 = 0;
    /* the two integrals share one copy of the solve + dqags code (a rolled loop: the kernel is bound by instruction
     * fetch where many warps are in different phases) */
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            for (int k = k0; k < k1; k++)
                sc.y[k] *= (ns_cknot(j1, j2, by, n_ai, k) - mui); /* w*(j-mu)^2*pmf */
            __syncwarp();
        }
        ns_cspline_solve(P, sc, n_ai);
        __syncwarp();
        ns_cspline_coefs(sc, n_ai, j1, j2, by);
        int ier = 0;
        const double v = ns_cintegrate<ROLLED>(s, ier, np);
        __syncwarp();
        if (pass == 0) {
            v1 = v;
            ier1 = ier;
        } else {
            v2 = v;
            ier2 = ier;
        }
    }
    cnt.n_quad += (unsigned long long)np;
    if (ier1 == 7 || ier2 == 7)
        cnt.quad_overflow = 1;
    if (ier1 != 0 || ier2 != 0 || !(fabs(v1) <= NS_DBL_MAX) || !(fabs(v2) <= NS_DBL_MAX))
        cnt.quad_err = 1;
    *e1 = v1 / sqrtV;
    *e2 = v2 / sqrtV;
    __syncwarp();
}
#endif /* __CUDACC__ */
