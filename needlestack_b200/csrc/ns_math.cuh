/*
 * ns_math.cuh — FP64 numeric layer of the B200 error-model path (device code).
 *
 * Everything the robust NB regression of /root/reference/bin/glm_rob_nb.r and the
 * per-allele driver of /root/reference/bin/needlestack.r take from base R:
 * dnbinom (Loader saddle point), the NB upper tail, digamma, discrete quantiles,
 * Fisher 2x2, natural cubic spline, QUADPACK dqags (GK21 + epsilon table).
 *
 * All functions are lane-local (no warp collectives) and are compiled for the
 * device by nvcc.  They are also legal host C++ (NS_HD expands to nothing without
 * nvcc) so that tests/ can build a serial emulation of the kernels on a machine
 * without a GPU; the shipped library never runs them on the host.
 */
#pragma once
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define NS_HD __host__ __device__ __forceinline__
#define NS_HDN __host__ __device__ __noinline__
#else
#define NS_HD inline
#define NS_HDN
#endif

#define NS_LN_SQRT_2PI 0.918938533204672741780329736406
#define NS_LN_2PI 1.837877066409345483560659472811
#define NS_DBL_EPS 2.220446049250313e-16
#define NS_DBL_MIN 2.2250738585072014e-308
#define NS_DBL_MAX 1.7976931348623157e+308

/* Branch-free reciprocal and square root for operands that are known to be normal and positive
 * (variances, means, window widths): hardware seed (MUFU.RCP64H / RSQ64H, ~20 bits) and two
 * Newton steps, <= 1-2 ulp.  The IEEE-exact `/` and sqrt() cost ~3x as many instructions because of
 * their special-case paths (ncu: 25 % of k_fit's instructions were division sequences).  On the
 * host (emulation build) they are the plain operations. */
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double ns_frcp(double a)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double ns_fsqrt(double a)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    /* y ~ 1/sqrt(a): two Newton steps on y, then s = a*y with one correction */
    double h = 0.5 * a;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    double s = a * y;
    s = fma(fma(-s, s, a), 0.5 * y, s);
    return (a > 0.0) ? s : ((a == 0.0) ? 0.0 : NAN);
}
/* s = sqrt(a) and y = 1/sqrt(a) from one hardware seed (a normal and positive): every reciprocal the
 * window sums need (1/sd, 1/(c sd), 1/(mu+1/sig) = sig mu y^2, 1/(V sd) = y^3) is a product of y */
__device__ __forceinline__ void ns_fsqrt2(double a, double &s, double &y)
{
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    s = a * y;
    s = fma(fma(-s, s, a), 0.5 * y, s);
}
__device__ __forceinline__ int ns_d2i_floor(double a) { return __double2int_rd(a); }
#else
static inline double ns_frcp(double a) { return 1.0 / a; }
static inline double ns_fsqrt(double a) { return sqrt(a); }
static inline void ns_fsqrt2(double a, double &s, double &y)
{
    s = sqrt(a);
    y = 1.0 / s;
}
static inline int ns_d2i_floor(double a) { return (int)floor(a); }
#endif

/* ---- constant tables: __constant__ on the device, plain statics on the host ---- */
#define NS_SFERR_INIT                                                                              \
    {0.0, 0.1534264097200273452913839, 0.08106146679532725821967026,                               \
     0.0548141210519176538961387, 0.04134069595540929409382208, 0.03316287351993628748511051,      \
     0.02767792568499833914878929, 0.02374616365629749597133028, 0.02079067210376509311152277,     \
     0.01848845053267318523077936, 0.01664469118982119216319487, 0.01513497322191737887351384,     \
     0.01387612882307074799874573, 0.01281046524292022692425066, 0.01189670994589177009505572,     \
     0.01110455975820691732663076, 0.01041126526197209649747857, 0.009799416126158803298390373,    \
     0.009255462182712732917728637, 0.008768700134139385462955047, 0.008330563433362871256469319,  \
     0.007934114564314020547249562, 0.007573675487951840794972024, 0.007244554301320383179546197,  \
     0.006942840107209529865664153, 0.006665247032707682442356181, 0.006408994188004207068439631,  \
     0.006171712263039457647534605, 0.005951370112758847735624416, 0.005746216513010115682026102,  \
     0.00555473355196280137103869}
/* Gauss-Kronrod 21: nodes (descending), Kronrod weights, 10-point Gauss weights */
#define NS_XGK_INIT                                                                                \
    {0.995657163025808080735527280689003, 0.973906528517171720077964012084452,                     \
     0.930157491355708226001207180059508, 0.865063366688984510732096688423493,                     \
     0.780817726586416897063717578345042, 0.679409568299024406234327365114874,                     \
     0.562757134668604683339000099272694, 0.433395394129247190799265943165784,                     \
     0.294392862701460198131126603103866, 0.148874338981631210884826001129720, 0.0}
#define NS_WGK_INIT                                                                                \
    {0.011694638867371874278064396062192, 0.032558162307964727478818972459390,                     \
     0.054755896574351996031381300244580, 0.075039674810919952767043140916190,                     \
     0.093125454583697605535065465083366, 0.109387158802297641899210590325805,                     \
     0.123491976262065851077958109831074, 0.134709217311473325928054001771707,                     \
     0.142775938577060080797094273138717, 0.147739104901338491374841515972068,                     \
     0.149445554002916905664936468389821}
#define NS_WG_INIT                                                                                 \
    {0.066671344308688137593568809893332, 0.149451349150580593145776339657697,                     \
     0.219086362515982043995534934228163, 0.269266719309996355091226921569469,                     \
     0.295524224714752870173892994651338}

#define NS_ODDRCP_INIT {1.0 / 1.0, 1.0 / 3.0, 1.0 / 5.0, 1.0 / 7.0, 1.0 / 9.0, 1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0, 1.0 / 17.0, 1.0 / 19.0, 1.0 / 21.0, 1.0 / 23.0, 1.0 / 25.0, 1.0 / 27.0, 1.0 / 29.0, 1.0 / 31.0, 1.0 / 33.0, 1.0 / 35.0, 1.0 / 37.0, 1.0 / 39.0, 1.0 / 41.0, 1.0 / 43.0, 1.0 / 45.0, 1.0 / 47.0, 1.0 / 49.0, 1.0 / 51.0, 1.0 / 53.0, 1.0 / 55.0, 1.0 / 57.0, 1.0 / 59.0, 1.0 / 61.0, 1.0 / 63.0, 1.0 / 65.0, 1.0 / 67.0, 1.0 / 69.0, 1.0 / 71.0, 1.0 / 73.0, 1.0 / 75.0, 1.0 / 77.0, 1.0 / 79.0, 1.0 / 81.0, 1.0 / 83.0, 1.0 / 85.0, 1.0 / 87.0, 1.0 / 89.0, 1.0 / 91.0, 1.0 / 93.0, 1.0 / 95.0, 1.0 / 97.0, 1.0 / 99.0, 1.0 / 101.0, 1.0 / 103.0, 1.0 / 105.0, 1.0 / 107.0, 1.0 / 109.0, 1.0 / 111.0, 1.0 / 113.0, 1.0 / 115.0, 1.0 / 117.0, 1.0 / 119.0, 1.0 / 121.0, 1.0 / 123.0, 1.0 / 125.0, 1.0 / 127.0}

#if defined(__CUDACC__)
static __constant__ double ns_d_oddrcp[64] = NS_ODDRCP_INIT;
static __constant__ double ns_d_sferr[31] = NS_SFERR_INIT;
static __constant__ double ns_d_xgk[11] = NS_XGK_INIT;
static __constant__ double ns_d_wgk[11] = NS_WGK_INIT;
static __constant__ double ns_d_wg[5] = NS_WG_INIT;
#endif
static const double ns_h_oddrcp[64] = NS_ODDRCP_INIT;
static const double ns_h_sferr[31] = NS_SFERR_INIT;
static const double ns_h_xgk[11] = NS_XGK_INIT;
static const double ns_h_wgk[11] = NS_WGK_INIT;
static const double ns_h_wg[5] = NS_WG_INIT;
#if defined(__CUDA_ARCH__)
#define NS_TBL(name) ns_d_##name
#else
#define NS_TBL(name) ns_h_##name
#endif

/* exp() and log() for the hot loops.  Same argument reduction and polynomial degree as any libm, but
 * the coefficients are operands from constant memory: the CUDA math library materialises every double
 * constant with two move instructions, which made ~2/3 of the ~75 instructions of each exp/log call
 * (SASS of the first builds; UMOV + IMAD.MOV were 20 % of k_fit's instructions).
 *   ns_fexp: finite x; underflows to 0 below -708 (no denormals), Taylor to r^13 on |r| <= ln2/2: < 1 ulp
 *   ns_flog: normal positive x; the fdlibm scheme  log(1+f) = f - (f^2/2 - s*(f^2/2 + R(s^2))), s = f/(2+f) */
#define NS_EXPC_INIT {1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06, 2.48015873015873e-05, 0.0001984126984126984, 0.001388888888888889, 0.008333333333333333, 0.041666666666666664, 0.16666666666666666, 0.5}
#define NS_LOGC_INIT                                                                               \
    {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,                 \
     2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,                 \
     1.479819860511658591e-01}
#if defined(__CUDACC__)
static __constant__ double ns_d_expc[12] = NS_EXPC_INIT;
static __constant__ double ns_d_logc[7] = NS_LOGC_INIT;
#endif
static const double ns_h_expc[12] = NS_EXPC_INIT;
static const double ns_h_logc[7] = NS_LOGC_INIT;

NS_HD double ns_fexp(double x)
{
    if (!(x > -708.0))
        return (x != x) ? x : 0.0;
    const double MAGIC = 6755399441055744.0; /* 1.5 * 2^52: rounds to nearest integer in the low bits */
    double kd = fma(x, 1.4426950408889634074, MAGIC);
    long long kb;
    memcpy(&kb, &kd, sizeof kb);
    const int k = (int)(kb & 0xffffffffLL); /* low 32 bits hold the integer (two's complement) */
    kd -= MAGIC;
    double r = fma(-kd, 6.93147180369123816490e-01, x);
    r = fma(-kd, 1.90821492927058770002e-10, r);
    double p = NS_TBL(expc)[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 1; i < 12; i++)
        p = fma(p, r, NS_TBL(expc)[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    long long pb;
    memcpy(&pb, &p, sizeof pb);
    pb += (long long)k << 52; /* p in [0.7, 1.42], k >= -1022: stays normal */
    memcpy(&p, &pb, sizeof p);
    return p;
}

/* ns_fexp without the early return (x finite or NaN): straight-line code, so that two calls interleave */
NS_HD double ns_fexp_bf(double x)
{ /* |x| < 2^31; below -708 the scaled bits are meaningless and the select returns 0 */
    const double MAGIC = 6755399441055744.0;
    double kd = fma(x, 1.4426950408889634074, MAGIC);
    long long kb;
    memcpy(&kb, &kd, sizeof kb);
    const int k = (int)(kb & 0xffffffffLL);
    kd -= MAGIC;
    double r = fma(-kd, 6.93147180369123816490e-01, x);
    r = fma(-kd, 1.90821492927058770002e-10, r);
    double p = NS_TBL(expc)[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 1; i < 12; i++)
        p = fma(p, r, NS_TBL(expc)[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    unsigned long long pb;
    memcpy(&pb, &p, sizeof pb);
    pb += (unsigned long long)(unsigned)k << 52;
    memcpy(&p, &pb, sizeof p);
    return (x < -708.0) ? 0.0 : p;
}

NS_HD double ns_flog(double x)
{
    long long xb;
    memcpy(&xb, &x, sizeof xb);
    int k = (int)(xb >> 52) - 1023;
    long long mb = (xb & 0x000fffffffffffffLL) | 0x3ff0000000000000LL; /* m in [1, 2) */
    double m;
    memcpy(&m, &mb, sizeof m);
    if (m > 1.41421356237309504880) {
        m *= 0.5;
        k += 1;
    }
    const double f = m - 1.0;
    const double s = f * ns_frcp(2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, NS_TBL(logc)[5], NS_TBL(logc)[3]), NS_TBL(logc)[1]);
    const double t2 = z * fma(w, fma(w, fma(w, NS_TBL(logc)[6], NS_TBL(logc)[4]), NS_TBL(logc)[2]), NS_TBL(logc)[0]);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    return dk * 6.93147180369123816490e-01 - ((hfsq - fma(s, hfsq + R, dk * 1.90821492927058770002e-10)) - f);
}

/* log(x) for x >= 1 (normal) from a 128-entry table: x = 2^k m, entry i = top 7 mantissa bits holds
 * rc ~ 1/c_i (c_i the interval centre, c_0 = 1) and L = -log(rc); r = m*rc - 1 is one FMA with |r| < 2^-7 and
 * log(x) = k ln2 + L + log1p(r), log1p by its Taylor series to r^8 (truncation below 2^-59 relative to r).  All
 * terms are non-negative for x >= 1 (no cancellation; x -> 1 gives k = 0, L = 0 and the series alone).  Half the
 * instructions of ns_flog: no reciprocal, a shorter polynomial.  The table lives in shared memory on the device
 * (lanes index it divergently, which constant memory would serialise). */
#include "ns_logtab.h"
#define NS_LOG1PC_INIT {-1.0 / 8, 1.0 / 7, -1.0 / 6, 1.0 / 5, -1.0 / 4, 1.0 / 3, -0.5}
#if defined(__CUDACC__)
static __device__ const double ns_g_logtab[256] = NS_LOGTAB_INIT;
static __constant__ double ns_d_log1pc[7] = NS_LOG1PC_INIT;
#endif
static const double ns_h_logtab[256] = NS_LOGTAB_INIT;
static const double ns_h_log1pc[7] = NS_LOG1PC_INIT;

NS_HD double ns_flog_tab(double x, const double *tab)
{
    long long xb;
    memcpy(&xb, &x, sizeof xb);
    const int hi = (int)(xb >> 32);
    const int k = (hi >> 20) - 1023;
    const int i = (hi >> 13) & 127;
    const long long mb = (xb & 0x000fffffffffffffLL) | 0x3ff0000000000000LL;
    double m;
    memcpy(&m, &mb, sizeof m);
    double rc, L;
#if defined(__CUDA_ARCH__)
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rc), "=d"(L) : "r"((unsigned)__cvta_generic_to_shared(tab) + 16u * (unsigned)i));
#else
    rc = tab[2 * i];
    L = tab[2 * i + 1];
#endif
    const double r = fma(m, rc, -1.0);
    double p = fma(r, NS_TBL(log1pc)[0], NS_TBL(log1pc)[1]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 2; j < 7; j++)
        p = fma(p, r, NS_TBL(log1pc)[j]);
    const double l1 = fma(r * r, p, r);
    const double dk = (double)k;
    return fma(dk, 6.93147180369123816490e-01, L) + fma(dk, 1.90821492927058770002e-10, l1);
}

/* ---------------------------------------------------------------------------------
 * Loader's saddle-point building blocks (R nmath stirlerr / bd0 / dbinom_raw), the
 * arithmetic behind every dnbinom() of glm_rob_nb.r:117,121,133,137,153,157,221.
 * ------------------------------------------------------------------------------- */
NS_HD double ns_stirlerr(double n)
{
    const double S0 = 1.0 / 12.0, S1 = 1.0 / 360.0, S2 = 1.0 / 1260.0, S3 = 1.0 / 1680.0,
                 S4 = 1.0 / 1188.0;
    if (n <= 15.0) {
        double nn = n + n;
        if (nn == (double)(int)nn)
            return NS_TBL(sferr)[(int)nn];
        return lgamma(n + 1.0) - (n + 0.5) * log(n) + n - NS_LN_SQRT_2PI;
    }
    double nn = n * n;
    if (n > 500)
        return (S0 - S1 / nn) / n;
    if (n > 80)
        return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35)
        return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

NS_HD double ns_bd0(double x, double np)
{
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        if (fabs(s) < NS_DBL_MIN)
            return s;
        double ej = 2 * x * v;
        v = v * v;
        for (int j = 1; j < 1000; j++) {
            ej *= v;
            /* R divides by 2j+1; the correctly rounded reciprocal differs by at most one ulp per term */
            double s1 = s + (j < 64 ? ej * NS_TBL(oddrcp)[j] : ej / (double)((j << 1) + 1));
            if (s1 == s)
                return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

/* log dbinom_raw(x, n, p, q) for 0 < x < n, 0 < p,q; nmx = n - x is passed explicitly */
NS_HD double ns_dbinom_raw_log_mid(double x, double nmx, double n, double p, double q)
{
    double lc = ns_stirlerr(n) - ns_stirlerr(x) - ns_stirlerr(nmx) - ns_bd0(x, n * p) -
                ns_bd0(nmx, n * q);
    double lf = NS_LN_2PI + log(x) + log1p(-x / n);
    return lc - 0.5 * lf;
}

/* general log dbinom_raw (used by the hypergeometric density) */
NS_HD double ns_dbinom_raw_log(double x, double n, double p, double q)
{
    if (p == 0)
        return (x == 0) ? 0.0 : -INFINITY;
    if (q == 0)
        return (x == n) ? 0.0 : -INFINITY;
    if (x == 0) {
        if (n == 0)
            return 0.0;
        return (p < 0.1) ? -ns_bd0(n, n * q) - n * p : n * log(q);
    }
    if (x == n)
        return (q < 0.1) ? -ns_bd0(n, n * p) - n * q : n * log(p);
    if (x < 0 || x > n)
        return -INFINITY;
    return ns_dbinom_raw_log_mid(x, n - x, n, p, q);
}

/* dnbinom(x, size=size, mu=mu) at an integer x >= 0 (R nmath dnbinom_mu).
 * size in [1/maxsig, 1/minsig] = [0.1, 1000] so the x < 1e-10*size branch is dead. */
NS_HD double ns_dnbinom_mu(double x, double size, double mu)
{
    if (x == 0)
        return exp(size * (size < mu ? log(size / (size + mu)) : log1p(-mu / (size + mu))));
    double q = mu / (size + mu);
    if (q == 0)
        return 0.0; /* mu == 0 and x > 0 */
    double p = size / (size + mu);
    double n = x + size;
    /* dbinom_raw(size, x+size, p, q) * size/(size+x) */
    double l = ns_dbinom_raw_log_mid(size, x, n, p, q);
    return (size / (size + x)) * exp(l);
}

/* digamma(x), x > 0 (arguments here are >= 1/maxsig = 0.1): upward recurrence to
 * x >= 12 then the asymptotic series; glm_rob_nb.r:102,143 */
NS_HD double ns_digamma(double x)
{
    /* sum_{k<m} 1/(x+k) as one fraction num/den (m <= 12 terms, den <= 12^12): one division */
    double num = 0.0, den = 1.0;
    while (x < 12.0) {
        num = num * x + den;
        den *= x;
        x += 1.0;
    }
    double r = 1.0 / x, r2 = r * r;
    double s = r2 * (1.0 / 12 -
                     r2 * (1.0 / 120 -
                           r2 * (1.0 / 252 -
                                 r2 * (1.0 / 240 -
                                       r2 * (1.0 / 132 -
                                             r2 * (691.0 / 32760 - r2 * (1.0 / 12)))))));
    double head = log(x) - 0.5 * r - s;
    return (num == 0.0) ? head : head - num / den;
}

/* ---------------------------------------------------------------------------------
 * Regularised incomplete beta with full relative accuracy in the far tail:
 *   I_x(a,b) = dbinom_raw(a, a+b, x, y) * b/(a+b) * CF(a,b,x)      (modified Lentz)
 * R reaches the same function through TOMS 708 bratio(); pnbinom(..., lower=F) of
 * glm_rob_nb.r:221 and needlestack.r:245,251.
 * ------------------------------------------------------------------------------- */
NS_HD double ns_betacf(double a, double b, double x)
{
    const double FPMIN = 1e-300, EPS = 1e-16;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < FPMIN)
        d = FPMIN;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 200000; m++) {
        double dm = (double)m, m2 = 2.0 * dm;
        double aa = dm * (b - dm) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < FPMIN)
            d = FPMIN;
        c = 1.0 + aa / c;
        if (fabs(c) < FPMIN)
            c = FPMIN;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + dm) * (qab + dm) * x / ((a + m2) * (qap + m2));
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

NS_HD double ns_dbinom_ab(double a, double b, double p, double q)
{
    double n = a + b;
    double lc = ns_stirlerr(n) - ns_stirlerr(a) - ns_stirlerr(b) - ns_bd0(a, n * p) -
                ns_bd0(b, n * q);
    double lf = NS_LN_2PI + log(a) + log(b / n);
    return exp(lc - 0.5 * lf);
}

/* lower ? I_x(a,b) : 1 - I_x(a,b), with y = 1 - x supplied */
NS_HD double ns_pbeta_raw(double x, double y, double a, double b, int lower)
{
    if (x <= 0)
        return lower ? 0.0 : 1.0;
    if (y <= 0)
        return lower ? 1.0 : 0.0;
    if (x < (a + 1.0) / (a + b + 2.0)) {
        double I = ns_dbinom_ab(a, b, x, y) * (b / (a + b)) * ns_betacf(a, b, x);
        return lower ? I : 1.0 - I;
    }
    double J = ns_dbinom_ab(b, a, y, x) * (a / (a + b)) * ns_betacf(b, a, y);
    return lower ? 1.0 - J : J;
}

/* P(Y > q), Y ~ NB(size, mu) : pnbinom(q, size, mu=mu, lower.tail=FALSE) */
NS_HD double ns_pnbinom_mu_upper(double q, double size, double mu)
{
    if (q < 0)
        return 1.0;
    q = floor(q + 1e-7);
    return ns_pbeta_raw(mu / (size + mu), size / (size + mu), q + 1.0, size, 1);
}
NS_HD double ns_pnbinom_mu_lower(double q, double size, double mu)
{
    if (q < 0)
        return 0.0;
    q = floor(q + 1e-7);
    return ns_pbeta_raw(mu / (size + mu), size / (size + mu), q + 1.0, size, 0);
}
NS_HD double ns_pbinom_lower(double q, double n, double p)
{
    if (q < 0)
        return 0.0;
    q = floor(q + 1e-7);
    if (n <= q)
        return 1.0;
    return ns_pbeta_raw(p, 1.0 - p, q + 1.0, n - q, 0);
}

/* continued-fraction tail kept out of line: the callers' hot path is the direct sum below */
#if defined(__CUDACC__)
static __host__ __device__ __noinline__ double ns_nb_pvalue_cf(double y, double size, double mu)
#else
static double ns_nb_pvalue_cf(double y, double size, double mu)
#endif
{
    return ns_dnbinom_mu(y, size, mu) + ns_pnbinom_mu_upper(y, size, mu);
}

/* P(Y >= y) = dnbinom(y) + pnbinom(y, lower=F): the p-value of glm_rob_nb.r:221 and the power
 * p-value of needlestack.r:244-245,250-251.
 * When the pmf ratio (j+size)/(j+1) * mu/(mu+size) is safely below 1 from y on (error-level
 * means: every sample of an exome panel) the tail is summed directly by the pmf recurrence from
 * pmf(0) = (1+mu/size)^(-size): 1 - sum_{j<y} when that is not a cancellation, else sum_{j>=y}
 * until the terms drop below 1e-17 of the sum.  Otherwise the continued fraction above.
 * rt: optional table of 1/j (j < rtn). */
NS_HD double ns_nb_pvalue(double y, double size, double mu, const double *rt = nullptr, int rtn = 0)
{
    if (y <= 0)
        return 1.0;
    if (!(mu > 0))
        return (mu == 0) ? 0.0 : NAN; /* mu == 0 (depth 0) and y > 0; NaN propagates */
    const double q = mu / (mu + size);
    const double r_y = ((y + size) / (y + 1)) * q;
    if (mu < 500.0 && y < 4096.0 && fmax(r_y, q) < 0.9) {
        double pm = exp(-size * log1p(mu / size));
        double S = 0.0, j = 0.0;
        for (; j < y; j += 1) {
            S += pm;
            double ij = (rt && j + 1 < (double)rtn) ? rt[(int)j + 1] : 1.0 / (j + 1);
            pm *= ((j + size) * ij) * q;
        }
        if (S < 0.5)
            return 1.0 - S;
        double T = 0.0;
        for (int it = 0; it < 4096; it++) {
            T += pm;
            const double t = pm;
            double ij = (rt && j + 1 < (double)rtn) ? rt[(int)j + 1] : 1.0 / (j + 1);
            pm *= ((j + size) * ij) * q;
            j += 1;
            if (t <= 1e-17 * T)
                break;
        }
        return T;
    }
    return ns_nb_pvalue_cf(y, size, mu);
}

/* A lower bound of ns_nb_pvalue at a fraction of its cost: the same value whenever the direct sum
 * 1 - sum_{j<y} pmf(j) applies, else the single term pmf(y) <= P(Y >= y) (no tail summation).  The post kernel
 * uses it to prove that no sample of a unit can reach the call threshold; units that might are re-done exactly. */
NS_HD double ns_nb_pvalue_lb(double y, double size, double mu, const double *rt = nullptr, int rtn = 0)
{
    if (y <= 0)
        return 1.0;
    if (!(mu > 0))
        return (mu == 0) ? 0.0 : NAN;
    const double q = mu / (mu + size);
    const double r_y = ((y + size) / (y + 1)) * q;
    if (mu < 500.0 && y < 4096.0 && fmax(r_y, q) < 0.9) {
        double pm = exp(-size * log1p(mu / size));
        double S = 0.0;
        for (double j = 0.0; j < y; j += 1) {
            S += pm;
            double ij = (rt && j + 1 < (double)rtn) ? rt[(int)j + 1] : 1.0 / (j + 1);
            pm *= ((j + size) * ij) * q;
        }
        return (S < 0.5) ? 1.0 - S : pm;
    }
    return ns_nb_pvalue_cf(y, size, mu);
}

/* ---------------------------------------------------------------------------------
 * Discrete quantiles qnbinom(p,size,mu=) / qbinom(p,n,prob) (needlestack.r:243,249):
 * the smallest k with CDF(k) >= p*(1-64 eps).  The Cornish-Fisher guess only chooses
 * where the search starts; a bracketing + bisection search returns the same k.
 * ------------------------------------------------------------------------------- */
NS_HD double ns_qnorm_lower(double p)
{
    /* Acklam's rational approximation (|rel err| < 1.2e-9): starting point only */
    const double a0 = -3.969683028665376e+01, a1 = 2.209460984245205e+02, a2 = -2.759285104469687e+02,
                 a3 = 1.383577518672690e+02, a4 = -3.066479806614716e+01, a5 = 2.506628277459239e+00;
    const double b0 = -5.447609879822406e+01, b1 = 1.615858368580409e+02, b2 = -1.556989798598866e+02,
                 b3 = 6.680131188771972e+01, b4 = -1.328068155288572e+01;
    const double c0 = -7.784894002430293e-03, c1 = -3.223964580411365e-01, c2 = -2.400758277161838e+00,
                 c3 = -2.549732539343734e+00, c4 = 4.374664141464968e+00, c5 = 2.938163982698783e+00;
    const double d0 = 7.784695709041462e-03, d1 = 3.224671290700398e-01, d2 = 2.445134137142996e+00,
                 d3 = 3.754408661907416e+00;
    if (p < 0.02425) {
        double q = sqrt(-2 * log(p));
        return (((((c0 * q + c1) * q + c2) * q + c3) * q + c4) * q + c5) /
               ((((d0 * q + d1) * q + d2) * q + d3) * q + 1);
    }
    if (p <= 1 - 0.02425) {
        double q = p - 0.5, r = q * q;
        return (((((a0 * r + a1) * r + a2) * r + a3) * r + a4) * r + a5) * q /
               (((((b0 * r + b1) * r + b2) * r + b3) * r + b4) * r + 1);
    }
    double q = sqrt(-2 * log(1 - p));
    return -(((((c0 * q + c1) * q + c2) * q + c3) * q + c4) * q + c5) /
           ((((d0 * q + d1) * q + d2) * q + d3) * q + 1);
}

/* KIND 0: NB(size=p1, mu=p2); KIND 1: Binom(n=p1, prob=p2) */
template <int KIND> NS_HD double ns_cdf(double k, double p1, double p2)
{
    return KIND == 0 ? ns_pnbinom_mu_lower(k, p1, p2) : ns_pbinom_lower(k, p1, p2);
}

template <int KIND> NS_HD double ns_discrete_quantile(double y0, double p, double p1, double p2, double ymax)
{
    p *= 1 - 64 * NS_DBL_EPS;
    double y = y0;
    if (y < 0)
        y = 0;
    if (y > ymax)
        y = ymax;
    /* invariant sought: cdf(lo) < p <= cdf(hi); lo = -1 stands for "below the support" */
    double lo, hi;
    if (ns_cdf<KIND>(y, p1, p2) >= p) {
        hi = y;
        double step = 1;
        lo = y - step;
        while (lo >= 0 && ns_cdf<KIND>(lo, p1, p2) >= p) {
            hi = lo;
            step *= 2;
            lo = hi - step;
        }
        if (lo < 0)
            lo = -1;
    } else {
        lo = y;
        double step = 1;
        hi = y + step;
        while (hi < ymax && ns_cdf<KIND>(hi, p1, p2) < p) {
            lo = hi;
            step *= 2;
            hi = lo + step;
        }
        if (hi > ymax)
            hi = ymax;
    }
    while (hi - lo > 1) {
        double mid = floor(0.5 * (lo + hi));
        if (ns_cdf<KIND>(mid, p1, p2) >= p)
            hi = mid;
        else
            lo = mid;
    }
    return hi;
}

NS_HD double ns_qnbinom_mu(double p, double size, double mu)
{
    double prob = size / (size + mu);
    if (prob == 1 || size == 0 || p <= 0)
        return 0;
    double Q = 1.0 / prob, P = (1.0 - prob) * Q;
    double m = size * P, sigma = sqrt(size * P * Q), gamma = (Q + P) / sigma;
    double z = ns_qnorm_lower(p);
    double y = nearbyint(m + sigma * (z + gamma * (z * z - 1) / 6));
    return ns_discrete_quantile<0>(y, p, size, mu, 1e15);
}

NS_HD double ns_qbinom(double p, double n, double pr)
{
    if (pr == 0 || n == 0 || p <= 0)
        return 0;
    double q = 1 - pr;
    if (q == 0 || p >= 1)
        return n;
    double m = n * pr, sigma = sqrt(n * pr * q), gamma = (q - pr) / sigma;
    double z = ns_qnorm_lower(p);
    double y = nearbyint(m + sigma * (z + gamma * (z * z - 1) / 6));
    return ns_discrete_quantile<1>(y, p, n, pr, n);
}

/* ---------------------------------------------------------------------------------
 * Fisher exact test, 2x2, two-sided (R fisher.test; needlestack.r:234,236 with
 * matrix(c(Vp,Vm,Rp,Rm),nrow=2)):  p = sum of dhyper over the support where
 * d <= d_obs*(1+1e-7), densities normalised by their own sum.
 * Two passes over the support, no storage.
 * ------------------------------------------------------------------------------- */
NS_HD double ns_dhyper_log(double x, double r, double b, double n)
{
    if (x < 0 || n < x || r < x || n - x > b)
        return -INFINITY;
    if (n == 0)
        return (x == 0) ? 0.0 : -INFINITY;
    double p = n / (r + b), q = (r + b - n) / (r + b);
    return ns_dbinom_raw_log(x, r, p, q) + ns_dbinom_raw_log(n - x, b, p, q) -
           ns_dbinom_raw_log(n, r + b, p, q);
}

/* hypergeometric pmf ratio d(s+1)/d(s) for dhyper(s, m, n, k) */
NS_HD double ns_hyper_up(double s, double m, double n, double k)
{
    return ((m - s) * (k - s)) / ((s + 1) * (n - k + s + 1));
}

/* sum of the densities (relative to the mode) over [a, b] (a <= b inside the support), walking
 * away from `from` whose relative density is d_from; returns also the density at x if visited.
 * thr < 0: plain sum; thr >= 0: only terms <= thr are added. */
NS_HD double ns_hyper_walk(double a, double b, double from, double d_from, double m, double n, double k, double thr,
                           double x, double &d_at_x)
{
    double acc = 0;
    if (from <= a) { /* upward a..b starting at a == from */
        double d = d_from;
        for (double s = a; s <= b; s += 1) {
            if (s == x)
                d_at_x = d;
            if (thr < 0 || d <= thr)
                acc += d;
            d *= ns_hyper_up(s, m, n, k);
        }
    } else { /* downward b..a starting at b == from */
        double d = d_from;
        for (double s = b; s >= a; s -= 1) {
            if (s == x)
                d_at_x = d;
            if (thr < 0 || d <= thr)
                acc += d;
            d /= ns_hyper_up(s - 1, m, n, k);
        }
    }
    return acc;
}

NS_HD double ns_hyper_mode(double m, double n, double k, double lo, double hi)
{
    double mode = floor((k + 1) * (m + 1) / (m + n + 2));
    return fmin(fmax(mode, lo), hi);
}

/* lane-local version: densities relative to the mode by the ratio recurrence (the observed
 * table's density and every term the 1e-7 slack of R's comparison can matter for are exact to
 * ~support*eps); terms that underflow relative to the mode are below 1e-308 of the total */
NS_HD double ns_fisher_2x2(double a11, double a21, double a12, double a22)
{
    const double m = a11 + a21, n = a12 + a22, k = a11 + a12, x = a11;
    const double lo = fmax(0.0, k - n), hi = fmin(k, m);
    if (hi <= lo)
        return 1.0;
    const double mode = ns_hyper_mode(m, n, k, lo, hi);
    double dobs = (x == mode) ? 1.0 : 0.0, dummy = 0;
    double tot = ns_hyper_walk(mode, hi, mode, 1.0, m, n, k, -1.0, x, dobs);
    if (mode > lo)
        tot += ns_hyper_walk(lo, mode - 1, mode - 1, 1.0 / ns_hyper_up(mode - 1, m, n, k), m, n, k, -1.0, x, dobs);
    const double thr = dobs * (1 + 1e-7);
    double pv = ns_hyper_walk(mode, hi, mode, 1.0, m, n, k, thr, -1.0, dummy);
    if (mode > lo)
        pv += ns_hyper_walk(lo, mode - 1, mode - 1, 1.0 / ns_hyper_up(mode - 1, m, n, k), m, n, k, thr, -1.0, dummy);
    return pv / tot;
}

/* SOR() of needlestack.r:206-215; IEEE arithmetic reproduces R's NaN / Inf cases */
NS_HD double ns_sor(double RO_f, double AO_f, double RO_r, double AO_r)
{
    double refRatio = fmax(RO_f, RO_r) / fmin(RO_f, RO_r);
    double altRatio = fmax(AO_f, AO_r) / fmin(AO_f, AO_r);
    double R = (RO_f / RO_r) * (AO_r / AO_f);
    return log((R + 1 / R) / (refRatio / altRatio));
}
/* R's pmax/pmin propagate NaN (fmax/fmin do not): used where an operand can be NaN */
NS_HD double ns_rmax(double a, double b) { return (a != a || b != b) ? NAN : fmax(a, b); }
NS_HD double ns_rmin(double a, double b) { return (a != a || b != b) ? NAN : fmin(a, b); }

/* out-of-line entry points for code that is rarely reached from a hot kernel */
#if defined(__CUDACC__)
#define NS_OOL static __host__ __device__ __noinline__
#else
#define NS_OOL static
#endif
NS_OOL double ns_fisher_2x2_ool(double a, double b, double c, double d) { return ns_fisher_2x2(a, b, c, d); }
NS_OOL double ns_qnbinom_mu_ool(double p, double size, double mu) { return ns_qnbinom_mu(p, size, mu); }
NS_OOL double ns_qbinom_ool(double p, double n, double pr) { return ns_qbinom(p, n, pr); }
