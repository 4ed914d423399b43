/*
 * ns_quad.cuh — natural cubic spline + QUADPACK dqags on the device.
 *
 * glm_rob_nb.r:115-118, 131-134, 151-154 switch, for windows wider than
 * 2*n_ai.sig.tukey lattice points, from the lattice sum to
 *     integrate(splinefun(x = j12, y = yj12, method = "natural"), min(j12), max(j12)+1)
 * R's integrate() is QUADPACK dqags (21-point Gauss-Kronrod, Wynn epsilon
 * extrapolation, limit = 100, epsabs = epsrel = eps^0.25) and its tolerance is loose
 * enough (1.2e-4) that the quadrature has to be reproduced step for step to stay
 * within 1e-6 of the reference.  This is that algorithm, lane
 * local, with the subdivision list bounded by NS_QLIMIT entries held in the lane's
 * local memory; a quadrature that would need more intervals reports ier = 7
 * (overflow of the on-chip list) instead of silently deviating.
 */
#pragma once
#include "ns_math.cuh"

#define NS_MAX_KNOTS 128 /* n_ai.sig.tukey <= 128 (reference default 100) */
#define NS_QLIMIT 24     /* on-chip subdivision list; R's limit is 100 */

/* The spline through (t_k, y_k), k = 0..n-1, t_k = round(seq(j1, j2, length.out = n))
 * (glm_rob_nb.r:116).  x, y and the second-derivative array c are stored per lane. */
struct ns_spline {
    int n;
    double j1, j2, by;
    double x[NS_MAX_KNOTS];
    double y[NS_MAX_KNOTS];
    double c[NS_MAX_KNOTS];  /* R's c[] before the final 3x scaling */
    double bm[NS_MAX_KNOTS]; /* eliminated diagonal (depends on the knots only) */
};

NS_HD double ns_knot(const ns_spline &s, int k) { return s.x[k]; }

NS_HD void ns_spline_init(ns_spline &s, int n, double j1, double j2)
{
    s.n = n;
    s.j1 = j1;
    s.j2 = j2;
    s.by = (j2 - j1) / (double)(n - 1);
    /* seq(j1, j2, length.out = n): c(from, from + (1:(n-2))*by, to), then round() */
    s.x[0] = j1;
    for (int k = 1; k < n - 1; k++)
        s.x[k] = nearbyint(j1 + (double)k * s.by);
    s.x[n - 1] = j2;
}

/* forward elimination of the diagonal: only the knot spacing enters (R splines.c
 * natural_spline, the b[] updates).  1-based i = 2..n-1 maps to 0-based 1..n-2. */
NS_HD void ns_spline_diag(ns_spline &s)
{
    const int n = s.n;
    double dprev = s.x[1] - s.x[0]; /* d[1] */
    double bprev = 0;
    for (int i = 1; i <= n - 2; i++) {
        double di = s.x[i + 1] - s.x[i];
        double bi = 2 * (dprev + di);
        if (i >= 2) {
            double t = dprev / bprev;
            bi = bi - t * dprev;
        }
        s.bm[i] = bi;
        bprev = bi;
        dprev = di;
    }
}

/* right-hand side + elimination + back substitution for the current y[] */
NS_HD void ns_spline_solve(ns_spline &s)
{
    const int n = s.n;
    /* c[i] = (y[i+1]-y[i])/d[i] - (y[i]-y[i-1])/d[i-1] */
    double dprev = s.x[1] - s.x[0];
    double sprev = (s.y[1] - s.y[0]) / dprev;
    double cprev = 0;
    for (int i = 1; i <= n - 2; i++) {
        double di = s.x[i + 1] - s.x[i];
        double snext = (s.y[i + 1] - s.y[i]) / di;
        double ci = snext - sprev;
        if (i >= 2) {
            double t = dprev / s.bm[i - 1];
            ci = ci - t * cprev;
        }
        s.c[i] = ci;
        cprev = ci;
        sprev = snext;
        dprev = di;
    }
    s.c[n - 2] = s.c[n - 2] / s.bm[n - 2];
    for (int i = n - 3; i >= 1; i--) {
        double di = s.x[i + 1] - s.x[i];
        s.c[i] = (s.c[i] - di * s.c[i + 1]) / s.bm[i];
    }
    s.c[0] = 0.0;
    s.c[n - 1] = 0.0;
}

/* spline_eval(): y[i] + dx*(b[i] + dx*(c[i] + dx*d[i])) with b, c, d of R's
 * natural_spline rebuilt from (y, c) of the interval; linear beyond the last knot */
NS_HD double ns_spline_eval(const ns_spline &s, double u)
{
    const int n = s.n;
    int i = (int)((u - s.j1) / s.by);
    if (i < 0)
        i = 0;
    if (i > n - 1)
        i = n - 1;
    while (i > 0 && u < s.x[i])
        i--;
    while (i < n - 1 && u >= s.x[i + 1])
        i++;
    double xi = s.x[i];
    double dx = u - xi;
    if (i == n - 1) {
        double dl = xi - s.x[n - 2];
        double bl = (s.y[n - 1] - s.y[n - 2]) / dl + dl * s.c[n - 2];
        return s.y[i] + dx * bl;
    }
    double di = s.x[i + 1] - xi;
    double c0 = s.c[i], c1 = s.c[i + 1];
    double b = (s.y[i + 1] - s.y[i]) / di - di * (c1 + 2.0 * c0);
    double d = (c1 - c0) / di;
    if (u < s.j1)
        d = 0.0;
    return s.y[i] + dx * (b + dx * (3.0 * c0 + dx * d));
}

/* ---- QUADPACK ------------------------------------------------------------------ */
struct ns_quad_out {
    double result, abserr;
    int neval, ier, last;
};

template <class F>
NS_HD void ns_dqk21(const F &f, double a, double b, double &result, double &abserr, double &resabs,
                    double &resasc)
{
    double fv1[10], fv2[10];
    const double epmach = NS_DBL_EPS, uflow = NS_DBL_MIN;
    double centr = (a + b) * .5, hlgth = (b - a) * .5, dhlgth = fabs(hlgth);
    double resg = 0.0;
    double fc = f(centr);
    double resk = NS_TBL(wgk)[10] * fc;
    resabs = fabs(resk);
    for (int j = 0; j < 5; j++) {
        int jtw = 2 * j + 1;
        double absc = hlgth * NS_TBL(xgk)[jtw];
        double fval1 = f(centr - absc), fval2 = f(centr + absc);
        fv1[jtw] = fval1;
        fv2[jtw] = fval2;
        double fsum = fval1 + fval2;
        resg += NS_TBL(wg)[j] * fsum;
        resk += NS_TBL(wgk)[jtw] * fsum;
        resabs += NS_TBL(wgk)[jtw] * (fabs(fval1) + fabs(fval2));
    }
    for (int j = 0; j < 5; j++) {
        int jtwm1 = 2 * j;
        double absc = hlgth * NS_TBL(xgk)[jtwm1];
        double fval1 = f(centr - absc), fval2 = f(centr + absc);
        fv1[jtwm1] = fval1;
        fv2[jtwm1] = fval2;
        double fsum = fval1 + fval2;
        resk += NS_TBL(wgk)[jtwm1] * fsum;
        resabs += NS_TBL(wgk)[jtwm1] * (fabs(fval1) + fabs(fval2));
    }
    double reskh = resk * .5;
    resasc = NS_TBL(wgk)[10] * fabs(fc - reskh);
    for (int j = 0; j < 10; j++)
        resasc += NS_TBL(wgk)[j] * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
    result = resk * hlgth;
    resabs *= dhlgth;
    resasc *= dhlgth;
    abserr = fabs((resk - resg) * hlgth);
    if (resasc != 0.0 && abserr != 0.0) {
        double r = abserr * 200.0 / resasc;
        abserr = resasc * fmin(1.0, r * sqrt(r)); /* pow(r, 1.5) */
    }
    if (resabs > uflow / (epmach * 50.0))
        abserr = fmax(epmach * 50.0 * resabs, abserr);
}

/* keeps iord (1-based) sorted by decreasing elist; QUADPACK dqpsrt */
NS_HD void ns_dqpsrt(int limit, int last, int &maxerr, double &ermax, const double *elist, int *iord,
                     int &nrmax)
{
    if (last <= 2) {
        iord[1] = 1;
        iord[2] = 2;
    } else {
        double errmax = elist[maxerr];
        if (nrmax > 1) {
            int ido = nrmax - 1;
            for (int i = 1; i <= ido; ++i) {
                int isucc = iord[nrmax - 1];
                if (errmax <= elist[isucc])
                    break;
                iord[nrmax] = isucc;
                --nrmax;
            }
        }
        int jupbn = (last > limit / 2 + 2) ? limit + 3 - last : last;
        double errmin = elist[last];
        int jbnd = jupbn - 1;
        int ibeg = nrmax + 1;
        bool done = false;
        for (int i = ibeg; i <= jbnd && !done; ++i) {
            int isucc = iord[i];
            if (errmax >= elist[isucc]) {
                iord[i - 1] = maxerr;
                int k = jbnd;
                bool placed = false;
                for (int j = i; j <= jbnd; j++, k--) {
                    isucc = iord[k];
                    if (errmin < elist[isucc]) {
                        iord[k + 1] = last;
                        placed = true;
                        break;
                    }
                    iord[k + 1] = isucc;
                }
                if (!placed)
                    iord[i] = last;
                done = true;
                break;
            }
            iord[i - 1] = isucc;
        }
        if (!done) {
            iord[jbnd] = maxerr;
            iord[jupbn] = last;
        }
    }
    maxerr = iord[nrmax];
    ermax = elist[maxerr];
}

/* Wynn epsilon algorithm; epstab / res3la 1-based; QUADPACK dqelg */
NS_HD void ns_dqelg(int &n, double *epstab, double &result, double &abserr, double *res3la, int &nres)
{
    const double epmach = NS_DBL_EPS, oflow = NS_DBL_MAX;
    const int limexp = 50;
    ++nres;
    abserr = oflow;
    result = epstab[n];
    if (n >= 3) {
        epstab[n + 2] = epstab[n];
        int newelm = (n - 1) / 2;
        epstab[n] = oflow;
        int num = n, k1 = n;
        for (int i = 1; i <= newelm; ++i) {
            int k2 = k1 - 1, k3 = k1 - 2;
            double res = epstab[k1 + 2];
            double e0 = epstab[k3], e1 = epstab[k2], e2 = res;
            double e1abs = fabs(e1);
            double delta2 = e2 - e1, err2 = fabs(delta2);
            double tol2 = fmax(fabs(e2), e1abs) * epmach;
            double delta3 = e1 - e0, err3 = fabs(delta3);
            double tol3 = fmax(e1abs, fabs(e0)) * epmach;
            if (err2 <= tol2 && err3 <= tol3) {
                result = res;
                abserr = err2 + err3;
                abserr = fmax(abserr, epmach * 5. * fabs(result));
                return;
            }
            double e3 = epstab[k1];
            epstab[k1] = e1;
            double delta1 = e1 - e3, err1 = fabs(delta1);
            double tol1 = fmax(e1abs, fabs(e3)) * epmach;
            double ss = 0;
            bool ok = false;
            if (err1 > tol1 && err2 > tol2 && err3 > tol3) {
                ss = 1. / delta1 + 1. / delta2 - 1. / delta3;
                double epsinf = fabs(ss * e1);
                if (epsinf > 1e-4)
                    ok = true;
            }
            if (!ok) {
                n = i + i - 1;
                break;
            }
            res = e1 + 1. / ss;
            epstab[k1] = res;
            k1 -= 2;
            double errA = err2 + fabs(res - e2) + err3;
            if (errA <= abserr) {
                abserr = errA;
                result = res;
            }
        }
        if (n == limexp)
            n = (limexp / 2 << 1) - 1;
        int ib = ((num / 2 << 1) == num) ? 2 : 1;
        int ie = newelm + 1;
        for (int i = 1; i <= ie; ++i) {
            int ib2 = ib + 2;
            epstab[ib] = epstab[ib2];
            ib = ib2;
        }
        if (num != n) {
            int indx = num - n + 1;
            for (int i = 1; i <= n; ++i) {
                epstab[i] = epstab[indx];
                ++indx;
            }
        }
        if (nres >= 4) {
            abserr = fabs(result - res3la[3]) + fabs(result - res3la[2]) + fabs(result - res3la[1]);
            res3la[1] = res3la[2];
            res3la[2] = res3la[3];
            res3la[3] = result;
        } else {
            res3la[nres] = result;
            abserr = oflow;
        }
    }
    abserr = fmax(abserr, epmach * 5. * fabs(result));
}

/* dqagse with R's limit (100) semantics but an on-chip list of NS_QLIMIT intervals.
 * ier: 0 ok, 1..6 as QUADPACK (an R error), 7 = on-chip list exhausted. */
struct ns_k21_serial {
    template <class F>
    static NS_HD void eval(const F &f, double a, double b, double &result, double &abserr, double &resabs,
                           double &resasc)
    {
        ns_dqk21(f, a, b, result, abserr, resabs, resasc);
    }
};

/* ROLLED: the two halves of a bisection go through ONE copy of the 21-point rule (a loop that is not unrolled) instead
 * of two inlined ones that the compiler interleaves: smaller code for the throughput form of the cluster kernel, which
 * is bound by instruction fetch; the latency form keeps the interleaved pair (C3 444 -> 429 ms rolled, C1 10.3 -> 10.8) */
template <class K21, class F, bool ROLLED = false>
NS_HD void ns_dqagse_t(const F &f, double a, double b, double epsabs, double epsrel, ns_quad_out &out)
{
    const int limit = 100; /* R's integrate() default; decides jupbnd / ier=1 */
    double alist[NS_QLIMIT + 2], blist[NS_QLIMIT + 2], rlist[NS_QLIMIT + 2], elist[NS_QLIMIT + 2];
    int iord[NS_QLIMIT + 3];
    double rlist2[NS_QLIMIT + 6], res3la[4];
    const double epmach = NS_DBL_EPS, uflow = NS_DBL_MIN, oflow = NS_DBL_MAX;
    double result = 0, abserr = 0, defabs, resabs, dres, errbnd;
    double area = 0, area1, area2, area12, a1, a2, b1, b2, error1, error2, erro12, errmax = 0,
           errsum = 0, erlast, erlarg = 0, ertest = 0, defab1, defab2, correc = 0, small = 0,
           reseps, abseps;
    int ier = 0, ierro = 0, last = 0, maxerr = 1, nrmax = 1, nres = 0, numrl2 = 2, ktmin = 0,
        iroff1 = 0, iroff2 = 0, iroff3 = 0, ksgn, extrap = 0, noext = 0, k, id, jupbnd;
    alist[1] = a;
    blist[1] = b;
    rlist[1] = 0;
    elist[1] = 0;
    res3la[1] = res3la[2] = res3la[3] = 0;
    K21::eval(f, a, b, result, abserr, defabs, resabs);
    dres = fabs(result);
    errbnd = fmax(epsabs, epsrel * dres);
    last = 1;
    rlist[1] = result;
    elist[1] = abserr;
    iord[1] = 1;
    if (abserr <= epmach * 100. * defabs && abserr > errbnd)
        ier = 2;
    if (ier != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.) {
        out.result = result;
        out.abserr = abserr;
        out.neval = 21;
        out.ier = ier;
        out.last = 1;
        return;
    }
    rlist2[1] = result;
    errmax = abserr;
    maxerr = 1;
    area = result;
    errsum = abserr;
    abserr = oflow;
    ksgn = -1;
    if (dres >= (1. - epmach * 50.) * defabs)
        ksgn = 1;
    bool sum_exit = false; /* QUADPACK label 115 */
    bool overflow = false;
    for (last = 2; last <= limit; ++last) {
        if (last > NS_QLIMIT) {
            overflow = true;
            break;
        }
        a1 = alist[maxerr];
        b1 = (alist[maxerr] + blist[maxerr]) * .5;
        a2 = b1;
        b2 = blist[maxerr];
        erlast = errmax;
        if (ROLLED) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int h = 0; h < 2; h++) {
                double ar, er, df;
                K21::eval(f, h ? a2 : a1, h ? b2 : b1, ar, er, resabs, df);
                if (h == 0) {
                    area1 = ar;
                    error1 = er;
                    defab1 = df;
                } else {
                    area2 = ar;
                    error2 = er;
                    defab2 = df;
                }
            }
        } else {
            K21::eval(f, a1, b1, area1, error1, resabs, defab1);
            K21::eval(f, a2, b2, area2, error2, resabs, defab2);
        }
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
        ns_dqpsrt(limit, last, maxerr, errmax, elist, iord, nrmax);
        if (errsum <= errbnd) {
            sum_exit = true;
            break;
        }
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
            bool found = false;
            for (k = id; k <= jupbnd; ++k) {
                maxerr = iord[nrmax];
                errmax = elist[maxerr];
                if (fabs(blist[maxerr] - alist[maxerr]) > small) {
                    found = true;
                    break;
                }
                ++nrmax;
            }
            if (found)
                continue;
        }
        ++numrl2;
        rlist2[numrl2] = area;
        ns_dqelg(numrl2, rlist2, reseps, abseps, res3la, nres);
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
    if (overflow) {
        last = NS_QLIMIT;
        sum_exit = true;
        ier = 7;
    }
    if (last > limit)
        last = limit;
    bool to130 = false;
    if (!sum_exit) {
        /* label 100 */
        if (abserr == oflow) {
            sum_exit = true;
        } else if (ier + ierro == 0) {
            /* label 110 */
            if (!(ksgn == -1 && fmax(fabs(result), fabs(area)) <= defabs * .01))
                if (.01 > result / area || result / area > 100. || errsum > fabs(area))
                    ier = 6;
            to130 = true;
        } else {
            if (ierro == 3)
                abserr += correc;
            if (ier == 0)
                ier = 3;
            bool to110 = false;
            if (result == 0. || area == 0.) {
                if (abserr > errsum)
                    sum_exit = true;
                else if (area == 0.)
                    to130 = true;
                else
                    to110 = true;
            } else if (abserr / fabs(result) > errsum / fabs(area)) {
                sum_exit = true;
            } else {
                to110 = true;
            }
            if (to110) {
                if (!(ksgn == -1 && fmax(fabs(result), fabs(area)) <= defabs * .01))
                    if (.01 > result / area || result / area > 100. || errsum > fabs(area))
                        ier = 6;
                to130 = true;
            }
        }
    }
    if (sum_exit) {
        result = 0.;
        for (k = 1; k <= last; ++k)
            result += rlist[k];
        abserr = errsum;
        to130 = true;
    }
    if (to130 && ier > 2 && ier != 7)
        --ier;
    out.result = result;
    out.abserr = abserr;
    out.neval = last * 42 - 21;
    out.ier = ier;
    out.last = last;
}

template <class F>
NS_HD void ns_dqagse(const F &f, double a, double b, double epsabs, double epsrel, ns_quad_out &out)
{
    ns_dqagse_t<ns_k21_serial>(f, a, b, epsabs, epsrel, out);
}

struct ns_spline_fn {
    const ns_spline *s;
    NS_HD double operator()(double u) const { return ns_spline_eval(*s, u); }
};

/* integrate(splinefun(...), lower = min(j12), upper = max(j12) + 1) with R's defaults */
NS_HD double ns_integrate_spline(const ns_spline &s, int &ier, int &npanels)
{
    const double tol = 1.220703125e-4; /* .Machine$double.eps^0.25 */
    ns_spline_fn f = {&s};
    ns_quad_out q;
    ns_dqagse(f, s.j1, s.j2 + 1.0, tol, tol, q);
    ier = q.ier;
    npanels += q.neval / 21;
    return q.result;
}
