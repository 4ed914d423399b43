/*
 * tests/emul/ns_emul.cpp — TEST INFRASTRUCTURE ONLY.
 * Serial (one-lane) build of the kernel bodies in needlestack_b200/csrc/ns_core.cuh so
 * that the control flow and numerics of the CUDA path can be checked against the oracle
 * on a machine without a GPU.  Never linked into libns_b200.so.
 */
#define NS_FIT_TRACE 1
#include "../../needlestack_b200/csrc/ns_core.cuh"
#include <stdlib.h>
#include <string.h>
#include <vector>

extern "C" {

/* the same defaults as the library (duplicated on purpose: the emulation must not link it) */
int emul_unit(int n, const int32_t *dp, const int32_t *vp, const int32_t *vm, const int32_t *rp,
              const int32_t *rm, int is_indel, int plain, int fast, const ns_params *prm_in,
              const uint8_t *role, double *pvalue, double *qvalue, double *gq, double *qval_minaf,
              double *sor, double *rvsb, double *fs, uint8_t *gt, uint8_t *status, double *pooled,
              double *sigma, double *slope, double *max_gq, uint32_t *flags_out, uint32_t *iters,
              uint64_t *counters)
{
    ns_params prm = *prm_in;
    prm.role = role;
    ns_rows r;
    r.dp = dp;
    r.vp = vp;
    r.vm = vm;
    r.rp = rp;
    r.rm = rm;
    r.is_indel = is_indel;
    r.valid = 1;
    std::vector<double> buf(ns_work_doubles(n) + 3 * (size_t)n + 16);
    static double rtab[NS_RTAB];
    for (int j = 1; j < NS_RTAB; j++)
        rtab[j] = 1.0 / (double)j;
    ns_work w;
    ns_work_bind(w, buf.data(), n, rtab, ns_h_logtab);
    uint32_t flags = ns_gate_unit<SerialG>(r, n, prm, plain, w.scratch);
    *sigma = NAN;
    *slope = NAN;
    *iters = 0;
    counters[0] = counters[1] = counters[2] = counters[3] = 0;
    if (!(flags & NS_F_GATED)) {
        const bool inv = flags & NS_F_REF_INV, extra = flags & NS_F_EXTRA_ROB;
        double maxmu = -INFINITY;
        for (int i = 0; i < n; i++) {
            double x = ns_ld(dp, i);
            double y = inv ? ns_ld(rp, i) + ns_ld(rm, i) : ns_ld(vp, i) + ns_ld(vm, i);
            maxmu = fmax(maxmu, x);
            bool keep = !(extra && y / x > prm.min_af_extra_rob);
            w.y[i] = y;
            w.x[i] = keep ? x : 0.0;
        }
        ns_fit_result res;
        if (fast)
            ns_fit_unit<SerialG, 1>(w, n, maxmu, prm, res);
        if (!fast || res.deferred) { /* the warp kernel hands such units to the heavy kernel */
            for (int i = 0; i < n; i++) {
                double x = ns_ld(dp, i);
                double y = inv ? ns_ld(rp, i) + ns_ld(rm, i) : ns_ld(vp, i) + ns_ld(vm, i);
                bool keep = !(extra && y / x > prm.min_af_extra_rob);
                w.y[i] = y;
                w.x[i] = keep ? x : 0.0;
            }
            int was_deferred = fast ? res.deferred : 0;
            ns_fit_unit<SerialG, 0>(w, n, maxmu, prm, res);
            res.deferred = was_deferred;
        }
        counters[3] = (uint64_t)res.deferred;
        flags |= res.flags;
        *sigma = res.sigma;
        *slope = res.slope;
        *iters = (uint32_t)res.n_outer | ((uint32_t)res.n_irls << 8) | ((uint32_t)res.n_obj << 18);
        counters[0] = res.cnt.n_seed;
        counters[1] = res.cnt.n_lattice;
        counters[2] = res.cnt.n_quad;
    }
    ns_unit_out o = {pvalue, qvalue, gq, qval_minaf, sor, rvsb, fs, gt, status, pooled};
    double mg = 0;
    flags = ns_post_unit<SerialG>(r, n, prm, flags, *sigma, *slope, nullptr, 0, o,
                                  buf.data() + ns_work_doubles(n), mg);
    *max_gq = mg;
    *flags_out = flags;
    return 0;
}

/* the iterates (sigma, beta after every outer pass) of the kernel's fit of one glmrob.nb(y, x) unit; fast != 0: the
 * warp kernel's table / recurrence path; returns the number of outer passes */
int emul_fit_trace(int n, const int32_t *y, const int32_t *x, int fast, const ns_params *prm, int cap, double *sig_trace,
                   double *beta_trace)
{
    std::vector<double> buf(ns_work_doubles(n) + 16);
    static double rtab[NS_RTAB];
    for (int j = 1; j < NS_RTAB; j++)
        rtab[j] = 1.0 / (double)j;
    ns_work w;
    ns_work_bind(w, buf.data(), n, rtab, ns_h_logtab);
    double maxmu = -INFINITY;
    for (int i = 0; i < n; i++) {
        w.y[i] = (double)y[i];
        w.x[i] = (double)x[i];
        maxmu = fmax(maxmu, (double)x[i]);
    }
    ns_fit_trace_sig = sig_trace;
    ns_fit_trace_beta = beta_trace;
    ns_fit_trace_cap = cap;
    ns_fit_result res;
    if (fast)
        ns_fit_unit<SerialG, 1>(w, n, maxmu, *prm, res);
    else
        ns_fit_unit<SerialG, 0>(w, n, maxmu, *prm, res);
    ns_fit_trace_sig = ns_fit_trace_beta = nullptr;
    ns_fit_trace_cap = 0;
    return res.deferred ? -1 : res.n_outer;
}

/* numeric layer probes */
double emul_dnbinom_mu(double x, double size, double mu) { return ns_dnbinom_mu(x, size, mu); }
double emul_nb_pvalue(double y, double size, double mu) { return ns_nb_pvalue(y, size, mu); }
double emul_digamma(double x) { return ns_digamma(x); }
double emul_fexp(double x) { return ns_fexp(x); }
double emul_flog(double x) { return ns_flog(x); }
double emul_nb_pvalue_lb(double y, double size, double mu) { return ns_nb_pvalue_lb(y, size, mu); }
double emul_flog_tab(double x) { return ns_flog_tab(x, ns_h_logtab); }
double emul_fexp_bf(double x) { return ns_fexp_bf(x); }
double emul_qnbinom_mu(double p, double size, double mu) { return ns_qnbinom_mu(p, size, mu); }
double emul_qbinom(double p, double n, double pr) { return ns_qbinom(p, n, pr); }
double emul_fisher(double a, double b, double c, double d) { return ns_fisher_2x2(a, b, c, d); }
double emul_ai(double mu, double sig, double c, int n_ai)
{
    ns_fit_counters cnt = {0, 0, 0, 0, 0};
    return ns_ai_sig_tukey(mu, sig, c, n_ai, ns_digamma(1 / sig), cnt);
}
void emul_E12(double mu, double sig, double c, int n_ai, double *e1, double *e2)
{
    ns_fit_counters cnt = {0, 0, 0, 0, 0};
    ns_E_tukeypsi_12(mu, sig, c, n_ai, cnt, *e1, *e2);
}
}
