/* ns_k_gate.cu — streaming gate kernel, row packer, scan tail, DFMA probe (sm_100a) */
#include "ns_kernels.cuh"

extern __shared__ double ns_smem[];

/* y/x > t with R's semantics for integer counts (0/0 = NaN compares false, y/0 = Inf compares true).
 * The product test decides whenever y and t*x are separated by far more than rounding; the division
 * only runs for the (practically never occurring) near-tie, so the result is exactly that of y/x > t. */
__device__ __forceinline__ bool ns_ratio_gt(int y, int x, double t)
{
    if (x == 0)
        return y > 0;
    const double p = t * (double)x, yd = (double)y;
    const double diff = yd - p;
    if (fabs(diff) > 1e-6 * fabs(p) + 1e-300)
        return diff > 0;
    return yd / (double)x > t;
}

/* y/x >= t, same conventions */
__device__ __forceinline__ bool ns_ratio_ge(int y, int x, double t)
{
    if (x == 0)
        return y > 0;
    const double p = t * (double)x, yd = (double)y;
    const double diff = yd - p;
    if (fabs(diff) > 1e-6 * fabs(p) + 1e-300)
        return diff > 0;
    return yd / (double)x >= t;
}

__device__ __forceinline__ void ns_gate_commit(const ns_unit_arrays &ua, ns_counters *cnt, const ns_params &prm, int u,
                                               uint32_t flags, int is_indel, double mu_hint)
{
    const bool gated = flags & NS_F_GATED, skipped = flags & NS_F_SKIPPED;
    const bool out_all = (!is_indel) && prm.output_all_SNVs;
    const bool need = !skipped && (!gated || prm.emit_mode == NS_EMIT_ALL || out_all);
    ua.flags[u] = flags;
    ua.sigma[u] = NAN;
    ua.slope[u] = NAN;
    ua.max_gq[u] = 0.0;
    ua.iters[u] = 0;
    ua.cycles[u] = 0;
    ua.needrow[u] = need;
    ua.emitflag[u] = 0;
    if (!gated) {
        atomicAdd(&cnt->n_fit, 1);
        /* A window of the first sigma objective evaluation (sigma = maxsig, c = c.tukey.sig) exceeds
         * 2*n_ai.sig.tukey lattice points once mu*(1 + c*sqrt(maxsig + 1/mu)) does: such units go
         * straight to the cluster-per-unit kernel, which runs next to the warp kernel on its own
         * stream.  Over-prediction only costs efficiency. */
        const double kwin = mu_hint * (1.0 + prm.c_tukey_sig * sqrt(prm.maxsig + 1.0 / fmax(mu_hint, 1e-300)));
        if (kwin > 1.0 * prm.n_ai_sig_tukey) /* half the switch point: the slope usually grows from its first estimate */
            ua.pre_list[atomicAdd(&cnt->n_pre, 1)] = u;
        else
            ua.fit_list[atomicAdd(&cnt->n_light, 1)] = u;
    }
    if (need)
        ua.post_list[atomicAdd(&cnt->n_post, 1)] = u;
}

/* Streaming gate for SNV planes without the extra-robust pre-filter (the caller's default):
 * warp = position; the depth row is scanned once for the three alleles; all tests are integer
 * compares and warp vote / reduce instructions; medians come from counts, not from a sort:
 *   median(x) < c  <=>  both middle order statistics < c  (count(x < c) decides, and when it equals
 *   n/2 for even n the two middle values are max{x < c} and min{x >= c}). */
__device__ __forceinline__ void ns_gate_position(const ns_unit_src &src, const ns_params &prm, int64_t p, int u_first,
                                                 int n_units, const ns_unit_arrays &ua, ns_counters *cnt)
{
    const int n = src.n, lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int ref = src.ref[p];
    if (ref > 3) {
        if (lane < 3 && u_first + lane < n_units)
            ns_gate_commit(ua, cnt, prm, u_first + lane, NS_F_SKIPPED | NS_F_GATED, 0, 0.0);
        return;
    }
    const int32_t *base = src.counts + p * 9 * (int64_t)n;
    const double mc = prm.min_coverage;
    /* depth row: samples above min_coverage, order statistics around it, max depth */
    int n_cov = 0, n_less = 0, max_less = -1, min_geq = 0x7fffffff, maxx = 0;
    for (int i = lane; i < n; i += 32) {
        const int x = base[i];
        n_cov += (double)x > mc;
        const bool less = (double)x < mc;
        n_less += less;
        max_less = less ? max(max_less, x) : max_less;
        min_geq = less ? min_geq : min(min_geq, x);
        maxx = max(maxx, x);
    }
    n_cov = __reduce_add_sync(FULL, n_cov);
    n_less = __reduce_add_sync(FULL, n_less);
    max_less = __reduce_max_sync(FULL, max_less);
    min_geq = __reduce_min_sync(FULL, min_geq);
    maxx = __reduce_max_sync(FULL, maxx);
    bool med_low; /* median(x) < min_coverage */
    if (n & 1)
        med_low = n_less >= (n + 1) / 2;
    else if (n_less >= n / 2 + 1)
        med_low = true;
    else if (n_less <= n / 2 - 1)
        med_low = false;
    else
        med_low = ((double)max_less + (double)min_geq) / 2 < mc;
    const bool pos_gate = med_low || n_cov < prm.size_min || n == 0;
    const int32_t *rp = base + (1 + ref) * (int64_t)n, *rm = base + (5 + ref) * (int64_t)n;
    for (int k = 0; k < 3; k++) {
        const int u = u_first + k;
        if (u >= n_units)
            break;
        const int alt = k + (k >= ref ? 1 : 0);
        const int32_t *vp = base + (1 + alt) * (int64_t)n, *vm = base + (5 + alt) * (int64_t)n;
        int cnt_inv = 0, pos_a = 0, pos_r = 0, max_a = 0, max_r = 0, af_a = 0, af_r = 0;
        for (int i = lane; i < n; i += 32) {
            const int x = base[i], ma = vp[i] + vm[i], rr = rp[i] + rm[i];
            cnt_inv += ns_ratio_gt(ma, x, 0.8);
            pos_a += ma > 0;
            pos_r += rr > 0;
            max_a = max(max_a, ma);
            max_r = max(max_r, rr);
            if (prm.min_af > 0) { /* max(y/x, na.rm=T) < min_af  <=>  no defined ratio reaches min_af */
                af_a |= ns_ratio_ge(ma, x, prm.min_af);
                af_r |= ns_ratio_ge(rr, x, prm.min_af);
            }
        }
        cnt_inv = __reduce_add_sync(FULL, cnt_inv);
        const bool inv = (double)cnt_inv > 0.5 * (double)n; /* needlestack.r:330 */
        const int npos = __reduce_add_sync(FULL, inv ? pos_r : pos_a);
        const int maxy = __reduce_max_sync(FULL, inv ? max_r : max_a);
        bool gate = pos_gate || (double)maxy < prm.min_reads || npos == 0;
        if (prm.min_af > 0)
            gate = gate || !__any_sync(FULL, inv ? af_r : af_a);
        uint32_t flags = (inv ? NS_F_REF_INV : 0u) | (gate ? NS_F_GATED : 0u);
        double mu_hint = 0;
        if (!gate) { /* mean allelic fraction over the samples in the fit (x > 0) times the largest depth */
            float s = 0.f;
            int nf = 0;
            for (int i = lane; i < n; i += 32) {
                const int x = base[i];
                const int y = inv ? rp[i] + rm[i] : vp[i] + vm[i];
                if (x > 0) {
                    s += __fdividef((float)y, (float)x);
                    nf++;
                }
            }
            for (int o = 16; o > 0; o >>= 1)
                s += __shfl_xor_sync(FULL, s, o);
            nf = __reduce_add_sync(FULL, nf);
            mu_hint = nf > 0 ? 1.001 * (double)s / (double)nf * (double)maxx : 0.0;
        }
        if (lane == 0)
            ns_gate_commit(ua, cnt, prm, u, flags, 0, mu_hint);
    }
}

__global__ void __launch_bounds__(NS_WPB_GATE * 32)
k_gate(ns_unit_src src, ns_params prm, int64_t u0, int n_units, ns_unit_arrays ua, ns_counters *cnt)
{
    const int n = src.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wstride = gridDim.x * NS_WPB_GATE;
    if (src.mode == 0 && !prm.extra_rob && (u0 % 3) == 0) {
        /* fast path: warp = position */
        const int n_pos = (n_units + 2) / 3;
        for (int q = blockIdx.x * NS_WPB_GATE + warp; q < n_pos; q += wstride)
            ns_gate_position(src, prm, u0 / 3 + q, 3 * q, n_units, ua, cnt);
        return;
    }
    double *scratch = ns_smem + (size_t)warp * (size_t)n;
    for (int u = blockIdx.x * NS_WPB_GATE + warp; u < n_units; u += wstride) {
        ns_rows r = ns_unit_rows(src, u0 + u);
        double mu_hint = 0;
        uint32_t flags = ns_gate_unit<WarpG>(r, n, prm, src.plain, scratch, &mu_hint);
        if (lane == 0)
            ns_gate_commit(ua, cnt, prm, u, flags, r.is_indel, mu_hint);
    }
}

__global__ void k_scan_total(const int *flag, const int *excl, int n, int *total)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *total = n > 0 ? excl[n - 1] + flag[n - 1] : 0;
}

/* rows of emitted units, in unit order, into the packed planes */
__global__ void k_pack(int n, int n_units, int64_t u0, int64_t emit_off, int cap, ns_unit_arrays ua,
                       ns_planes src, ns_planes dst, int64_t *emit_unit, int *emit_index_out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int u = blockIdx.x * wpb + warp; u < n_units; u += gridDim.x * wpb) {
        if (!ua.emitflag[u]) {
            if (lane == 0)
                emit_index_out[u] = -1;
            continue;
        }
        const int d = ua.emitidx[u];
        if (lane == 0)
            emit_index_out[u] = (int)(emit_off + d);
        if (d >= cap)
            continue;
        const size_t so = (size_t)ua.row[u] * n, dof = (size_t)d * n;
        for (int i = lane; i < n; i += 32) {
            dst.pvalue[dof + i] = src.pvalue[so + i];
            dst.qvalue[dof + i] = src.qvalue[so + i];
            dst.gq[dof + i] = src.gq[so + i];
            dst.qval_minaf[dof + i] = src.qval_minaf[so + i];
            dst.sor[dof + i] = src.sor[so + i];
            dst.rvsb[dof + i] = src.rvsb[so + i];
            dst.fs[dof + i] = src.fs[so + i];
            dst.gt[dof + i] = src.gt[so + i];
            dst.status[dof + i] = src.status[so + i];
        }
        if (lane < 8)
            dst.pooled[(size_t)d * 8 + lane] = src.pooled[(size_t)ua.row[u] * 8 + lane];
        if (lane == 0)
            emit_unit[d] = u0 + u;
    }
}

/* DFMA throughput probe: 8 independent chains per thread */
__global__ void k_dfma(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b);
        x1 = fma(x1, a, b);
        x2 = fma(x2, a, b);
        x3 = fma(x3, a, b);
        x4 = fma(x4, a, b);
        x5 = fma(x5, a, b);
        x6 = fma(x6, a, b);
        x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

