/* ns_k_gate.cu — streaming gate kernel, row packer, scan tail, DFMA probe (sm_100a) */
#include "ns_kernels.cuh"

extern __shared__ double ns_smem[];

__global__ void __launch_bounds__(NS_WPB_GATE * 32)
k_gate(ns_unit_src src, ns_params prm, int64_t u0, int n_units, ns_unit_arrays ua, ns_counters *cnt)
{
    const int n = src.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *scratch = ns_smem + (size_t)warp * (size_t)n;
    const int wstride = gridDim.x * NS_WPB_GATE;
    const bool emit_all = prm.emit_mode == NS_EMIT_ALL;
    for (int u = blockIdx.x * NS_WPB_GATE + warp; u < n_units; u += wstride) {
        ns_rows r = ns_unit_rows(src, u0 + u);
        double mu_hint = 0;
        uint32_t flags = ns_gate_unit<WarpG>(r, n, prm, src.plain, scratch, &mu_hint);
        if (lane == 0) {
            const bool gated = flags & NS_F_GATED, skipped = flags & NS_F_SKIPPED;
            const bool out_all = (!r.is_indel) && prm.output_all_SNVs;
            const bool need = !skipped && (!gated || emit_all || out_all);
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
                /* A window of the first sigma objective evaluation (sigma = maxsig, c = c.tukey.sig)
                 * exceeds 2*n_ai.sig.tukey lattice points once mu*(1 + c*sqrt(maxsig + 1/mu)) does:
                 * such units go straight to the CTA-per-unit kernel, which runs next to the warp
                 * kernel on its own stream.  Over-prediction only costs efficiency. */
                const double kwin = mu_hint * (1.0 + prm.c_tukey_sig * sqrt(prm.maxsig + 1.0 / fmax(mu_hint, 1e-300)));
                if (kwin > 2.0 * prm.n_ai_sig_tukey)
                    ua.pre_list[atomicAdd(&cnt->n_pre, 1)] = u;
                else
                    ua.fit_list[atomicAdd(&cnt->n_light, 1)] = u;
            }
            if (need)
                ua.post_list[atomicAdd(&cnt->n_post, 1)] = u;
        }
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

