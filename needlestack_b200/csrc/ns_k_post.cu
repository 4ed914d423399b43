/* ns_k_post.cu — p/q/GQ, Holm power q-values, T/N status, strand bias, genotypes; y* tables (sm_100a) */
#include "ns_kernels.cuh"

extern __shared__ double ns_smem[];

__global__ void __launch_bounds__(128, 4)
k_post(ns_unit_src src, ns_params prm, int64_t u0, ns_unit_arrays ua, ns_planes pl,
                       const double *qtab, int qtab_len, ns_counters *cnt)
{
    const int n = src.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *rtab = ns_smem; /* CTA-shared 1/j, j < NS_RTAB */
    for (int j = threadIdx.x; j < NS_RTAB; j += blockDim.x)
        rtab[j] = j > 0 ? 1.0 / (double)j : 0.0;
    __syncthreads();
    double *sm = ns_smem + NS_RTAB + (size_t)warp * (3 * (size_t)n + 8);
    const int n_post = cnt->n_post;
    for (;;) {
        int idx = 0;
        if (lane == 0)
            idx = atomicAdd(&cnt->next_post, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= n_post)
            break;
        const int u = ua.post_list[idx];
        if (ua.flags[u] & NS_F_POOLED)
            continue; /* fitted in the call-wide pool: its rows are computed at the end of the call */
        const size_t row = (size_t)ua.row[u];
        ns_rows r = ns_unit_rows(src, u0 + u);
        ns_unit_out o;
        o.pvalue = pl.pvalue + row * n;
        o.qvalue = pl.qvalue + row * n;
        o.gq = pl.gq + row * n;
        o.qval_minaf = pl.qval_minaf + row * n;
        o.sor = pl.sor + row * n;
        o.rvsb = pl.rvsb + row * n;
        o.fs = pl.fs + row * n;
        o.gt = pl.gt + row * n;
        o.status = pl.status + row * n;
        o.pooled = pl.pooled + row * 8;
        if (lane < 8)
            o.pooled[lane] = NAN;
        double mg = 0;
        uint32_t flags = ns_post_unit<WarpG>(r, n, prm, ua.flags[u], ua.sigma[u], ua.slope[u], qtab, qtab_len,
                                             o, sm, mg, rtab, NS_RTAB);
        if (lane == 0) {
            ua.flags[u] = flags;
            ua.max_gq[u] = mg;
            bool emit = prm.emit_mode == NS_EMIT_ALL || (prm.emit_mode == NS_EMIT_FITTED && !(flags & NS_F_GATED)) ||
                        (flags & NS_F_GATE2);
            ua.emitflag[u] = emit;
        }
        __syncwarp();
    }
}

/* y* tables of toQvalueN / toQvalueT (needlestack.r:243,249): they depend on the depth only */
__global__ void k_qtab(double *tabN, double *tabT, int len, double sigma_normal, double afmin)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= len)
        return;
    tabN[x] = ns_qnbinom_mu(0.01, 1 / sigma_normal, 0.5 * (double)x);
    if (tabT)
        tabT[x] = (afmin >= 0) ? ns_qbinom(0.01, (double)x, afmin) : 0.0;
}


/* toQvalue(x, y) of plot_rob_nb.r:83-87 on a grid: the Holm-adjusted tail probability of a new point
 * (x = DP, y = AO) added to the n fitted observations, as a Phred value.  Every CTA rebuilds the
 * unit's ascending p-values and Holm prefix maxima in shared memory (n <= ~9000), then its threads
 * take grid points; each point is one tail evaluation + one binary search (the reference re-sorts
 * n+1 p-values per grid point). */
__global__ void __launch_bounds__(256)
k_qgrid(int n, const int32_t *cov, const int32_t *ma, double sigma, double slope, int64_t m, const int32_t *gx,
        const int32_t *gy, double *out)
{
    double *rtab = ns_smem, *p = ns_smem + NS_RTAB, *srt = p + n, *hm = srt + n;
    for (int j = threadIdx.x; j < NS_RTAB; j += blockDim.x)
        rtab[j] = j > 0 ? 1.0 / (double)j : 0.0;
    __syncthreads();
    const double size = 1 / sigma;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        p[i] = ns_nb_pvalue((double)ma[i], size, slope * (double)cov[i], rtab, NS_RTAB);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double pi = p[i];
        int rank = 0;
        for (int j = 0; j < n; j++) {
            const double pj = p[j];
            rank += (pj < pi) || (pj == pi && j < i);
        }
        srt[rank] = pi;
    }
    __syncthreads();
    if (threadIdx.x < 32) { /* hm[k] = max_{j<=k} (n+1-j) srt[j]: chunk per lane, warp scan of the chunk maxima */
        const int lane = threadIdx.x, per = (n + 31) / 32;
        const int b0 = lane * per < n ? lane * per : n, b1 = b0 + per < n ? b0 + per : n;
        double loc = -INFINITY;
        for (int k = b0; k < b1; k++)
            loc = fmax(loc, (double)(n + 1 - k) * srt[k]);
        double run = WarpG::excl_prefix_max(loc);
        for (int k = b0; k < b1; k++) {
            run = fmax(run, (double)(n + 1 - k) * srt[k]);
            hm[k] = run;
        }
    }
    __syncthreads();
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < m; g += (int64_t)gridDim.x * blockDim.x) {
        const double x = (double)gx[g], y = (double)gy[g];
        double q = 0.0;
        if (!(y > x)) {
            const double pstar = ns_nb_pvalue(y, size, slope * x, rtab, NS_RTAB);
            q = -10 * log10(ns_holm_last(pstar, srt, hm, n)) + 0.0;
        }
        out[g] = q;
    }
}
