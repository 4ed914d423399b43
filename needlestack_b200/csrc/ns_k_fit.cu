/* ns_k_fit.cu — the robust NB regression kernels: persistent warp kernel (fast path) and one-CTA-per-unit heavy kernel (sm_100a) */
#include "ns_kernels.cuh"
#include "ns_coop.cuh"

extern __shared__ double ns_smem[];

/* fill the CTA-shared table of reciprocals 1/j */
__device__ __forceinline__ void ns_fill_rtab(double *rtab)
{
    for (int j = threadIdx.x; j < NS_RTAB; j += blockDim.x)
        rtab[j] = j > 0 ? 1.0 / (double)j : 0.0;
    __syncthreads();
}

__device__ __forceinline__ void ns_load_unit(const ns_rows &r, const ns_work &w, int n, uint32_t flags,
                                             const ns_params &prm, int lane, int nl, double &maxmu)
{
    const bool inv = flags & NS_F_REF_INV, extra = flags & NS_F_EXTRA_ROB;
    maxmu = -INFINITY;
    for (int i = lane; i < n; i += nl) {
        double x = ns_ld(r.dp, i);
        double y = inv ? ns_ld(r.rp, i) + ns_ld(r.rm, i) : ns_ld(r.vp, i) + ns_ld(r.vm, i);
        maxmu = fmax(maxmu, x);
        bool keep = !(extra && y / x > prm.min_af_extra_rob);
        w.y[i] = y;
        w.x[i] = keep ? x : 0.0;
    }
}

/* The warp kernel's variant: samples are stored by decreasing depth.  The fit does not depend on the order of the
 * samples, but the joint window loop of a pair-round runs to the longest window among its 64 samples, and the
 * window grows with mu = slope * depth: with the deep samples together in the first round, the second round's
 * loop is ~1/3 shorter (C2: -17 % lattice iterations). */
__device__ __forceinline__ void ns_load_unit_sorted(const ns_rows &r, const ns_work &w, int n, uint32_t flags,
                                                    const ns_params &prm, int lane, double &maxmu)
{
    const bool inv = flags & NS_F_REF_INV, extra = flags & NS_F_EXTRA_ROB;
    int *xs = reinterpret_cast<int *>(w.scratch);
    for (int i = lane; i < n; i += 32)
        xs[i] = r.dp[i];
    __syncwarp();
    maxmu = -INFINITY;
    for (int i = lane; i < n; i += 32) {
        const int xi = xs[i];
        int rank = 0;
        for (int j = 0; j < n; j++) {
            const int xj = xs[j];
            rank += (xj > xi) || (xj == xi && j < i);
        }
        const double x = (double)xi;
        const double y = inv ? ns_ld(r.rp, i) + ns_ld(r.rm, i) : ns_ld(r.vp, i) + ns_ld(r.vm, i);
        maxmu = fmax(maxmu, x);
        const bool keep = !(extra && y / x > prm.min_af_extra_rob);
        w.y[rank] = y;
        w.x[rank] = keep ? x : 0.0;
    }
}

__device__ __forceinline__ void ns_store_fit(const ns_unit_arrays &ua, int u, uint32_t flags, const ns_fit_result &res,
                                             long long t_start)
{
    ua.sigma[u] = res.sigma;
    ua.slope[u] = res.slope;
    ua.flags[u] = flags | res.flags;
    unsigned no = res.n_outer > 255 ? 255 : res.n_outer, ni = res.n_irls > 1023 ? 1023 : res.n_irls,
             nj = res.n_obj > 16383 ? 16383 : res.n_obj;
    ua.iters[u] = no | (ni << 8) | (nj << 18);
    ua.cycles[u] = (unsigned long long)(clock64() - t_start);
}

/* persistent warps, one unit per warp at a time, fast (table/recurrence) path */
#ifndef NS_FIT_MINB
#define NS_FIT_MINB 5
#endif
__global__ void __launch_bounds__(128, NS_FIT_MINB)
k_fit(ns_unit_src src, const __grid_constant__ ns_params prm, int64_t u0, ns_unit_arrays ua, ns_counters *cnt)
{
    const int n = src.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *rtab = ns_smem, *ltab = ns_smem + NS_RTAB;
    for (int j = threadIdx.x; j < 256; j += blockDim.x)
        ltab[j] = ns_g_logtab[j];
    ns_fill_rtab(rtab);
    ns_work w;
    ns_work_bind(w, ns_smem + NS_RTAB + 256 + (size_t)warp * ns_work_doubles(n), n, rtab, ltab);
    unsigned long long tot_seed = 0, tot_lat = 0;
    const int n_fit = cnt->n_light;
    for (;;) {
        int idx = 0;
        if (lane == 0)
            idx = atomicAdd(&cnt->next_fit, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= n_fit)
            break;
        const int u = ua.fit_list[idx];
        const long long t_start = clock64();
        const uint32_t flags = ua.flags[u];
        ns_rows r = ns_unit_rows(src, u0 + u);
        double maxmu;
        ns_load_unit_sorted(r, w, n, flags, prm, lane, maxmu);
        maxmu = WarpG::max(maxmu);
        __syncwarp();
        ns_fit_result res;
        ns_fit_unit<WarpG, 1>(w, n, maxmu, prm, res);
        if (res.deferred) {
            /* needs the spline + quadrature branch: one CTA per unit in k_fit_heavy */
            if (lane == 0)
                ua.heavy_list[atomicAdd(&cnt->n_heavy, 1)] = u;
        } else {
            tot_seed += res.cnt.n_seed;
            tot_lat += res.cnt.n_lattice;
            if (lane == 0)
                ns_store_fit(ua, u, flags, res, t_start);
        }
        __syncwarp();
    }
    for (int o = 16; o > 0; o >>= 1) {
        tot_seed += __shfl_xor_sync(0xffffffffu, tot_seed, o);
        tot_lat += __shfl_xor_sync(0xffffffffu, tot_lat, o);
    }
    if (lane == 0 && (tot_seed | tot_lat)) {
        atomicAdd(&cnt->n_seed, tot_seed);
        atomicAdd(&cnt->n_lattice, tot_lat);
        atomicAdd(&cnt->fast_seed, tot_seed);
        atomicAdd(&cnt->fast_lattice, tot_lat);
    }
}

/* One thread-block cluster per unit: the CTAs of the cluster split the unit's samples, every
 * warp takes one sample at a time and its lanes share that sample's window (ns_coop.cuh); sums
 * over samples are exchanged through distributed shared memory (ClusterG). */
#ifndef NS_HEAVY_MINB
#define NS_HEAVY_MINB 1 /* 128 registers; 64 (to leave half the register file to the warp kernel) spills and is 1.8x slower */
#endif
template <class CG>
static __device__ __forceinline__ void ns_fit_heavy_body(const ns_unit_src &src, const ns_params &prm, int64_t u0,
                                                         const ns_unit_arrays &ua, ns_counters *cnt, int which, int *started)
{
    const int n = src.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ int s_idx;
    /* CTA-shared reciprocal table 1/j, j < NS_RTAB_HEAVY */
    double *rt = ns_smem;
    for (int j = threadIdx.x; j < NS_RTAB_HEAVY; j += blockDim.x)
        rt[j] = j > 0 ? 1.0 / (double)j : 0.0;
    if (threadIdx.x == 0)
        ClusterG::phase() = 0;
    ns_work w;
    ns_work_bind(w, ns_smem + NS_RTAB_HEAVY, n, rt);
    double *cs = ns_smem + NS_RTAB_HEAVY + ns_work_doubles(n) + (size_t)warp * NS_COOP_DOUBLES;
    w.coop.rt = rt;
    w.coop.rtn = NS_RTAB_HEAVY;
    w.coop.y = cs;
    w.coop.c = cs + NS_MAX_KNOTS;
    w.coop.ib = cs + 2 * NS_MAX_KNOTS;
    w.coop.m = cs + 3 * NS_MAX_KNOTS;
    unsigned long long tot_seed = 0, tot_lat = 0, tot_quad = 0;
    /* which = 0: units the gate kernel predicted heavy; 1: units the warp kernel handed over */
    const int n_heavy = which ? cnt->n_heavy : cnt->n_pre;
    int *next = which ? &cnt->next_heavy : &cnt->next_pre;
    const int *list = which ? ua.heavy_list : ua.pre_list;
    const unsigned crank = ClusterG::crank(), csize = ClusterG::csize();
    /* tell the host that this cluster is resident (it then launches the warp kernel next to us) */
    if (started && crank == 0 && threadIdx.x == 0) {
        atomicAdd_system(started, 1);
        __threadfence_system();
    }
    __syncthreads();
    for (;;) {
        ClusterG::csync(); /* the previous unit is finished everywhere, s_idx may be overwritten */
        if (crank == 0 && threadIdx.x == 0) {
            int idx = atomicAdd(next, 1);
            for (unsigned r = 0; r < csize; r++)
                ClusterG::remote_store_int(&s_idx, r, idx);
        }
        ClusterG::csync();
        const int idx = s_idx;
        if (idx >= n_heavy)
            break;
        const int u = list[idx];
        const long long t_start = clock64();
        const uint32_t flags = ua.flags[u];
        ns_rows r = ns_unit_rows(src, u0 + u);
        double maxmu;
        ns_load_unit(r, w, n, flags, prm, threadIdx.x, blockDim.x, maxmu);
        maxmu = BlockG::max(maxmu);
        __syncthreads();
        ns_fit_result res;
        ns_fit_unit<CG, 2>(w, n, maxmu, prm, res);
        if (lane == 0) { /* counters are uniform within a warp, disjoint across warps and CTAs */
            tot_seed += res.cnt.n_seed;
            tot_lat += res.cnt.n_lattice;
            tot_quad += res.cnt.n_quad;
        }
        if (crank == 0 && threadIdx.x == 0)
            ns_store_fit(ua, u, flags, res, t_start);
    }
    tot_seed = (unsigned long long)BlockG::sum((double)tot_seed);
    tot_lat = (unsigned long long)BlockG::sum((double)tot_lat);
    tot_quad = (unsigned long long)BlockG::sum((double)tot_quad);
    if (threadIdx.x == 0 && (tot_seed | tot_lat | tot_quad)) {
        atomicAdd(&cnt->n_seed, tot_seed);
        atomicAdd(&cnt->n_lattice, tot_lat);
        atomicAdd(&cnt->n_quad, tot_quad);
    }
}

/* latency form: 16 warps per CTA at 128 registers (a unit's critical path is what a small batch waits for) */
__global__ void __launch_bounds__(NS_HEAVY_THREADS, NS_HEAVY_MINB)
k_fit_heavy(ns_unit_src src, const __grid_constant__ ns_params prm, int64_t u0, ns_unit_arrays ua, ns_counters *cnt, int which, int *started)
{
    ns_fit_heavy_body<ClusterG>(src, prm, u0, ua, cnt, which, started);
}

/* throughput form, for deep panels with thousands of queued units: 20 warps per CTA at 96 registers (C3 +10 %) */
__global__ void __launch_bounds__(NS_HEAVY_THREADS_TP, 1)
k_fit_heavy_tp(ns_unit_src src, const __grid_constant__ ns_params prm, int64_t u0, ns_unit_arrays ua, ns_counters *cnt, int which, int *started)
{
    ns_fit_heavy_body<ClusterGT>(src, prm, u0, ua, cnt, which, started);
}
