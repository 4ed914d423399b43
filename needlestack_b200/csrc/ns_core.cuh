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
    static constexpr bool rolled = false; /* see ns_dqagse_t */
    __device__ static __forceinline__ int nl() { return 32; }
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
    /* per-sigma tables, j < J (J = 64): tt[j] = 1/(a+j) (increments of digamma(a+j) - digamma(a)),
     * r2[j] = (a+j)/(j+1) (pmf ratio without the mu/(mu+size) factor) */
    __device__ static __forceinline__ void fill_dtab(double *tt, double *r2, const double *rt, double a, int J)
    {
        __syncwarp();
        for (int j = lane(); j < J; j += 32) {
            const double aj = a + (double)j;
            r2[j] = aj * rt[j + 1];
            tt[j] = ns_frcp(aj);
        }
        __syncwarp();
    }
    __device__ static __forceinline__ void fill_r2tab(double *r2, const double *rt, double a, int J)
    {
        __syncwarp();
        for (int j = lane(); j < J; j += 32)
            r2[j] = (a + (double)j) * rt[j + 1];
        __syncwarp();
    }
    /* replica-local variants (identical to the group-wide ones unless the group spans CTAs) */
    __device__ static __forceinline__ int lnl() { return nl(); }
    __device__ static __forceinline__ int llane() { return lane(); }
    __device__ static __forceinline__ double lsum(double v) { return sum(v); }
    __device__ static __forceinline__ int lisum(int v) { return isum(v); }
    __device__ static __forceinline__ void lsync() { sync(); }
};

/* one CTA = one unit, one WARP = one "lane" of the group (the heavy kernel): the values handed to
 * the collectives are uniform across each warp; reductions go through shared memory */
struct CoopG {
    static constexpr bool rolled = false; /* see ns_dqagse_t */
    __device__ static __forceinline__ int nl() { return blockDim.x >> 5; }
    __device__ static __forceinline__ int lane() { return threadIdx.x >> 5; }
    __device__ static __forceinline__ double *red()
    {
        __shared__ double r[34];
        return r;
    }
    __device__ static __forceinline__ double sum(double v)
    {
        double *r = red();
        __syncthreads();
        if ((threadIdx.x & 31) == 0)
            r[threadIdx.x >> 5] = v;
        __syncthreads();
        double s = 0;
        const int nw = blockDim.x >> 5;
        for (int k = 0; k < nw; k++)
            s += r[k];
        return s;
    }
    __device__ static __forceinline__ double max(double v)
    {
        double *r = red();
        __syncthreads();
        if ((threadIdx.x & 31) == 0)
            r[threadIdx.x >> 5] = v;
        __syncthreads();
        double s = -INFINITY;
        const int nw = blockDim.x >> 5;
        for (int k = 0; k < nw; k++)
            s = fmax(s, r[k]);
        return s;
    }
    __device__ static __forceinline__ int isum(int v) { return (int)sum((double)v); }
    __device__ static __forceinline__ int any(int p) { return __syncthreads_or(p); }
    __device__ static __forceinline__ void sync() { __syncthreads(); }
    __device__ static __forceinline__ void fill_dtab(double *, double *, const double *, double, int) {}
    __device__ static __forceinline__ void fill_r2tab(double *, const double *, double, int) {}
    /* replica-local variants (identical to the group-wide ones unless the group spans CTAs) */
    __device__ static __forceinline__ int lnl() { return nl(); }
    __device__ static __forceinline__ int llane() { return lane(); }
    __device__ static __forceinline__ double lsum(double v) { return sum(v); }
    __device__ static __forceinline__ int lisum(int v) { return isum(v); }
    __device__ static __forceinline__ void lsync() { sync(); }
};

/* one thread-block CLUSTER = one unit: every CTA keeps a replica of the unit's per-sample state in
 * its own shared memory, the samples of the expensive window loops are split over all warps of
 * the cluster, and the partial sums are exchanged through distributed shared memory
 * (cluster.map_shared_rank) with one barrier.cluster per reduction (double-buffered slots).
 * Values handed to the collectives are uniform within each warp. */
struct ClusterG {
    static constexpr bool rolled = false; /* see ns_dqagse_t */
    __device__ static __forceinline__ unsigned crank()
    {
        unsigned r;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
        return r;
    }
    __device__ static __forceinline__ unsigned csize()
    {
        unsigned r;
        asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
        return r;
    }
    __device__ static __forceinline__ void csync()
    {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    /* store v into slot `slot` of the exchange array of CTA `rank` */
    __device__ static __forceinline__ void remote_store(double *local_ptr, unsigned rank, double v)
    {
        unsigned a = (unsigned)__cvta_generic_to_shared(local_ptr), ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
        asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
    }
    __device__ static __forceinline__ void remote_store_int(int *local_ptr, unsigned rank, int v)
    {
        unsigned a = (unsigned)__cvta_generic_to_shared(local_ptr), ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
    }
    __device__ static __forceinline__ double *xch()
    {
        __shared__ double x[2][16];
        return &x[0][0];
    }
    __device__ static __forceinline__ int &phase()
    {
        __shared__ int ph;
        return ph;
    }
    __device__ static __forceinline__ int nl() { return (int)csize() * (blockDim.x >> 5); }
    __device__ static __forceinline__ int lane() { return (int)crank() * (blockDim.x >> 5) + (threadIdx.x >> 5); }
    __device__ static __forceinline__ int lnl() { return blockDim.x >> 5; }
    __device__ static __forceinline__ int llane() { return threadIdx.x >> 5; }
    __device__ static __forceinline__ double lsum(double v) { return CoopG::sum(v); }
    __device__ static __forceinline__ int lisum(int v) { return CoopG::isum(v); }
    __device__ static __forceinline__ void lsync() { __syncthreads(); }
    __device__ static __forceinline__ void sync() { __syncthreads(); }
    /* OP 0: sum, 1: max */
    template <int OP> __device__ static __forceinline__ double allreduce(double v)
    {
        double p = OP == 0 ? CoopG::sum(v) : CoopG::max(v); /* CTA partial, uniform in the CTA */
        const unsigned cs = csize();
        if (cs == 1)
            return p;
        /* every thread knows the parity: it toggles once per call, uniformly */
        int &ph = phase();
        const int par = ph & 1;
        double *x = xch() + par * 16;
        if (threadIdx.x < cs)
            remote_store(x + crank(), threadIdx.x, p);
        csync();
        double t = OP == 0 ? 0.0 : -INFINITY;
        for (unsigned k = 0; k < cs; k++)
            t = OP == 0 ? t + x[k] : fmax(t, x[k]);
        __syncthreads();
        if (threadIdx.x == 0)
            ph = par ^ 1;
        __syncthreads();
        return t;
    }
    __device__ static __forceinline__ double sum(double v) { return allreduce<0>(v); }
    __device__ static __forceinline__ double max(double v) { return allreduce<1>(v); }
    __device__ static __forceinline__ int isum(int v) { return (int)sum((double)v); }
    __device__ static __forceinline__ int any(int p) { return sum(__syncthreads_or(p) ? 1.0 : 0.0) > 0.0; }
    __device__ static __forceinline__ void fill_dtab(double *, double *, const double *, double, int) {}
    __device__ static __forceinline__ void fill_r2tab(double *, const double *, double, int) {}
};

/* the same for the throughput form of the kernel (k_fit_heavy_tp): smaller code where it costs latency (ns_dqagse_t) */
struct ClusterGT : ClusterG {
    static constexpr bool rolled = true;
};

/* one CTA = one unit (the heavy kernel): reductions through shared memory */
struct BlockG {
    static constexpr bool rolled = false; /* see ns_dqagse_t */
    __device__ static __forceinline__ int nl() { return blockDim.x; }
    __device__ static __forceinline__ int lane() { return threadIdx.x; }
    __device__ static __forceinline__ double *red() 
    {
        __shared__ double r[34];
        return r;
    }
    __device__ static __forceinline__ double sum(double v)
    {
        double *r = red();
        v = WarpG::sum(v);
        __syncthreads();
        if ((threadIdx.x & 31) == 0)
            r[threadIdx.x >> 5] = v;
        __syncthreads();
        double s = 0;
        const int nw = blockDim.x >> 5;
        for (int k = 0; k < nw; k++)
            s += r[k];
        return s;
    }
    __device__ static __forceinline__ double max(double v)
    {
        double *r = red();
        v = WarpG::max(v);
        __syncthreads();
        if ((threadIdx.x & 31) == 0)
            r[threadIdx.x >> 5] = v;
        __syncthreads();
        double s = -INFINITY;
        const int nw = blockDim.x >> 5;
        for (int k = 0; k < nw; k++)
            s = fmax(s, r[k]);
        return s;
    }
    __device__ static __forceinline__ int isum(int v) { return (int)sum((double)v); }
    __device__ static __forceinline__ int any(int p) { return __syncthreads_or(p); }
    __device__ static __forceinline__ void sync() { __syncthreads(); }
    __device__ static __forceinline__ void fill_dtab(double *, double *, const double *, double, int) {}
    __device__ static __forceinline__ void fill_r2tab(double *, const double *, double, int) {}
    /* replica-local variants (identical to the group-wide ones unless the group spans CTAs) */
    __device__ static __forceinline__ int lnl() { return nl(); }
    __device__ static __forceinline__ int llane() { return lane(); }
    __device__ static __forceinline__ double lsum(double v) { return sum(v); }
    __device__ static __forceinline__ int lisum(int v) { return isum(v); }
    __device__ static __forceinline__ void lsync() { sync(); }
};
#define NS_GD __device__
#else
#define NS_GD
#endif

struct SerialG {
    static constexpr bool rolled = false; /* see ns_dqagse_t */
    static NS_HD int nl() { return 1; }
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
    static NS_HD void fill_dtab(double *tt, double *r2, const double *rt, double a, int J)
    {
        for (int j = 0; j < J; j++) {
            r2[j] = (a + (double)j) * rt[j + 1];
            tt[j] = 1.0 / (a + (double)j);
        }
    }
    static NS_HD void fill_r2tab(double *r2, const double *rt, double a, int J)
    {
        for (int j = 0; j < J; j++)
            r2[j] = (a + (double)j) * rt[j + 1];
    }
    /* replica-local variants (identical to the group-wide ones unless the group spans CTAs) */
    static NS_HD int lnl() { return nl(); }
    static NS_HD int llane() { return lane(); }
    static NS_HD double lsum(double v) { return sum(v); }
    static NS_HD int lisum(int v) { return isum(v); }
    static NS_HD void lsync() { sync(); }
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
NS_GF inline uint32_t ns_gate_unit(const ns_rows &r, int n, const ns_params &prm, int plain, double *scratch,
                                   double *mu_hint = nullptr)
{
    const int lane = G::lane();
    if (!r.valid)
        return NS_F_SKIPPED | NS_F_GATED;
    uint32_t flags = 0;
    /* ref inversion: sum(ma/DP > 0.8, na.rm=T) > 0.5*n  (NaN compares false) */
    int cnt = 0;
    if (!plain) {
        for (int i = lane; i < n; i += G::nl()) {
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
        for (int i = lane; i < n; i += G::nl()) {
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
    int n0 = 0, n_cov = 0, n_pos = 0, n_fit = 0;
    double maxy = -INFINITY, maxaf = -INFINITY, maxx = 0, sum_af = 0;
    for (int i = lane; i < n; i += G::nl()) {
        double x = ns_ld(r.dp, i);
        double y = inv ? ns_ld(r.rp, i) + ns_ld(r.rm, i) : ns_ld(r.vp, i) + ns_ld(r.vm, i);
        double a = y / x;
        bool keep = !(extra && a > prm.min_af_extra_rob);
        maxx = fmax(maxx, x);
        if (keep && x > 0) {
            sum_af += a;
            n_fit++;
        }
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
    if (mu_hint) {
        /* upper bound of the first fitted mean max_i(inislope * x_i): the trimmed mean of
         * glm_rob_nb.r:77 never exceeds the plain mean of the allelic fractions */
        maxx = G::max(maxx);
        sum_af = G::sum(sum_af);
        n_fit = G::isum(n_fit);
        *mu_hint = n_fit > 0 ? (sum_af / (double)n_fit) * maxx : 0.0;
    }
    bool gate = (n0 == 0) || (n_cov < prm.size_min) || (maxy < min_reads) || (n_pos == 0 && !extra) ||
                (maxaf < min_af);
    if (!gate) {
        /* median(x) < min_coverage over the kept samples: the two middle order statistics
         * by rank counting (integers; ties broken by index) */
        G::sync();
        for (int i = lane; i < n; i += G::nl()) {
            double x = ns_ld(r.dp, i);
            double y = inv ? ns_ld(r.rp, i) + ns_ld(r.rm, i) : ns_ld(r.vp, i) + ns_ld(r.vm, i);
            bool keep = !(extra && y / x > prm.min_af_extra_rob);
            scratch[i] = keep ? x : -1.0; /* depths are >= 0 */
        }
        G::sync();
        const int k_lo = (n0 - 1) / 2, k_hi = n0 / 2;
        double vlo = 0, vhi = 0;
        for (int i = lane; i < n; i += G::nl()) {
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

#include "ns_fit.cuh"

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
    for (int i = G::lane(); i < n; i += G::nl()) {
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

/* fisher.test(2x2) for the POOLED table (needlestack.r:236): the support can hold thousands of
 * terms, so the lanes of the group split it into contiguous chunks; each lane seeds its chunk
 * with one Loader evaluation relative to the mode and walks it by the ratio recurrence. */
template <class G> NS_GF inline double ns_fisher_pooled(double a11, double a21, double a12, double a22)
{
    const double m = a11 + a21, n = a12 + a22, k = a11 + a12, x = a11;
    const double lo = fmax(0.0, k - n), hi = fmin(k, m);
    if (hi <= lo)
        return 1.0;
    const double ns = hi - lo + 1;
    if (G::nl() == 1 || ns < 256)
        return ns_fisher_2x2_ool(a11, a21, a12, a22);
    const double mode = ns_hyper_mode(m, n, k, lo, hi);
    const double lmode = ns_dhyper_log(mode, m, n, k);
    const double per = ceil(ns / (double)G::nl());
    const double a = lo + (double)G::lane() * per, b = fmin(hi, a + per - 1);
    double dobs = 0, dummy = 0, tot = 0, dstart = 0;
    const bool have = a <= hi;
    if (have) {
        dstart = exp(ns_dhyper_log(a, m, n, k) - lmode);
        tot = ns_hyper_walk(a, b, a, dstart, m, n, k, -1.0, x, dobs);
    }
    tot = G::sum(tot);
    dobs = G::sum(dobs); /* exactly one lane visited x */
    const double thr = dobs * (1 + 1e-7);
    double pv = have ? ns_hyper_walk(a, b, a, dstart, m, n, k, thr, -1.0, dummy) : 0.0;
    pv = G::sum(pv);
    return pv / tot;
}

/* Per-unit tail: p-values, BH q-values, GQ (glm_rob_nb.r:219-224); power q-values and
 * tumour/normal status (needlestack.r:340-380); gate 1, strand bias, gate 2, genotype
 * (needlestack.r:381-409).  sm: group-shared scratch of >= 3n+8 doubles.
 * Returns the unit's flag word (input flags | NS_F_GATE1 | NS_F_GATE2). */
template <class G>
NS_GF inline uint32_t ns_post_unit(const ns_rows &r, int n, const ns_params &prm, uint32_t flags,
                                   double sigma, double slope, const double *qtab, int qtab_len,
                                   const ns_unit_out &o, double *sm, double &max_gq_out,
                                   const double *rt = nullptr, int rtn = 0)
{
    const int lane = G::lane();
    const bool gated = flags & NS_F_GATED;
    const bool inv = flags & NS_F_REF_INV;
    const double thr = prm.GQ_threshold;
    double *srt = sm, *hm = sm + n, *tmp = sm + 2 * (size_t)n;
    const double size = 1 / sigma;
    const int per = (n + G::nl() - 1) / G::nl(); /* contiguous block of sorted ranks per lane */
    const int b0 = (lane * per < n) ? lane * per : n, b1 = (b0 + per < n) ? b0 + per : n;
    /* p-values on the full vectors (:221) */
    if (gated) {
        for (int i = lane; i < n; i += G::nl()) {
            o.pvalue[i] = 1;
            o.qvalue[i] = 1;
            o.gq[i] = 0;
        }
    } else {
        /* Early out when only the reference's VCF calls are wanted: every BH q-value is >= the smallest
         * p-value (the factors n/rank are >= 1), so if -10 log10(min p) cannot reach GQ_threshold no sample
         * passes needlestack.r:381 and nothing of this unit is written.  Its rows and max_gq are then
         * not computed (max_gq is reported as NaN).  The test runs on lower bounds of the p-values (no tail
         * summation: 99.8 % of the fitted units of an exome panel end here); a unit that survives it is
         * decided on the exact values below. */
        const bool calls_only = prm.emit_mode == NS_EMIT_CALLS && !((!r.is_indel) && prm.output_all_SNVs);
        if (calls_only) {
            double plb = INFINITY;
            for (int i = lane; i < n; i += G::nl()) {
                double x = ns_ld(r.dp, i);
                double y = inv ? ns_ld(r.rp, i) + ns_ld(r.rm, i) : ns_ld(r.vp, i) + ns_ld(r.vm, i);
                plb = fmin(plb, ns_nb_pvalue_lb(y, size, slope * x, rt, rtn));
            }
            plb = -G::max(-plb);
            if (-10 * log10(plb) < thr - 1e-9) {
                max_gq_out = NAN;
                return flags;
            }
        }
        double pmin = INFINITY;
        for (int i = lane; i < n; i += G::nl()) {
            double x = ns_ld(r.dp, i);
            double y = inv ? ns_ld(r.rp, i) + ns_ld(r.rm, i) : ns_ld(r.vp, i) + ns_ld(r.vm, i);
            double p = ns_nb_pvalue(y, size, slope * x, rt, rtn);
            tmp[i] = p;
            pmin = fmin(pmin, p);
        }
        if (calls_only) {
            pmin = -G::max(-pmin);
            if (-10 * log10(pmin) < thr - 1e-9) {
                max_gq_out = NAN;
                return flags;
            }
        }
        for (int i = lane; i < n; i += G::nl())
            o.pvalue[i] = tmp[i];
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
        for (int i = lane; i < n; i += G::nl()) {
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
    for (int i = lane; i < n; i += G::nl()) {
        mg = fmax(mg, o.gq[i]);
        anyq |= o.gq[i] >= thr;
    }
    mg = G::max(mg);
    anyq = G::any(anyq);
    max_gq_out = mg;

    /* qval_minAF (needlestack.r:340-380) */
    for (int i = lane; i < n; i += G::nl()) {
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
        for (int k = lane; k < prm.n_pairs; k += G::nl())
            germ |= o.gq[prm.nindex[k]] > thr;
        germ = G::any(germ);
    }
    for (int i = lane; i < n; i += G::nl()) {
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
            ystar = (qtab && xi < qtab_len) ? qtab[xi] : ns_qnbinom_mu_ool(0.01, 1 / prm.sigma_normal, 0.5 * x);
        } else {
            ystar = ns_qbinom_ool(0.01, x, afmin);
        }
        double pstar = ns_nb_pvalue(ystar, size, slope * x, rt, rtn);
        o.qval_minaf[i] = -10 * log10(ns_holm_last(pstar, srt, hm, n));
    }
    G::sync();
    /* tumour/normal status, needlestack.r:349-373, in the reference's assignment order */
    if (tn) {
        const int32_t *T = prm.tindex, *N = prm.nindex;
        for (int k = lane; k < prm.n_only_t; k += G::nl())
            if (o.gq[prm.only_t[k]] > thr)
                o.status[prm.only_t[k]] = NS_ST_UNKNOWN;
        for (int k = lane; k < prm.n_only_n; k += G::nl())
            if (o.gq[prm.only_n[k]] > thr)
                o.status[prm.only_n[k]] = NS_ST_GERMLINE_UNCONFIRMABLE;
        G::sync();
        /* one pair may appear once; a sample belonging to several pairs is resolved by
         * running the rules sequentially per rule over pairs, as R's vector assignment does */
        for (int rule = 0; rule < 5; rule++) {
            for (int pass = 0; pass < 2; pass++) { /* 0: T side, 1: N side */
                if (rule == 4 && pass == 1)
                    break;
                for (int k = lane; k < prm.n_pairs; k += G::nl()) {
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
        for (int i = lane; i < n; i += G::nl()) {
            uint8_t s = o.status[i];
            ngerm |= (s == NS_ST_GERMLINE_CONFIRMED || s == NS_ST_GERMLINE_UNCONFIRMED ||
                      s == NS_ST_GERMLINE_UNCONFIRMABLE || s == NS_ST_UNKNOWN);
        }
        ngerm = G::any(ngerm);
        if (ngerm)
            for (int k = lane; k < prm.n_pairs; k += G::nl())
                if (o.status[T[k]] == NS_ST_SOMATIC)
                    o.status[T[k]] = NS_ST_POSSIBLE_CONTAMINATION;
        G::sync();
    }

    /* gate 1 (needlestack.r:381; :534, :698 for indels, which have no output_all) */
    const bool out_all = (!r.is_indel) && prm.output_all_SNVs;
    const bool gate1 = out_all || (!gated && anyq);
    for (int i = lane; i < n; i += G::nl()) {
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
    for (int i = lane; i < n; i += G::nl()) {
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
        o.fs[i] = -10 * log10(ns_fisher_2x2_ool(Vp, Vm, Rp, Rm)); /* the cap at :235 is a no-op */
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
        const double fs_all = -10 * log10(ns_fisher_pooled<G>(sVp, sVm, sRp, sRm));
        if (lane == 0) {
            o.pooled[0] = as;
            o.pooled[1] = ar;
            o.pooled[2] = fs_all;
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
    for (int i = lane; i < n; i += G::nl())
        any2 |= (sbs[i] <= sb_thr && o.gq[i] >= thr);
    any2 = G::any(any2);
    if (out_all || any2)
        flags |= NS_F_GATE2;
    /* genotype, needlestack.r:403-409 (:555-561 and :718-724 also use SB_threshold_SNV) */
    for (int i = lane; i < n; i += G::nl()) {
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
