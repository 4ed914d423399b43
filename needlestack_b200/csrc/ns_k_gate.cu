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

/* per-unit outputs of the gate (written by one lane) and the unit's work-list class, identical in
 * all lanes: bits 0-1 = 1 warp kernel, 2 cluster kernel, 0 none; bit 2 = needs result rows */
/* lean: nobody reads the per-unit sigma / slope / max(GQ) / iteration words of a GATED unit (the caller asked for
 * the emitted rows only and gated units cannot be emitted): 36 of the 48 bytes the gate writes per unit are skipped */
__device__ __forceinline__ int ns_gate_commit(const ns_unit_arrays &ua, const ns_params &prm, int u, uint32_t flags,
                                              int is_indel, double mu_hint, bool writer, bool lean = false)
{
    const bool gated = flags & NS_F_GATED, skipped = flags & NS_F_SKIPPED;
    const bool out_all = (!is_indel) && prm.output_all_SNVs;
    const bool need = !skipped && (!gated || prm.emit_mode == NS_EMIT_ALL || out_all);
    if (writer) {
        ua.flags[u] = flags;
        if (!(lean && gated)) {
            ua.sigma[u] = NAN;
            ua.slope[u] = NAN;
            ua.max_gq[u] = 0.0;
            ua.iters[u] = 0;
            ua.cycles[u] = 0;
        }
        ua.needrow[u] = need;
        ua.emitflag[u] = 0;
    }
    int cls = need ? 4 : 0;
    if (!gated) {
        /* A window of the first sigma objective evaluation (sigma = maxsig, c = c.tukey.sig) exceeds
         * 2*n_ai.sig.tukey lattice points once mu*(1 + c*sqrt(maxsig + 1/mu)) does: such units go
         * straight to the cluster-per-unit kernel, which runs next to the warp kernel on its own
         * stream.  Over-prediction only costs efficiency; the threshold is half the switch point
         * because the slope usually grows from its first estimate. */
        const double kwin = mu_hint * (1.0 + prm.c_tukey_sig * sqrt(prm.maxsig + 1.0 / fmax(mu_hint, 1e-300)));
        cls |= (kwin > 1.0 * prm.n_ai_sig_tukey) ? 2 : 1;
    }
    return cls;
}

/* Appends the units of K positions per warp (packed classes, 3 bits per allele; lane = 3*j + k is allele k of the
 * warp's j-th position, unit u_of(j) + k) to the work lists: slots are reserved with shared-memory atomics inside
 * the CTA and ONE global atomic per list per CTA iteration (per-unit global atomics on three addresses were the
 * gate kernel's bottleneck; K > 1 amortises the three CTA barriers).  Must be called by all threads of the CTA. */
template <int K, class UOF>
__device__ __forceinline__ void ns_gate_insert(const int (&packed)[K], UOF u_of, int n_units, const ns_unit_arrays &ua,
                                               ns_counters *cnt)
{
    __shared__ int s_cnt[3], s_base[3];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 3)
        s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int cls = 0, off_l = 0, off_p = 0, u = 0;
    if (lane < 3 * K) {
        const int j = lane / 3, k = lane - 3 * j;
        int pk = packed[0];
#pragma unroll
        for (int t = 1; t < K; t++)
            pk = (j == t) ? packed[t] : pk;
        u = u_of(j) + k;
        if (u < n_units) {
            cls = (pk >> (3 * k)) & 7;
            if (cls & 3)
                off_l = atomicAdd(&s_cnt[(cls & 3) - 1], 1);
            if (cls & 4)
                off_p = atomicAdd(&s_cnt[2], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int *g = threadIdx.x == 0 ? &cnt->n_light : (threadIdx.x == 1 ? &cnt->n_pre : &cnt->n_post);
        s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(g, s_cnt[threadIdx.x]) : 0;
    }
    __syncthreads();
    if ((cls & 3) == 1)
        ua.fit_list[s_base[0] + off_l] = u;
    else if ((cls & 3) == 2)
        ua.pre_list[s_base[1] + off_l] = u;
    if (cls & 4)
        ua.post_list[s_base[2] + off_p] = u;
}

/* SH: the position's planes sit in a shared-memory stage filled by cp.async.bulk (ns_gate_staged), else in global memory */
template <int VEC, bool SH> struct ns_ldvec;
template <> struct ns_ldvec<4, false> {
    static __device__ __forceinline__ void ld(const int32_t *p, int *v)
    {
        const int4 t = __ldg(reinterpret_cast<const int4 *>(p));
        v[0] = t.x;
        v[1] = t.y;
        v[2] = t.z;
        v[3] = t.w;
    }
};
template <> struct ns_ldvec<2, false> {
    static __device__ __forceinline__ void ld(const int32_t *p, int *v)
    {
        const int2 t = __ldg(reinterpret_cast<const int2 *>(p));
        v[0] = t.x;
        v[1] = t.y;
    }
};
template <> struct ns_ldvec<4, true> {
    static __device__ __forceinline__ void ld(const int32_t *p, int *v)
    {
        const unsigned a = (unsigned)__cvta_generic_to_shared(p);
        asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(a));
    }
};
template <> struct ns_ldvec<2, true> {
    static __device__ __forceinline__ void ld(const int32_t *p, int *v)
    {
        const unsigned a = (unsigned)__cvta_generic_to_shared(p);
        asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(a));
    }
};
template <bool SH> __device__ __forceinline__ int ns_ld1(const int32_t *p)
{
    if (SH) {
        int v;
        const unsigned a = (unsigned)__cvta_generic_to_shared(p);
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
        return v;
    }
    return __ldg(p);
}

/* Streaming gate for SNV planes without the extra-robust pre-filter (the caller's default):
 * warp = position; the depth row is scanned once for the three alleles; all tests are integer
 * compares and warp vote / reduce instructions; medians come from counts, not from a sort:
 *   median(x) < c  <=>  both middle order statistics < c  (count(x < c) decides, and when it equals
 *   n/2 for even n the two middle values are max{x < c} and min{x >= c}). */
template <bool SH>
__device__ __forceinline__ int ns_gate_position(const int32_t *base, int n, int ref, const ns_params &prm, int u_first,
                                                int n_units, const ns_unit_arrays &ua, bool lean)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    if (ref > 3) {
        if (lane < 3 && u_first + lane < n_units)
            ns_gate_commit(ua, prm, u_first + lane, NS_F_SKIPPED | NS_F_GATED, 0, 0.0, true, lean);
        return 0;
    }
    int packed = 0;
    const double mc = prm.min_coverage;
    /* depth row: samples above min_coverage, order statistics around it, max depth */
    int n_cov = 0, n_less = 0, max_less = -1, min_geq = 0x7fffffff, maxx = 0;
    for (int i = lane; i < n; i += 32) {
        const int x = ns_ld1<SH>(base + i);
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
            const int x = ns_ld1<SH>(base + i), ma = ns_ld1<SH>(vp + i) + ns_ld1<SH>(vm + i), rr = ns_ld1<SH>(rp + i) + ns_ld1<SH>(rm + i);
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
                const int x = ns_ld1<SH>(base + i);
                const int y = inv ? ns_ld1<SH>(rp + i) + ns_ld1<SH>(rm + i) : ns_ld1<SH>(vp + i) + ns_ld1<SH>(vm + i);
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
        packed |= ns_gate_commit(ua, prm, u, flags, 0, mu_hint, lane == 0, lean) << (3 * k);
    }
    return packed;
}

/* Same tests with the position's nine rows held in registers (n <= 32*R): all 9*R loads of a warp are
 * in flight at once (3.6 KB per warp at n = 100), which is what a streaming kernel needs to approach
 * the HBM rate; the per-base statistics are computed once and shared by the three alleles. */
template <int R, bool SH>
__device__ __forceinline__ int ns_gate_position_reg(const int32_t *base, int n, int ref, const ns_params &prm,
                                                    int u_first, int n_units, const ns_unit_arrays &ua, bool lean)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    /* issue every load of the position before anything depends on one of them */
    int v[9][R];
#pragma unroll
    for (int r = 0; r < 9; r++)
#pragma unroll
        for (int k = 0; k < R; k++) {
            const int i = lane + 32 * k;
            v[r][k] = (i < n) ? ns_ld1<SH>(base + (int64_t)r * n + i) : 0;
        }
    if (ref > 3) {
        if (lane < 3 && u_first + lane < n_units)
            ns_gate_commit(ua, prm, u_first + lane, NS_F_SKIPPED | NS_F_GATED, 0, 0.0, true, lean);
        return 0;
    }
    int packed = 0;
    const double mc = prm.min_coverage;
    int n_cov = 0, n_less = 0, max_less = -1, min_geq = 0x7fffffff, maxx = 0;
    int cinv[4] = {0, 0, 0, 0}, npos[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0}, afok[4] = {0, 0, 0, 0};
    int sy[4] = {0, 0, 0, 0}; /* sum of the base's counts over the lane's samples (fits: <= 4 depths) */
    long long sx = 0;
#pragma unroll
    for (int k = 0; k < R; k++) {
        const bool in = lane + 32 * k < n;
        const int x = v[0][k];
        n_cov += in && (double)x > mc;
        const bool less = in && (double)x < mc;
        n_less += less;
        max_less = less ? max(max_less, x) : max_less;
        min_geq = (in && !less) ? min(min_geq, x) : min_geq;
        maxx = max(maxx, x);
        sx += x;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int y = v[1 + b][k] + v[5 + b][k];
            /* y/x > 0.8 (needlestack.r:330) for integers, exactly: the quotient rounds above the double
             * 0.8 iff 5y > 4x (a rational y/x != 4/5 is at least 1/(5x) away from it), and y/0 is
             * Inf (true) for y > 0, NaN (false) for y == 0 — also 5y > 0 */
            cinv[b] += in && (5LL * (long long)y > 4LL * (long long)x);
            npos[b] += y > 0;
            mx[b] = max(mx[b], y);
            sy[b] += y;
            if (prm.min_af > 0)
                afok[b] |= in && ns_ratio_ge(y, x, prm.min_af);
        }
    }
    n_cov = __reduce_add_sync(FULL, n_cov);
    n_less = __reduce_add_sync(FULL, n_less);
    max_less = __reduce_max_sync(FULL, max_less);
    min_geq = __reduce_min_sync(FULL, min_geq);
    maxx = __reduce_max_sync(FULL, maxx);
    for (int o = 16; o > 0; o >>= 1)
        sx += __shfl_xor_sync(FULL, sx, o);
#pragma unroll
    for (int b = 0; b < 4; b++) {
        cinv[b] = __reduce_add_sync(FULL, cinv[b]);
        npos[b] = __reduce_add_sync(FULL, npos[b]);
        mx[b] = __reduce_max_sync(FULL, mx[b]);
        if (prm.min_af > 0)
            afok[b] = __any_sync(FULL, afok[b]);
    }
    bool med_low; /* median(x) < min_coverage */
    if (n & 1)
        med_low = n_less >= (n + 1) / 2;
    else if (n_less >= n / 2 + 1)
        med_low = true;
    else if (n_less <= n / 2 - 1)
        med_low = false;
    else
        med_low = ((double)max_less + (double)min_geq) / 2 < mc;
    const bool pos_gate = med_low || n_cov < prm.size_min;
    /* static selection of a base's statistics (no dynamically indexed register arrays) */
#define NS_SEL4(a, b) ((b) == 0 ? a[0] : (b) == 1 ? a[1] : (b) == 2 ? a[2] : a[3])
    for (int k = 0; k < 3; k++) {
        const int u = u_first + k;
        if (u >= n_units)
            break;
        const int alt = k + (k >= ref ? 1 : 0);
        const bool inv = (double)NS_SEL4(cinv, alt) > 0.5 * (double)n; /* needlestack.r:330 */
        const int yb = inv ? ref : alt;
        const int np_ = NS_SEL4(npos, yb), maxy = NS_SEL4(mx, yb);
        bool gate = pos_gate || (double)maxy < prm.min_reads || np_ == 0;
        if (prm.min_af > 0)
            gate = gate || !NS_SEL4(afok, yb);
        const uint32_t flags = (inv ? NS_F_REF_INV : 0u) | (gate ? NS_F_GATED : 0u);
        double mu_hint = 0;
        if (!gate) { /* pooled allelic fraction times the largest depth: first-fit mean of the deepest sample */
            const long long tot = (long long)__reduce_add_sync(FULL, (unsigned)NS_SEL4(sy, yb));
            mu_hint = sx > 0 ? 1.1 * (double)tot / (double)sx * (double)maxx : 0.0;
        }
        packed |= ns_gate_commit(ua, prm, u, flags, 0, mu_hint, lane == 0, lean) << (3 * k);
    }
#undef NS_SEL4
    return packed;
}

/* The streaming form of the gate (n a multiple of VEC, n <= 32 * VEC, 16- or 8-byte aligned planes):
 *   - lane l owns samples VEC*l .. VEC*l+VEC-1: nine 128-bit (VEC = 4) or 64-bit (VEC = 2) loads per lane cover the
 *     position, all in flight before the first use (the scalar form issued 36 loads and their address arithmetic);
 *   - min_coverage is compared as an integer (for integer x: x > mc <=> x > floor(mc), x < mc <=> x < ceil(mc));
 *   - the per-base statistics are reduced once with REDUX and shared by the three alleles;
 *   - lanes 0..2 decide and commit the three units of the position side by side (one copy of the decision code and
 *     eight store instructions per position instead of three copies and twenty-four single-lane stores).
 * ncu (profiles/): 987 -> ~430 warp instructions per position at n = 100; the kernel was issue-bound, not
 * HBM-bound, at 1.9 TB/s. */
/* the position's nine rows, VEC consecutive samples per lane (zeros in the lanes beyond n) */
template <int VEC, bool SH>
__device__ __forceinline__ void ns_gate_load_vec(const int32_t *pos_base, int n, int (&v)[9][VEC])
{
    const int lane = threadIdx.x & 31;
    const bool act = VEC * lane < n;
    const int32_t *base = pos_base + VEC * lane;
#pragma unroll
    for (int r = 0; r < 9; r++) {
#pragma unroll
        for (int j = 0; j < VEC; j++)
            v[r][j] = 0;
        if (act)
            ns_ldvec<VEC, SH>::ld(base + (int64_t)r * n, v[r]);
    }
}

/* The gate of one position from its rows in registers.  The kernel is bound by instruction issue, not by HBM, so the
 * warp reductions are packed: the two coverage counts share one REDUX (16 bits each: n <= 128), the four inversion
 * counts another (8 bits each), the yes/no facts (some count > 0, some count >= min_reads, some value >= 2^28) one
 * OR-reduction; the sums that only feed the cluster-path prediction are reduced only when a unit of the position
 * passes the gate (3 reductions per position instead of 16 for the gated majority). */
template <int VEC>
__device__ __forceinline__ int ns_gate_compute_vec(const int (&v)[9][VEC], int n, int ref, const ns_params &prm, int t_gt,
                                                   int t_lt, int t_reads, int u_first, int n_units,
                                                   const ns_unit_arrays &ua, bool lean)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const bool act = VEC * lane < n;
    if (ref > 3) {
        if (lane < 3 && u_first + lane < n_units)
            ns_gate_commit(ua, prm, u_first + lane, NS_F_SKIPPED | NS_F_GATED, 0, 0.0, true, lean);
        return 0;
    }
    int n_cov = 0, n_less = 0, maxx = 0;
    unsigned sx = 0; /* only feeds the cluster-path prediction: a wrapped sum (depths > 2^30) costs efficiency, not results */
#pragma unroll
    for (int j = 0; j < VEC; j++) {
        const int x = v[0][j];
        n_cov += x > t_gt;
        n_less += x < t_lt;
        maxx = max(maxx, x);
        sx += (unsigned)x;
    }
    if (!act) { /* the zeros of a lane beyond n are not samples */
        n_cov = 0;
        n_less = 0;
    }
    /* per base: #(y/x > 0.8) (needlestack.r:330: exactly 5y > 4x for integers, see ns_gate_position_reg; 0/0 and the
     * zeros of lanes beyond n are false), max(y), sum(y).  sum(y > 0) == 0 (glm_rob_nb.r:38) is max(y) == 0.
     * Depths and counts below 2^28 (always, in practice) keep 4x - 5y in 32 bits. */
    int cinv[4], mx[4], afok[4];
    unsigned sy[4];
    unsigned cpack = 0, facts = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        cinv[b] = mx[b] = afok[b] = 0;
        sy[b] = 0;
#pragma unroll
        for (int j = 0; j < VEC; j++) {
            const int x = v[0][j], y = v[1 + b][j] + v[5 + b][j];
            cinv[b] += (unsigned)(4 * x - 5 * y) >> 31; /* sign bit of 4x - 5y; redone below when it could wrap */
            mx[b] = max(mx[b], y);
            sy[b] += (unsigned)y;
        }
        cpack |= (unsigned)cinv[b] << (8 * b);
        facts |= (mx[b] > 0 ? 1u : 0u) << b;
        facts |= (mx[b] >= t_reads ? 1u : 0u) << (4 + b);
        facts |= (mx[b] >= (1 << 28) ? 1u : 0u) << 8;
    }
    facts |= (maxx >= (1 << 28) ? 1u : 0u) << 8;
    if (prm.min_af > 0) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
#pragma unroll
            for (int j = 0; j < VEC; j++)
                afok[b] |= ns_ratio_ge(v[1 + b][j] + v[5 + b][j], v[0][j], prm.min_af);
            facts |= (afok[b] ? 1u : 0u) << (9 + b);
        }
    }
    const unsigned cov = __reduce_add_sync(FULL, (unsigned)n_cov | ((unsigned)n_less << 16));
    cpack = __reduce_add_sync(FULL, cpack);
    facts = __reduce_or_sync(FULL, facts);
    n_cov = (int)(cov & 0xffffu);
    n_less = (int)(cov >> 16);
#pragma unroll
    for (int b = 0; b < 4; b++)
        cinv[b] = (int)((cpack >> (8 * b)) & 0xffu);
    if (facts & (1u << 8)) {
        /* depths or counts of 2^28 and more (the latter only in malformed input): the inversion test in 64 bits */
#pragma unroll
        for (int b = 0; b < 4; b++) {
            int c = 0;
#pragma unroll
            for (int j = 0; j < VEC; j++)
                c += 5LL * (long long)(v[1 + b][j] + v[5 + b][j]) > 4LL * (long long)v[0][j];
            cinv[b] = __reduce_add_sync(FULL, c);
        }
    }
    bool med_low; /* median(x) < min_coverage */
    if (n & 1)
        med_low = n_less >= (n + 1) / 2;
    else if (n_less >= n / 2 + 1)
        med_low = true;
    else if (n_less <= n / 2 - 1)
        med_low = false;
    else { /* the two middle order statistics straddle the threshold (rare): they are max{x < c} and min{x >= c} */
        int max_less = -1, min_geq = 0x7fffffff;
        if (act) {
#pragma unroll
            for (int j = 0; j < VEC; j++) {
                const int x = v[0][j];
                if (x < t_lt)
                    max_less = max(max_less, x);
                else
                    min_geq = min(min_geq, x);
            }
        }
        max_less = __reduce_max_sync(FULL, max_less);
        min_geq = __reduce_min_sync(FULL, min_geq);
        med_low = ((double)max_less + (double)min_geq) / 2 < prm.min_coverage;
    }
    const bool pos_gate = med_low || n_cov < prm.size_min;
#define NS_SEL4(a, b) ((b) == 0 ? a[0] : (b) == 1 ? a[1] : (b) == 2 ? a[2] : a[3])
    /* lane k < 3 = allele k of setdiff(c(A,T,C,G), ref) */
    const int k = lane < 3 ? lane : 0;
    const int alt = k + (k >= ref ? 1 : 0);
    const bool inv = (double)NS_SEL4(cinv, alt) > 0.5 * (double)n; /* needlestack.r:330 */
    const int yb = inv ? ref : alt;
    /* max(y) < min_reads (y integer: max(y) < ceil(min_reads)) or sum(y > 0) == 0, glm_rob_nb.r:38 */
    bool gate = pos_gate || !((facts >> (4 + yb)) & 1u) || !((facts >> yb) & 1u);
    if (prm.min_af > 0)
        gate = gate || !((facts >> (9 + yb)) & 1u);
    const uint32_t flags = (inv ? NS_F_REF_INV : 0u) | (gate ? NS_F_GATED : 0u);
    const bool writer = lane < 3 && u_first + lane < n_units;
    /* pooled allelic fraction times the largest depth = first-fit mean of the deepest sample: needed (and reduced)
     * only when a unit of the position goes on to the fit */
    double mu_hint = 0.0;
    if (__any_sync(FULL, writer && !gate)) {
        maxx = __reduce_max_sync(FULL, maxx);
        sx = __reduce_add_sync(FULL, sx);
#pragma unroll
        for (int b = 0; b < 4; b++)
            sy[b] = __reduce_add_sync(FULL, sy[b]);
        mu_hint = (!gate && sx > 0) ? 1.1 * (double)NS_SEL4(sy, yb) / (double)sx * (double)maxx : 0.0;
    }
#undef NS_SEL4
    int cls = ns_gate_commit(ua, prm, u_first + k, flags, 0, mu_hint, writer, lean);
    cls = writer ? cls << (3 * lane) : 0;
    return (int)__reduce_or_sync(FULL, (unsigned)cls);
}

/* integer form of max(y) < min_reads for integer counts: max(y) >= t_reads passes */
__device__ __forceinline__ int ns_gate_t_reads(const ns_params &prm)
{
    const double c = ceil(prm.min_reads);
    return c <= 0.0 ? 0 : (c >= 2147483647.0 ? 2147483647 : (int)c);
}

template <int VEC, bool SH>
__device__ __forceinline__ int ns_gate_position_vec(const int32_t *pos_base, int n, int ref, const ns_params &prm, int t_gt,
                                                    int t_lt, int u_first, int n_units, const ns_unit_arrays &ua, bool lean)
{
    int v[9][VEC];
    ns_gate_load_vec<VEC, SH>(pos_base, n, v);
    return ns_gate_compute_vec<VEC>(v, n, ref, prm, t_gt, t_lt, ns_gate_t_reads(prm), u_first, n_units, ua, lean);
}

/* one position through the form that fits n and the alignment of its planes; SH: planes in a shared-memory stage */
template <bool SH>
__device__ __forceinline__ int ns_gate_dispatch(const int32_t *base, int n, int ref, int vec, const ns_params &prm, int t_gt,
                                                int t_lt, int u_first, int n_units, const ns_unit_arrays &ua, bool lean)
{
    if (vec == 4)
        return ns_gate_position_vec<4, SH>(base, n, ref, prm, t_gt, t_lt, u_first, n_units, ua, lean);
    if (vec == 2)
        return ns_gate_position_vec<2, SH>(base, n, ref, prm, t_gt, t_lt, u_first, n_units, ua, lean);
    if (n <= 64)
        return ns_gate_position_reg<2, SH>(base, n, ref, prm, u_first, n_units, ua, lean);
    if (n <= 128)
        return ns_gate_position_reg<4, SH>(base, n, ref, prm, u_first, n_units, ua, lean);
    return ns_gate_position<SH>(base, n, ref, prm, u_first, n_units, ua, lean);
}

/* ---- bulk-asynchronous staging (sm_90+: cp.async.bulk + mbarrier) -----------------------------------------------
 * The planes of K * NS_WPB_GATE consecutive positions are one contiguous piece of global memory (36 n bytes per
 * position).  One thread asks the copy engine for the whole piece with ONE cp.async.bulk per stage; completion is
 * counted in bytes on the stage's mbarrier; the warps wait on the barrier and read their positions from shared
 * memory.  `stages` pieces per CTA are in flight while one is being consumed, whatever n is - with register
 * loads the bytes in flight shrank with n (3.6 TB/s at n = 50 against 5.0 TB/s at n = 100) and every 16 bytes cost a
 * load instruction plus its address arithmetic in an issue-bound kernel. */
__device__ __forceinline__ void ns_mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ns_mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tNS_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra NS_DONE;\n\t"
                 "bra NS_WAIT;\n\tNS_DONE:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(parity)
                 : "memory");
}
/* arm the barrier with the byte count, then hand the copy to the async proxy (issued by one thread) */
__device__ __forceinline__ void ns_bulk_load(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar)
{
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                 "l"(gmem_src), "r"(bytes), "r"(b)
                 : "memory");
}

#define NS_GATE_SMEM_HDR 128 /* mbarriers in front of the 128-byte aligned stages */

template <int K>
__device__ __forceinline__ void ns_gate_staged(const ns_unit_src &src, const ns_params &prm, int64_t p_first, int n_full_blocks,
                                               int n_units, int stages, int vec, int t_gt, int t_lt, bool lean,
                                               const ns_unit_arrays &ua, ns_counters *cnt)
{
    constexpr int PB = K * NS_WPB_GATE;
    const int n = src.n;
    const int warp = threadIdx.x >> 5;
    const size_t pos_ints = (size_t)9 * n;
    const unsigned stage_bytes = (unsigned)(PB * pos_ints * sizeof(int32_t));
    uint64_t *bar = reinterpret_cast<uint64_t *>(ns_smem);
    char *stage0 = reinterpret_cast<char *>(ns_smem) + NS_GATE_SMEM_HDR;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++)
            ns_mbar_init(bar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < stages; s++) {
            const int blk = blockIdx.x + s * gridDim.x;
            if (blk < n_full_blocks)
                ns_bulk_load(stage0 + (size_t)s * stage_bytes, src.counts + (p_first + (int64_t)blk * PB) * pos_ints, stage_bytes, bar + s);
        }
    }
    __syncthreads();
    int it = 0;
    for (int blk = blockIdx.x; blk < n_full_blocks; blk += gridDim.x, it++) {
        const int s = it % stages;
        const unsigned parity = (unsigned)(it / stages) & 1u;
        int refc[K];
#pragma unroll
        for (int j = 0; j < K; j++)
            refc[j] = src.ref[p_first + (int64_t)blk * PB + j * NS_WPB_GATE + warp]; /* issued before the wait */
        ns_mbar_wait(bar + s, parity);
        const int32_t *st = reinterpret_cast<const int32_t *>(stage0 + (size_t)s * stage_bytes);
        int packed[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int q = blk * PB + j * NS_WPB_GATE + warp; /* position within the chunk */
            packed[j] = ns_gate_dispatch<true>(st + (size_t)(j * NS_WPB_GATE + warp) * pos_ints, n, refc[j], vec, prm, t_gt, t_lt,
                                               3 * q, n_units, ua, lean);
        }
        /* the list insertion is a CTA collective with barriers: behind it every warp has finished reading the stage */
        ns_gate_insert<K>(packed, [&](int j) { return 3 * (blk * PB + j * NS_WPB_GATE + warp); }, n_units, ua, cnt);
        const int nxt = blk + stages * gridDim.x;
        if (threadIdx.x == 0 && nxt < n_full_blocks)
            ns_bulk_load(stage0 + (size_t)s * stage_bytes, src.counts + (p_first + (int64_t)nxt * PB) * pos_ints, stage_bytes, bar + s);
    }
}

/* The bulk-asynchronous form of the gate for whole blocks of stage_k * 8 positions of SNV planes (positions of a stage
 * are 36 n bytes apart: 16-byte aligned rows when n % 4 == 0, 8-byte aligned when n is even); the positions behind
 * the last whole block are left to k_gate(q_begin).  Two CTAs per SM (shared memory bounds the occupancy, so the
 * kernel may use up to 128 registers).  lean: see ns_gate_commit. */
__global__ void __launch_bounds__(NS_WPB_GATE * 32, 2)
k_gate_staged(ns_unit_src src, ns_params prm, int64_t u0, int n_units, ns_unit_arrays ua, ns_counters *cnt, int stage_k,
              int stages, int n_full_blocks, int lean_i)
{
    const int n = src.n;
    const double mcv = prm.min_coverage;
    const int t_gt = mcv < 0 ? -1 : (int)floor(mcv), t_lt = mcv <= 0 ? 0 : (int)ceil(mcv);
    const bool small_mc = mcv < 2.0e9;
    const int vec_s = (small_mc && (n & 3) == 0 && n > 32 && n <= 128) ? 4 : (small_mc && (n & 1) == 0 && n > 16 && n <= 64) ? 2 : 0;
    if (stage_k == 2)
        ns_gate_staged<2>(src, prm, u0 / 3, n_full_blocks, n_units, stages, vec_s, t_gt, t_lt, lean_i != 0, ua, cnt);
    else
        ns_gate_staged<1>(src, prm, u0 / 3, n_full_blocks, n_units, stages, vec_s, t_gt, t_lt, lean_i != 0, ua, cnt);
}

/* The gate over the positions q_begin .. of a chunk, one instantiation per load form so that each gets its own
 * register allocation under the 64-register cap (four CTAs per SM):
 *   FORM 4  n % 4 == 0, 32 < n <= 128, 16-byte aligned planes: nine 128-bit loads per lane
 *   FORM 2  n even, 16 < n <= 64, 8-byte aligned: nine 64-bit loads per lane, the NEXT position's rows requested
 *           before the current one is evaluated (a position is only 2.3 KB per warp: too little in flight otherwise)
 *   FORM 0  everything else: scalar register / strided loads, explicit-unit sources, the extra-robust pre-filter
 * The host picks the form with ns_gate_form() (same conditions). */
template <int FORM>
__device__ __forceinline__ void ns_gate_body(const ns_unit_src &src, const ns_params &prm, int64_t u0, int n_units,
                                             const ns_unit_arrays &ua, ns_counters *cnt, int q_begin, bool lean)
{
    const int n = src.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wstride = gridDim.x * NS_WPB_GATE;
    if (src.mode == 0 && !prm.extra_rob && (u0 % 3) == 0) {
        /* fast path: warp = position (uniform trip count per CTA: the list insertion is a CTA collective) */
        const int n_pos = (n_units + 2) / 3;
        /* integer forms of the min_coverage tests */
        const double mcv = prm.min_coverage;
        const int t_gt = mcv < 0 ? -1 : (int)floor(mcv), t_lt = mcv <= 0 ? 0 : (int)ceil(mcv);
        constexpr int K = 4; /* positions per warp between two list insertions */
        if (FORM == 2) {
            const int t_reads = ns_gate_t_reads(prm);
            for (int q0 = q_begin + blockIdx.x * NS_WPB_GATE * K; q0 < n_pos; q0 += wstride * K) {
                int packed[K];
                int va[9][2], vb[9][2];
                int ra = 4, rb = 4;
                {
                    const int q = q0 + warp;
                    if (q < n_pos) {
                        ns_gate_load_vec<2, false>(src.counts + (u0 / 3 + q) * 9 * (int64_t)n, n, va);
                        ra = src.ref[u0 / 3 + q];
                    }
                }
#pragma unroll
                for (int j = 0; j < K; j++) {
                    const int q = q0 + j * NS_WPB_GATE + warp, qn = q + NS_WPB_GATE;
                    int(&cur)[9][2] = (j & 1) ? vb : va;
                    int(&nxt)[9][2] = (j & 1) ? va : vb;
                    int &rcur = (j & 1) ? rb : ra;
                    int &rnxt = (j & 1) ? ra : rb;
                    if (j + 1 < K && qn < n_pos) {
                        ns_gate_load_vec<2, false>(src.counts + (u0 / 3 + qn) * 9 * (int64_t)n, n, nxt);
                        rnxt = src.ref[u0 / 3 + qn];
                    }
                    packed[j] = q < n_pos ? ns_gate_compute_vec<2>(cur, n, rcur, prm, t_gt, t_lt, t_reads, 3 * q, n_units, ua, lean) : 0;
                }
                ns_gate_insert<K>(packed, [&](int j) { return 3 * (q0 + j * NS_WPB_GATE + warp); }, n_units, ua, cnt);
            }
            return;
        }
        for (int q0 = q_begin + blockIdx.x * NS_WPB_GATE * K; q0 < n_pos; q0 += wstride * K) {
            int packed[K];
#pragma unroll 1
            for (int j = 0; j < K; j++) {
                const int q = q0 + j * NS_WPB_GATE + warp;
                int pk = 0;
                if (q < n_pos) {
                    const int64_t p = u0 / 3 + q;
                    const int32_t *base = src.counts + p * 9 * (int64_t)n;
                    /* vector form: the loads are in flight before the REF code is needed */
                    if (FORM == 4)
                        pk = ns_gate_position_vec<4, false>(base, n, src.ref[p], prm, t_gt, t_lt, 3 * q, n_units, ua, lean);
                    else
                        pk = ns_gate_dispatch<false>(base, n, src.ref[p], 0, prm, t_gt, t_lt, 3 * q, n_units, ua, lean);
                }
#pragma unroll
                for (int t = 0; t < K; t++)
                    if (t == j)
                        packed[t] = pk;
            }
            ns_gate_insert<K>(packed, [&](int j) { return 3 * (q0 + j * NS_WPB_GATE + warp); }, n_units, ua, cnt);
        }
        return;
    }
    if (FORM != 0)
        return; /* the host never launches a vector form for these sources */
    double *scratch = ns_smem + (size_t)warp * (size_t)n;
    for (int v0 = blockIdx.x * NS_WPB_GATE; v0 < n_units; v0 += wstride) {
        const int u = v0 + warp;
        int packed = 0;
        if (u < n_units) {
            ns_rows r = ns_unit_rows(src, u0 + u);
            double mu_hint = 0;
            uint32_t flags = ns_gate_unit<WarpG>(r, n, prm, src.plain, scratch, &mu_hint);
            packed = ns_gate_commit(ua, prm, u, flags, r.is_indel, mu_hint, lane == 0, lean);
        }
        const int pk1[1] = {packed};
        ns_gate_insert<1>(pk1, [&](int) { return u; }, n_units, ua, cnt);
    }
}

/* q_begin: first position of the chunk this launch handles (the whole blocks before it went through k_gate_staged);
 * lean: see ns_gate_commit */
__global__ void __launch_bounds__(NS_WPB_GATE * 32, 4)
k_gate(ns_unit_src src, ns_params prm, int64_t u0, int n_units, ns_unit_arrays ua, ns_counters *cnt, int q_begin, int lean_i)
{
    ns_gate_body<0>(src, prm, u0, n_units, ua, cnt, q_begin, lean_i != 0);
}
__global__ void __launch_bounds__(NS_WPB_GATE * 32, 4)
k_gate_v4(ns_unit_src src, ns_params prm, int64_t u0, int n_units, ns_unit_arrays ua, ns_counters *cnt, int q_begin, int lean_i)
{
    ns_gate_body<4>(src, prm, u0, n_units, ua, cnt, q_begin, lean_i != 0);
}
__global__ void __launch_bounds__(NS_WPB_GATE * 32, 4)
k_gate_v2(ns_unit_src src, ns_params prm, int64_t u0, int n_units, ns_unit_arrays ua, ns_counters *cnt, int q_begin, int lean_i)
{
    ns_gate_body<2>(src, prm, u0, n_units, ua, cnt, q_begin, lean_i != 0);
}

__global__ void k_scan_total(const int *flag, const int *excl, int n, int *total)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *total = n > 0 ? excl[n - 1] + flag[n - 1] : 0;
}

/* Rows of the emitted units, in unit order, into the packed planes.  One THREAD per unit does the bookkeeping (row
 * index or -1, skipped-position count, insertion point of a unit whose fit runs in the call-wide pool) with coalesced
 * accesses - a WGS-like chunk has millions of units and a few hundred emitted rows - and the warp then copies the rows
 * of its emitting lanes together. */
__global__ void k_pack(int n, int n_units, int64_t u0, int64_t emit_off, int cap, ns_unit_arrays ua,
                       ns_planes src, ns_planes dst, int64_t *emit_unit, int *emit_index_out, int64_t *pool_lightpos,
                       ns_counters *cnt)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    int n_skip = 0;
    const int stride = gridDim.x * blockDim.x;
    for (int base = blockIdx.x * blockDim.x + threadIdx.x - lane; base < n_units; base += stride) {
        const int u = base + lane;
        int d = 0, row = 0;
        bool copy = false;
        if (u < n_units) {
            const uint32_t fl = ua.flags[u];
            n_skip += (fl & NS_F_SKIPPED) != 0;
            if (fl & NS_F_POOLED) {
                pool_lightpos[ua.row[u]] = emit_off + ua.emitidx[u];
                emit_index_out[u] = -1;
            } else if (!ua.emitflag[u]) {
                emit_index_out[u] = -1;
            } else {
                d = ua.emitidx[u];
                emit_index_out[u] = (int)(emit_off + d);
                copy = d < cap;
                row = ua.row[u];
            }
        }
        unsigned m = __ballot_sync(FULL, copy);
        while (m) {
            const int sl = __ffs(m) - 1;
            m &= m - 1;
            const int dd = __shfl_sync(FULL, d, sl), rr = __shfl_sync(FULL, row, sl), uu = base + sl;
            const size_t so = (size_t)rr * n, dof = (size_t)dd * n;
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
                dst.pooled[(size_t)dd * 8 + lane] = src.pooled[(size_t)rr * 8 + lane];
            if (lane == 0) {
                emit_unit[dd] = u0 + uu;
                dst.rsigma[dd] = ua.sigma[uu];
                dst.rslope[dd] = ua.slope[uu];
                dst.rmaxgq[dd] = ua.max_gq[uu];
                dst.rflags[dd] = ua.flags[uu];
            }
        }
    }
    n_skip = __reduce_add_sync(FULL, n_skip);
    if (lane == 0 && n_skip)
        atomicAdd(&cnt->n_skipped, n_skip);
}

/* Copies the count rows of `count` listed units of the current chunk into the pool slots base .. base+count-1
 * (warp per unit), marks the units NS_F_POOLED in the chunk's arrays (ua.row[u] then holds the pool slot) and
 * initialises the counter block of the cluster-kernel launch that will fit them. */
__global__ void k_pool_gather(ns_unit_src src, int64_t u0_src, int64_t u_off, const int *list, int count,
                              ns_unit_arrays ua, ns_pool_dev pool, int base, ns_counters *launch_cnt)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int n = src.n;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ns_counters z;
        memset(&z, 0, sizeof(z));
        z.n_pre = count;
        *launch_cnt = z;
    }
    for (int i = blockIdx.x * wpb + warp; i < count; i += gridDim.x * wpb) {
        const int u = list[i];
        const int slot = base + i;
        const ns_rows r = ns_unit_rows(src, u0_src + u);
        const size_t o = (size_t)slot * n;
        for (int k = lane; k < n; k += 32) {
            pool.dp[o + k] = r.dp[k];
            pool.vp[o + k] = r.vp[k];
            pool.vm[o + k] = r.vm ? r.vm[k] : 0;
            pool.rp[o + k] = r.rp ? r.rp[k] : 0;
            pool.rm[o + k] = r.rm ? r.rm[k] : 0;
        }
        if (lane == 0) {
            const uint32_t fl = ua.flags[u];
            pool.flags[slot] = fl;
            pool.is_indel[slot] = (uint8_t)r.is_indel;
            pool.unit[slot] = u_off + u;
            pool.lightpos[slot] = 0;
            ua.flags[u] = fl | NS_F_POOLED;
            ua.row[u] = slot;
        }
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

