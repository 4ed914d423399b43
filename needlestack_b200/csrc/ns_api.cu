/*
 * ns_api.cu — CUDA kernels (sm_100a) and the C ABI of include/ns_b200.h.
 *
 * Pipeline per chunk of units (one unit = one glmrob.nb() call of the reference):
 *   k_gate   warp per unit, streams the count rows (HBM bound): ref-inversion test,
 *            extra-robust pre-filter, NA gate (needlestack.r:330-337, glm_rob_nb.r:20-49);
 *            appends survivors to the fit list
 *   k_fit    persistent warps pull units from the fit list (dynamic load balance) and run
 *            the robust NB regression (glm_rob_nb.r:58-205) — FP64 pipe bound
 *   k_post   p/q/GQ, Holm power q-values, T/N status, strand bias, genotypes
 *            (glm_rob_nb.r:219-224, needlestack.r:340-409)
 *   k_pack   gathers the rows of the emitted units, in unit order, for the D2H copy
 * Units that take the cluster-per-unit kernel (spline + quadrature windows) leave the chunk: they go to a
 * call-wide pool (HeavyPool) that runs on side streams next to later chunks and is joined once per call.
 * There is no CPU implementation behind this file: without a CUDA device ns_create() fails.
 */
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <vector>

#include "ns_kernels.cuh"
#include "ns_hostmerge.hpp"
#include "ns_api_internal.hpp"

#define NS_GATE_SMEM_HDR 128
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->err = b_;                                                                         \
            return NS_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)


/* ---- page-locked growable host planes ----------------------------------------------------------- */
void HostRows::release()
{
    double **dpl[] = {&p.pvalue, &p.qvalue, &p.gq, &p.qval_minaf, &p.sor, &p.rvsb, &p.fs, &p.pooled,
                      &p.row_sigma, &p.row_slope, &p.row_max_gq};
    for (double **d : dpl) {
        if (*d)
            cudaFreeHost(*d);
        *d = nullptr;
    }
    if (p.gt)
        cudaFreeHost(p.gt);
    if (p.status)
        cudaFreeHost(p.status);
    if (p.emit_unit)
        cudaFreeHost(p.emit_unit);
    if (p.row_flags)
        cudaFreeHost(p.row_flags);
    p.gt = p.status = nullptr;
    p.emit_unit = nullptr;
    p.row_flags = nullptr;
    cap = 0;
}

bool HostRows::reserve(int64_t rows, int n_samples, int64_t keep)
{
    if (rows <= cap && n_samples == n)
        return true;
    if (n_samples != n)
        keep = 0;
    int64_t want = rows + rows / 2 + 256;
    ns_row_planes q = {};
    const size_t rb = (size_t)n_samples * sizeof(double);
    bool ok = true;
    auto al = [&](void **dst, size_t bytes) {
        if (ok && cudaHostAlloc(dst, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
            *dst = nullptr;
            ok = false;
        }
    };
    al((void **)&q.pvalue, want * rb), al((void **)&q.qvalue, want * rb), al((void **)&q.gq, want * rb);
    al((void **)&q.qval_minaf, want * rb), al((void **)&q.sor, want * rb), al((void **)&q.rvsb, want * rb);
    al((void **)&q.fs, want * rb), al((void **)&q.gt, want * (size_t)n_samples), al((void **)&q.status, want * (size_t)n_samples);
    al((void **)&q.pooled, want * 8 * sizeof(double)), al((void **)&q.emit_unit, want * sizeof(int64_t));
    al((void **)&q.row_sigma, want * sizeof(double)), al((void **)&q.row_slope, want * sizeof(double));
    al((void **)&q.row_max_gq, want * sizeof(double)), al((void **)&q.row_flags, want * sizeof(uint32_t));
    if (!ok) {
        HostRows tmp;
        tmp.p = q;
        tmp.release();
        return false;
    }
    if (keep > 0 && cap > 0)
        ns_rows_copy(q, 0, p, 0, keep, n_samples);
    release();
    p = q;
    cap = want;
    n = n_samples;
    return true;
}

RowSink ns_sink_from_out(ns_out *out)
{
    RowSink s;
    s.p.pvalue = out->pvalue, s.p.qvalue = out->qvalue, s.p.gq = out->gq, s.p.qval_minaf = out->qval_minaf;
    s.p.sor = out->sor, s.p.rvsb = out->rvsb, s.p.fs = out->fs, s.p.gt = out->gt, s.p.status = out->status;
    s.p.pooled = out->pooled, s.p.emit_unit = out->emit_unit;
    s.p.row_sigma = out->row_sigma, s.p.row_slope = out->row_slope, s.p.row_max_gq = out->row_max_gq;
    s.p.row_flags = out->row_flags;
    s.cap = out->emit_cap;
    s.want_rows = out->pvalue || out->qvalue || out->gq || out->qval_minaf || out->sor || out->rvsb || out->fs ||
                  out->gt || out->status || out->pooled || out->emit_unit || out->row_sigma || out->row_slope ||
                  out->row_max_gq || out->row_flags;
    return s;
}

/* Grant a kernel its dynamic shared memory: the attribute belongs to the FUNCTION on a device, not to a context, and
 * setting it again with a smaller size lowers it - so the largest size granted so far is kept per (device, kernel)
 * for the whole process and only ever raised (it was re-issued for every chunk before). */
#include <mutex>
static std::mutex g_smem_mu;
static size_t g_smem_set[64][8];
template <class K> static cudaError_t ns_need_smem(ns_ctx *ctx, int slot, K kern, size_t bytes)
{
    if (bytes <= 48 * 1024)
        return cudaSuccess;
    std::lock_guard<std::mutex> lk(g_smem_mu);
    size_t &have = g_smem_set[ctx->device & 63][slot];
    if (bytes <= have)
        return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess)
        have = bytes;
    return e;
}

/* wait for everything this context has in flight: called on every error return, so that no copy is still reading
 * the caller's host buffers (or writing the caller's outputs) when the call hands control back */
void ns_drain(ns_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    if (ctx->copy_stream)
        cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->heavy_stream)
        cudaStreamSynchronize(ctx->heavy_stream);
    for (int i = 0; i < NS_POOL_STREAMS; i++)
        if (ctx->pool.streams[i])
            cudaStreamSynchronize(ctx->pool.streams[i]);
    if (ctx->stream)
        cudaStreamSynchronize(ctx->stream);
    ctx->pool.n_used = ctx->pool.n_launch = 0;
    for (int i = 0; i < NS_POOL_STREAMS; i++)
        ctx->pool.used[i] = false;
    cudaGetLastError();
}

extern "C" {

int ns_abi_version(void) { return NS_ABI_VERSION; }

size_t ns_sizeof(int which)
{
    return which == 0 ? sizeof(ns_params) : which == 1 ? sizeof(ns_out) : which == 2 ? sizeof(ns_timings) : 0;
}

void ns_params_glm_default(ns_params *p)
{ /* bin/glm_rob_nb.r:16-18 */
    memset(p, 0, sizeof(*p));
    p->c_tukey_beta = 5;
    p->c_tukey_sig = 3;
    p->minsig = 1e-3;
    p->maxsig = 10;
    p->minmu = 1e-10;
    p->sig_prec = 1e-8;
    p->tol = 1e-6;
    p->maxit = 30;
    p->maxit_sig = 50;
    p->n_ai_sig_tukey = 100;
    p->min_coverage = 1;
    p->min_reads = 1;
    p->min_af = 0;
    p->size_min = 10;
    p->extra_rob = 1;
    p->min_af_extra_rob = 0.2;
    p->min_prop_extra_rob = 0.1;
    p->max_prop_extra_rob = 0.5;
    p->GQ_threshold = 50;
    p->SB_threshold_SNV = 100;
    p->SB_threshold_indel = 100;
    p->sigma_normal = 0.1;
    p->afmin_power = -1;
    p->emit_mode = NS_EMIT_ALL;
}

void ns_params_caller_default(ns_params *p)
{ /* bin/needlestack.r:64-92 */
    ns_params_glm_default(p);
    p->min_coverage = 30;
    p->min_reads = 3;
    p->min_af = 0;
    p->extra_rob = 0;
    p->max_prop_extra_rob = 0.8;
    p->SB_type = 0;
    p->emit_mode = NS_EMIT_CALLS;
}

ns_ctx *ns_create(int device, char *err, size_t errlen)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        if (err && errlen)
            snprintf(err, errlen, "ns_create: no usable CUDA device %d (%s, %d devices); this library has no CPU path",
                     device, cudaGetErrorString(e), ndev);
        return nullptr;
    }
    cudaSetDevice(device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10) {
        if (err && errlen)
            snprintf(err, errlen, "ns_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                     device, prop.major, prop.minor);
        return nullptr;
    }
    ns_ctx *ctx = new ns_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->heavy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreate(&ctx->ev_h0) == cudaSuccess && cudaEventCreate(&ctx->ev_h1) == cudaSuccess;
    for (int i = 0; i < 2 && ok; i++)
        ok = cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 10 && ok; i++)
        ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
    HeavyPool &pl = ctx->pool;
    for (int i = 0; i < NS_POOL_STREAMS && ok; i++)
        ok = cudaStreamCreateWithFlags(&pl.streams[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&pl.ev_done[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreate(&pl.ev_t0[i]) == cudaSuccess && cudaEventCreate(&pl.ev_t1[i]) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&pl.ev_ready, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreate(&pl.ev_tail0) == cudaSuccess && cudaEventCreate(&pl.ev_tail1) == cudaSuccess &&
         cudaEventCreate(&pl.ev_post0) == cudaSuccess && cudaEventCreate(&pl.ev_post1) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&pl.h_count, 4 * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&ctx->h_counters, sizeof(ns_counters)) == cudaSuccess;
    ok = ok && cudaHostAlloc((void **)&ctx->h_started, sizeof(int), cudaHostAllocMapped) == cudaSuccess &&
         cudaHostGetDevicePointer((void **)&ctx->d_started, ctx->h_started, 0) == cudaSuccess;
    ok = ok && ctx->counters.reserve(1) == cudaSuccess;
    if (!ok) {
        if (err && errlen)
            snprintf(err, errlen, "ns_create: CUDA resource creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        ns_destroy(ctx); /* releases whatever was created */
        return nullptr;
    }
    ctx->stream = ctx->own_stream;
    const char *cu = getenv("NS_CHUNK_UNITS");
    if (cu) {
        /* whole positions only (the streaming gate takes one warp per position), at least one */
        int64_t v = atoll(cu);
        v -= v % 3;
        ctx->chunk_units = v < 3 ? 3 : v;
    }
    const char *ep = getenv("NS_POOL");
    if (ep && atoi(ep) == 0)
        pl.enabled = false;
    return ctx;
}

void ns_destroy(ns_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; i++) {
        ctx->in_counts[i].release();
        ctx->in_ref[i].release();
        if (ctx->ev_h2d[i])
            cudaEventDestroy(ctx->ev_h2d[i]);
        if (ctx->ev_free[i])
            cudaEventDestroy(ctx->ev_free[i]);
    }
    for (int i = 0; i < 5; i++)
        ctx->in_rows[i].release();
    ctx->in_indel.release();
    ctx->in_counts16[0].release();
    ctx->in_counts16[1].release();
    ctx->grid_out.release();
    ctx->flags.release(), ctx->iters.release(), ctx->cycles.release(), ctx->sigma.release(), ctx->slope.release(), ctx->max_gq.release();
    ctx->needrow.release(), ctx->row.release(), ctx->emitflag.release(), ctx->emitidx.release();
    ctx->fit_list.release(), ctx->post_list.release(), ctx->heavy_list.release(), ctx->pre_list.release(), ctx->emit_index_out.release(), ctx->emit_unit.release();
    ctx->counters.release(), ctx->cub_tmp.release(), ctx->rows.release(), ctx->packed.release();
    ctx->pair_idx.release(), ctx->role.release(), ctx->qtabN.release(), ctx->qtabT.release();
    HeavyPool &pl = ctx->pool;
    for (int i = 0; i < 5; i++)
        pl.rows[i].release();
    pl.is_indel.release(), pl.flags.release(), pl.iters.release(), pl.cycles.release(), pl.sigma.release(), pl.slope.release();
    pl.max_gq.release(), pl.unit.release(), pl.lightpos.release(), pl.ident.release(), pl.emitflag.release(), pl.cnt.release();
    pl.planes.release();
    pl.h_rows.release();
    if (pl.h_units)
        cudaFreeHost(pl.h_units);
    if (pl.h_count)
        cudaFreeHost(pl.h_count);
    for (int i = 0; i < NS_POOL_STREAMS; i++) {
        if (pl.streams[i])
            cudaStreamDestroy(pl.streams[i]);
        if (pl.ev_done[i])
            cudaEventDestroy(pl.ev_done[i]);
        if (pl.ev_t0[i])
            cudaEventDestroy(pl.ev_t0[i]);
        if (pl.ev_t1[i])
            cudaEventDestroy(pl.ev_t1[i]);
    }
    cudaEvent_t pev[] = {pl.ev_ready, pl.ev_tail0, pl.ev_tail1, pl.ev_post0, pl.ev_post1, ctx->ev_fork, ctx->ev_join, ctx->ev_h0, ctx->ev_h1};
    for (cudaEvent_t e : pev)
        if (e)
            cudaEventDestroy(e);
    for (int i = 0; i < 10; i++)
        if (ctx->ev[i])
            cudaEventDestroy(ctx->ev[i]);
    if (ctx->h_counters)
        cudaFreeHost(ctx->h_counters);
    if (ctx->h_started)
        cudaFreeHost(ctx->h_started);
    if (ctx->own_stream)
        cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream)
        cudaStreamDestroy(ctx->copy_stream);
    if (ctx->heavy_stream)
        cudaStreamDestroy(ctx->heavy_stream);
    delete ctx;
}

const char *ns_last_error(ns_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int ns_set_stream(ns_ctx *ctx, void *cuda_stream)
{
    if (!ctx)
        return NS_ERR_ARG;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return NS_OK;
}

void *ns_alloc_pinned(size_t bytes)
{
    void *p = nullptr;
    /* portable: page-locked for every device of the process (a multi-GPU call copies from one host buffer) */
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess)
        return nullptr;
    return p;
}
void ns_free_pinned(void *p)
{
    if (p)
        cudaFreeHost(p);
}

int ns_get_timings(ns_ctx *ctx, ns_timings *t)
{
    if (!ctx || !t)
        return NS_ERR_ARG;
    *t = ctx->tm;
    return NS_OK;
}

int ns_measure_fp64_peak(ns_ctx *ctx, double *tflops)
{
    if (!ctx || !tflops)
        return NS_ERR_ARG;
    cudaSetDevice(ctx->device);
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 1 << 14;
    DevBuf<double> out;
    CK(out.reserve((size_t)blocks * threads));
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(ctx->ev[0], ctx->stream));
        k_dfma<<<blocks, threads, 0, ctx->stream>>>(out.p, iters, 0.999999, 1e-9);
        CK(cudaEventRecord(ctx->ev[1], ctx->stream));
        CK(cudaEventSynchronize(ctx->ev[1]));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        double fl = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
        double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best)
            best = tf;
    }
    out.release();
    *tflops = best;
    return NS_OK;
}

} /* extern "C" */

/* upload pair index sets, build the role array and the y* tables; returns device params */
int ns_prepare_params(ns_ctx *ctx, const ns_params *in, int n, ns_params *dprm, const double **qtab)
{
    *dprm = *in;
    if (in->n_ai_sig_tukey < 3 || in->n_ai_sig_tukey > NS_MAX_KNOTS) {
        ctx->err = "n_ai.sig.tukey must be in [3, 128]";
        return NS_ERR_ARG;
    }
    dprm->tindex = dprm->nindex = dprm->only_t = dprm->only_n = nullptr;
    dprm->role = nullptr;
    if (in->is_tn) {
        const int np = in->n_pairs, nt = in->n_only_t, nn = in->n_only_n;
        if (np < 0 || nt < 0 || nn < 0 || (np > 0 && (!in->tindex || !in->nindex)) || (nt > 0 && !in->only_t) ||
            (nn > 0 && !in->only_n)) {
            ctx->err = "tumour/normal index sets missing";
            return NS_ERR_ARG;
        }
        std::vector<int32_t> h((size_t)2 * np + nt + nn + 4);
        std::vector<uint8_t> role((size_t)n, 0);
        auto chk = [&](const int32_t *a, int m) {
            for (int i = 0; i < m; i++)
                if (a[i] < 0 || a[i] >= n)
                    return false;
            return true;
        };
        if (!chk(in->tindex, np) || !chk(in->nindex, np) || !chk(in->only_t, nt) || !chk(in->only_n, nn)) {
            ctx->err = "tumour/normal sample index out of range";
            return NS_ERR_ARG;
        }
        /* assignment order of needlestack.r:343-347: N of pairs, T of pairs, only N, only T */
        for (int i = 0; i < np; i++)
            h[i] = in->tindex[i], h[np + i] = in->nindex[i];
        for (int i = 0; i < nt; i++)
            h[2 * np + i] = in->only_t[i];
        for (int i = 0; i < nn; i++)
            h[2 * np + nt + i] = in->only_n[i];
        for (int i = 0; i < np; i++)
            role[in->nindex[i]] = 1;
        for (int i = 0; i < np; i++)
            role[in->tindex[i]] = 2;
        for (int i = 0; i < nn; i++)
            role[in->only_n[i]] = 3;
        for (int i = 0; i < nt; i++)
            role[in->only_t[i]] = 4;
        CK(ctx->pair_idx.reserve(h.size()));
        CK(ctx->role.reserve((size_t)n));
        CK(cudaMemcpyAsync(ctx->pair_idx.p, h.data(), h.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->role.p, role.data(), (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        dprm->tindex = ctx->pair_idx.p;
        dprm->nindex = ctx->pair_idx.p + np;
        dprm->only_t = ctx->pair_idx.p + 2 * np;
        dprm->only_n = ctx->pair_idx.p + 2 * np + nt;
        dprm->role = ctx->role.p;
    }
    double afmin = in->afmin_power;
    if (in->is_tn && afmin == -1)
        afmin = 0.01;
    if (ctx->qtab_sigma != in->sigma_normal || ctx->qtab_afmin != afmin || !ctx->qtabN.p) {
        CK(ctx->qtabN.reserve(NS_QTAB_LEN));
        k_qtab<<<(NS_QTAB_LEN + 127) / 128, 128, 0, ctx->stream>>>(ctx->qtabN.p, nullptr, NS_QTAB_LEN, in->sigma_normal,
                                                                  afmin);
        CK(cudaGetLastError());
        ctx->qtab_sigma = in->sigma_normal;
        ctx->qtab_afmin = afmin;
        ctx->tm.kernel_launches++;
    }
    *qtab = ctx->qtabN.p;
    return NS_OK;
}

static cudaError_t ns_reserve_units(ns_ctx *ctx, size_t U)
{
    cudaError_t e;
    if ((e = ctx->flags.reserve(U)) || (e = ctx->iters.reserve(U)) || (e = ctx->cycles.reserve(U)) || (e = ctx->sigma.reserve(U)) ||
        (e = ctx->slope.reserve(U)) || (e = ctx->max_gq.reserve(U)) || (e = ctx->needrow.reserve(U)) ||
        (e = ctx->row.reserve(U)) || (e = ctx->emitflag.reserve(U)) || (e = ctx->emitidx.reserve(U)) ||
        (e = ctx->fit_list.reserve(U)) || (e = ctx->post_list.reserve(U)) || (e = ctx->heavy_list.reserve(U)) || (e = ctx->pre_list.reserve(U)) || (e = ctx->emit_index_out.reserve(U)))
        return e;
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (int *)nullptr, (int *)nullptr, (int)U);
    return ctx->cub_tmp.reserve(tmp + 256);
}

/* geometry of the cluster-per-unit kernel: threads per CTA and CTAs per cluster */
static void ns_heavy_geometry(int n, int n_units, int sm_count, int *threads, int *cl)
{
    int th = NS_HEAVY_THREADS;
    const char *eth = getenv("NS_HEAVY_THREADS");
    if (eth && (atoi(eth) == 128 || atoi(eth) == 256 || atoi(eth) == 512 || atoi(eth) == NS_HEAVY_THREADS_TP))
        th = atoi(eth); /* tests and experiments; NS_HEAVY_THREADS_TP selects the throughput form of the kernel */
    const int wpc = th / 32;
    /* just enough CTAs for one sample per warp, up to the portable cluster size 8 (n = 100: 7 CTAs of 16 warps -
     * a power of two would hold an eighth SM whose warps have no sample; C2 step 371 -> 361 ms) */
    int c = (n + wpc - 1) / wpc;
    if (c > 8)
        c = 8;
    if (c < 1)
        c = 1;
    /* many queued units (deep panels: every unit takes this kernel): throughput matters more than the
     * latency of one unit, and smaller clusters waste fewer warps at the barriers */
    const int c_latency = c;
    while (c > 1 && n_units > 8 * (sm_count / c))
        c >>= 1;
    /* down to one CTA per unit it is the throughput form of the kernel (k_fit_heavy_tp: 20 warps, 96 registers) */
    if (c == 1 && c_latency > 1 && !eth)
        th = NS_HEAVY_THREADS_TP;
    const char *ecl = getenv("NS_HEAVY_CLUSTER");
    if (ecl && atoi(ecl) >= 1 && atoi(ecl) <= 16)
        c = atoi(ecl);
    *threads = th;
    *cl = c;
}

static size_t ns_heavy_smem(int n, int threads)
{
    return (NS_RTAB_HEAVY + ns_work_doubles(n) + (size_t)(threads / 32) * NS_COOP_DOUBLES) * sizeof(double);
}

/* cluster launch of k_fit_heavy: CL CTAs per unit, as many clusters as can be co-resident (at most max_clusters
 * when > 0).  cp: the counter block holding this launch's unit count and list cursor. */
static int ns_launch_heavy(ns_ctx *ctx, cudaStream_t st, int n, int n_units_max, const ns_unit_src &src,
                           const ns_params &dprm, int64_t u0, const ns_unit_arrays &ua, ns_counters *cp, int which,
                           int *started, int max_clusters, int *n_clusters_out)
{
    int threads, cl;
    ns_heavy_geometry(n, n_units_max, ctx->sm_count, &threads, &cl);
    const size_t smh = ns_heavy_smem(n, threads);
    if (smh > 227 * 1024) {
        ctx->err = "too many samples for the cluster kernel's working set";
        return NS_ERR_ARG;
    }
    const bool tp = threads == NS_HEAVY_THREADS_TP;
    auto kern = tp ? k_fit_heavy_tp : k_fit_heavy;
    CK(ns_need_smem(ctx, tp ? 1 : 0, kern, smh));
    if (cl > 8)
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smh;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3(cl);
    int ncl = 0;
    CK(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
    if (ncl < 1)
        ncl = 1;
    if (ncl > n_units_max)
        ncl = n_units_max;
    if (max_clusters > 0 && ncl > max_clusters)
        ncl = max_clusters;
    cfg.gridDim = dim3(cl * ncl);
    if (n_clusters_out)
        *n_clusters_out = ncl;
    CK(cudaLaunchKernelEx(&cfg, kern, src, dprm, u0, ua, cp, which, started));
    ctx->tm.kernel_launches++;
    return NS_OK;
}

static float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

/* ---- call-wide pool of cluster-path units --------------------------------------------------------- */
static int ns_pool_begin(ns_ctx *ctx, int n)
{
    HeavyPool &pl = ctx->pool;
    pl.n_used = pl.n_launch = 0;
    pl.next_stream = 0;
    for (int i = 0; i < NS_POOL_STREAMS; i++)
        pl.used[i] = false;
    if (!pl.enabled)
        return NS_OK;
    /* ~80 B per sample per pooled unit (count rows + result planes): 8192 units up to n = 400, fewer beyond */
    int cap = (int)std::min<double>(8192.0, 256.0e6 / (80.0 * n));
    const char *ec = getenv("NS_POOL_CAP");
    if (ec && atoi(ec) > 0)
        cap = atoi(ec);
    if (cap < 16)
        cap = 16;
    if (pl.cap == cap && pl.n == n)
        return NS_OK;
    const size_t m = (size_t)cap * n;
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 5 && e == cudaSuccess; k++)
        e = pl.rows[k].reserve(m);
    if (e || (e = pl.is_indel.reserve(cap)) || (e = pl.flags.reserve(cap)) || (e = pl.iters.reserve(cap)) ||
        (e = pl.cycles.reserve(cap)) || (e = pl.sigma.reserve(cap)) || (e = pl.slope.reserve(cap)) ||
        (e = pl.max_gq.reserve(cap)) || (e = pl.unit.reserve(cap)) || (e = pl.lightpos.reserve(cap)) ||
        (e = pl.ident.reserve(cap)) || (e = pl.emitflag.reserve(cap)) || (e = pl.cnt.reserve(NS_POOL_MAX_LAUNCH + 1)) ||
        (e = pl.planes.reserve(cap, n))) {
        ctx->err = std::string("device allocation (cluster-path pool) failed: ") + cudaGetErrorString(e);
        return NS_ERR_NOMEM;
    }
    std::vector<int> id((size_t)cap);
    for (int i = 0; i < cap; i++)
        id[i] = i;
    CK(cudaMemcpyAsync(pl.ident.p, id.data(), (size_t)cap * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    pl.cap = cap;
    pl.n = n;
    return NS_OK;
}

static ns_pool_dev ns_pool_view(HeavyPool &pl)
{
    ns_pool_dev v = {pl.rows[0].p, pl.rows[1].p, pl.rows[2].p, pl.rows[3].p, pl.rows[4].p,
                     pl.is_indel.p, pl.flags.p, pl.unit.p, pl.lightpos.p};
    return v;
}

static ns_unit_src ns_pool_src(HeavyPool &pl, int n, int plain)
{
    ns_unit_src s;
    memset(&s, 0, sizeof(s));
    s.mode = 1;
    s.n = n;
    s.dp = pl.rows[0].p, s.vp = pl.rows[1].p, s.vm = pl.rows[2].p, s.rp = pl.rows[3].p, s.rm = pl.rows[4].p;
    s.is_indel = pl.is_indel.p;
    s.plain = plain;
    return s;
}

/* unit arrays of the pool: "unit" = pool slot; the lists are slices of the identity */
static ns_unit_arrays ns_pool_ua(HeavyPool &pl, int first)
{
    ns_unit_arrays ua = {pl.flags.p, pl.iters.p, pl.cycles.p, pl.sigma.p, pl.slope.p, pl.max_gq.p, pl.emitflag.p,
                         pl.ident.p, pl.emitflag.p, pl.emitflag.p, pl.ident.p, pl.ident.p, pl.ident.p, pl.ident.p + first};
    return ua;
}

/* can `count` more units of this chunk go to the pool? */
static bool ns_pool_room(ns_ctx *ctx, int count)
{
    HeavyPool &pl = ctx->pool;
    int max_per_chunk = 1024; /* beyond this the chunk is a deep panel: the in-chunk throughput regime is the better fit */
    const char *em = getenv("NS_POOL_MAX_PER_CHUNK");
    if (em && atoi(em) > 0)
        max_per_chunk = atoi(em);
    return pl.enabled && pl.cap > 0 && count > 0 && count <= max_per_chunk && pl.n_used + count <= pl.cap &&
           pl.n_launch < NS_POOL_MAX_LAUNCH;
}

/* moves `count` units of the current chunk (listed in `list`) into the pool and starts the cluster kernel for them on
 * one of the pool's streams; wait_resident: let the clusters take their SMs before the caller launches k_fit */
static int ns_pool_submit(ns_ctx *ctx, const ns_params &dprm, const ns_unit_src &src, int64_t u0_src, int64_t u_off,
                          const int *list, int count, const ns_unit_arrays &ua, bool wait_resident)
{
    HeavyPool &pl = ctx->pool;
    const int n = src.n;
    cudaStream_t st = ctx->stream;
    const int base = pl.n_used, k = pl.n_launch;
    {
        const int wpb = 8;
        int blocks = (count + wpb - 1) / wpb;
        if (blocks > ctx->sm_count * 4)
            blocks = ctx->sm_count * 4;
        k_pool_gather<<<blocks, wpb * 32, 0, st>>>(src, u0_src, u_off, list, count, ua, ns_pool_view(pl), base, pl.cnt.p + k);
        CK(cudaGetLastError());
        ctx->tm.kernel_launches++;
    }
    CK(cudaEventRecord(pl.ev_ready, st));
    const int si = pl.next_stream;
    pl.next_stream = (pl.next_stream + 1) % NS_POOL_STREAMS;
    cudaStream_t hs = pl.streams[si];
    CK(cudaStreamWaitEvent(hs, pl.ev_ready, 0));
    if (!pl.used[si]) {
        CK(cudaEventRecord(pl.ev_t0[si], hs));
        pl.used[si] = true;
    }
    int ncl = 0, max_cl = 0;
    const char *emc = getenv("NS_POOL_MAX_CLUSTERS");
    if (emc && atoi(emc) > 0)
        max_cl = atoi(emc);
    /* Earlier pool launches still running: their clusters hold SMs, the new ones may or may not find room at once.
     * Streaming shapes (a chunk takes less time than one cluster-path unit) must not stall on that, but where a chunk
     * and a unit take about as long (n = 400) skipping the wait lets the warp kernel take every SM first and pushes the
     * unit behind it, chunk after chunk (C4 end to end: 190 ms of tail): wait, but only briefly. */
    double wait_cap = 0.005;
    if (wait_resident)
        for (int i = 0; i < NS_POOL_STREAMS; i++)
            if (pl.used[i] && pl.n_launch > 0 && cudaEventQuery(pl.ev_t1[i]) == cudaErrorNotReady)
                wait_cap = 0.0005;
    cudaGetLastError();
    int *started = wait_resident ? ctx->d_started : nullptr;
    if (wait_resident)
        *(volatile int *)ctx->h_started = 0;
    int rc = ns_launch_heavy(ctx, hs, n, count, ns_pool_src(pl, n, src.plain), dprm, 0, ns_pool_ua(pl, base), pl.cnt.p + k, 0,
                             started, max_cl, &ncl);
    if (rc)
        return rc;
    CK(cudaEventRecord(pl.ev_t1[si], hs));
    if (wait_resident) {
        /* a cluster needs whole SMs and the persistent warp kernel would never give them back: bounded host-side
         * wait on a mapped flag until the clusters are resident (not a kernel waiting on a kernel) */
        const auto t0 = std::chrono::steady_clock::now();
        while (*(volatile int *)ctx->h_started < ncl) {
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > wait_cap)
                break;
        }
    }
    pl.n_used += count;
    pl.n_launch++;
    return NS_OK;
}

/* end of the call: join the pool's streams, run k_post for the pooled units, bring their results to the host and
 * merge them into the per-unit arrays and the emitted rows (which stay in unit order) */
static int ns_pool_finish(ns_ctx *ctx, const ns_params &dprm, const double *qtab, int n, int plain, ns_out *out,
                          RowSink *sink, int64_t *emit_off, int64_t n_units_total)
{
    HeavyPool &pl = ctx->pool;
    if (pl.n_used == 0)
        return NS_OK;
    cudaStream_t st = ctx->stream;
    const int M = pl.n_used;
    CK(cudaEventRecord(pl.ev_tail0, st));
    for (int i = 0; i < NS_POOL_STREAMS; i++)
        if (pl.used[i]) {
            CK(cudaEventRecord(pl.ev_done[i], pl.streams[i]));
            CK(cudaStreamWaitEvent(st, pl.ev_done[i], 0));
        }
    CK(cudaEventRecord(pl.ev_tail1, st));
    /* k_post over the pool slots */
    ns_counters *pc = pl.cnt.p + NS_POOL_MAX_LAUNCH;
    CK(cudaMemsetAsync(pc, 0, sizeof(ns_counters), st));
    pl.h_count[0] = M;
    CK(cudaMemcpyAsync(&pc->n_post, &pl.h_count[0], sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(pl.emitflag.p, 0, (size_t)M * sizeof(int), st));
    CK(cudaEventRecord(pl.ev_post0, st));
    {
        int wpb = 4;
        size_t per_warp = (3 * (size_t)n + 8) * sizeof(double);
        while (wpb > 1 && wpb * per_warp > 200 * 1024)
            wpb >>= 1;
        size_t sm = NS_RTAB * sizeof(double) + wpb * per_warp;
        CK(ns_need_smem(ctx, 3, k_post, sm));
        int blocks = (M + wpb - 1) / wpb;
        if (blocks > ctx->sm_count * 4)
            blocks = ctx->sm_count * 4;
        k_post<<<blocks, wpb * 32, sm, st>>>(ns_pool_src(pl, n, plain), dprm, 0, ns_pool_ua(pl, 0), pl.planes.view(), qtab,
                                            NS_QTAB_LEN, pc);
        CK(cudaGetLastError());
        ctx->tm.kernel_launches++;
    }
    CK(cudaEventRecord(pl.ev_post1, st));
    /* per-unit results and the launches' work counters */
    const size_t per_unit = sizeof(int64_t) * 2 + sizeof(double) * 3 + sizeof(uint32_t) * 2 + sizeof(unsigned long long) + sizeof(int);
    const size_t need = (size_t)M * per_unit + (size_t)(NS_POOL_MAX_LAUNCH + 1) * sizeof(ns_counters) + 256;
    if (need > pl.h_units_cap) {
        if (pl.h_units)
            cudaFreeHost(pl.h_units);
        pl.h_units = nullptr;
        pl.h_units_cap = 0;
        if (cudaMallocHost((void **)&pl.h_units, need * 2) != cudaSuccess) {
            ctx->err = "page-locked allocation (pool staging) failed";
            return NS_ERR_NOMEM;
        }
        pl.h_units_cap = need * 2;
    }
    char *hp = pl.h_units;
    int64_t *h_unit = (int64_t *)hp;
    hp += (size_t)M * sizeof(int64_t);
    int64_t *h_lightpos = (int64_t *)hp;
    hp += (size_t)M * sizeof(int64_t);
    double *h_sigma = (double *)hp;
    hp += (size_t)M * sizeof(double);
    double *h_slope = (double *)hp;
    hp += (size_t)M * sizeof(double);
    double *h_maxgq = (double *)hp;
    hp += (size_t)M * sizeof(double);
    unsigned long long *h_cycles = (unsigned long long *)hp;
    hp += (size_t)M * sizeof(unsigned long long);
    ns_counters *h_cnt = (ns_counters *)hp;
    hp += (size_t)(NS_POOL_MAX_LAUNCH + 1) * sizeof(ns_counters);
    uint32_t *h_flags = (uint32_t *)hp;
    hp += (size_t)M * sizeof(uint32_t);
    uint32_t *h_iters = (uint32_t *)hp;
    hp += (size_t)M * sizeof(uint32_t);
    int *h_emit = (int *)hp;
#define PD2H(dst, srcp, count, type) CK(cudaMemcpyAsync((dst), (srcp), (size_t)(count) * sizeof(type), cudaMemcpyDeviceToHost, st))
    PD2H(h_unit, pl.unit.p, M, int64_t);
    PD2H(h_lightpos, pl.lightpos.p, M, int64_t);
    PD2H(h_sigma, pl.sigma.p, M, double);
    PD2H(h_slope, pl.slope.p, M, double);
    PD2H(h_maxgq, pl.max_gq.p, M, double);
    PD2H(h_cycles, pl.cycles.p, M, unsigned long long);
    PD2H(h_cnt, pl.cnt.p, pl.n_launch, ns_counters);
    PD2H(h_flags, pl.flags.p, M, uint32_t);
    PD2H(h_iters, pl.iters.p, M, uint32_t);
    PD2H(h_emit, pl.emitflag.p, M, int);
    if (sink->want_rows) {
        if (!pl.h_rows.reserve(M, n, 0)) {
            ctx->err = "page-locked allocation (pool rows) failed";
            return NS_ERR_NOMEM;
        }
        const size_t m = (size_t)M * n;
        ns_planes pk = pl.planes.view();
        const ns_row_planes &h = pl.h_rows.p;
        if (sink->p.pvalue)
            PD2H(h.pvalue, pk.pvalue, m, double);
        if (sink->p.qvalue)
            PD2H(h.qvalue, pk.qvalue, m, double);
        if (sink->p.gq)
            PD2H(h.gq, pk.gq, m, double);
        if (sink->p.qval_minaf)
            PD2H(h.qval_minaf, pk.qval_minaf, m, double);
        if (sink->p.sor)
            PD2H(h.sor, pk.sor, m, double);
        if (sink->p.rvsb)
            PD2H(h.rvsb, pk.rvsb, m, double);
        if (sink->p.fs)
            PD2H(h.fs, pk.fs, m, double);
        if (sink->p.gt)
            PD2H(h.gt, pk.gt, m, uint8_t);
        if (sink->p.status)
            PD2H(h.status, pk.status, m, uint8_t);
        if (sink->p.pooled)
            PD2H(h.pooled, pk.pooled, (size_t)M * 8, double);
    }
#undef PD2H
    CK(cudaStreamSynchronize(st));
    for (int i = 0; i < NS_POOL_STREAMS; i++)
        if (pl.used[i])
            ctx->tm.heavy_pre_ms += ev_ms(pl.ev_t0[i], pl.ev_t1[i]);
    ctx->tm.pool_tail_ms += ev_ms(pl.ev_tail0, pl.ev_tail1);
    ctx->tm.post_ms += ev_ms(pl.ev_post0, pl.ev_post1);
    ctx->tm.pool_units += M;
    for (int k = 0; k < pl.n_launch; k++) {
        out->n_seed += h_cnt[k].n_seed;
        out->n_lattice += h_cnt[k].n_lattice;
        out->n_quad += h_cnt[k].n_quad;
    }
    /* per-unit arrays */
    for (int i = 0; i < M; i++) {
        const int64_t u = h_unit[i];
        if (out->sigma)
            out->sigma[u] = h_sigma[i];
        if (out->slope)
            out->slope[u] = h_slope[i];
        if (out->max_gq)
            out->max_gq[u] = h_maxgq[i];
        if (out->flags)
            out->flags[u] = h_flags[i] & ~NS_F_POOLED;
        if (out->iters)
            out->iters[u] = h_iters[i];
        if (out->unit_cycles)
            out->unit_cycles[u] = h_cycles[i];
    }
    if (sink->want_rows)
        for (int i = 0; i < M; i++) { /* staging row i = pool slot i */
            pl.h_rows.p.row_sigma[i] = h_sigma[i];
            pl.h_rows.p.row_slope[i] = h_slope[i];
            pl.h_rows.p.row_max_gq[i] = h_maxgq[i];
            pl.h_rows.p.row_flags[i] = h_flags[i] & ~NS_F_POOLED;
        }
    /* emitted rows: the pooled units that emit, in unit order, each with the number of rows that precede it */
    std::vector<ns_late_row> late;
    for (int i = 0; i < M; i++)
        if (h_emit[i])
            late.push_back(ns_late_row{h_unit[i], h_lightpos[i], (int64_t)i});
    std::sort(late.begin(), late.end(), [](const ns_late_row &a, const ns_late_row &b) { return a.unit < b.unit; });
    const int64_t E1 = *emit_off, E2 = (int64_t)late.size();
    if (E2 > 0) {
        if (sink->grow && E1 + E2 > sink->cap) {
            if (!sink->grow->reserve(E1 + E2, n, E1)) {
                ctx->err = "page-locked allocation (emitted rows) failed";
                return NS_ERR_NOMEM;
            }
            sink->p = sink->grow->p;
            sink->cap = sink->grow->cap;
        }
        if (E1 + E2 <= sink->cap || !sink->want_rows) {
            ns_row_planes dst = sink->p;
            if (!sink->want_rows)
                memset(&dst, 0, sizeof(dst));
            ns_merge_pool_rows(dst, E1, pl.h_rows.p, late.data(), E2, n, out->emit_index, n_units_total);
        }
        *emit_off = E1 + E2;
    }
    pl.n_used = pl.n_launch = 0;
    for (int i = 0; i < NS_POOL_STREAMS; i++)
        pl.used[i] = false;
    return NS_OK;
}

/* one chunk of units whose inputs are on the device; results to host at unit offset u_off */
static int ns_run_chunk(ns_ctx *ctx, const ns_params &dprm, const double *qtab, const ns_unit_src &src, int64_t u0_src,
                        int64_t u_off, int U, ns_out *out, RowSink *sink, int64_t *emit_off)
{
    const int n = src.n;
    cudaStream_t st = ctx->stream;
    HeavyPool &pl = ctx->pool;
    cudaError_t ce = ns_reserve_units(ctx, (size_t)U);
    if (ce != cudaSuccess) {
        ctx->err = std::string("device allocation (unit arrays) failed: ") + cudaGetErrorString(ce);
        return NS_ERR_NOMEM;
    }
    ns_unit_arrays ua = {ctx->flags.p,    ctx->iters.p,    ctx->cycles.p,   ctx->sigma.p,   ctx->slope.p,    ctx->max_gq.p,  ctx->needrow.p,
                         ctx->row.p,      ctx->emitflag.p, ctx->emitidx.p, ctx->fit_list.p, ctx->post_list.p, ctx->heavy_list.p, ctx->pre_list.p};
    CK(cudaMemsetAsync(ctx->counters.p, 0, sizeof(ns_counters), st));
    /* gate */
    CK(cudaEventRecord(ctx->ev[0], st));
    {
        /* nobody reads the per-unit words of gated units when only rows are wanted and gated units cannot be emitted */
        const int lean = !out->sigma && !out->slope && !out->max_gq && !out->iters && !out->unit_cycles &&
                         dprm.emit_mode != NS_EMIT_ALL && !dprm.output_all_SNVs;
        /* bulk-asynchronous staging for SNV planes (see ns_gate_staged): 16 or 8 positions per stage, 2-4 stages,
         * at most ~96 KB per CTA so that two CTAs share an SM; needs 16-byte aligned planes */
        int stage_k = 0, stages = 0, q_begin = 0;
        const char *eg = getenv("NS_GATE_STAGED");
        /* measured on B200 (profiles/README.md, round 2): 0.53 of the HBM copy rate at n = 100 against 0.74 for the
         * register-load form - the kernel is bound by instruction issue and the register file already holds as many
         * bytes in flight as the stages do - so the staged form is opt-in (NS_GATE_STAGED=1) */
        if (src.mode == 0 && !dprm.extra_rob && (u0_src % 3) == 0 && eg && atoi(eg) == 1) {
            const size_t pos_bytes = (size_t)36 * n;
            const uintptr_t addr = (uintptr_t)(src.counts + (u0_src / 3) * 9 * (int64_t)n);
            stage_k = 16 * pos_bytes <= 32 * 1024 ? 2 : (8 * pos_bytes <= 48 * 1024 ? 1 : 0);
            if ((addr & 15) != 0 || U / 3 < 64 * NS_WPB_GATE)
                stage_k = 0;
            if (stage_k) {
                const size_t sb = (size_t)stage_k * NS_WPB_GATE * pos_bytes;
                stages = (int)((100 * 1024 - NS_GATE_SMEM_HDR) / sb);
                stages = stages < 2 ? 2 : (stages > 4 ? 4 : stages);
                const char *es = getenv("NS_GATE_STAGES");
                if (es && atoi(es) >= 1 && atoi(es) <= 8)
                    stages = atoi(es);
                const size_t sms = NS_GATE_SMEM_HDR + (size_t)stages * sb;
                if (sms > 200 * 1024) {
                    stage_k = 0;
                } else {
                    const int PB = stage_k * NS_WPB_GATE;
                    const int n_full = (U / 3) / PB;
                    int per_sm = (int)((220 * 1024) / (sms + 1024));
                    per_sm = per_sm > 2 ? 2 : (per_sm < 1 ? 1 : per_sm);
                    int blocks = ctx->sm_count * per_sm;
                    if (blocks > n_full)
                        blocks = n_full;
                    CK(ns_need_smem(ctx, 6, k_gate_staged, sms));
                    k_gate_staged<<<blocks, NS_WPB_GATE * 32, sms, st>>>(src, dprm, u0_src, U, ua, ctx->counters.p, stage_k, stages,
                                                                        n_full, lean);
                    CK(cudaGetLastError());
                    ctx->tm.kernel_launches++;
                    q_begin = n_full * PB;
                }
            }
        }
        if (3 * (int64_t)q_begin < U) { /* everything (register-load form), or the positions behind the last whole block */
            const int rest = U - 3 * q_begin;
            size_t sm = (size_t)NS_WPB_GATE * n * sizeof(double);
            int blocks = (rest + NS_WPB_GATE - 1) / NS_WPB_GATE;
            int maxb = ctx->sm_count * 8;
            if (blocks > maxb)
                blocks = maxb;
            const int form = ns_gate_form(src.mode, dprm.extra_rob, u0_src, n, dprm.min_coverage,
                                          src.mode == 0 ? src.counts + (u0_src / 3) * 9 * (int64_t)n : nullptr);
            if (form == 4) {
                k_gate_v4<<<blocks, NS_WPB_GATE * 32, 0, st>>>(src, dprm, u0_src, U, ua, ctx->counters.p, q_begin, lean);
            } else if (form == 2) {
                k_gate_v2<<<blocks, NS_WPB_GATE * 32, 0, st>>>(src, dprm, u0_src, U, ua, ctx->counters.p, q_begin, lean);
            } else {
                CK(ns_need_smem(ctx, 2, k_gate, sm));
                k_gate<<<blocks, NS_WPB_GATE * 32, sm, st>>>(src, dprm, u0_src, U, ua, ctx->counters.p, q_begin, lean);
            }
            CK(cudaGetLastError());
        }
    }
    CK(cudaEventRecord(ctx->ev[1], st));
    /* rows for the units that need per-sample planes */
    {
        size_t tmp = ctx->cub_tmp.cap;
        CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, ctx->needrow.p, ctx->row.p, U, st));
        k_scan_total<<<1, 32, 0, st>>>(ctx->needrow.p, ctx->row.p, U, &ctx->counters.p->n_rows);
    }
    CK(cudaMemcpyAsync(ctx->h_counters, ctx->counters.p, sizeof(ns_counters), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int n_rows = ctx->h_counters->n_rows;
    const int n_fit = ctx->h_counters->n_light + ctx->h_counters->n_pre;
    ce = ctx->rows.reserve((size_t)(n_rows > 0 ? n_rows : 1), n);
    if (ce != cudaSuccess) {
        ctx->err = std::string("device allocation (result planes) failed: ") + cudaGetErrorString(ce);
        return NS_ERR_NOMEM;
    }
    /* fit.  Units predicted for the cluster kernel: into the call-wide pool when it has room (they then run next to
     * this and later chunks and are joined at the end of the call), else - deep panels, where every unit is one - on
     * the side stream within this chunk.  Units the warp kernel hands over go the same two ways. */
    const int n_pre = ctx->h_counters->n_pre, n_light = ctx->h_counters->n_light;
    CK(cudaEventRecord(ctx->ev[2], st));
    bool heavy_forked = false;
    int n_pooled = 0;
    if (n_pre > 0 && ns_pool_room(ctx, n_pre)) {
        int rc = ns_pool_submit(ctx, dprm, src, u0_src, u_off, ctx->pre_list.p, n_pre, ua, n_light > 0);
        if (rc)
            return rc;
        n_pooled += n_pre;
    } else if (n_pre > 0) {
        CK(cudaEventRecord(ctx->ev_fork, st));
        CK(cudaStreamWaitEvent(ctx->heavy_stream, ctx->ev_fork, 0));
        CK(cudaEventRecord(ctx->ev_h0, ctx->heavy_stream));
        int ncl = 0;
        *(volatile int *)ctx->h_started = 0;
        int rc = ns_launch_heavy(ctx, ctx->heavy_stream, n, n_pre, src, dprm, u0_src, ua, ctx->counters.p, 0, ctx->d_started, 0, &ncl);
        if (rc)
            return rc;
        if (n_light > 0) {
            const auto t0 = std::chrono::steady_clock::now();
            while (*(volatile int *)ctx->h_started < ncl) {
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 0.02)
                    break;
            }
        }
        CK(cudaEventRecord(ctx->ev_h1, ctx->heavy_stream));
        CK(cudaEventRecord(ctx->ev_join, ctx->heavy_stream));
        heavy_forked = true;
    }
    if (n_light > 0) {
        const int wpb = 4;
        size_t per_warp = ns_work_doubles(n) * sizeof(double);
        size_t sm = (NS_RTAB + 256) * sizeof(double) + wpb * per_warp; /* 1/j table, log table, the warps' slices */
        if (sm > 227 * 1024) {
            ctx->err = "too many samples for the on-chip working set (n > ~1100)";
            return NS_ERR_ARG;
        }
        /* (four samples per lane in flight at 128 registers / 16 warps per SM was measured and dropped: C2 441 ms per
         * step against 353 ms for two samples per lane at 96 registers / 20 warps; profiles/README.md, round 2) */
        CK(ns_need_smem(ctx, 4, k_fit, sm));
        int bps = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_fit, wpb * 32, sm));
        if (bps < 1)
            bps = 1;
        int blocks = ctx->sm_count * bps;
        int need = (n_light + wpb - 1) / wpb;
        if (blocks > need)
            blocks = need;
        k_fit<<<blocks, wpb * 32, sm, st>>>(src, dprm, u0_src, ua, ctx->counters.p);
        CK(cudaGetLastError());
        ctx->tm.kernel_launches++;
    }
    CK(cudaEventRecord(ctx->ev[7], st));
    if (heavy_forked)
        CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    int n_def = 0;
    if (n_light > 0) {
        /* units the warp kernel handed over (rare): read their count, act only if there are any */
        CK(cudaMemcpyAsync(&ctx->h_counters->n_heavy, &ctx->counters.p->n_heavy, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        n_def = ctx->h_counters->n_heavy;
        if (n_def > 0 && ns_pool_room(ctx, n_def)) {
            int rc = ns_pool_submit(ctx, dprm, src, u0_src, u_off, ctx->heavy_list.p, n_def, ua, false);
            if (rc)
                return rc;
            n_pooled += n_def;
        } else if (n_def > 0) {
            int rc = ns_launch_heavy(ctx, st, n, n_def, src, dprm, u0_src, ua, ctx->counters.p, 1, nullptr, 0, nullptr);
            if (rc)
                return rc;
        }
    }
    CK(cudaEventRecord(ctx->ev[3], st));
    /* post */
    const int n_post = ctx->h_counters->n_post;
    if (n_post > 0) {
        int wpb = 4;
        size_t per_warp = (3 * (size_t)n + 8) * sizeof(double);
        while (wpb > 1 && wpb * per_warp > 200 * 1024)
            wpb >>= 1;
        size_t sm = NS_RTAB * sizeof(double) + wpb * per_warp;
        CK(ns_need_smem(ctx, 3, k_post, sm));
        int bps = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_post, wpb * 32, sm));
        if (bps < 1)
            bps = 1;
        int blocks = ctx->sm_count * bps;
        int need = (n_post + wpb - 1) / wpb;
        if (blocks > need)
            blocks = need;
        k_post<<<blocks, wpb * 32, sm, st>>>(src, dprm, u0_src, ua, ctx->rows.view(), qtab, NS_QTAB_LEN,
                                            ctx->counters.p);
        CK(cudaGetLastError());
        ctx->tm.kernel_launches++;
    }
    CK(cudaEventRecord(ctx->ev[4], st));
    /* emitted rows */
    {
        size_t tmp = ctx->cub_tmp.cap;
        CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, ctx->emitflag.p, ctx->emitidx.p, U, st));
        k_scan_total<<<1, 32, 0, st>>>(ctx->emitflag.p, ctx->emitidx.p, U, &ctx->counters.p->n_emit);
    }
    CK(cudaMemcpyAsync(ctx->h_counters, ctx->counters.p, sizeof(ns_counters), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int n_emit = ctx->h_counters->n_emit;
    if (sink->grow && *emit_off + n_emit > sink->cap) {
        if (!sink->grow->reserve(*emit_off + n_emit, n, *emit_off)) {
            ctx->err = "page-locked allocation (emitted rows) failed";
            return NS_ERR_NOMEM;
        }
        sink->p = sink->grow->p;
        sink->cap = sink->grow->cap;
    }
    int64_t room = sink->cap - *emit_off;
    if (room < 0)
        room = 0;
    const int n_copy = (int)((int64_t)n_emit < room ? n_emit : room);
    {
        /* buffers only ever grow; the row capacity depends on the current n */
        size_t cap = (size_t)n_emit + 1024;
        ce = ctx->packed.reserve(cap, n);
        if (ce == cudaSuccess)
            ce = ctx->emit_unit.reserve(cap);
        if (ce != cudaSuccess) {
            ctx->err = std::string("device allocation (packed planes) failed: ") + cudaGetErrorString(ce);
            return NS_ERR_NOMEM;
        }
        ctx->packed_cap = cap;
    }
    {
        int blocks = (U + 255) / 256;
        if (blocks > ctx->sm_count * 8)
            blocks = ctx->sm_count * 8;
        k_pack<<<blocks, 256, 0, st>>>(n, U, u_off, *emit_off, (int)ctx->packed_cap, ua, ctx->rows.view(),
                                            ctx->packed.view(), ctx->emit_unit.p, ctx->emit_index_out.p, pl.lightpos.p,
                                            ctx->counters.p);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ctx->ev[5], st));
    /* D2H */
#define D2H(dst, srcp, count, type)                                                                \
    if (dst)                                                                                       \
    CK(cudaMemcpyAsync((dst), (srcp), (size_t)(count) * sizeof(type), cudaMemcpyDeviceToHost, st))
    D2H(out->sigma ? out->sigma + u_off : nullptr, ctx->sigma.p, U, double);
    D2H(out->slope ? out->slope + u_off : nullptr, ctx->slope.p, U, double);
    D2H(out->max_gq ? out->max_gq + u_off : nullptr, ctx->max_gq.p, U, double);
    D2H(out->flags ? out->flags + u_off : nullptr, ctx->flags.p, U, uint32_t);
    D2H(out->iters ? out->iters + u_off : nullptr, ctx->iters.p, U, uint32_t);
    D2H(out->unit_cycles ? out->unit_cycles + u_off : nullptr, ctx->cycles.p, U, uint64_t);
    D2H(out->emit_index ? out->emit_index + u_off : nullptr, ctx->emit_index_out.p, U, int32_t);
    D2H(&ctx->h_counters->n_skipped, &ctx->counters.p->n_skipped, 1, int);
    if (sink->want_rows && n_copy > 0) {
        const size_t eo = (size_t)*emit_off;
        const size_t m = (size_t)n_copy * n;
        ns_planes pk = ctx->packed.view();
        const ns_row_planes &sp = sink->p;
        D2H(sp.pvalue ? sp.pvalue + eo * n : nullptr, pk.pvalue, m, double);
        D2H(sp.qvalue ? sp.qvalue + eo * n : nullptr, pk.qvalue, m, double);
        D2H(sp.gq ? sp.gq + eo * n : nullptr, pk.gq, m, double);
        D2H(sp.qval_minaf ? sp.qval_minaf + eo * n : nullptr, pk.qval_minaf, m, double);
        D2H(sp.sor ? sp.sor + eo * n : nullptr, pk.sor, m, double);
        D2H(sp.rvsb ? sp.rvsb + eo * n : nullptr, pk.rvsb, m, double);
        D2H(sp.fs ? sp.fs + eo * n : nullptr, pk.fs, m, double);
        D2H(sp.gt ? sp.gt + eo * n : nullptr, pk.gt, m, uint8_t);
        D2H(sp.status ? sp.status + eo * n : nullptr, pk.status, m, uint8_t);
        D2H(sp.pooled ? sp.pooled + eo * 8 : nullptr, pk.pooled, (size_t)n_copy * 8, double);
        D2H(sp.emit_unit ? sp.emit_unit + eo : nullptr, ctx->emit_unit.p, n_copy, int64_t);
        D2H(sp.row_sigma ? sp.row_sigma + eo : nullptr, pk.rsigma, n_copy, double);
        D2H(sp.row_slope ? sp.row_slope + eo : nullptr, pk.rslope, n_copy, double);
        D2H(sp.row_max_gq ? sp.row_max_gq + eo : nullptr, pk.rmaxgq, n_copy, double);
        D2H(sp.row_flags ? sp.row_flags + eo : nullptr, pk.rflags, n_copy, uint32_t);
    }
#undef D2H
    CK(cudaEventRecord(ctx->ev[6], st));
    CK(cudaStreamSynchronize(st));
    ctx->tm.gate_ms += ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->tm.fit_ms += ev_ms(ctx->ev[2], ctx->ev[7]);
    ctx->tm.heavy_ms += ev_ms(ctx->ev[7], ctx->ev[3]);
    if (heavy_forked)
        ctx->tm.heavy_pre_ms += ev_ms(ctx->ev_h0, ctx->ev_h1);
    ctx->tm.fit_phase_ms += ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->tm.n_heavy += n_def + n_pre;
    ctx->tm.n_deferred += n_def;
    ctx->tm.fast_seed += ctx->h_counters->fast_seed;
    ctx->tm.fast_lattice += ctx->h_counters->fast_lattice;
    ctx->tm.post_ms += ev_ms(ctx->ev[3], ctx->ev[4]);
    ctx->tm.pack_ms += ev_ms(ctx->ev[4], ctx->ev[5]);
    ctx->tm.d2h_ms += ev_ms(ctx->ev[5], ctx->ev[6]);
    ctx->tm.kernel_launches += 8; /* gate, 2 x (cub scan = init + scan kernels, total), pack */
    ctx->tm.chunks += 1;
    ctx->n_skipped_call += ctx->h_counters->n_skipped;
    *emit_off += n_emit;
    out->n_fitted += n_fit;
    out->n_seed += ctx->h_counters->n_seed;
    out->n_lattice += ctx->h_counters->n_lattice;
    out->n_quad += ctx->h_counters->n_quad;
    (void)n_pooled;
    return NS_OK;
}

int64_t ns_pick_chunk(ns_ctx *ctx, int n, int64_t total_units, bool device_resident)
{
    int64_t c;
    if (ctx->chunk_units > 0) {
        c = ctx->chunk_units;
    } else {
        /* Device-resident planes of small panels (WGS-like: almost every unit is gated, the chunk is a streaming pass with
         * a handful of cluster-path units): nothing to overlap with a copy, so chunks four times as large - the
         * cluster-path units of a chunk start when its gate has run, and fewer, larger chunks start them earlier (C5, 4 M
         * positions: 8 chunks 47 ms per step, 2 chunks 39.5 ms).  Not for
         * larger panels: at n = 100 a chunk's 25 such units already hold every SM for its first ~35 ms. */
        const bool big = device_resident && n <= 64;
        /* bound the per-chunk result planes (58 B per sample per row, twice with the packed copy): <= ~6 GB if every unit
         * were fitted (small panels: a fraction of a percent is) */
        c = (int64_t)((big ? 24000.0e6 : 6000.0e6) / (58.0 * n));
        /* ~1 GB of count planes per chunk, at least 262144 positions (the chunk is the unit of the H2D / compute
         * overlap: n = 100: 262144 positions, the measured end-to-end optimum; small panels get proportionally more
         * positions per chunk) */
        int64_t cap = 3 * (int64_t)((big ? 3.76e9 : 0.94e9) / (36.0 * n));
        if (cap < 3 * 262144)
            cap = 3 * 262144;
        if (cap > 3 * (big ? 4194304 : 2097152))
            cap = 3 * (big ? 4194304 : 2097152); /* per-unit arrays: ~70 B per unit */
        if (c > cap)
            c = cap;
        if (c < 3 * 256)
            c = 3 * 256;
    }
    c -= c % 3;
    if (c < 3)
        c = 3;
    if (c >= total_units)
        return total_units;
    if (ctx->chunk_units <= 0) {
        /* equal chunks: a short last chunk has too little fit work to hide the critical path of its cluster-path units
         * (the call then waits tens of ms for them); whole groups of 64 positions keep every chunk start 16-byte aligned */
        const int64_t n_chunks = (total_units + c - 1) / c;
        int64_t per = (total_units / 3 + n_chunks - 1) / n_chunks;
        per = (per + 63) / 64 * 64;
        if (3 * per < c)
            c = 3 * per;
    }
    return c;
}

void ns_reset_out(ns_ctx *ctx, ns_out *out)
{
    out->n_emitted = 0;
    out->n_fitted = out->n_gated = out->n_skipped = 0;
    out->n_seed = out->n_lattice = out->n_quad = 0;
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    ctx->n_skipped_call = 0;
}

static void ns_finish_out(ns_ctx *ctx, ns_out *out, int64_t total_units, int64_t emitted)
{
    out->n_emitted = emitted;
    out->n_skipped = ctx->n_skipped_call; /* counted on the device (k_pack), not by a host pass over the flags */
    out->n_gated = total_units - out->n_fitted - out->n_skipped;
}

static int ns_check_overflow(ns_ctx *ctx, ns_out *out, int64_t emitted)
{
    if (emitted > out->emit_cap && (out->pvalue || out->gq || out->emit_unit)) {
        ctx->err = "more emitted units than out->emit_cap";
        return NS_ERR_EMIT_OVERFLOW;
    }
    return NS_OK;
}

/* 16-bit planes widened on the device: WGS-like shapes are bound by the host link (1.8 KB per position at n = 50),
 * and exome / genome depths fit 16 bits, so the narrow wire format halves the transfer; the kernels keep int32 rows */
__global__ void k_widen16(const uint16_t *__restrict__ src, int32_t *__restrict__ dst, size_t n)
{
    const size_t n8 = n / 8;
    const uint4 *s8 = reinterpret_cast<const uint4 *>(src);
    int4 *d4 = reinterpret_cast<int4 *>(dst);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(s8 + i);
        d4[2 * i] = make_int4((int)(v.x & 0xffffu), (int)(v.x >> 16), (int)(v.y & 0xffffu), (int)(v.y >> 16));
        d4[2 * i + 1] = make_int4((int)(v.z & 0xffffu), (int)(v.z >> 16), (int)(v.w & 0xffffu), (int)(v.w >> 16));
    }
    for (size_t i = 8 * n8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = (int32_t)src[i];
}

static int ns_snv_device_body(ns_ctx *ctx, const ns_params *prm, int n, int64_t P, const uint8_t *d_ref,
                              const int32_t *d_counts, ns_out *out)
{
    CK(cudaSetDevice(ctx->device));
    ns_reset_out(ctx, out);
    ns_params dprm;
    const double *qtab = nullptr;
    int rc = ns_prepare_params(ctx, prm, n, &dprm, &qtab);
    if (rc)
        return rc;
    if ((rc = ns_pool_begin(ctx, n)))
        return rc;
    cudaEvent_t t0 = ctx->ev[8], t1 = ctx->ev[9];
    CK(cudaEventRecord(t0, ctx->stream));
    const int64_t U = 3 * P;
    const int64_t chunk = U > 0 ? ns_pick_chunk(ctx, n, U, true) : 0;
    int64_t emit_off = 0;
    RowSink sink = ns_sink_from_out(out);
    ns_unit_src src;
    memset(&src, 0, sizeof(src));
    src.mode = 0;
    src.n = n;
    src.counts = d_counts;
    src.ref = d_ref;
    for (int64_t u0 = 0; u0 < U; u0 += chunk) {
        int Uc = (int)((U - u0 < chunk) ? (U - u0) : chunk);
        rc = ns_run_chunk(ctx, dprm, qtab, src, u0, u0, Uc, out, &sink, &emit_off);
        if (rc)
            return rc;
    }
    if ((rc = ns_pool_finish(ctx, dprm, qtab, n, 0, out, &sink, &emit_off, U)))
        return rc;
    CK(cudaEventRecord(t1, ctx->stream));
    CK(cudaEventSynchronize(t1));
    ctx->tm.total_ms = ev_ms(t0, t1);
    ns_finish_out(ctx, out, U, emit_off);
    return ns_check_overflow(ctx, out, emit_off);
}

/* SNV planes on the host.  Chunks of positions are pulled from `q` (one context: its own queue, every chunk; a
 * multi-GPU call: one queue shared by the host threads of all devices); H2D of the next chunk runs on
 * the copy stream under the kernels of the current one (two staging buffers). */
int ns_snv_host_queue(ns_ctx *ctx, const ns_params *prm, int n, int64_t P, const uint8_t *ref, const int32_t *counts,
                      const uint16_t *counts16, ns_out *out, RowSink *sink, ChunkQueue *q, int slot, int64_t *emitted)
{
    CK(cudaSetDevice(ctx->device));
    ns_params dprm;
    const double *qtab = nullptr;
    int rc = ns_prepare_params(ctx, prm, n, &dprm, &qtab);
    if (rc)
        return rc;
    if ((rc = ns_pool_begin(ctx, n)))
        return rc;
    cudaEvent_t t0 = ctx->ev[8], t1 = ctx->ev[9];
    CK(cudaEventRecord(t0, ctx->stream));
    const int64_t pchunk = q->pchunk;
    int64_t emit_off = 0;
    const size_t row_elems = (size_t)9 * n;
    auto stage = [&](int64_t ci, int buf) -> int {
        const int64_t p0 = ci * pchunk;
        int64_t pc = (P - p0 < pchunk) ? (P - p0) : pchunk;
        cudaError_t e;
        if ((e = ctx->in_counts[buf].reserve((size_t)pc * row_elems)) || (e = ctx->in_ref[buf].reserve((size_t)pc))) {
            ctx->err = std::string("device allocation (input staging) failed: ") + cudaGetErrorString(e);
            return NS_ERR_NOMEM;
        }
        if (counts16 && (e = ctx->in_counts16[buf].reserve((size_t)pc * row_elems))) {
            ctx->err = std::string("device allocation (input staging) failed: ") + cudaGetErrorString(e);
            return NS_ERR_NOMEM;
        }
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[buf], 0));
        if (counts16) {
            const size_t ne = (size_t)pc * row_elems;
            CK(cudaMemcpyAsync(ctx->in_counts16[buf].p, counts16 + (size_t)p0 * row_elems, ne * sizeof(uint16_t),
                               cudaMemcpyHostToDevice, ctx->copy_stream));
            k_widen16<<<ctx->sm_count * 4, 256, 0, ctx->copy_stream>>>(ctx->in_counts16[buf].p, ctx->in_counts[buf].p, ne);
            CK(cudaGetLastError());
            ctx->tm.kernel_launches++;
        } else
            CK(cudaMemcpyAsync(ctx->in_counts[buf].p, counts + (size_t)p0 * row_elems, (size_t)pc * row_elems * sizeof(int32_t),
                               cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaMemcpyAsync(ctx->in_ref[buf].p, ref + p0, (size_t)pc, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->ev_h2d[buf], ctx->copy_stream));
        return NS_OK;
    };
    int64_t taken = 0;
    auto take = [&]() -> int64_t {
        /* max_per_slot: the queue prefetches one chunk ahead, so with as many chunks as devices the threads that
         * start first would take two each and leave the last devices idle */
        if (q->failed.load() || (q->max_per_slot > 0 && taken >= q->max_per_slot))
            return -1;
        taken++;
        const int64_t ci = q->next.fetch_add(1);
        if (ci >= q->n_chunks)
            return -1;
        if (!q->owner.empty())
            q->owner[(size_t)ci] = slot;
        return ci;
    };
    /* the staging buffers start out free */
    CK(cudaEventRecord(ctx->ev_free[0], ctx->stream));
    CK(cudaEventRecord(ctx->ev_free[1], ctx->stream));
    int64_t cur = take();
    if (cur >= 0 && (rc = stage(cur, 0)))
        return rc;
    int64_t k = 0;
    while (cur >= 0) {
        const int buf = (int)(k & 1);
        const int64_t nxt = take();
        if (nxt >= 0 && (rc = stage(nxt, buf ^ 1)))
            return rc;
        const int64_t p0 = cur * pchunk;
        const int64_t pc = (P - p0 < pchunk) ? (P - p0) : pchunk;
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[buf], 0));
        ns_unit_src src;
        memset(&src, 0, sizeof(src));
        src.mode = 0;
        src.n = n;
        src.counts = ctx->in_counts[buf].p;
        src.ref = ctx->in_ref[buf].p;
        rc = ns_run_chunk(ctx, dprm, qtab, src, 0, 3 * p0, (int)(3 * pc), out, sink, &emit_off);
        if (rc)
            return rc;
        CK(cudaEventRecord(ctx->ev_free[buf], ctx->stream));
        cur = nxt;
        k++;
    }
    if ((rc = ns_pool_finish(ctx, dprm, qtab, n, 0, out, sink, &emit_off, 3 * P)))
        return rc;
    CK(cudaEventRecord(t1, ctx->stream));
    CK(cudaEventSynchronize(t1));
    ctx->tm.total_ms = ev_ms(t0, t1);
    *emitted = emit_off;
    return NS_OK;
}

static int ns_call_snv_host(ns_ctx *ctx, const ns_params *prm, int n, int64_t P, const uint8_t *ref, const int32_t *counts,
                            const uint16_t *counts16, ns_out *out)
{
    if (!ctx)
        return NS_ERR_ARG;
    if (!prm || !out || n < 1 || P < 0 || (P > 0 && (!ref || (!counts && !counts16)))) {
        ctx->err = "ns_call_snv: bad argument";
        return NS_ERR_ARG;
    }
    ns_reset_out(ctx, out);
    const int64_t U = 3 * P;
    ChunkQueue q;
    q.pchunk = U > 0 ? ns_pick_chunk(ctx, n, U) / 3 : 1;
    q.n_chunks = P > 0 ? (P + q.pchunk - 1) / q.pchunk : 0;
    RowSink sink = ns_sink_from_out(out);
    int64_t emitted = 0;
    int rc = ns_snv_host_queue(ctx, prm, n, P, ref, counts, counts16, out, &sink, &q, 0, &emitted);
    if (rc) {
        ns_drain(ctx);
        return rc;
    }
    ns_finish_out(ctx, out, U, emitted);
    return ns_check_overflow(ctx, out, emitted);
}

static int ns_call_units_body(ns_ctx *ctx, const ns_params *prm, int n, int64_t U, const int32_t *dp, const int32_t *vp,
                              const int32_t *vm, const int32_t *rp, const int32_t *rm, const uint8_t *is_indel, int plain,
                              ns_out *out)
{
    CK(cudaSetDevice(ctx->device));
    ns_reset_out(ctx, out);
    ns_params dprm;
    const double *qtab = nullptr;
    int rc = ns_prepare_params(ctx, prm, n, &dprm, &qtab);
    if (rc)
        return rc;
    if ((rc = ns_pool_begin(ctx, n)))
        return rc;
    cudaEvent_t t0 = ctx->ev[8], t1 = ctx->ev[9];
    CK(cudaEventRecord(t0, ctx->stream));
    const int64_t chunk = U > 0 ? ns_pick_chunk(ctx, n, U) : 0;
    int64_t emit_off = 0;
    RowSink sink = ns_sink_from_out(out);
    const int32_t *hrows[5] = {dp, vp, vm, rp, rm};
    for (int64_t u0 = 0; u0 < U; u0 += chunk) {
        const int Uc = (int)((U - u0 < chunk) ? (U - u0) : chunk);
        ns_unit_src src;
        memset(&src, 0, sizeof(src));
        src.mode = 1;
        src.n = n;
        src.plain = plain;
        const int32_t *drows[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        for (int k = 0; k < 5; k++) {
            if (!hrows[k])
                continue;
            cudaError_t e = ctx->in_rows[k].reserve((size_t)Uc * n);
            if (e != cudaSuccess) {
                ctx->err = std::string("device allocation (unit rows) failed: ") + cudaGetErrorString(e);
                return NS_ERR_NOMEM;
            }
            CK(cudaMemcpyAsync(ctx->in_rows[k].p, hrows[k] + (size_t)u0 * n, (size_t)Uc * n * sizeof(int32_t),
                               cudaMemcpyHostToDevice, ctx->stream));
            drows[k] = ctx->in_rows[k].p;
        }
        if (is_indel) {
            cudaError_t e = ctx->in_indel.reserve((size_t)Uc);
            if (e != cudaSuccess) {
                ctx->err = std::string("device allocation failed: ") + cudaGetErrorString(e);
                return NS_ERR_NOMEM;
            }
            CK(cudaMemcpyAsync(ctx->in_indel.p, is_indel + u0, (size_t)Uc, cudaMemcpyHostToDevice, ctx->stream));
            src.is_indel = ctx->in_indel.p;
        }
        src.dp = drows[0];
        src.vp = drows[1];
        src.vm = drows[2];
        src.rp = drows[3];
        src.rm = drows[4];
        rc = ns_run_chunk(ctx, dprm, qtab, src, 0, u0, Uc, out, &sink, &emit_off);
        if (rc)
            return rc;
    }
    if ((rc = ns_pool_finish(ctx, dprm, qtab, n, plain, out, &sink, &emit_off, U)))
        return rc;
    CK(cudaEventRecord(t1, ctx->stream));
    CK(cudaEventSynchronize(t1));
    ctx->tm.total_ms = ev_ms(t0, t1);
    ns_finish_out(ctx, out, U, emit_off);
    return ns_check_overflow(ctx, out, emit_off);
}

extern "C" {

int ns_call_snv_device(ns_ctx *ctx, const ns_params *prm, int n, int64_t P, const uint8_t *d_ref,
                       const int32_t *d_counts, ns_out *out)
{
    if (!ctx)
        return NS_ERR_ARG;
    if (!prm || !out || n < 1 || P < 0 || (P > 0 && (!d_ref || !d_counts))) {
        ctx->err = "ns_call_snv_device: bad argument";
        return NS_ERR_ARG;
    }
    int rc = ns_snv_device_body(ctx, prm, n, P, d_ref, d_counts, out);
    if (rc && rc != NS_ERR_EMIT_OVERFLOW)
        ns_drain(ctx);
    return rc;
}

int ns_call_snv(ns_ctx *ctx, const ns_params *prm, int n, int64_t P, const uint8_t *ref, const int32_t *counts,
                ns_out *out)
{
    return ns_call_snv_host(ctx, prm, n, P, ref, counts, nullptr, out);
}

int ns_call_snv_u16(ns_ctx *ctx, const ns_params *prm, int n, int64_t P, const uint8_t *ref, const uint16_t *counts,
                    ns_out *out)
{
    return ns_call_snv_host(ctx, prm, n, P, ref, nullptr, counts, out);
}

int ns_call_units(ns_ctx *ctx, const ns_params *prm, int n, int64_t U, const int32_t *dp, const int32_t *vp,
                  const int32_t *vm, const int32_t *rp, const int32_t *rm, const uint8_t *is_indel, int plain,
                  ns_out *out)
{
    if (!ctx)
        return NS_ERR_ARG;
    if (!prm || !out || n < 1 || U < 0 || (U > 0 && (!dp || !vp))) {
        ctx->err = "ns_call_units: bad argument";
        return NS_ERR_ARG;
    }
    int rc = ns_call_units_body(ctx, prm, n, U, dp, vp, vm, rp, rm, is_indel, plain, out);
    if (rc && rc != NS_ERR_EMIT_OVERFLOW)
        ns_drain(ctx);
    return rc;
}

int ns_glmrob_nb(ns_ctx *ctx, const ns_params *prm, int n, const int32_t *y, const int32_t *x, double *sigma,
                 double *slope, double *pvalues, double *qvalues, double *gq, int32_t *extra_rob, uint32_t *flags)
{
    if (!ctx)
        return NS_ERR_ARG;
    if (!prm || !y || !x || n < 1) {
        ctx->err = "ns_glmrob_nb: bad argument";
        return NS_ERR_ARG;
    }
    ns_params p = *prm;
    p.emit_mode = NS_EMIT_ALL;
    p.is_tn = 0;
    p.output_all_SNVs = 0;
    ns_out out;
    memset(&out, 0, sizeof(out));
    double sg = NAN, sl = NAN;
    uint32_t fl = 0;
    out.sigma = &sg;
    out.slope = &sl;
    out.flags = &fl;
    out.emit_cap = 1;
    out.pvalue = pvalues;
    out.qvalue = qvalues;
    out.gq = gq;
    int rc = ns_call_units(ctx, &p, n, 1, x, y, nullptr, nullptr, nullptr, nullptr, 1, &out);
    if (rc)
        return rc;
    if (sigma)
        *sigma = sg;
    if (slope)
        *slope = sl;
    if (extra_rob)
        *extra_rob = (fl & NS_F_EXTRA_ROB) ? 1 : 0;
    if (flags)
        *flags = fl;
    return NS_OK;
}

int ns_qvalue_grid(ns_ctx *ctx, double sigma, double slope, int n, const int32_t *coverage, const int32_t *ma_count,
                   int64_t m, const int32_t *gx, const int32_t *gy, double *qval)
{
    if (!ctx)
        return NS_ERR_ARG;
    if (n < 1 || m < 0 || !coverage || !ma_count || (m > 0 && (!gx || !gy || !qval)) || !(sigma > 0) || !(slope > 0)) {
        ctx->err = "ns_qvalue_grid: bad argument (the reference draws no contours when the fit is NA)";
        return NS_ERR_ARG;
    }
    if (m == 0)
        return NS_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t sm = (NS_RTAB + 3 * (size_t)n) * sizeof(double);
    if (sm > 227 * 1024) {
        ctx->err = "ns_qvalue_grid: too many samples for the on-chip working set";
        return NS_ERR_ARG;
    }
    cudaError_t e;
    if ((e = ctx->in_rows[0].reserve((size_t)n)) || (e = ctx->in_rows[1].reserve((size_t)n)) ||
        (e = ctx->in_rows[2].reserve((size_t)m)) || (e = ctx->in_rows[3].reserve((size_t)m)) ||
        (e = ctx->grid_out.reserve((size_t)m))) {
        ctx->err = std::string("device allocation (q-value grid) failed: ") + cudaGetErrorString(e);
        return NS_ERR_NOMEM;
    }
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->in_rows[0].p, coverage, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->in_rows[1].p, ma_count, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->in_rows[2].p, gx, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->in_rows[3].p, gy, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(ns_need_smem(ctx, 5, k_qgrid, sm));
    int blocks = (int)((m + 255) / 256);
    if (blocks > 2 * ctx->sm_count)
        blocks = 2 * ctx->sm_count;
    k_qgrid<<<blocks, 256, sm, st>>>(n, ctx->in_rows[0].p, ctx->in_rows[1].p, sigma, slope, m, ctx->in_rows[2].p,
                                     ctx->in_rows[3].p, ctx->grid_out.p);
    CK(cudaGetLastError());
    ctx->tm.kernel_launches++;
    CK(cudaMemcpyAsync(qval, ctx->grid_out.p, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return NS_OK;
}

} /* extern "C" */
