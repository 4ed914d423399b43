/*
 * ns_api_internal.hpp — host-side state of one GPU context (ns_ctx) shared by ns_api.cu (single-GPU entry points)
 * and ns_multi.cu (one workload over several GPUs from one host process).  Not part of the ABI.
 */
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <string>
#include <vector>

#include "ns_kernels.cuh"
#include "ns_hostmerge.hpp"

template <class T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap)
            return cudaSuccess;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess)
            cap = want;
        return e;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PlaneSet {
    DevBuf<double> pvalue, qvalue, gq, qval_minaf, sor, rvsb, fs, pooled, rsigma, rslope, rmaxgq;
    DevBuf<uint8_t> gt, status;
    DevBuf<uint32_t> rflags;
    cudaError_t reserve(size_t rows, int n)
    {
        cudaError_t e;
        size_t m = rows * (size_t)n;
        if ((e = pvalue.reserve(m)) || (e = qvalue.reserve(m)) || (e = gq.reserve(m)) || (e = qval_minaf.reserve(m)) ||
            (e = sor.reserve(m)) || (e = rvsb.reserve(m)) || (e = fs.reserve(m)) || (e = pooled.reserve(rows * 8)) ||
            (e = gt.reserve(m)) || (e = status.reserve(m)) || (e = rsigma.reserve(rows)) || (e = rslope.reserve(rows)) ||
            (e = rmaxgq.reserve(rows)) || (e = rflags.reserve(rows)))
            return e;
        return cudaSuccess;
    }
    ns_planes view()
    {
        ns_planes v = {pvalue.p, qvalue.p, gq.p, qval_minaf.p, sor.p, rvsb.p, fs.p, pooled.p, gt.p, status.p,
                       rsigma.p, rslope.p, rmaxgq.p, rflags.p};
        return v;
    }
    void release()
    {
        pvalue.release(), qvalue.release(), gq.release(), qval_minaf.release(), sor.release(), rvsb.release(),
            fs.release(), pooled.release(), gt.release(), status.release(), rsigma.release(), rslope.release(),
            rmaxgq.release(), rflags.release();
    }
};

/* page-locked host planes that grow: the pool's staging area and the private rows of a device in a multi-GPU call */
struct HostRows {
    ns_row_planes p = {};
    int64_t cap = 0;
    int n = 0;
    void release();
    /* at least `rows` rows of n samples; the first `keep` rows survive a re-allocation */
    bool reserve(int64_t rows, int n_samples, int64_t keep);
};

#define NS_POOL_STREAMS 4
#define NS_POOL_MAX_LAUNCH 256

/* Call-wide pool of cluster-path units: chunks append the units the gate kernel predicted for
 * the cluster kernel (and the ones the warp kernel handed over) together with a copy of their count rows; the
 * cluster kernel runs on side streams next to the gate / fit / post kernels of LATER chunks and is joined once, at
 * the end of the call, where k_post runs for all pooled units and their rows are merged into the caller's planes. */
struct HeavyPool {
    bool enabled = true;
    int cap = 0, n = 0;
    int n_used = 0, n_launch = 0, next_stream = 0;
    DevBuf<int32_t> rows[5];
    DevBuf<uint8_t> is_indel;
    DevBuf<uint32_t> flags, iters;
    DevBuf<unsigned long long> cycles;
    DevBuf<double> sigma, slope, max_gq;
    DevBuf<int64_t> unit, lightpos;
    DevBuf<int> ident, emitflag;
    DevBuf<ns_counters> cnt; /* one per launch + one for the post pass */
    PlaneSet planes;
    cudaStream_t streams[NS_POOL_STREAMS] = {};
    cudaEvent_t ev_ready = nullptr, ev_done[NS_POOL_STREAMS] = {}, ev_t0[NS_POOL_STREAMS] = {}, ev_t1[NS_POOL_STREAMS] = {};
    cudaEvent_t ev_tail0 = nullptr, ev_tail1 = nullptr, ev_post0 = nullptr, ev_post1 = nullptr;
    bool used[NS_POOL_STREAMS] = {};
    /* host staging of the per-unit results (page-locked, grows) */
    char *h_units = nullptr;
    size_t h_units_cap = 0;
    HostRows h_rows;
    int *h_count = nullptr; /* pinned */
};

/* where the emitted rows of a call go: the caller's planes (fixed capacity) or growable private planes */
struct RowSink {
    ns_row_planes p = {};
    int64_t cap = 0;
    HostRows *grow = nullptr; /* non-null: p/cap mirror *grow and may be enlarged */
    bool want_rows = false;
};

/* one queue of position chunks shared by the host threads of a multi-GPU call (a single context has its own) */
struct ChunkQueue {
    std::atomic<int64_t> next{0};
    int64_t n_chunks = 0, pchunk = 0;
    int64_t max_per_slot = 0;    /* most chunks one device may take (0: no limit) */
    std::atomic<int> failed{0};
    std::vector<int> owner;      /* device slot that took chunk c (multi-GPU calls; empty otherwise) */
};

struct ns_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr, own_stream = nullptr, copy_stream = nullptr, heavy_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_h0 = nullptr, ev_h1 = nullptr;
    std::string err;
    /* staging of host inputs, double buffered */
    DevBuf<int32_t> in_counts[2];
    DevBuf<uint8_t> in_ref[2];
    DevBuf<int32_t> in_rows[5];
    DevBuf<uint8_t> in_indel;
    DevBuf<uint16_t> in_counts16[2];
    DevBuf<double> grid_out;
    cudaEvent_t ev_h2d[2] = {}, ev_free[2] = {};
    /* per-chunk unit arrays */
    DevBuf<uint32_t> flags, iters;
    DevBuf<unsigned long long> cycles;
    DevBuf<double> sigma, slope, max_gq;
    DevBuf<int> needrow, row, emitflag, emitidx, fit_list, post_list, heavy_list, pre_list, emit_index_out;
    DevBuf<int64_t> emit_unit;
    DevBuf<ns_counters> counters;
    DevBuf<unsigned char> cub_tmp;
    PlaneSet rows, packed;
    size_t packed_cap = 0;
    HeavyPool pool;
    /* parameter-dependent device state */
    DevBuf<int32_t> pair_idx;
    DevBuf<uint8_t> role;
    DevBuf<double> qtabN, qtabT;
    double qtab_sigma = -1, qtab_afmin = -2;
    /* timings */
    cudaEvent_t ev[10] = {};
    ns_timings tm;
    ns_counters *h_counters = nullptr; /* pinned */
    int *h_started = nullptr, *d_started = nullptr; /* mapped pinned flag written by k_fit_heavy */
    int64_t chunk_units = 0;           /* 0 = auto */
    int64_t n_skipped_call = 0;
};

/* ---- shared between ns_api.cu and ns_multi.cu ---- */
int ns_prepare_params(ns_ctx *ctx, const ns_params *in, int n, ns_params *dprm, const double **qtab);
int64_t ns_pick_chunk(ns_ctx *ctx, int n, int64_t total_units, bool device_resident = false);
void ns_reset_out(ns_ctx *ctx, ns_out *out);
void ns_drain(ns_ctx *ctx);
RowSink ns_sink_from_out(ns_out *out);
/* SNV planes on the host, chunks pulled from `q`; per-unit results into `out` at their global offsets, emitted rows
 * into `sink` in the order this context processed its chunks; *emitted = rows this context wrote */
int ns_snv_host_queue(ns_ctx *ctx, const ns_params *prm, int n, int64_t P, const uint8_t *ref, const int32_t *counts,
                      const uint16_t *counts16, ns_out *out, RowSink *sink, ChunkQueue *q, int slot, int64_t *emitted);
