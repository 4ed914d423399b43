/*
 * ns_kernels.cuh — device-side bookkeeping structs and kernel declarations shared by the
 * translation units of libns_b200.so (kernels are split over several .cu files so that the
 * in-tree build compiles them in parallel).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ns_core.cuh"

#define NS_COOP_DOUBLES (4 * NS_MAX_KNOTS)
#define NS_RTAB_HEAVY 4096
#define NS_HEAVY_THREADS_TP 640 /* throughput form of the cluster kernel: 20 warps per CTA, 96 registers */
#ifndef NS_HEAVY_THREADS
#define NS_HEAVY_THREADS 512 /* 16 warps per CTA x 8 CTAs per cluster = one sample per warp at n <= 128 */
#endif

#define NS_WPB_GATE 8
#define NS_QTAB_LEN 65536

/* internal unit flag (never leaves the library): the unit's fit runs in the call-wide pool of cluster-path units;
 * the chunk it came from skips it in k_post / k_pack and the pool delivers its results at the end of the call */
#define NS_F_POOLED 0x80000000u

/* ---- device-side bookkeeping ------------------------------------------------------- */
struct ns_counters {
    int n_fit, n_post, n_rows, n_emit;
    int next_fit, next_post, n_heavy, next_heavy;
    int n_pre, next_pre, n_light, n_skipped; /* n_pre: units predicted heavy by the gate kernel; n_light: the rest */
    unsigned long long n_seed, n_lattice, n_quad, pad2;
    unsigned long long fast_seed, fast_lattice;
};

struct ns_planes {
    double *pvalue, *qvalue, *gq, *qval_minaf, *sor, *rvsb, *fs, *pooled;
    uint8_t *gt, *status;
    double *rsigma, *rslope, *rmaxgq; /* per row: the unit's sigma, slope, max(GQ) (packed planes only) */
    uint32_t *rflags;
};

/* device side of the call-wide pool: the five count rows of every pooled unit (copied out of the chunk's input
 * buffers, which are recycled), its flag word, global unit id and the number of emitted rows that precede it */
struct ns_pool_dev {
    int32_t *dp, *vp, *vm, *rp, *rm; /* [cap][n] */
    uint8_t *is_indel;               /* [cap] */
    uint32_t *flags;                 /* [cap] */
    int64_t *unit, *lightpos;        /* [cap] */
};

struct ns_unit_arrays {
    uint32_t *flags, *iters;
    unsigned long long *cycles;
    double *sigma, *slope, *max_gq;
    int *needrow, *row, *emitflag, *emitidx;
    int *fit_list, *post_list, *heavy_list, *pre_list;
};


__global__ void k_gate(ns_unit_src src, ns_params prm, int64_t u0, int n_units, ns_unit_arrays ua, ns_counters *cnt, int q_begin,
                       int lean);
__global__ void k_gate_v4(ns_unit_src src, ns_params prm, int64_t u0, int n_units, ns_unit_arrays ua, ns_counters *cnt, int q_begin,
                          int lean);
__global__ void k_gate_v2(ns_unit_src src, ns_params prm, int64_t u0, int n_units, ns_unit_arrays ua, ns_counters *cnt, int q_begin,
                          int lean);
/* load form of the streaming gate for SNV planes at `counts` (host and device agree on it): 4, 2 or 0 */
static inline int ns_gate_form(int mode, int extra_rob, int64_t u0, int n, double min_coverage, const void *counts_at_u0)
{
    if (mode != 0 || extra_rob || (u0 % 3) != 0 || !(min_coverage < 2.0e9))
        return 0;
    const uintptr_t addr = (uintptr_t)counts_at_u0;
    if ((n & 3) == 0 && n > 32 && n <= 128 && (addr & 15) == 0)
        return 4;
    if ((n & 1) == 0 && n > 16 && n <= 64 && (addr & 7) == 0)
        return 2;
    return 0;
}
__global__ void k_gate_staged(ns_unit_src src, ns_params prm, int64_t u0, int n_units, ns_unit_arrays ua, ns_counters *cnt,
                              int stage_k, int stages, int n_full_blocks, int lean);
__global__ void k_fit(ns_unit_src src, const __grid_constant__ ns_params prm, int64_t u0, ns_unit_arrays ua, ns_counters *cnt);
__global__ void k_fit_heavy(ns_unit_src src, const __grid_constant__ ns_params prm, int64_t u0, ns_unit_arrays ua, ns_counters *cnt,
                            int which, int *started);
__global__ void k_fit_heavy_tp(ns_unit_src src, const __grid_constant__ ns_params prm, int64_t u0, ns_unit_arrays ua, ns_counters *cnt,
                               int which, int *started);
__global__ void k_post(ns_unit_src src, ns_params prm, int64_t u0, ns_unit_arrays ua, ns_planes pl,
                       const double *qtab, int qtab_len, ns_counters *cnt);
__global__ void k_scan_total(const int *flag, const int *excl, int n, int *total);
__global__ void k_pack(int n, int n_units, int64_t u0, int64_t emit_off, int cap, ns_unit_arrays ua,
                       ns_planes src, ns_planes dst, int64_t *emit_unit, int *emit_index_out, int64_t *pool_lightpos,
                       ns_counters *cnt);
__global__ void k_pool_gather(ns_unit_src src, int64_t u0_src, int64_t u_off, const int *list, int count,
                              ns_unit_arrays ua, ns_pool_dev pool, int base, ns_counters *launch_cnt);
__global__ void k_qtab(double *tabN, double *tabT, int len, double sigma_normal, double afmin);
__global__ void k_qgrid(int n, const int32_t *cov, const int32_t *ma, double sigma, double slope, int64_t m, const int32_t *gx,
                        const int32_t *gy, double *out);
__global__ void k_dfma(double *out, int iters, double a, double b);
