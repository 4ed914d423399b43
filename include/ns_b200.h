/*
 * ns_b200.h — C ABI of the B200 error-model library (libns_b200.so).
 *
 * This is the drop-in boundary for needlestack's per-site hot path: plain C, plain
 * pointers and sizes, no R / torch / C++ types.  Each entry point names the reference
 * interface it replaces (paths under the needlestack tree, bin/...).
 *
 *   reference today (interpreted R)                      this library
 *   ------------------------------------------------     -----------------------------------
 *   glmrob.nb(y, x, ...)        bin/glm_rob_nb.r:16-228   ns_glmrob_nb(), ns_call_units(plain=1)
 *   per-allele SNV body         bin/needlestack.r:321-413 ns_call_snv(), ns_call_snv_device()
 *   per-allele DEL / INS body   bin/needlestack.r:461-567,
 *                               :624-730                  ns_call_units(is_indel=1)
 *   annotate_vcf.r DP/AD loop   bin/annotate_vcf.r:57-69  ns_call_units()
 *   toQvalueN / toQvalueT       bin/needlestack.r:242-252 (fused into the calls above)
 *   SOR / common_annot          bin/needlestack.r:206-239 (fused)
 *
 * One "unit" = one (position, candidate allele) = one glmrob.nb() call of the reference.
 * All count inputs are int32.  All entry points are synchronous: when they return, every
 * requested output is in the caller's host memory.  Return 0 on success, a negative
 * NS_ERR_* otherwise (message via ns_last_error).  The library has no CPU path: without a
 * CUDA device ns_create() fails.
 */
#ifndef NS_B200_H
#define NS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NS_ABI_VERSION 2

/* ---- error codes ---------------------------------------------------------------- */
#define NS_OK 0
#define NS_ERR_CUDA (-1)          /* a CUDA runtime call failed */
#define NS_ERR_ARG (-2)           /* bad argument (null pointer, n < 1, unsupported option) */
#define NS_ERR_EMIT_OVERFLOW (-3) /* more emitted units than out->emit_cap; out->n_emitted = needed */
#define NS_ERR_NOMEM (-4)

/* ---- unit flag word (out->flags[u]) ---------------------------------------------- */
#define NS_F_GATED 0x001u      /* glm_rob_nb.r:38-49 returned the NA result */
#define NS_F_EXTRA_ROB 0x002u  /* extra-robust pre-filter removed samples (glm_rob_nb.r:22-36) */
#define NS_F_REF_INV 0x004u    /* regression on the REF counts (needlestack.r:330-337) */
#define NS_F_SIG_AT_MIN 0x008u /* last uniroot() failed -> sigma = minsig (glm_rob_nb.r:175-178) */
#define NS_F_MAXIT 0x010u      /* outer loop stopped by maxit */
#define NS_F_QUAD_ERROR 0x020u /* integrate() error inside the IRLS loop: the reference would stop() */
#define NS_F_NAN_ROWS 0x040u   /* NaN weights/responses dropped in the WLS step */
#define NS_F_GATE1 0x080u      /* needlestack.r:381 passed (strand-bias annotations valid) */
#define NS_F_GATE2 0x100u      /* needlestack.r:387 passed (the reference writes a VCF line) */
#define NS_F_SKIPPED 0x200u    /* REF not in A,T,C,G: position skipped (needlestack.r:319) */
#define NS_F_QUAD_OVERFLOW 0x400u /* quadrature needed more than the on-chip interval list */

/* genotype codes (needlestack.r:403-409) and somatic status (needlestack.r:342-373) */
enum { NS_GT_00 = 0, NS_GT_01 = 1, NS_GT_11 = 2, NS_GT_MISSING = 3 };
enum {
    NS_ST_NONE = 0, /* "." */
    NS_ST_SOMATIC = 1,
    NS_ST_GERMLINE_CONFIRMED = 2,
    NS_ST_GERMLINE_UNCONFIRMED = 3,
    NS_ST_GERMLINE_UNCONFIRMABLE = 4,
    NS_ST_UNKNOWN = 5,
    NS_ST_POSSIBLE_CONTAMINATION = 6
};

/* which units get per-sample rows in the emitted planes */
enum {
    NS_EMIT_CALLS = 0,  /* units the reference writes to the VCF (gate 2) */
    NS_EMIT_FITTED = 1, /* every unit that was fitted (not gated) */
    NS_EMIT_ALL = 2     /* every unit that was not skipped (gated ones: p=q=1, GQ=0) */
};

/* ---- parameters ------------------------------------------------------------------ */
typedef struct ns_params {
    /* glmrob.nb() arguments, bin/glm_rob_nb.r:16-18 (maxmu is ignored there, :20) */
    double c_tukey_beta, c_tukey_sig, minsig, maxsig, minmu, sig_prec, tol;
    double min_coverage, min_reads, min_af;
    double min_af_extra_rob, min_prop_extra_rob, max_prop_extra_rob;
    int32_t maxit, maxit_sig, n_ai_sig_tukey, size_min, extra_rob;
    /* caller level, bin/needlestack.r:64-92 */
    int32_t SB_type; /* 0 SOR, 1 RVSB, 2 FS */
    int32_t output_all_SNVs;
    int32_t is_tn; /* a pairs file was supplied (needlestack.r:162-185) */
    double GQ_threshold, SB_threshold_SNV, SB_threshold_indel;
    double sigma_normal; /* --sigma_normal, toQvalueN */
    double afmin_power;  /* -1: unset */
    /* tumour/normal index sets, 0-based sample indices, HOST pointers (copied per call) */
    int32_t n_pairs, n_only_t, n_only_n;
    int32_t emit_mode; /* NS_EMIT_* */
    const int32_t *tindex, *nindex, *only_t, *only_n;
    const uint8_t *role; /* internal (device side); ignored on input */
} ns_params;

/* glmrob.nb()'s own defaults (bin/glm_rob_nb.r:16-18); caller fields zeroed */
void ns_params_glm_default(ns_params *p);
/* needlestack.r's defaults (bin/needlestack.r:64-92) on top of the above */
void ns_params_caller_default(ns_params *p);

/* ---- outputs (all caller-allocated HOST memory; any pointer may be NULL = not wanted) */
typedef struct ns_out {
    /* per unit, length n_units */
    double *sigma, *slope; /* NaN when gated (R: NA) */
    double *max_gq;        /* max(reg_res$GQ): the VCF QUAL column.  With NS_EMIT_CALLS it is NaN for units whose
                              smallest p-value already rules out a call (their multiple-testing step is skipped) */
    uint32_t *flags;       /* NS_F_* */
    uint32_t *iters;       /* n_outer | n_irls << 8 | n_objective_evals << 18 */
    int32_t *emit_index;   /* row in the emitted planes, or -1 */
    uint64_t *unit_cycles; /* diagnostic: SM clock cycles the fit of each unit took (0 if gated) */
    /* emitted rows, in increasing unit order */
    int64_t emit_cap;  /* in: rows available in the planes below */
    int64_t n_emitted; /* out */
    int64_t *emit_unit; /* [emit_cap] unit id of each row */
    double *pvalue, *qvalue, *gq, *qval_minaf, *sor, *rvsb, *fs; /* [emit_cap][n] */
    uint8_t *gt, *status;                                        /* [emit_cap][n] */
    double *pooled; /* [emit_cap][8]: all_sor, all_rvsb, FisherStrand_all, sum Rp, Rm, Vp, Vm, all_DP */
    /* batch counters (out) */
    int64_t n_fitted, n_gated, n_skipped;
    uint64_t n_seed, n_lattice, n_quad; /* work counters of SURVEY.md section 8(d) */
    /* per emitted row (ABI 2; any may be NULL): sigma, slope, max(GQ) and the flag word of the row's unit.  A caller
     * that only wants the variants leaves the per-unit arrays above NULL: a WGS-like chunk of 12 M units is 0.5 GB of
     * per-unit device-to-host traffic against a few hundred emitted rows. */
    double *row_sigma, *row_slope, *row_max_gq; /* [emit_cap] */
    uint32_t *row_flags;                        /* [emit_cap] */
} ns_out;

/* per-call device timings of the last ns_call_* on this context, milliseconds */
typedef struct ns_timings {
    double h2d_ms, gate_ms, fit_ms, heavy_ms, post_ms, pack_ms, d2h_ms, total_ms;
    double heavy_pre_ms; /* CTA-per-unit kernel on its own stream, concurrent with fit_ms */
    double fit_phase_ms; /* wall time of the whole fit phase (both streams) */
    int64_t kernel_launches;
    int64_t n_heavy; /* units that took the spline + quadrature kernel */
    int64_t n_deferred; /* ... of which the warp kernel handed over (not predicted by the gate kernel) */
    uint64_t fast_seed, fast_lattice; /* work counters of the warp kernel alone (k_fit) */
    /* call-wide pool of cluster-path units (ABI 2) */
    int64_t pool_units;  /* units whose fit ran in the pool, joined once at the end of the call */
    int64_t chunks;      /* chunks this context processed */
    double pool_tail_ms; /* time the end of the call waited for the pool's last units */
} ns_timings;

typedef struct ns_ctx ns_ctx;

/* ---- life cycle ------------------------------------------------------------------ */
int ns_abi_version(void);
/* sizeof(ns_params), sizeof(ns_out), sizeof(ns_timings) for which = 0, 1, 2: lets a binding check its layout */
size_t ns_sizeof(int which);
/* one context per GPU and host thread; device = CUDA ordinal */
ns_ctx *ns_create(int device, char *err, size_t errlen);
void ns_destroy(ns_ctx *ctx);
const char *ns_last_error(ns_ctx *ctx);
/* run the context's kernels on a caller-owned CUDA stream (cudaStream_t), 0 = own stream */
int ns_set_stream(ns_ctx *ctx, void *cuda_stream);
/* page-locked host buffers for the columnar loader */
void *ns_alloc_pinned(size_t bytes);
void ns_free_pinned(void *p);
int ns_get_timings(ns_ctx *ctx, ns_timings *t);

/* ---- the hot path ---------------------------------------------------------------- */
/*
 * SNV units of P positions.  counts: int32 [P][9][n], planes in the table's column order
 * depth, A, T, C, G, a, t, c, g (bin/needlestack.r:188-198); ref: uint8 [P], 0..3 = A,T,C,G,
 * anything else = position skipped (:319).  Unit u = 3*p + k is the k-th allele of
 * setdiff(c("A","T","C","G"), ref) (:321).  Replaces the body of needlestack.r:321-413
 * up to (not including) the text of the VCF line.
 */
int ns_call_snv(ns_ctx *ctx, const ns_params *prm, int n_samples, int64_t n_positions,
                const uint8_t *ref, const int32_t *counts, ns_out *out);
/* same with 16-bit count planes (uint16 [P][9][n]; every count must fit): half the host-to-device bytes, widened
 * to the int32 rows of the kernels on the device.  For WGS-like shapes, where the host link bounds the throughput. */
int ns_call_snv_u16(ns_ctx *ctx, const ns_params *prm, int n_samples, int64_t n_positions,
                    const uint8_t *ref, const uint16_t *counts, ns_out *out);
/* same, inputs already resident on the context's device */
int ns_call_snv_device(ns_ctx *ctx, const ns_params *prm, int n_samples, int64_t n_positions,
                       const uint8_t *d_ref, const int32_t *d_counts, ns_out *out);
/*
 * Explicit units: int32 rows [n_units][n] of depth, alt fwd, alt rev, ref fwd, ref rev
 * (vm, rp, rm may be NULL = zeros).  is_indel: uint8 [n_units] or NULL.  plain != 0 runs
 * glmrob.nb(y = vp (+vm), x = dp) without the caller's ref-inversion test.
 * Replaces needlestack.r:461-567 (DEL), :624-730 (INS), annotate_vcf.r:57-69.
 */
int ns_call_units(ns_ctx *ctx, const ns_params *prm, int n_samples, int64_t n_units,
                  const int32_t *dp, const int32_t *vp, const int32_t *vm, const int32_t *rp,
                  const int32_t *rm, const uint8_t *is_indel, int plain, ns_out *out);
/*
 * One glmrob.nb(y, x, ...) call (bin/glm_rob_nb.r:16): outputs are the fields of its
 * result list: coef = (sigma, slope), pvalues, qvalues, GQ (length n), extra_rob.
 */
int ns_glmrob_nb(ns_ctx *ctx, const ns_params *prm, int n_samples, const int32_t *y,
                 const int32_t *x, double *sigma, double *slope, double *pvalues, double *qvalues,
                 double *gq, int32_t *extra_rob, uint32_t *flags);

/*
 * The contour / power grid of plot_rob_nb (bin/plot_rob_nb.r:80-121): toQvalue(x, y) (:83-87) for m
 * points (gx = DP, gy = AO) against one fitted unit (coverage, ma_count, sigma, slope as returned by
 * glmrob.nb).  qval[g] = -10 log10 of the Holm-adjusted (p.adjust's default) tail probability of the
 * new point among the n + 1; 0 where gy > gx.  The reference evaluates up to ~3500 such points per
 * plotted variant, each with n + 1 tail evaluations and a sort.
 */
int ns_qvalue_grid(ns_ctx *ctx, double sigma, double slope, int n_samples, const int32_t *coverage,
                   const int32_t *ma_count, int64_t m, const int32_t *gx, const int32_t *gy, double *qval);

/* ---- columnar loader for the mpileup2readcounts text table ------------------------------------
 * Replaces the per-line strsplit / column gathers of bin/needlestack.r:316-329 (SNV), :461-479 (DEL),
 * :624-643 (INS).  Host code, multi-threaded; the count planes are page-locked when a CUDA device is
 * present so they can go straight into ns_call_snv.  The table owns everything it returns. */
typedef struct ns_table ns_table;
/* text: the whole table (3 + 11 n tab-separated columns per line; chr, loc, ref, then per sample
 * depth,A,T,C,G,a,t,c,g,INS,DEL); has_header != 0 skips the first line (bin/needlestack.r:317). */
ns_table *ns_table_parse(const char *text, size_t len, int n_samples, int has_header, int n_threads, char *err,
                         size_t errlen);
void ns_table_free(ns_table *t);
int64_t ns_table_positions(const ns_table *t);
int ns_table_samples(const ns_table *t);
int ns_table_counts_pinned(const ns_table *t);
const int32_t *ns_table_counts(const ns_table *t);  /* [P][9][n], the `counts` argument of ns_call_snv */
const uint8_t *ns_table_ref(const ns_table *t);     /* [P] 0..3 = A,T,C,G, 4 = skipped (bin/needlestack.r:319) */
const char *ns_table_ref_text(const ns_table *t);   /* [P] first character of the REF column as written */
const int64_t *ns_table_loc(const ns_table *t);     /* [P] column 2 */
const int32_t *ns_table_chrom(const ns_table *t);   /* [P] index into the chromosome dictionary */
int ns_table_n_chrom(const ns_table *t);
const char *ns_table_chrom_name(const ns_table *t, int id);
const int64_t *ns_table_indel_cov(const ns_table *t); /* [P][2] sum of all deletion / insertion counts */
/* indel units: one per (position, unique upper-cased sequence), deletions before insertions per line */
int64_t ns_table_indel_units(const ns_table *t);
const int64_t *ns_table_indel_pos(const ns_table *t);  /* [Ui] row of the position */
const uint8_t *ns_table_indel_type(const ns_table *t); /* [Ui] 1 = deletion, 2 = insertion */
const char *ns_table_indel_seq(const ns_table *t, int64_t u);
const int32_t *ns_table_indel_vp(const ns_table *t);   /* [Ui][n] forward-strand counts */
const int32_t *ns_table_indel_vm(const ns_table *t);   /* [Ui][n] reverse-strand counts */

/* ---- VCF record writer --------------------------------------------------------------------------
 * Replaces the per-variant cat(..., append=T) calls of bin/needlestack.r:396-413 (SNV), :549-567 (DEL), :712-730 (INS):
 * same fields in the same order, numbers as R's cat() prints them under options(digits=7, scipen=100) (:153), the
 * 3-bp CONT= context from an in-memory FASTA instead of two `samtools faidx` processes per variant (:389-394).
 * Host code, multi-threaded.  `out` is the HOST result of ns_call_snv on the planes of `t` (ns_vcf_format_snv) or of
 * ns_call_units on the indel units of `t` (ns_vcf_format_indel) with every emitted plane and emit_unit present; one
 * record per emitted row, in row order.  sample_names: n strings (the POSSIBLE_CONTAMINATION_FROM_ text, :372-373). */
typedef struct ns_fasta ns_fasta;
typedef struct ns_vcf_text ns_vcf_text;
ns_fasta *ns_fasta_load(const char *path); /* NULL when the file cannot be read */
void ns_fasta_free(ns_fasta *fa);
ns_vcf_text *ns_vcf_format_snv(const ns_table *t, const ns_out *out, const char *const *sample_names, const ns_fasta *fa,
                               double gq_threshold, int n_threads);
ns_vcf_text *ns_vcf_format_indel(const ns_table *t, const ns_out *out, const char *const *sample_names,
                                 const ns_fasta *fa, double gq_threshold, int n_threads);
const char *ns_vcf_text_data(const ns_vcf_text *v);
size_t ns_vcf_text_size(const ns_vcf_text *v);
const int64_t *ns_vcf_text_offsets(const ns_vcf_text *v); /* [records + 1]: record e = data[off[e] .. off[e+1]) */
void ns_vcf_text_free(ns_vcf_text *v);
/* one number as the writer prints it; na_as_nan != 0: NaN prints as "NaN" (a 0/0 of the reference), else "NA";
 * returns the length of the full text */
int ns_vcf_format_number(double x, int na_as_nan, char *buf, size_t buflen);

/* ---- one workload over several GPUs from one host process -------------------------------------
 * Replaces the reference's region split + final merge: `--nsplit` / bin/bed_cut.r:19-58 cut the BED into pieces, one
 * single-threaded Rscript per piece (needlestack.nf:421-503), per-piece VCFs concatenated and sorted at the end
 * (needlestack.nf:546-553).  Here ONE call takes the whole table: a queue of position chunks, one host thread per
 * device pulling from it, results gathered in chunk (= genomic) order on the host; outputs are exactly those of
 * ns_call_snv on one device.  devices: CUDA ordinals (an ordinal may be listed twice: two contexts on one GPU).
 * This is how a single-threaded R process reaches every GPU of the box with one .Call. */
typedef struct ns_multi ns_multi;
ns_multi *ns_create_multi(const int *devices, int n_devices, char *err, size_t errlen);
void ns_destroy_multi(ns_multi *m);
int ns_multi_devices(const ns_multi *m);
const char *ns_multi_last_error(ns_multi *m);
int ns_multi_call_snv(ns_multi *m, const ns_params *prm, int n_samples, int64_t n_positions, const uint8_t *ref,
                      const int32_t *counts, ns_out *out);
int ns_multi_call_snv_u16(ns_multi *m, const ns_params *prm, int n_samples, int64_t n_positions, const uint8_t *ref,
                          const uint16_t *counts, ns_out *out);
/* device timings of slot `slot` in the last multi call (t->chunks = chunks it pulled) and the wall time of its host
 * thread, ms */
int ns_multi_get_timings(ns_multi *m, int slot, ns_timings *t, double *wall_ms);

/* FP64 FMA peak of the context's device, TFLOP/s (the denominator of the fit roofline) */
int ns_measure_fp64_peak(ns_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
