/*
 * oracle/orc_model.h — TEST INFRASTRUCTURE ONLY (CPU oracle, model layer).
 *
 * Literal CPU restatement of the reference's per-site error-model hot path:
 *   glmrob.nb()               /root/reference/bin/glm_rob_nb.r:16-228
 *   per-allele driver body    /root/reference/bin/needlestack.r:321-409 (SNV),
 *                             :481-561 (DEL), :644-724 (INS)
 *   toQvalueN / toQvalueT     /root/reference/bin/needlestack.r:242-252
 *   SOR / common_annot        /root/reference/bin/needlestack.r:206-239
 * PARITY UNPINNED: R is not available, the reference holds no golden vectors
 * (SURVEY.md §4/§8c).  Never linked into or called from the product path.
 */
#ifndef ORC_MODEL_H
#define ORC_MODEL_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    /* glmrob.nb arguments, glm_rob_nb.r:16-18 */
    double c_tukey_beta, c_tukey_sig, minsig, maxsig, minmu, sig_prec, tol;
    double min_coverage, min_reads, min_af;
    double min_af_extra_rob, min_prop_extra_rob, max_prop_extra_rob;
    int32_t maxit, maxit_sig, n_ai_sig_tukey, size_min, extra_rob, _pad;
} orc_glm_params;

typedef struct {
    double sigma, slope; /* NaN when gated (R: NA) */
    int32_t gated, extra_rob, n_outer, n_irls, n_obj_eval, sig_at_min, maxit_hit, quad_error,
        nan_rows, _pad;
    int64_t n_seed, n_lattice, n_quad;
} orc_glm_info;

void orc_glm_params_default(orc_glm_params *p); /* glmrob.nb's own defaults */

/* glmrob.nb(y, x, ...) -> coverage/ma_count are the inputs; outputs length N */
int orc_glmrob_nb(int N, const double *y, const double *x, const orc_glm_params *prm,
                  double *pvalues, double *qvalues, double *gq, orc_glm_info *info);
/* same fit, recording sigma and beta after each of the first `cap` outer passes (test diagnostics) */
int orc_glmrob_nb_trace(int N, const double *y, const double *x, const orc_glm_params *prm, int cap, double *sig_trace,
                        double *beta_trace, orc_glm_info *info);

/* caller-level parameters, needlestack.r:64-92 */
typedef struct {
    orc_glm_params glm;
    double GQ_threshold, SB_threshold_SNV, SB_threshold_indel, sigma_normal, afmin_power;
    int32_t SB_type;         /* 0 SOR, 1 RVSB, 2 FS */
    int32_t output_all_SNVs; /* needlestack.r:72 */
    int32_t is_tn;           /* pairs file supplied */
    int32_t n_pairs, n_only_t, n_only_n;
    const int32_t *tindex, *nindex, *only_t, *only_n; /* 0-based sample indices */
} orc_call_params;

void orc_call_params_default(orc_call_params *p); /* needlestack.r:64-92 */

enum { ORC_GT_00 = 0, ORC_GT_01 = 1, ORC_GT_11 = 2, ORC_GT_MISSING = 3 };
enum {
    ORC_ST_NONE = 0,
    ORC_ST_SOMATIC = 1,
    ORC_ST_GERMLINE_CONFIRMED = 2,
    ORC_ST_GERMLINE_UNCONFIRMED = 3,
    ORC_ST_GERMLINE_UNCONFIRMABLE = 4,
    ORC_ST_UNKNOWN = 5,
    ORC_ST_POSSIBLE_CONTAMINATION = 6
};

typedef struct {
    orc_glm_info glm;
    int32_t ref_inv, gate1, gate2, _pad;
    double max_gq;
    /* pooled annotations (valid when gate1): all_sor, all_rvsb, FisherStrand_all,
     * sum Rp, sum Rm, sum Vp, sum Vm, all_DP */
    double pooled[8];
} orc_unit_info;

/* One (position, allele) unit exactly as the per-allele body of needlestack.r.
 * Per-sample outputs have length n; sor/rvsb/fs/gt are only meaningful when gate1. */
int orc_call_unit(int n, const int32_t *dp, const int32_t *vp, const int32_t *vm,
                  const int32_t *rp, const int32_t *rm, int is_indel, const orc_call_params *prm,
                  double *pvalue, double *qvalue, double *gq, double *qval_minaf, double *sor,
                  double *rvsb, double *fs, uint8_t *gt, uint8_t *status, orc_unit_info *info);

/* SNV batch: counts plane-major [9][P][n] (DP,A,T,C,G,a,t,c,g), ref code 0..3 = A,T,C,G
 * (anything else: position skipped, needlestack.r:319).  Units = 3 per position in
 * setdiff(c(A,T,C,G),ref) order.  Dense outputs [3P][n]; info[3P]. n_threads>1 uses pthreads. */
int orc_call_snv_batch(int64_t P, int n, const uint8_t *ref, const int32_t *counts,
                       const orc_call_params *prm, int n_threads, double *pvalue, double *qvalue,
                       double *gq, double *qval_minaf, double *sor, double *rvsb, double *fs,
                       uint8_t *gt, uint8_t *status, orc_unit_info *info);

/* toQvalue(x, y) of plot_rob_nb.r:83-87 for m grid points against one fitted unit */
int orc_plot_toqvalue(int n, const int32_t *coverage, const int32_t *ma_count, double sigma, double slope,
                      int64_t m, const int32_t *gx, const int32_t *gy, double *qval);

#ifdef __cplusplus
}
#endif
#endif
