/*
 * ns_rshim.c — thin .Call shim between R and libns_b200.so (include/ns_b200.h).
 *
 * NOT COMPILED IN THIS IMAGE (no R headers here); build where R is installed with
 *     R CMD SHLIB ns_rshim.c -I../../include -L../lib -lns_b200
 * No logic lives here: it unpacks R vectors, calls the C ABI and wraps the outputs.
 * R integer matrices are column-major, so a [n x U] integer matrix IS the [U][n] row layout the
 * library takes (one column = one unit).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>
#include "ns_b200.h"

static void ctx_finalizer(SEXP xp)
{
    ns_ctx *c = (ns_ctx *)R_ExternalPtrAddr(xp);
    if (c) {
        ns_destroy(c);
        R_ClearExternalPtr(xp);
    }
}

static ns_ctx *get_ctx(SEXP xp)
{
    ns_ctx *c = (ns_ctx *)R_ExternalPtrAddr(xp);
    if (!c)
        error("needlestack-b200: context was destroyed");
    return c;
}

SEXP C_ns_create(SEXP device)
{
    char err[512] = "";
    ns_ctx *c = ns_create(asInteger(device), err, sizeof err);
    if (!c)
        error("%s", err);
    SEXP xp = PROTECT(R_MakeExternalPtr(c, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(xp, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return xp;
}

/* named numeric list -> ns_params; unknown names are an error so that typos cannot pass silently */
static void fill_params(SEXP lst, ns_params *p, int caller_defaults)
{
    if (caller_defaults)
        ns_params_caller_default(p);
    else
        ns_params_glm_default(p);
    SEXP nm = getAttrib(lst, R_NamesSymbol);
    for (R_xlen_t i = 0; i < XLENGTH(lst); i++) {
        const char *k = CHAR(STRING_ELT(nm, i));
        double v = asReal(VECTOR_ELT(lst, i));
#define D(f) if (!strcmp(k, #f)) { p->f = v; continue; }
#define I(f) if (!strcmp(k, #f)) { p->f = (int32_t)v; continue; }
        D(c_tukey_beta) D(c_tukey_sig) D(minsig) D(maxsig) D(minmu) D(sig_prec) D(tol) D(min_coverage) D(min_reads)
        D(min_af) D(min_af_extra_rob) D(min_prop_extra_rob) D(max_prop_extra_rob) D(GQ_threshold) D(SB_threshold_SNV)
        D(SB_threshold_indel) D(sigma_normal) D(afmin_power)
        I(maxit) I(maxit_sig) I(n_ai_sig_tukey) I(size_min) I(extra_rob) I(SB_type) I(output_all_SNVs) I(emit_mode)
#undef D
#undef I
        error("needlestack-b200: unknown parameter '%s'", k);
    }
}

/* glmrob.nb(y, x, ...): y, x integer vectors; returns list(sigma, slope, pvalues, qvalues, GQ, extra_rob, flags) */
SEXP C_ns_glmrob_nb(SEXP ctx, SEXP y, SEXP x, SEXP params)
{
    ns_ctx *c = get_ctx(ctx);
    ns_params p;
    fill_params(params, &p, 0);
    int n = (int)XLENGTH(y);
    if (XLENGTH(x) != n)
        error("length(y) does not match dim(X)[1].");
    SEXP pv = PROTECT(allocVector(REALSXP, n)), qv = PROTECT(allocVector(REALSXP, n)),
         gq = PROTECT(allocVector(REALSXP, n));
    double sigma, slope;
    int32_t er;
    uint32_t fl;
    int rc = ns_glmrob_nb(c, &p, n, INTEGER(y), INTEGER(x), &sigma, &slope, REAL(pv), REAL(qv), REAL(gq), &er, &fl);
    if (rc)
        error("needlestack-b200: %s", ns_last_error(c));
    const char *names[] = {"sigma", "slope", "pvalues", "qvalues", "GQ", "extra_rob", "flags", ""};
    SEXP res = PROTECT(mkNamed(VECSXP, names));
    SET_VECTOR_ELT(res, 0, ScalarReal(fl & NS_F_GATED ? NA_REAL : sigma));
    SET_VECTOR_ELT(res, 1, ScalarReal(fl & NS_F_GATED ? NA_REAL : slope));
    SET_VECTOR_ELT(res, 2, pv);
    SET_VECTOR_ELT(res, 3, qv);
    SET_VECTOR_ELT(res, 4, gq);
    SET_VECTOR_ELT(res, 5, ScalarLogical(er));
    SET_VECTOR_ELT(res, 6, ScalarInteger((int)fl));
    UNPROTECT(4);
    return res;
}

/* One chunk of positions of the readcounts table.
 * counts: integer array dim = c(n, 9, P) (so that memory is [P][9][n]); ref: raw/integer codes 0..3.
 * tn: NULL or list(tindex, nindex, only_t, only_n) of 0-based integer vectors.
 * Returns per-unit vectors and the emitted rows as [n x E] matrices. */
SEXP C_ns_call_snv(SEXP ctx, SEXP counts, SEXP ref, SEXP params, SEXP tn, SEXP emit_cap)
{
    ns_ctx *c = get_ctx(ctx);
    ns_params p;
    fill_params(params, &p, 1);
    SEXP dim = getAttrib(counts, R_DimSymbol);
    int n = INTEGER(dim)[0];
    int64_t P = INTEGER(dim)[2];
    if (tn != R_NilValue) {
        p.is_tn = 1;
        p.tindex = INTEGER(VECTOR_ELT(tn, 0));
        p.nindex = INTEGER(VECTOR_ELT(tn, 1));
        p.n_pairs = (int32_t)XLENGTH(VECTOR_ELT(tn, 0));
        p.only_t = INTEGER(VECTOR_ELT(tn, 2));
        p.n_only_t = (int32_t)XLENGTH(VECTOR_ELT(tn, 2));
        p.only_n = INTEGER(VECTOR_ELT(tn, 3));
        p.n_only_n = (int32_t)XLENGTH(VECTOR_ELT(tn, 3));
    }
    int64_t U = 3 * P, E = (int64_t)asReal(emit_cap);
    ns_out o;
    memset(&o, 0, sizeof o);
    const char *names[] = {"sigma", "slope", "max_gq", "flags", "emit_index", "n_emitted", "emit_unit", "pvalue",
                           "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs", "gt", "status", "pooled", ""};
    SEXP res = PROTECT(mkNamed(VECSXP, names));
    SEXP v;
#define VEC(i, type, len) SET_VECTOR_ELT(res, i, v = allocVector(type, len))
    VEC(0, REALSXP, U); o.sigma = REAL(v);
    VEC(1, REALSXP, U); o.slope = REAL(v);
    VEC(2, REALSXP, U); o.max_gq = REAL(v);
    VEC(3, INTSXP, U);  o.flags = (uint32_t *)INTEGER(v);
    VEC(4, INTSXP, U);  o.emit_index = INTEGER(v);
    SEXP eu = PROTECT(allocVector(REALSXP, E)); /* int64 ids do not fit R integers: returned as doubles below */
    int64_t *ids = (int64_t *)R_alloc(E > 0 ? E : 1, sizeof(int64_t));
    o.emit_unit = ids;
    o.emit_cap = E;
#define MAT(i, field) SET_VECTOR_ELT(res, i, v = allocMatrix(REALSXP, n, (int)E)); o.field = REAL(v)
    MAT(7, pvalue); MAT(8, qvalue); MAT(9, gq); MAT(10, qval_minaf); MAT(11, sor); MAT(12, rvsb); MAT(13, fs);
    SET_VECTOR_ELT(res, 14, v = allocMatrix(RAWSXP, n, (int)E)); o.gt = RAW(v);
    SET_VECTOR_ELT(res, 15, v = allocMatrix(RAWSXP, n, (int)E)); o.status = RAW(v);
    SET_VECTOR_ELT(res, 16, v = allocMatrix(REALSXP, 8, (int)E)); o.pooled = REAL(v);
    const uint8_t *refp = TYPEOF(ref) == RAWSXP ? RAW(ref) : NULL;
    if (!refp)
        error("ref must be a raw vector of base codes (0..3 = A,T,C,G)");
    int rc = ns_call_snv(c, &p, n, P, refp, INTEGER(counts), &o);
    if (rc && rc != NS_ERR_EMIT_OVERFLOW)
        error("needlestack-b200: %s", ns_last_error(c));
    for (int64_t i = 0; i < E && i < o.n_emitted; i++)
        REAL(eu)[i] = (double)ids[i];
    SET_VECTOR_ELT(res, 5, ScalarReal((double)o.n_emitted)); /* > emit_cap: call again with a larger cap */
    SET_VECTOR_ELT(res, 6, eu);
    UNPROTECT(2);
    return res;
}

/* Explicit units (indel alleles of needlestack.r:461-567 / :624-730, the DP / AD units of annotate_vcf.r:57-69):
 * dp, vp and the optional vm, rp, rm are integer matrices dim = c(n, U) (memory [U][n]); is_indel raw(U) or NULL.
 * Returns per-unit sigma, slope, flags and the [n x U] GQ / q-value / p-value matrices of every unit (emit_mode ALL). */
SEXP C_ns_call_units(SEXP ctx, SEXP dp, SEXP vp, SEXP vm, SEXP rp, SEXP rm, SEXP is_indel, SEXP params)
{
    ns_ctx *c = get_ctx(ctx);
    ns_params p;
    fill_params(params, &p, 1);
    p.emit_mode = NS_EMIT_ALL;
    SEXP dim = getAttrib(dp, R_DimSymbol);
    const int n = INTEGER(dim)[0];
    const int64_t U = INTEGER(dim)[1];
    ns_out o;
    memset(&o, 0, sizeof o);
    const char *names[] = {"sigma", "slope", "flags", "pvalue", "qvalue", "gq", ""};
    SEXP res = PROTECT(mkNamed(VECSXP, names));
    SEXP v;
    SET_VECTOR_ELT(res, 0, v = allocVector(REALSXP, U)); o.sigma = REAL(v);
    SET_VECTOR_ELT(res, 1, v = allocVector(REALSXP, U)); o.slope = REAL(v);
    SET_VECTOR_ELT(res, 2, v = allocVector(INTSXP, U));  o.flags = (uint32_t *)INTEGER(v);
    SET_VECTOR_ELT(res, 3, v = allocMatrix(REALSXP, n, (int)U)); o.pvalue = REAL(v);
    SET_VECTOR_ELT(res, 4, v = allocMatrix(REALSXP, n, (int)U)); o.qvalue = REAL(v);
    SET_VECTOR_ELT(res, 5, v = allocMatrix(REALSXP, n, (int)U)); o.gq = REAL(v);
    o.emit_cap = U;
#define OPT(x) ((x) == R_NilValue ? NULL : INTEGER(x))
    int rc = ns_call_units(c, &p, n, U, INTEGER(dp), INTEGER(vp), OPT(vm), OPT(rp), OPT(rm),
                           is_indel == R_NilValue ? NULL : RAW(is_indel), 0, &o);
#undef OPT
    if (rc)
        error("needlestack-b200: %s", ns_last_error(c));
    UNPROTECT(1);
    return res;
}

/* toQvalue(x, y) of plot_rob_nb.r:83-87 on a grid: gx, gy integer vectors of equal length */
SEXP C_ns_qvalue_grid(SEXP ctx, SEXP sigma, SEXP slope, SEXP coverage, SEXP ma_count, SEXP gx, SEXP gy)
{
    ns_ctx *c = get_ctx(ctx);
    const R_xlen_t m = XLENGTH(gx);
    if (XLENGTH(gy) != m || XLENGTH(coverage) != XLENGTH(ma_count))
        error("needlestack-b200: coverage/ma_count and gx/gy must have equal lengths");
    SEXP q = PROTECT(allocVector(REALSXP, m));
    int rc = ns_qvalue_grid(c, asReal(sigma), asReal(slope), (int)XLENGTH(coverage), INTEGER(coverage), INTEGER(ma_count),
                            (int64_t)m, INTEGER(gx), INTEGER(gy), REAL(q));
    if (rc)
        error("needlestack-b200: %s", ns_last_error(c));
    UNPROTECT(1);
    return q;
}

static const R_CallMethodDef call_methods[] = {{"C_ns_create", (DL_FUNC)&C_ns_create, 1},
                                               {"C_ns_glmrob_nb", (DL_FUNC)&C_ns_glmrob_nb, 4},
                                               {"C_ns_call_snv", (DL_FUNC)&C_ns_call_snv, 6},
                                               {"C_ns_call_units", (DL_FUNC)&C_ns_call_units, 8},
                                               {"C_ns_qvalue_grid", (DL_FUNC)&C_ns_qvalue_grid, 7},
                                               {NULL, NULL, 0}};

void R_init_ns_rshim(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
