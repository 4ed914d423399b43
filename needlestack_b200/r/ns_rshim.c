/*
 * ns_rshim.c — thin .Call shim between R and libns_b200.so (include/ns_b200.h).
 *
 * Build where R is installed with
 *     R CMD SHLIB ns_rshim.c -I../../include -L../lib -lns_b200
 * R is not installed in this image: here the file is compiled against a stub of the R API (tests/rstub/, test
 * infrastructure) and every routine below is driven through it on a GPU (tests/test_rshim_stub.py).
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

/* all GPUs of the box from the single R thread: devices = integer vector of CUDA ordinals */
static void multi_finalizer(SEXP xp)
{
    ns_multi *m = (ns_multi *)R_ExternalPtrAddr(xp);
    if (m) {
        ns_destroy_multi(m);
        R_ClearExternalPtr(xp);
    }
}

SEXP C_ns_create_multi(SEXP devices)
{
    char err[512] = "";
    ns_multi *m = ns_create_multi(INTEGER(devices), (int)XLENGTH(devices), err, sizeof err);
    if (!m)
        error("%s", err);
    SEXP xp = PROTECT(R_MakeExternalPtr(m, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(xp, multi_finalizer, TRUE);
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

/* tn: NULL or list(tindex, nindex, only_t, only_n) of 0-based integer vectors (bin/needlestack.r:162-185) */
static void fill_tn(SEXP tn, ns_params *p)
{
    if (tn == R_NilValue)
        return;
    p->is_tn = 1;
    p->tindex = INTEGER(VECTOR_ELT(tn, 0));
    p->nindex = INTEGER(VECTOR_ELT(tn, 1));
    p->n_pairs = (int32_t)XLENGTH(VECTOR_ELT(tn, 0));
    p->only_t = INTEGER(VECTOR_ELT(tn, 2));
    p->n_only_t = (int32_t)XLENGTH(VECTOR_ELT(tn, 2));
    p->only_n = INTEGER(VECTOR_ELT(tn, 3));
    p->n_only_n = (int32_t)XLENGTH(VECTOR_ELT(tn, 3));
}

/* result list of the row-returning routines: per-unit vectors (length U) and the emitted rows as [n x E] matrices;
 * fills `o` with pointers into the freshly allocated R vectors.  The list is returned PROTECTed (+1), *ids is the
 * int64 staging of emit_unit (R has no 64-bit integers: converted to doubles by finish_rows). */
static SEXP alloc_rows(int n, int64_t U, int64_t E, ns_out *o, int64_t **ids)
{
    memset(o, 0, sizeof *o);
    const char *names[] = {"sigma", "slope", "max_gq", "flags", "emit_index", "n_emitted", "emit_unit", "pvalue",
                           "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs", "gt", "status", "pooled", ""};
    SEXP res = PROTECT(mkNamed(VECSXP, names));
    SEXP v;
#define VEC(i, type, len) SET_VECTOR_ELT(res, i, v = allocVector(type, len))
    VEC(0, REALSXP, U); o->sigma = REAL(v);
    VEC(1, REALSXP, U); o->slope = REAL(v);
    VEC(2, REALSXP, U); o->max_gq = REAL(v);
    VEC(3, INTSXP, U);  o->flags = (uint32_t *)INTEGER(v);
    VEC(4, INTSXP, U);  o->emit_index = INTEGER(v);
    VEC(6, REALSXP, E);
#undef VEC
    *ids = (int64_t *)R_alloc(E > 0 ? E : 1, sizeof(int64_t));
    o->emit_unit = *ids;
    o->emit_cap = E;
#define MAT(i, field) SET_VECTOR_ELT(res, i, v = allocMatrix(REALSXP, n, (int)E)); o->field = REAL(v)
    MAT(7, pvalue); MAT(8, qvalue); MAT(9, gq); MAT(10, qval_minaf); MAT(11, sor); MAT(12, rvsb); MAT(13, fs);
#undef MAT
    SET_VECTOR_ELT(res, 14, v = allocMatrix(RAWSXP, n, (int)E)); o->gt = RAW(v);
    SET_VECTOR_ELT(res, 15, v = allocMatrix(RAWSXP, n, (int)E)); o->status = RAW(v);
    SET_VECTOR_ELT(res, 16, v = allocMatrix(REALSXP, 8, (int)E)); o->pooled = REAL(v);
    return res;
}

static void finish_rows(SEXP res, const ns_out *o, const int64_t *ids)
{
    double *eu = REAL(VECTOR_ELT(res, 6));
    for (int64_t i = 0; i < o->emit_cap && i < o->n_emitted; i++)
        eu[i] = (double)ids[i];
    SET_VECTOR_ELT(res, 5, ScalarReal((double)o->n_emitted)); /* > emit_cap: call again with a larger cap */
}

/* One chunk of positions of the readcounts table.
 * counts: integer array dim = c(n, 9, P) (so that memory is [P][9][n]); ref: raw codes 0..3 (4 = skipped line).
 * Returns per-unit vectors and the emitted rows as [n x E] matrices. */
static SEXP call_snv_impl(SEXP ctx, SEXP counts, SEXP ref, SEXP params, SEXP tn, SEXP emit_cap, int multi)
{
    ns_ctx *c = multi ? NULL : get_ctx(ctx);
    ns_multi *m = multi ? (ns_multi *)R_ExternalPtrAddr(ctx) : NULL;
    if (multi && !m)
        error("needlestack-b200: context was destroyed");
    ns_params p;
    fill_params(params, &p, 1);
    fill_tn(tn, &p);
    SEXP dim = getAttrib(counts, R_DimSymbol);
    if (dim == R_NilValue || XLENGTH(dim) != 3 || INTEGER(dim)[1] != 9)
        error("counts must be an integer array of dim c(n, 9, P)");
    const int n = INTEGER(dim)[0];
    const int64_t P = INTEGER(dim)[2];
    if (TYPEOF(ref) != RAWSXP || XLENGTH(ref) != P)
        error("ref must be a raw vector of P base codes (0..3 = A,T,C,G)");
    ns_out o;
    int64_t *ids;
    SEXP res = alloc_rows(n, 3 * P, (int64_t)asReal(emit_cap), &o, &ids);
    int rc = multi ? ns_multi_call_snv(m, &p, n, P, RAW(ref), INTEGER(counts), &o)
                   : ns_call_snv(c, &p, n, P, RAW(ref), INTEGER(counts), &o);
    if (rc && rc != NS_ERR_EMIT_OVERFLOW)
        error("needlestack-b200: %s", multi ? ns_multi_last_error(m) : ns_last_error(c));
    finish_rows(res, &o, ids);
    UNPROTECT(1);
    return res;
}

SEXP C_ns_call_snv(SEXP ctx, SEXP counts, SEXP ref, SEXP params, SEXP tn, SEXP emit_cap)
{
    return call_snv_impl(ctx, counts, ref, params, tn, emit_cap, 0);
}

/* the same chunk over every device of a C_ns_create_multi() context (one queue of position chunks inside the library) */
SEXP C_ns_multi_call_snv(SEXP mctx, SEXP counts, SEXP ref, SEXP params, SEXP tn, SEXP emit_cap)
{
    return call_snv_impl(mctx, counts, ref, params, tn, emit_cap, 1);
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

/* The same units with the caller-level outputs of the SNV routine: rows (GQ, QVAL_minAF, strand bias, GT, STATUS, pooled
 * annotations) of the units the reference would write (emit_mode of `params`, default: the VCF calls), tumour/normal
 * pairs honoured.  What the indel blocks of bin/needlestack.r:461-567 / :624-730 need. */
SEXP C_ns_call_units_rows(SEXP ctx, SEXP dp, SEXP vp, SEXP vm, SEXP rp, SEXP rm, SEXP is_indel, SEXP params, SEXP tn)
{
    ns_ctx *c = get_ctx(ctx);
    ns_params p;
    fill_params(params, &p, 1);
    fill_tn(tn, &p);
    SEXP dim = getAttrib(dp, R_DimSymbol);
    if (dim == R_NilValue || XLENGTH(dim) != 2)
        error("dp must be an integer matrix of dim c(n, U)");
    const int n = INTEGER(dim)[0];
    const int64_t U = INTEGER(dim)[1];
    ns_out o;
    int64_t *ids;
    SEXP res = alloc_rows(n, U, U, &o, &ids);
#define OPT(x) ((x) == R_NilValue ? NULL : INTEGER(x))
    int rc = ns_call_units(c, &p, n, U, INTEGER(dp), INTEGER(vp), OPT(vm), OPT(rp), OPT(rm),
                           is_indel == R_NilValue ? NULL : RAW(is_indel), 0, &o);
#undef OPT
    if (rc)
        error("needlestack-b200: %s", ns_last_error(c));
    finish_rows(res, &o, ids);
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
                                               {"C_ns_call_units_rows", (DL_FUNC)&C_ns_call_units_rows, 9},
                                               {"C_ns_qvalue_grid", (DL_FUNC)&C_ns_qvalue_grid, 7},
                                               {"C_ns_create_multi", (DL_FUNC)&C_ns_create_multi, 1},
                                               {"C_ns_multi_call_snv", (DL_FUNC)&C_ns_multi_call_snv, 6},
                                               {NULL, NULL, 0}};

void R_init_ns_rshim(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
