/*
 * tests/rstub/rstub.c — TEST INFRASTRUCTURE ONLY: implementation of the stub R API of Rinternals.h plus the harness
 * entry points tests/test_rshim_stub.py calls through ctypes (constructors / accessors of stub SEXPs and
 * rstub_call(), which runs one registered .Call routine under setjmp so that error() comes back as NULL + message).
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

static struct SEXPREC nil_rec = {NILSXP, 0, 0, 0, 0, 0}, names_rec = {NILSXP, 0, 0, 0, 0, 0}, dim_rec = {NILSXP, 0, 0, 0, 0, 0};
SEXP R_NilValue = &nil_rec, R_NamesSymbol = &names_rec, R_DimSymbol = &dim_rec;
double R_NaReal;
int rstub_protect_depth = 0;
jmp_buf rstub_jmp;
char rstub_errmsg[1024];
static DllInfo g_dll;

static size_t elt(int type)
{
    return type == INTSXP || type == LGLSXP ? sizeof(int) : type == REALSXP ? sizeof(double) : type == RAWSXP || type == CHARSXP ? 1
           : sizeof(SEXP);
}
SEXP Rf_allocVector(int type, R_xlen_t n)
{
    SEXP s = (SEXP)calloc(1, sizeof(*s));
    s->type = type;
    s->len = n;
    s->data = calloc((size_t)(n > 0 ? n : 1) + (type == CHARSXP), elt(type));
    s->names = s->dim = R_NilValue;
    if (type == VECSXP || type == STRSXP)
        for (R_xlen_t i = 0; i < n; i++)
            ((SEXP *)s->data)[i] = R_NilValue;
    return s;
}
SEXP Rf_allocMatrix(int type, int nrow, int ncol)
{
    SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    s->dim = Rf_allocVector(INTSXP, 2);
    INTEGER(s->dim)[0] = nrow;
    INTEGER(s->dim)[1] = ncol;
    return s;
}
static SEXP mkchar(const char *t)
{
    SEXP c = Rf_allocVector(CHARSXP, (R_xlen_t)strlen(t));
    strcpy((char *)c->data, t);
    return c;
}
SEXP Rf_mkNamed(int type, const char **names)
{
    R_xlen_t n = 0;
    while (names[n][0])
        n++;
    SEXP s = Rf_allocVector(type, n);
    s->names = Rf_allocVector(STRSXP, n);
    for (R_xlen_t i = 0; i < n; i++)
        ((SEXP *)s->names->data)[i] = mkchar(names[i]);
    return s;
}
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); INTEGER(s)[0] = v != 0; return s; }
SEXP Rf_getAttrib(SEXP x, SEXP what) { return what == R_NamesSymbol ? x->names : what == R_DimSymbol ? x->dim : R_NilValue; }
int Rf_asInteger(SEXP x) { return x->type == REALSXP ? (int)REAL(x)[0] : INTEGER(x)[0]; }
double Rf_asReal(SEXP x) { return x->type == REALSXP ? REAL(x)[0] : (double)INTEGER(x)[0]; }
void Rf_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(rstub_errmsg, sizeof rstub_errmsg, fmt, ap);
    va_end(ap);
    longjmp(rstub_jmp, 1);
}
SEXP Rf_protect(SEXP x) { rstub_protect_depth++; return x; }
void Rf_unprotect(int n) { rstub_protect_depth -= n; }
char *R_alloc(size_t n, int size) { return (char *)calloc(n ? n : 1, (size_t)size); }
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot) { (void)tag; (void)prot; SEXP s = Rf_allocVector(EXTPTRSXP, 0); free(s->data); s->data = p; return s; }
void *R_ExternalPtrAddr(SEXP x) { return x->data; }
void R_ClearExternalPtr(SEXP x) { x->data = NULL; }
void R_RegisterCFinalizerEx(SEXP x, void (*fin)(SEXP), Rboolean onexit) { (void)onexit; x->fin = fin; }
int R_registerRoutines(DllInfo *info, const void *c, const R_CallMethodDef *call, const void *f, const void *e)
{
    (void)c; (void)f; (void)e;
    info->call = call;
    info->n_call = 0;
    while (call && call[info->n_call].name)
        info->n_call++;
    return 1;
}
int R_useDynamicSymbols(DllInfo *info, int value) { info->dynamic = value; return 1; }

/* ---- harness ------------------------------------------------------------------------------ */
void R_init_ns_rshim(DllInfo *dll);
int rstub_init(void) { R_NaReal = NAN; memset(&g_dll, 0, sizeof g_dll); g_dll.dynamic = 1; R_init_ns_rshim(&g_dll); return g_dll.n_call; }
int rstub_dynamic_symbols(void) { return g_dll.dynamic; }
const char *rstub_routine_name(int i) { return i < g_dll.n_call ? g_dll.call[i].name : NULL; }
int rstub_routine_nargs(int i) { return i < g_dll.n_call ? g_dll.call[i].numArgs : -1; }
const char *rstub_last_error(void) { return rstub_errmsg; }
int rstub_depth(void) { return rstub_protect_depth; }
SEXP rstub_nil(void) { return R_NilValue; }
SEXP rstub_ints(const int *v, R_xlen_t n) { SEXP s = Rf_allocVector(INTSXP, n); memcpy(s->data, v, (size_t)n * sizeof(int)); return s; }
SEXP rstub_reals(const double *v, R_xlen_t n) { SEXP s = Rf_allocVector(REALSXP, n); memcpy(s->data, v, (size_t)n * sizeof(double)); return s; }
SEXP rstub_raws(const unsigned char *v, R_xlen_t n) { SEXP s = Rf_allocVector(RAWSXP, n); memcpy(s->data, v, (size_t)n); return s; }
void rstub_set_dim(SEXP x, const int *d, int nd) { x->dim = rstub_ints(d, nd); }
SEXP rstub_list(R_xlen_t n) { SEXP s = Rf_allocVector(VECSXP, n); s->names = Rf_allocVector(STRSXP, n); return s; }
void rstub_list_set(SEXP l, R_xlen_t i, const char *name, SEXP v) { ((SEXP *)l->data)[i] = v; ((SEXP *)l->names->data)[i] = mkchar(name ? name : ""); }
SEXP rstub_list_get(SEXP l, const char *name)
{
    for (R_xlen_t i = 0; i < l->len; i++)
        if (!strcmp(CHAR(((SEXP *)l->names->data)[i]), name))
            return ((SEXP *)l->data)[i];
    return NULL;
}
int rstub_type(SEXP x) { return x->type; }
R_xlen_t rstub_len(SEXP x) { return x->len; }
void *rstub_data(SEXP x) { return x->data; }
int rstub_dim(SEXP x, int k) { return x->dim == R_NilValue || k >= x->dim->len ? -1 : INTEGER(x->dim)[k]; }
void rstub_finalize(SEXP x) { if (x->fin) x->fin(x); }
/* .Call(name, args...): looks the routine up in the registered table, checks the argument count, runs it */
SEXP rstub_call(const char *name, int nargs, SEXP *a)
{
    for (int i = 0; i < g_dll.n_call; i++)
        if (!strcmp(g_dll.call[i].name, name)) {
            if (g_dll.call[i].numArgs != nargs) {
                snprintf(rstub_errmsg, sizeof rstub_errmsg, "Incorrect number of arguments (%d), expecting %d for '%s'", nargs,
                         g_dll.call[i].numArgs, name);
                return NULL;
            }
            rstub_errmsg[0] = 0;
            const int depth0 = rstub_protect_depth;
            if (setjmp(rstub_jmp)) {
                rstub_protect_depth = depth0; /* R unwinds the protect stack on error */
                return NULL;
            }
            DL_FUNC f = g_dll.call[i].fun;
            switch (nargs) {
            case 1: return ((SEXP(*)(SEXP))f)(a[0]);
            case 4: return ((SEXP(*)(SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3]);
            case 6: return ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4], a[5]);
            case 7: return ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]);
            case 8: return ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
            case 9: return ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]);
            default: snprintf(rstub_errmsg, sizeof rstub_errmsg, "harness: unsupported arity %d", nargs); return NULL;
            }
        }
    snprintf(rstub_errmsg, sizeof rstub_errmsg, "C symbol name \"%s\" not in load table", name);
    return NULL;
}
