/*
 * tests/rstub/Rinternals.h — TEST INFRASTRUCTURE ONLY: the handful of R API entry points that
 * needlestack_b200/r/ns_rshim.c uses, re-implemented on a plain C struct so that the .Call shim can be compiled and
 * DRIVEN on a machine without R (tests/test_rshim_stub.py).  Semantics kept: vectors with length / names / dim
 * attributes, PROTECT depth accounting (the harness checks it returns to zero), error() does not return
 * (longjmp to the harness), external pointers with finalizers.  Nothing here is part of the product.
 */
#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include <setjmp.h>
#include <stddef.h>
#include <stdint.h>

#define NILSXP 0
#define CHARSXP 9
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define EXTPTRSXP 22
#define RAWSXP 24

typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
typedef struct SEXPREC {
    int type;
    R_xlen_t len;
    void *data; /* int / double / Rbyte / SEXP[] / char[] */
    struct SEXPREC *names, *dim;
    void (*fin)(struct SEXPREC *);
} *SEXP;
typedef enum { FALSE = 0, TRUE = 1 } Rboolean;

extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
extern double R_NaReal;
#define NA_REAL R_NaReal
extern int rstub_protect_depth;
extern jmp_buf rstub_jmp;
extern char rstub_errmsg[1024];

SEXP Rf_allocVector(int type, R_xlen_t n);
SEXP Rf_allocMatrix(int type, int nrow, int ncol);
SEXP Rf_mkNamed(int type, const char **names);
SEXP Rf_ScalarReal(double v);
SEXP Rf_ScalarInteger(int v);
SEXP Rf_ScalarLogical(int v);
SEXP Rf_getAttrib(SEXP x, SEXP what);
int Rf_asInteger(SEXP x);
double Rf_asReal(SEXP x);
void Rf_error(const char *fmt, ...) __attribute__((noreturn));
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
char *R_alloc(size_t n, int size);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP x);
void R_ClearExternalPtr(SEXP x);
void R_RegisterCFinalizerEx(SEXP x, void (*fin)(SEXP), Rboolean onexit);

#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define mkNamed Rf_mkNamed
#define ScalarReal Rf_ScalarReal
#define ScalarInteger Rf_ScalarInteger
#define ScalarLogical Rf_ScalarLogical
#define getAttrib Rf_getAttrib
#define asInteger Rf_asInteger
#define asReal Rf_asReal
#define error Rf_error
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
#define INTEGER(x) ((int *)(x)->data)
#define REAL(x) ((double *)(x)->data)
#define RAW(x) ((Rbyte *)(x)->data)
#define XLENGTH(x) ((x)->len)
#define TYPEOF(x) ((x)->type)
#define VECTOR_ELT(x, i) (((SEXP *)(x)->data)[i])
#define SET_VECTOR_ELT(x, i, v) (((SEXP *)(x)->data)[i] = (v))
#define STRING_ELT(x, i) (((SEXP *)(x)->data)[i])
#define CHAR(x) ((const char *)(x)->data)
#endif
