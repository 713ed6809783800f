/* Minimal stand-in for R's <Rinternals.h>/<R_ext/Rdynload.h>: ONLY for compile-checking
 * r/rsoptsc_shim.c in an image without R.  Declares the documented R C API entry points the
 * shim uses, with R's signatures.  Never shipped; a real build includes R's own headers. */
#ifndef RSOPTSC_STUB_RINTERNALS_H
#define RSOPTSC_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define RAWSXP 24
extern SEXP R_NilValue, R_DimSymbol, R_NamesSymbol;
double* REAL(SEXP x);
int* INTEGER(SEXP x);
unsigned char* RAW(SEXP x);
SEXP Rf_getAttrib(SEXP x, SEXP name);
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val);
SEXP Rf_allocVector(unsigned int type, R_xlen_t n);
SEXP Rf_allocMatrix(unsigned int type, int nrow, int ncol);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
SEXP Rf_mkChar(const char* s);
SEXP Rf_ScalarReal(double x);
SEXP Rf_ScalarInteger(int x);
void SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
R_xlen_t Rf_xlength(SEXP x);
int Rf_isNull(SEXP x);
int Rf_isReal(SEXP x);
int Rf_isInteger(SEXP x);
int Rf_isMatrix(SEXP x);
double Rf_asReal(SEXP x);
int Rf_asInteger(SEXP x);
void Rf_error(const char* fmt, ...);
void R_CheckUserInterrupt(void);
Rboolean R_ToplevelExec(void (*fun)(void*), void* data);
void Rprintf(const char* fmt, ...);
char* R_alloc(size_t n, int size);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
int R_registerRoutines(DllInfo* info, const void* c, const R_CallMethodDef* call, const void* f, const void* e);
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value);
#endif
