/* minimal stand-in for <Rinternals.h>: declarations only, enough to type-check the rglue C sources (R is not installed here) */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
typedef struct SEXPREC* SEXP;
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
#define EXTPTRSXP 22
extern SEXP R_NilValue;
SEXP Rf_allocVector(unsigned int, long);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
double* REAL(SEXP);
int* INTEGER(SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_length(SEXP);
Rboolean Rf_isNull(SEXP);
Rboolean Rf_isMatrix(SEXP);
SEXP VECTOR_ELT(SEXP, long);
SEXP SET_VECTOR_ELT(SEXP, long, SEXP);
void Rf_error(const char*, ...) __attribute__((noreturn));
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
SEXP R_ExternalPtrProtected(SEXP);
int TYPEOF(SEXP);
SEXP Rf_coerceVector(SEXP, unsigned int);
SEXP Rf_ScalarLogical(int);
SEXP Rf_ScalarInteger(int);
char* R_alloc(unsigned long, int);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif
