/* MOCK of Rinternals.h: just enough declarations to type-check r/shim.c without R installed. */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
#define REALSXP 14
#define INTSXP 13
#define VECSXP 19
#define STRSXP 16
extern SEXP R_NamesSymbol;
int LENGTH(SEXP);
R_xlen_t XLENGTH(SEXP);
double *REAL(SEXP);
int *INTEGER(SEXP);
SEXP Rf_allocMatrix(int, int, int);
SEXP Rf_allocVector(int, R_xlen_t);
SEXP Rf_duplicate(SEXP);
SEXP Rf_mkChar(const char *);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
double Rf_asReal(SEXP);
int Rf_asInteger(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char *, ...) __attribute__((noreturn));
#endif
