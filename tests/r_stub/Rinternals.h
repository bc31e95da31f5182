/* tests/r_stub/Rinternals.h -- TEST INFRASTRUCTURE.  Declarations (only) of the part of R's C API that glue/init_gpu.cpp and the
 * R branch of glue/gpubart_shim.cpp use, with the types and signatures of R >= 4.0's <Rinternals.h>, so that the glue can be
 * type-checked with `g++ -fsyntax-only` in an image without R.  Nothing here is ever linked or run. */
#ifndef R_INTERNALS_H_
#define R_INTERNALS_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef unsigned int SEXPTYPE;

#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define RAWSXP 24

extern SEXP R_NilValue, R_GlobalEnv, R_NamesSymbol, R_DimSymbol, R_DimNamesSymbol, R_ClassSymbol, R_RowNamesSymbol;
extern int R_NaInt;
extern double R_NaReal;
#define NA_INTEGER R_NaInt
#define NA_LOGICAL R_NaInt
#define NA_REAL R_NaReal
int R_IsNA(double);
int R_IsNaN(double);
#define ISNA(x) R_IsNA(x)
#define ISNAN(x) (R_IsNA(x) || R_IsNaN(x))

SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void R_PreserveObject(SEXP);
void R_ReleaseObject(SEXP);

SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_coerceVector(SEXP, SEXPTYPE);
SEXP Rf_mkChar(const char*);
SEXP Rf_mkString(const char*);
SEXP Rf_ScalarLogical(int);
SEXP Rf_install(const char*);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_lang4(SEXP, SEXP, SEXP, SEXP);
SEXP Rf_eval(SEXP, SEXP);
SEXP R_do_slot(SEXP obj, SEXP name);
Rboolean Rf_inherits(SEXP, const char*);
Rboolean Rf_isNull(SEXP);
Rboolean Rf_isReal(SEXP);
Rboolean Rf_isInteger(SEXP);
Rboolean Rf_isString(SEXP);
Rboolean Rf_isFunction(SEXP);
Rboolean Rf_isEnvironment(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
R_xlen_t XLENGTH(SEXP);
int* INTEGER(SEXP);
double* REAL(SEXP);
Rbyte* RAW(SEXP);
const char* R_CHAR(SEXP);
#define CHAR(x) R_CHAR(x)
SEXP STRING_ELT(SEXP, R_xlen_t);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);

typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean onexit);

#ifdef __cplusplus
[[noreturn]]
#endif
void Rf_error(const char*, ...);
double unif_rand(void);

#ifdef __cplusplus
}
#endif
#endif
