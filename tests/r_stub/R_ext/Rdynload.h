/* stub of <R_ext/Rdynload.h>: see ../Rinternals.h */
#ifndef R_EXT_DYNLOAD_H_
#define R_EXT_DYNLOAD_H_
#include "../Rinternals.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
int R_registerRoutines(DllInfo* info, const void* cMethods, const R_CallMethodDef* callMethods, const void* fortranMethods, const void* externalMethods);
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value);
void R_RegisterCCallable(const char* package, const char* name, DL_FUNC fptr);
DL_FUNC R_GetCCallable(const char* package, const char* name);
#ifdef __cplusplus
}
#endif
#endif
