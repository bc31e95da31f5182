/* stub of <R_ext/Print.h>: see ../Rinternals.h */
#ifndef R_EXT_PRINT_H_
#define R_EXT_PRINT_H_
#ifdef __cplusplus
extern "C" {
#endif
void Rprintf(const char*, ...);
#ifdef __cplusplus
}
#endif
#endif
