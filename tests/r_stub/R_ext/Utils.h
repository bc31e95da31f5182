/* stub of <R_ext/Utils.h>: see ../Rinternals.h */
#ifndef R_EXT_UTILS_H_
#define R_EXT_UTILS_H_
#ifdef __cplusplus
extern "C" {
#endif
void R_CheckUserInterrupt(void);
#ifdef __cplusplus
}
#endif
#endif
