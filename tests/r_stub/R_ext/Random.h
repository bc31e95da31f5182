/* stub of <R_ext/Random.h>: see ../Rinternals.h */
#ifndef R_EXT_RANDOM_H_
#define R_EXT_RANDOM_H_
#ifdef __cplusplus
extern "C" {
#endif
void GetRNGstate(void);
void PutRNGstate(void);
#ifdef __cplusplus
}
#endif
#endif
