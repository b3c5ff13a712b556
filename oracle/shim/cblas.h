/* Minimal <cblas.h> stand-in for building the reference (src/matrix.cc:9 includes
 * <cblas.h>; OpenBLAS headers are not installed in this image).  TEST
 * INFRASTRUCTURE ONLY: declares the one entry point the reference calls
 * (src/matrix.cc:308-322).  The definition lives in oracle/ref_shim.cc and
 * dispatches to a deterministic in-order fp32 loop or to a dlopen()ed OpenBLAS. */
#ifndef CE_ORACLE_SHIM_CBLAS_H_
#define CE_ORACLE_SHIM_CBLAS_H_

#ifdef __cplusplus
extern "C" {
#endif

enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };

void cblas_sgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta,
                 enum CBLAS_TRANSPOSE tb, int m, int n, int k, float alpha,
                 const float *a, int lda, const float *b, int ldb, float beta,
                 float *c, int ldc);

#ifdef __cplusplus
}
#endif
#endif
