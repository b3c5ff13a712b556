// cblas_inorder.cc -- TEST INFRASTRUCTURE: the one external symbol the reference needs
// (cblas_sgemm, src/matrix.cc:308-322; OpenBLAS is not vendored) as a plain in-order fp32 loop,
// for the CPU build of the reference's own binary.  The GPU build links cblas_forbidden.cc instead.
#include <cblas.h>

extern "C" void cblas_sgemm(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE ta, const enum CBLAS_TRANSPOSE tb,
                            const int m, const int n, const int k, const float alpha, const float *a,
                            const int lda, const float *b, const int ldb, const float beta, float *c,
                            const int ldc) {
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < n; ++j) {
      float s = 0.0f;
      for (int l = 0; l < k; ++l) {
        const float av = ta == CblasNoTrans ? a[(long)i * lda + l] : a[(long)l * lda + i];
        const float bv = tb == CblasNoTrans ? b[(long)l * ldb + j] : b[(long)j * ldb + l];
        s += av * bv;
      }
      c[(long)i * ldc + j] = alpha * s + (beta == 0.0f ? 0.0f : beta * c[(long)i * ldc + j]);
    }
  }
}
