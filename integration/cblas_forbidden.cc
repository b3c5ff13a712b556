// cblas_forbidden.cc -- linked into the GPU build of the reference binary: the reference's CPU GEMM
// must never run there (the acoustic model is evaluated by libce_gpu.so), so reaching it aborts.
#include <cblas.h>
#include <stdio.h>
#include <stdlib.h>

extern "C" void cblas_sgemm(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE, const enum CBLAS_TRANSPOSE,
                            const int, const int, const int, const float, const float *, const int,
                            const float *, const int, const float, float *, const int) {
  fprintf(stderr, "cblas_sgemm was called in the GPU build: the CPU acoustic model must not run\n");
  abort();
}
