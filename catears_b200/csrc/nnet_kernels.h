// nnet_kernels.h -- launchers of the memory-bound kernels around the GEMMs (nnet_kernels.cu).
#ifndef CE_GPU_NNET_KERNELS_H_
#define CE_GPU_NNET_KERNELS_H_

#include "gemm.h"

namespace ce {

// Which rows of an utterance block take part in FindMinMax: rows in [lo, P - hi) that the
// consuming layer's Splice+Narrow reads (matters for very short utterances).
struct RowUse {
  int32_t lo, hi;
  int32_t next_n_taps;               // 0 = every row in range
  int32_t next_tap_off[kMaxTaps];
  int32_t next_lo, next_hi;
};

int InitMinMaxLaunch(uint32_t *mm, int n_pairs, cudaStream_t s);
int MinMaxLaunch(const float *x, int64_t ld, int C, int M, const int32_t *tile_utt,
                 const UttRows *utts, const RowUse &use, uint32_t *minmax, cudaStream_t s);
int QParamsLaunch(const uint32_t *minmax, QParam *q, int n, cudaStream_t s);
// Quantize of x [M x ld_in] (C columns) into q [M x c_pad] (+ row sums of the codes).  With
// `minmax` ([n_utts][2], the reduced FindMinMax) the parameters are computed in the same launch and
// written to qp; with minmax == nullptr qp is the input.
int QuantizeLaunch(const float *x, int64_t ld_in, int C, int M, int c_pad, const int32_t *tile_utt,
                   const uint32_t *minmax, int n_utts, QParam *qp, uint8_t *q, int32_t *rowsum,
                   cudaStream_t s);
// Counts disagreements between the production quantiser arithmetic and the plain IEEE form over n
// pseudo-random (value, scale, zero point) triples (adds to *mismatches_dev).
int QuantSelfTestLaunch(int64_t n, uint64_t seed, unsigned long long *mismatches_dev, cudaStream_t s);
int ConvertLaunch(const float *x, int64_t ld_in, int C, int64_t M, int c_pad,
                  __nv_bfloat16 *out_bf16, float *out_hi, float *out_lo, cudaStream_t s,
                  __nv_bfloat16 *out_x3 = nullptr);
// What a finished log-likelihood row is written as (include/ce_gpu.h, ce_gpu_model_set_output):
// all N columns, the columns ids[0..n) (device array), or the n best (loglik, pdf) pairs.
enum { kOutDense = 0, kOutSubset = 1, kOutTopK = 2 };
struct OutSel {
  int mode = kOutDense;
  int n = 0;
  const int32_t *ids = nullptr;
};
// Largest n of a top-n selection (the sort runs in shared memory, 8 bytes an entry per warp).
constexpr int kMaxTopK = 1024;
// `loglik` rows are ld_out 4-byte words apart: N floats, n floats, or n (float, int32) pairs.
int FinalizeLaunch(const float *logits, int64_t ld, int N, int M, const int32_t *tile_utt,
                   const UttRows *utts, const int64_t *out_row_off, int left, int right,
                   bool log_softmax, const float *log_prior, float *loglik, int64_t ld_out,
                   int32_t *argmax, cudaStream_t s, const OutSel &sel = OutSel(),
                   const float *lse_in = nullptr);   // [M] log-sum-exp of every row, already reduced (selecting kernels)

}  // namespace ce
#endif  // CE_GPU_NNET_KERNELS_H_
