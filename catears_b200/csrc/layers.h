// layers.h -- launchers of the stand-alone layer kernels (layers.cu) used by the general layer
// program: SpliceLayer with its clamp, and the row-wise ReLU / BatchNorm / Normalize / Softmax /
// LogSoftmax layers of src/nnet.cc.
#ifndef CE_GPU_LAYERS_H_
#define CE_GPU_LAYERS_H_

#include "gemm.h"

namespace ce {

enum RowOp { kRowReLU = 0, kRowBatchNorm, kRowNormalize, kRowSoftmax, kRowLogSoftmax };

// out[row][t * C + c] = in[clamp(row + idx[t])][c] for the valid rows [lo, P - hi) of every
// utterance block, clamped to that same range (src/nnet.cc:50-75).
int SpliceLaunch(const float *in, int64_t ld_in, int C, int M, const int32_t *tile_utt,
                 const UttRows *utts, int lo, int hi, const int32_t *idx_dev, int n_idx, float *out,
                 int64_t ld_out, cudaStream_t s);

// In place over the valid rows.  scale / offset: BatchNorm only.
int RowwiseLaunch(int op, float *x, int64_t ld, int C, int M, const int32_t *tile_utt,
                  const UttRows *utts, int lo, int hi, const float *scale_dev,
                  const float *offset_dev, cudaStream_t s);

}  // namespace ce
#endif  // CE_GPU_LAYERS_H_
