// layers.cu -- the layers of src/nnet.cc that do not fold into a tensor-core block, as stand-alone
// row-wise kernels over the padded row space (rows are never compacted: a NarrowLayer only shrinks
// the valid range [lo, P - hi) of every utterance block):
//
//   splice_kernel      SpliceLayer::Propagate     src/nnet.cc:50-75    (gather with clamp to the
//                                                                      matrix the reference holds =
//                                                                      the valid rows of the block)
//   rowwise_kernel     ReLULayer                  src/nnet.cc:149-160
//                      BatchNormLayer             src/nnet.cc:106-117  (x * scale, then + offset)
//                      NormalizeLayer             src/nnet.cc:162-175  (x *= (float)sqrt(D / sum x^2))
//                      SoftmaxLayer               src/nnet.cc:126-135 -> ApplySoftMax src/vector.cc:95-107
//                      LogSoftmaxLayer            src/nnet.cc:137-146 -> ApplyLogSoftMax src/vector.cc:110-122
//
// These serve layer stacks tool/convert_am.py never emits (nnet2-style Normalize / Softmax, a
// Splice without its Narrow): correctness first, one warp per row, fp32 in place.
#include "layers.h"

#include <float.h>

#include <algorithm>

namespace ce {
namespace {

// position of `row` inside its utterance block and the block's geometry; false for rows outside
// the valid range
__device__ __forceinline__ bool ValidRow(int row, const int32_t *__restrict__ tile_utt,
                                         const UttRows *__restrict__ utts, int lo, int hi, int *pos,
                                         UttRows *ur) {
  const int utt = tile_utt[row / kRowGran];
  *ur = utts[utt];
  *pos = row - ur->row_off;
  return *pos >= lo && *pos < ur->rows - hi;
}

__global__ void __launch_bounds__(256)
splice_kernel(const float *__restrict__ in, int64_t ld_in, int C, int M,
              const int32_t *__restrict__ tile_utt, const UttRows *__restrict__ utts, int lo, int hi,
              const int32_t *__restrict__ idx, int n_idx, float *__restrict__ out, int64_t ld_out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  int pos;
  UttRows ur;
  if (!ValidRow(row, tile_utt, utts, lo, hi, &pos, &ur)) return;
  const int first = lo, last = ur.rows - hi - 1;         // the reference's rows 0 .. NumRows() - 1
  float *o = out + (int64_t)row * ld_out;
  for (int t = 0; t < n_idx; ++t) {
    int src = pos + idx[t];                              // nnet.cc:64-66
    src = src < first ? first : (src > last ? last : src);
    const float *r = in + (int64_t)(ur.row_off + src) * ld_in;
    for (int c = lane; c < C; c += 32) o[(int64_t)t * C + c] = r[c];
  }
}

__device__ __forceinline__ float WarpSum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float WarpMax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(256)
rowwise_kernel(int op, float *__restrict__ x, int64_t ld, int C, int M,
               const int32_t *__restrict__ tile_utt, const UttRows *__restrict__ utts, int lo, int hi,
               const float *__restrict__ scale, const float *__restrict__ offset) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  int pos;
  UttRows ur;
  if (!ValidRow(row, tile_utt, utts, lo, hi, &pos, &ur)) return;
  float *r = x + (int64_t)row * ld;
  switch (op) {
    case kRowReLU:
      for (int c = lane; c < C; c += 32)
        if (r[c] < 0.0f) r[c] = 0.0f;                    // nnet.cc:156
      break;
    case kRowBatchNorm:
      for (int c = lane; c < C; c += 32)
        r[c] = __fadd_rn(__fmul_rn(r[c], scale[c]), offset[c]);   // nnet.cc:114-115
      break;
    case kRowNormalize: {
      float ss = 0.0f;
      for (int c = lane; c < C; c += 32) ss = fmaf(r[c], r[c], ss);
      ss = WarpSum(ss);                                  // VecVec, src/vector.cc:82-92 (fp32 sum)
      const float sc = (float)sqrt((double)(float)C / (double)ss);   // nnet.cc:171-172
      for (int c = lane; c < C; c += 32) r[c] = __fmul_rn(r[c], sc);
      break;
    }
    case kRowSoftmax: {
      // exp(x) / sum exp(x) (vector.cc:95-107), evaluated with the row maximum taken out: the same
      // value wherever the reference's own sum is finite
      float m = -FLT_MAX;
      for (int c = lane; c < C; c += 32) m = fmaxf(m, r[c]);
      m = WarpMax(m);
      float s = 0.0f;
      for (int c = lane; c < C; c += 32) s += expf(r[c] - m);
      s = WarpSum(s);
      for (int c = lane; c < C; c += 32) r[c] = __fdiv_rn(expf(r[c] - m), s);
      break;
    }
    case kRowLogSoftmax: {
      float m = -FLT_MAX;
      for (int c = lane; c < C; c += 32) m = fmaxf(m, r[c]);
      m = WarpMax(m);
      float s = 0.0f;
      for (int c = lane; c < C; c += 32) s += expf(r[c] - m);
      s = WarpSum(s);
      const float lse = m + logf(s);
      for (int c = lane; c < C; c += 32) r[c] = __fsub_rn(r[c], lse);   // vector.cc:120
      break;
    }
  }
}

}  // namespace

int SpliceLaunch(const float *in, int64_t ld_in, int C, int M, const int32_t *tile_utt,
                 const UttRows *utts, int lo, int hi, const int32_t *idx_dev, int n_idx, float *out,
                 int64_t ld_out, cudaStream_t s) {
  if (M <= 0) return CE_GPU_OK;
  ProfScope prof(kProfOther, s);
  splice_kernel<<<(M + 7) / 8, 256, 0, s>>>(in, ld_in, C, M, tile_utt, utts, lo, hi, idx_dev, n_idx, out,
                                           ld_out);
  CE_LAUNCHED();
  return CE_GPU_OK;
}

int RowwiseLaunch(int op, float *x, int64_t ld, int C, int M, const int32_t *tile_utt,
                  const UttRows *utts, int lo, int hi, const float *scale_dev,
                  const float *offset_dev, cudaStream_t s) {
  if (M <= 0) return CE_GPU_OK;
  ProfScope prof(kProfOther, s);
  rowwise_kernel<<<(M + 7) / 8, 256, 0, s>>>(op, x, ld, C, M, tile_utt, utts, lo, hi, scale_dev,
                                            offset_dev);
  CE_LAUNCHED();
  return CE_GPU_OK;
}

}  // namespace ce
