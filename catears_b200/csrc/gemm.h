// gemm.h -- the tcgen05 implicit-GEMM of the acoustic model's Linear layers (K3/K4).
#ifndef CE_GPU_GEMM_H_
#define CE_GPU_GEMM_H_

#include <cuda.h>
#include <cuda_bf16.h>

#include "common.h"

namespace ce {

// kKindBF16X3: bf16 hi/lo split operands, three products per K element (hi*hi + hi*lo + lo*hi):
// a 16-bit mantissa at the bf16 tensor rate.  Both operands are stored INTERLEAVED per 128-byte
// K atom -- [32 hi | 32 lo] bf16 for 32 consecutive channels -- so one staged atom feeds all three
// products (six K=16 MMAs) and the operand traffic is 2x, not 3x, that of plain bf16.
enum GemmKind { kKindI8 = 0, kKindBF16 = 1, kKindTF32 = 2, kKindBF16X3 = 3 };

constexpr int kTileM = 128;          // rows per CTA tile = UMMA M (cta_group::1)
constexpr int kTileN = 256;          // columns per CTA tile = UMMA N
constexpr int kTileKBytes = 128;     // one 128-byte swizzle atom of K per pipeline stage
constexpr int kMaxTaps = 8;
constexpr int kRowGran = 32;         // rows per entry of the row -> utterance table (one epilogue warp)

inline int KindEltBytes(int kind) { return kind == kKindI8 ? 1 : (kind == kKindBF16 || kind == kKindBF16X3) ? 2 : 4; }
// logical K elements (channels) per 128-byte atom
inline int KindTileK(int kind) { return kind == kKindBF16X3 ? 32 : kTileKBytes / KindEltBytes(kind); }
// stored elements per row for `c` (padded) channels: the hi/lo interleave doubles them
inline int KindPhysCols(int kind, int c) { return kind == kKindBF16X3 ? 2 * c : c; }

// Per-utterance affine u8 quantisation parameters (struct QuantizationParams, src/matrix.h:231-234).
struct QParam {
  float scale;
  int32_t zero_point;
};

// Geometry of one utterance inside the activation row space.  Every utterance owns a block of
// rows starting at a multiple of kTileM, so that a GEMM tile never spans two utterances -- or, for
// batches of short blocks (the micro-batches of live streams), at a multiple of kRowGran, the 32
// accumulator rows one epilogue warp drains (GemmArgs::gran).
struct UttRows {
  int32_t row_off;   // first row of the block
  int32_t rows;      // P = T + left_context + right_context
};

// Device-side arguments of one GEMM launch.  A is an activation matrix [M x c_pad] (row-major,
// K contiguous); the B operand is the packed weight matrix [N x (n_taps * c_pad)] (K contiguous).
//   out[m][n] = epilogue( sum_{pass} sum_{tap} sum_{c} A_pass[m + tap_off[tap]][c] *
//                                                       B_pass[n][tap * c_pad + c] )
// Rows outside [0, M) read as zero (TMA out-of-bounds fill).
struct GemmArgs {
  int32_t M, N;
  int32_t c_pad;                 // STORED elements per tap and row: multiple of 128 bytes (KindPhysCols)
  int32_t n_taps;
  int32_t tap_off[kMaxTaps];
  int32_t n_pass;                // 1, or 3 for the error-compensated 3xTF32 product
  int32_t pass_a[3], pass_b[3];  // operand selectors (0 = hi / only, 1 = lo)

  // ---- epilogue ----
  const float *bias;             // [N padded to kTileN], nullptr = none          nnet.cc:34
  const float *bn_scale;         // nullptr = no BatchNorm                        nnet.cc:114
  const float *bn_offset;        //                                               nnet.cc:115
  int32_t relu;                  //                                               nnet.cc:156
  // u8 x u8 -> s32 (gemmlowp contract)
  const int32_t *a_rowsum;       // [M] sum of the u8 codes of every A row (one tap)
  const int32_t *b_colsum;       // [N padded] sum over K of the u8 weight codes
  int32_t zp_b;
  float scale_b;
  int32_t k_true;                // un-padded K = n_taps * C
  const QParam *qa;              // [n_utts] activation quantisation of each utterance
  const int32_t *tile_utt;       // [ceil(M / kRowGran)] utterance of each 32-row granule; nullptr = 0
  int32_t gran;                  // 0: blocks start at multiples of kTileM (one utterance per tile);
                                 // 1: at multiples of kRowGran (int8: per-warp parameters)
  const UttRows *utts;           // [n_utts]; nullptr = one block of M rows

  // outputs (any may be nullptr)
  float *out_f32;                // fp32 result (for 3xTF32 consumers: the tf32-rounded "hi" part)
  float *out_lo;                 // v - hi
  __nv_bfloat16 *out_bf16;
  int32_t *out_acc;              // int32 accumulators after the zero-point corrections
  int64_t ld_out;                // row stride (elements) of out_f32 / out_lo / out_bf16 / out_acc
  int32_t n_store;               // columns written per row (>= N; columns [N, n_store) get 0) --
                                 // the zero padding of the next layer's K dimension.  kKindBF16X3 with
                                 // out_bf16: logical columns (multiple of 32); ld_out is in stored elements
  int32_t round_tf32;            // out_f32 = tf32-rounded value (so that out_lo is exact)
  int32_t debug;                 // CE_GPU_GEMM_DEBUG bits (timing probes only, results are WRONG):
                                 // 1 = epilogue drains TMEM but skips math and stores,
                                 // 2 = no tcgen05.mma is issued, 4 = no TMA loads are issued,
                                 // 8 = fused output layer without its stores, 16 = without its exponentials,
                                 // 32 = (not a probe: results stay right) plain stores instead of TMA stores

  // fused FindMinMax (src/matrix.cc:329-345) for the next layer's Quantize: only rows the next
  // layer's Splice+Narrow actually reads take part.
  uint32_t *minmax;              // [n_utts][2] order-preserving encodings; nullptr = off
  int32_t mm_lo, mm_hi;          // this layer's valid rows of an utterance: [mm_lo, P - mm_hi)
  int32_t next_n_taps;           // 0: every valid row is used
  int32_t next_tap_off[kMaxTaps];
  int32_t next_lo, next_hi;      // next layer's valid output rows: [next_lo, P - next_hi)

  // ---- LSM: the output layer fused with LogSoftmax + prior + argmax (no logits round trip) ----
  // One CTA group walks all column tiles of its row tile: a first sweep reduces every row's maximum
  // and sum of exponentials, a second sweep recomputes the accumulators and writes
  //   out_f32[out_row][n] = (x[n] - logsumexp(x)) - lsm_prior[n]      src/nnet.cc:137-146, src/am.cc:109-112
  // for the rows [lsm_left, P - lsm_right) of every utterance block, plus the row's argmax.
  int32_t lsm;                   // 1 = on (out_f32 / ld_out = destination; may be nullptr: argmax only)
  int32_t lsm_softmax;           // 0: no LogSoftmax layer (one sweep: prior + argmax only)
  int32_t lsm_left, lsm_right;
  int32_t lsm_rowspace;          // 1: out_row = the row itself (a workspace in row space), else compact:
  const int64_t *lsm_out_row_off;  //  out_row = lsm_out_row_off[utt] + (pos - lsm_left); nullptr = 0
  int64_t lsm_out_rows;          // rows of the destination matrix (the bound of the TMA stores)
  const float *lsm_prior;        // [N] log prior; nullptr = none
  int32_t *lsm_argmax;           // [out rows] first maximum of the finished row; nullptr = off
  int32_t lsm_single;            // 1: ONE sweep that writes the layer's plain result (row space) and the row's
  float *lsm_lse_out;            //    log-sum-exp to lsm_lse_out[row] -- for the selecting output kernels, which
                                 //    then subtract the very number the dense fused output subtracts
  int32_t lsm_zero;              // always 0 (an opaque -0.0 for the kernel's un-fused multiply, see lsm_value)

  unsigned long long *dbg;       // CE_GPU_GEMM_PROF: 8 cycle counters of the epilogue warps (gemm.cu), nullptr = off
};

struct GemmOperands {            // host-side description for the tensor maps
  const void *a[2];              // device pointers (hi, lo), [rows_a x c_pad]
  int64_t rows_a;
  const void *b[2];              // [n_rows_b x k_total]
  int64_t rows_b;                // N (rows beyond it read as zero)
  int64_t k_total;               // n_taps * c_pad
};

// Enqueues the GEMM.  `kind` selects the tensor-core data path.
int GemmLaunch(int kind, const GemmOperands &ops, const GemmArgs &args, cudaStream_t s);

// Order-preserving float <-> uint32 map used by the min/max atomics.
__host__ __device__ inline uint32_t OrderedFromFloat(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  uint32_t b;
  memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float FloatFromOrdered(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}

}  // namespace ce

#endif  // CE_GPU_GEMM_H_
