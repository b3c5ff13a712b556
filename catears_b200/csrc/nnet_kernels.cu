// nnet_kernels.cu -- the memory-bound kernels around the tensor-core GEMMs (K5, K6).
//
//   minmax_kernel     FindMinMax                    src/matrix.cc:329-345   (layer-0 input)
//   qparams_kernel    ComputeQuantizationParams     src/matrix.cc:348-362
//   quantize_kernel   Quantize                      src/matrix.cc:366-387   (+ row sums of the codes
//                     for the zero-point correction of internal/unpack.h:118-125)
//   convert_kernel    fp32 -> bf16 / (tf32 hi, lo) operand formats of the float paths
//   finalize_kernel   LogSoftmaxLayer src/nnet.cc:137-146 (ApplyLogSoftMax src/vector.cc:110-122)
//                     + "row -= log_prior" src/am.cc:109-112 + argmax, written as compact rows
#include "nnet_kernels.h"

#include <float.h>

#include <algorithm>

namespace ce {
namespace {

__device__ __forceinline__ bool RowUsed(int pos, int P, const RowUse &u) {
  if (pos < u.lo || pos >= P - u.hi) return false;
  if (u.next_n_taps == 0) return true;
  for (int t = 0; t < u.next_n_taps; ++t) {
    const int o = pos - u.next_tap_off[t];
    if (o >= u.next_lo && o < P - u.next_hi) return true;
  }
  return false;
}

__global__ void init_minmax_kernel(uint32_t *mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    mm[2 * i] = OrderedFromFloat(FLT_MAX);               // matrix.cc:330
    mm[2 * i + 1] = OrderedFromFloat(FLT_MIN);           // matrix.cc:331 (smallest positive normal)
  }
}

// One warp per row of x [M x ld] (first C columns).
__global__ void __launch_bounds__(256)
minmax_kernel(const float *__restrict__ x, int64_t ld, int C, int M,
              const int32_t *__restrict__ tile_utt, const UttRows *__restrict__ utts, RowUse use,
              uint32_t *__restrict__ minmax) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (blockDim.x >> 5);
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < M; row += warps) {
  const int utt = tile_utt ? tile_utt[row / kRowGran] : 0;
  int pos = row, P = M;
  if (utts) {
    pos = row - utts[utt].row_off;
    P = utts[utt].rows;
  }
  if (!RowUsed(pos, P, use)) continue;
  float vmin = FLT_MAX, vmax = -FLT_MAX;
  const float *r = x + (int64_t)row * ld;
  for (int c = lane; c < C; c += 32) {
    const float v = r[c];
    vmin = (v < vmin) ? v : vmin;
    vmax = (v > vmax) ? v : vmax;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if (lane == 0 && vmin <= vmax) {
    atomicMin(minmax + 2 * utt, OrderedFromFloat(vmin));
    atomicMax(minmax + 2 * utt + 1, OrderedFromFloat(vmax));
  }
  }
}

__global__ void qparams_kernel(const uint32_t *__restrict__ minmax, QParam *__restrict__ q, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float mn = FloatFromOrdered(minmax[2 * i]);
  const float mx = FloatFromOrdered(minmax[2 * i + 1]);
  const double scale = (double)__fsub_rn(mx, mn) / 255.0;             // matrix.cc:354
  const double fzp = (double)(-mn) / scale;                           // matrix.cc:357
  q[i].zero_point = (int32_t)round(fzp);                              // matrix.cc:358
  q[i].scale = (float)scale;                                          // matrix.cc:361
}

__device__ __forceinline__ uint32_t QuantOne(float v, float scale, float zp) {
  float q = __fadd_rn(__fdiv_rn(v, scale), zp);                       // matrix.cc:383
  q = (q < 255.0f) ? q : ((255.0f < q) ? 255.0f : q);                 // std::min(val, 255.0f)
  q = (0.0f < q) ? q : 0.0f;                                          // std::max(0.0f, .) (NaN -> 0)
  return (uint32_t)roundf(q);                                         // matrix.cc:385
}

// One warp per row: x [M x ld_in] fp32 (C columns) -> q [M x c_pad] u8 (zero padded) and
// rowsum[row] = sum of the C codes.  Four 16-byte loads per lane are in flight at a time.
__global__ void __launch_bounds__(256)
quantize_kernel(const float *__restrict__ x, int64_t ld_in, int C, int M, int c_pad,
                const int32_t *__restrict__ tile_utt, const QParam *__restrict__ qp,
                uint8_t *__restrict__ q, int32_t *__restrict__ rowsum) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (blockDim.x >> 5);
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < M; row += warps) {
  const int utt = tile_utt ? tile_utt[row / kRowGran] : 0;
  const QParam p = qp[utt];
  const float zp = (float)p.zero_point;
  const float *r = x + (int64_t)row * ld_in;
  uint32_t *o = reinterpret_cast<uint32_t *>(q + (int64_t)row * c_pad);
  int32_t sum = 0;
  const bool vec = ((ld_in & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  for (int base = 0; base < c_pad; base += 512) {
    float4 f[4];
    bool full[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c4 = base + u * 128 + lane * 4;
      full[u] = vec && c4 + 4 <= C;
      f[u] = full[u] ? __ldcs(reinterpret_cast<const float4 *>(r + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c4 = base + u * 128 + lane * 4;
      if (c4 >= c_pad) continue;
      uint32_t code[4] = {0, 0, 0, 0};
      if (full[u]) {
        code[0] = QuantOne(f[u].x, p.scale, zp);
        code[1] = QuantOne(f[u].y, p.scale, zp);
        code[2] = QuantOne(f[u].z, p.scale, zp);
        code[3] = QuantOne(f[u].w, p.scale, zp);
      } else {
        for (int j = 0; j < 4; ++j)
          if (c4 + j < C) code[j] = QuantOne(r[c4 + j], p.scale, zp);
      }
      sum += (int32_t)(code[0] + code[1] + code[2] + code[3]);
      o[c4 >> 2] = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
  if (lane == 0 && rowsum) rowsum[row] = sum;
  }
}

// ---------------------------------------------------------------------------------------------
// Quantize, the production form.  The kernel has to run BESIDE the persistent GEMM CTAs of the other
// chunk stream, which leave about 24 K registers and 1.7 K thread slots per SM: too few warps to
// cover HBM latency with one row per warp in flight.  So every warp keeps TWO rows (2 x NV x 16 B
// per lane, 8 KB per warp for 1024 columns) in flight, a block is 4 warps at <= 96 registers (two
// blocks fit beside a GEMM CTA = 64 KB in flight per SM), and the instruction stream is short:
// the division by the per-utterance scale is done with the correctly rounded reciprocal and two
// fused residual corrections (Markstein), which returns RN(v / scale) exactly as the reference's
// `val / scale` does, in 5 instructions instead of the ~11 of a general IEEE division.
// ComputeQuantizationParams (src/matrix.cc:348-362) is evaluated in the same kernel from the
// min/max the producing GEMM reduced, and published for the consuming GEMM's epilogue.
// ---------------------------------------------------------------------------------------------
constexpr int kQuantWarps = 4;
constexpr int kQuantRowsPerWarp = 4;
constexpr int kQuantUnitRows = kQuantWarps * kQuantRowsPerWarp;   // 16

__device__ __forceinline__ QParam QParamsFromMinMax(const uint32_t *__restrict__ mm) {
  const float mn = FloatFromOrdered(mm[0]);
  const float mx = FloatFromOrdered(mm[1]);
  const double scale = (double)__fsub_rn(mx, mn) / 255.0;             // matrix.cc:354
  const double fzp = (double)(-mn) / scale;                           // matrix.cc:357
  QParam q;
  q.zero_point = (int32_t)round(fzp);                                 // matrix.cc:358
  q.scale = (float)scale;                                             // matrix.cc:361
  return q;
}

struct QuantConst {
  float scale, inv, zp;
  bool fast;          // scale is comfortably normal: the reciprocal path is exact
};

__device__ __forceinline__ QuantConst MakeQuantConst(const QParam p) {
  QuantConst k;
  k.scale = p.scale;
  k.zp = (float)p.zero_point;
  k.fast = p.scale > 1e-18f && p.scale < 1e18f;                       // false for NaN / denormal / inf
  k.inv = k.fast ? __frcp_rn(p.scale) : 0.0f;
  return k;
}

// RN(v / scale): q0 = RN(v * RN(1/scale)) is within 2 ulp; one residual step makes it faithful and
// the second one correctly rounded (Markstein's theorem; the residuals are exact FMAs because
// |v| <= 255 * scale * (1 + eps) keeps every term in the normal range, and for |v / scale| < 0.4
// any last-bit error is absorbed by the rounding to an integer code).
__device__ __forceinline__ float DivByScale(float v, const QuantConst &k) {
  float q = __fmul_rn(v, k.inv);
  float r = __fmaf_rn(-q, k.scale, v);
  q = __fmaf_rn(r, k.inv, q);
  r = __fmaf_rn(-q, k.scale, v);
  return __fmaf_rn(r, k.inv, q);
}

// roundf(min(max(q, 0), 255)) as an int in [.., ..] BEFORE saturation: rounding half away from zero
// commutes with the clamp (monotone, 0 and 255 are fixed points), truncating q + 0.49999997f is
// round-half-away for 0 <= q < 2^23, negative / NaN inputs end at <= 0 and huge ones saturate.
template <bool FAST>
__device__ __forceinline__ int32_t CodeOf(float v, const QuantConst &k) {
  const float d = FAST ? DivByScale(v, k) : __fdiv_rn(v, k.scale);    // matrix.cc:383
  const float q = __fadd_rn(d, k.zp);
  return __float2int_rz(__fadd_rn(q, 0.49999997f));                   // matrix.cc:384-385
}

__device__ __forceinline__ uint32_t PackCodes(int32_t a, int32_t b, int32_t c, int32_t d) {
  uint32_t hi, w;   // bytes (low to high): a, b, c, d, each saturated to [0, 255]
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(b), "r"(a), "r"(hi));
  return w;
}

// One row held in registers -> NV words of codes per lane; returns this lane's share of the row sum.
template <int NV, bool FAST>
__device__ __forceinline__ int32_t QuantRow(const float4 (&v)[NV], const QuantConst &k, int C, int c_pad,
                                            int lane, uint32_t *__restrict__ o) {
  int32_t sum = 0;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c4 = (j * 32 + lane) * 4;
    uint32_t w = 0;
    if (c4 < C) {
      w = PackCodes(CodeOf<FAST>(v[j].x, k), CodeOf<FAST>(v[j].y, k), CodeOf<FAST>(v[j].z, k),
                    CodeOf<FAST>(v[j].w, k));
      sum = (int32_t)__dp4a(w, 0x01010101u, (uint32_t)sum);
    }
    if (c4 < c_pad) o[j * 32 + lane] = w;
  }
  return sum;
}

template <int NV>
__global__ void __maxnreg__(96)
quantize_rows_kernel(const float *__restrict__ x, int64_t ld_in, int C, int M, int c_pad,
                     const int32_t *__restrict__ tile_utt, const uint32_t *__restrict__ minmax,
                     QParam *__restrict__ qp, uint8_t *__restrict__ q, int32_t *__restrict__ rowsum) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int n_units = (M + kQuantUnitRows - 1) / kQuantUnitRows;
  // contiguous unit range of this block
  const int u_begin = (int)((int64_t)n_units * blockIdx.x / gridDim.x);
  const int u_end = (int)((int64_t)n_units * (blockIdx.x + 1) / gridDim.x);
  const int n_rows = (u_end - u_begin) * kQuantRowsPerWarp;           // rows of this warp
  if (n_rows <= 0) return;
  // Units are walked from the END of the matrix: the producing GEMM wrote its last ~100 MB of fp32
  // most recently (still in L2), and the consuming GEMM starts reading the codes at row 0, which are
  // then the ones written last.
  auto row_of = [&](int i) {
    return (n_units - 1 - (u_begin + i / kQuantRowsPerWarp)) * kQuantUnitRows + warp * kQuantRowsPerWarp +
           (i % kQuantRowsPerWarp);
  };
  auto load_row = [&](int row, float4 (&buf)[NV]) {
    const float *r = x + (int64_t)row * ld_in;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = (i * 32 + lane) * 4;
      buf[i] = (row < M && c4 < C) ? __ldcs(reinterpret_cast<const float4 *>(r + c4))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  float4 cur[NV], nxt[NV];
  load_row(row_of(0), cur);
  int utt_cached = -1;
  QuantConst k = {1.0f, 1.0f, 0.0f, false};
  for (int i = 0; i < n_rows; ++i) {
    const int row = row_of(i);
    if (i + 1 < n_rows) load_row(row_of(i + 1), nxt);
    if (row < M) {
      const int utt = tile_utt ? tile_utt[row / kRowGran] : 0;
      if (utt != utt_cached) {                                       // warp-uniform
        QParam p;
        if (minmax) {
          p = QParamsFromMinMax(minmax + 2 * utt);
          if (lane == 0) qp[utt] = p;                                // same value from every writer
        } else {
          p = qp[utt];
        }
        k = MakeQuantConst(p);
        utt_cached = utt;
      }
      uint32_t *o = reinterpret_cast<uint32_t *>(q + (int64_t)row * c_pad);
      const int32_t sum = k.fast ? QuantRow<NV, true>(cur, k, C, c_pad, lane, o)     // warp-uniform
                                 : QuantRow<NV, false>(cur, k, C, c_pad, lane, o);
      int32_t tot = sum;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, s);
      if (lane == 0 && rowsum) rowsum[row] = tot;
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) cur[j] = nxt[j];
  }
}

// Self-test of the reciprocal division against the IEEE division, and of the packed rounding
// against roundf: counts (v, scale) pairs whose codes differ.
__global__ void quant_selftest_kernel(int64_t n, uint64_t seed, unsigned long long *mismatches) {
  unsigned long long bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    // scale: random mantissa, exponent in [-40, 40]; v: scale * code-ish value with random low bits
    const uint32_t mant = (uint32_t)z & 0x7fffffu;
    const int e = (int)((z >> 23) % 81) - 40;
    const float scale = __uint_as_float(((uint32_t)(127 + e) << 23) | mant);
    const uint32_t r2 = (uint32_t)(z >> 32);
    const int mode = r2 & 3;
    float t;                                              // target quotient
    if (mode == 0) t = (float)(r2 >> 8) * (300.0f / 16777216.0f) - 20.0f;            // uniform [-20, 280)
    else if (mode == 1) t = (float)((r2 >> 8) & 255) + 0.5f;                          // near the ties
    else if (mode == 2) t = (float)((r2 >> 8) & 255) + 0.5f + ((int)((r2 >> 16) & 15) - 8) * 1e-5f;
    else t = __uint_as_float((r2 & 0x7fffffffu) % 0x43800000u);                        // any float in [0, 256)
    float v = __fmul_rn(t, scale);
    if (v != 0.0f) v = __uint_as_float(__float_as_uint(v) + ((r2 >> 5) & 3) - 1);    // +-1 ulp jitter
    QParam p;
    p.scale = scale;
    p.zero_point = (int32_t)((z >> 40) & 255);
    const QuantConst k = MakeQuantConst(p);
    const uint32_t got = (k.fast ? PackCodes(CodeOf<true>(v, k), 0, 0, 0)
                                : PackCodes(CodeOf<false>(v, k), 0, 0, 0)) & 255u;
    const uint32_t want = QuantOne(v, scale, (float)p.zero_point);
    if (got != want) ++bad;
    // the quotient itself, wherever it can matter for a code (|v / scale| >= 1/4)
    const float ref = __fdiv_rn(v, scale);
    if (k.fast && fabsf(ref) >= 0.25f && fabsf(ref) < 1e6f && DivByScale(v, k) != ref) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

__device__ __forceinline__ float RoundTf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// x [M x ld_in] fp32 (C columns) -> zero-padded [M x c_pad] bf16, or fp32 hi (+ lo), or the
// interleaved bf16 hi/lo operand of kKindBF16X3 ([M x 2 c_pad]: per 32 channels [32 hi | 32 lo]).
__global__ void __launch_bounds__(256)
convert_kernel(const float *__restrict__ x, int64_t ld_in, int C, int64_t M, int c_pad,
               __nv_bfloat16 *__restrict__ out_bf16, float *__restrict__ out_hi,
               float *__restrict__ out_lo, __nv_bfloat16 *__restrict__ out_x3) {
  const int64_t n = M * c_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / c_pad;
    const int c = (int)(i % c_pad);
    const float v = (c < C) ? x[row * ld_in + c] : 0.0f;
    if (out_bf16) out_bf16[i] = __float2bfloat16_rn(v);
    if (out_x3) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      __nv_bfloat16 *o = out_x3 + row * (2 * (int64_t)c_pad) + (c >> 5) * 64 + (c & 31);
      o[0] = h;
      o[32] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
    if (out_hi) {
      const float h = RoundTf32(v);
      out_hi[i] = h;
      if (out_lo) out_lo[i] = v - h;
    }
  }
}

// One warp per padded row; valid rows only.  Numerically stable log-sum-exp (the reference
// has no max subtraction, src/vector.cc:110-122; results agree to fp32 rounding while the
// reference's sum is finite).
__global__ void __launch_bounds__(256)
finalize_kernel(const float *__restrict__ logits, int64_t ld, int N, int M,
                const int32_t *__restrict__ tile_utt, const UttRows *__restrict__ utts,
                const int64_t *__restrict__ out_row_off, int left, int right, int log_softmax,
                const float *__restrict__ log_prior, float *__restrict__ loglik, int64_t ld_out,
                int32_t *__restrict__ argmax) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int utt = tile_utt[row / kRowGran];
  const UttRows ur = utts[utt];
  const int pos = row - ur.row_off;
  if (pos < left || pos >= ur.rows - right) return;
  const int64_t orow = out_row_off[utt] + (pos - left);
  const float *x = logits + (int64_t)row * ld;

  float lse = 0.0f;
  if (log_softmax) {
    float m = -FLT_MAX;
    for (int c = lane; c < N; c += 32) m = fmaxf(m, x[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.0f;
    for (int c = lane; c < N; c += 32) s += expf(x[c] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    lse = m + logf(s);
  }
  float best = -FLT_MAX;
  int best_i = 0x7fffffff;
  float *y = loglik ? loglik + orow * ld_out : nullptr;
  for (int c = lane; c < N; c += 32) {
    float v = x[c];
    if (log_softmax) v = __fsub_rn(v, lse);              // x -= log(sum)      vector.cc:120
    v = __fsub_rn(v, log_prior[c]);                      // AddVec(-1, log_prior_)  am.cc:111
    if (y) y[c] = v;
    if (v > best) {                                      // first maximum wins
      best = v;
      best_i = c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) {
      best = ob;
      best_i = oi;
    }
  }
  if (argmax && lane == 0) argmax[orow] = best_i == 0x7fffffff ? 0 : best_i;
}

__device__ __forceinline__ float Ex2(float x) {          // 2^x for x <= 0: MUFU.EX2, tiny results flush to 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Same, for N % 4 == 0 and N <= 128 * NV: the row stays in registers (NV float4 per lane), so the
// logits are read from HBM exactly once: 4N bytes in + 4N bytes out per frame.
template <int NV>
__global__ void __launch_bounds__(128, 3)
finalize_rowcache_kernel(const float *__restrict__ logits, int64_t ld, int N, int M,
                         const int32_t *__restrict__ tile_utt, const UttRows *__restrict__ utts,
                         const int64_t *__restrict__ out_row_off, int left, int right, int log_softmax,
                         const float *__restrict__ log_prior, float *__restrict__ loglik,
                         int64_t ld_out, int32_t *__restrict__ argmax) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (blockDim.x >> 5);
  // rows are walked from the end: the output-layer GEMM wrote those last (L2)
  for (int rr = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rr < M; rr += warps) {
  const int row = M - 1 - rr;
  const int utt = tile_utt[row / kRowGran];
  const UttRows ur = utts[utt];
  const int pos = row - ur.row_off;
  if (pos < left || pos >= ur.rows - right) continue;
  const int64_t orow = out_row_off[utt] + (pos - left);
  const float4 *x4 = reinterpret_cast<const float4 *>(logits + (int64_t)row * ld);
  const float4 *lp4 = reinterpret_cast<const float4 *>(log_prior);
  const int n4 = N >> 2;

  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 32 + lane;
    v[i] = (c < n4) ? __ldcs(x4 + c) : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
  }
  float lse = 0.0f;
  if (log_softmax) {
    float m = -FLT_MAX;
#pragma unroll
    for (int i = 0; i < NV; ++i) m = fmaxf(fmaxf(fmaxf(m, v[i].x), fmaxf(v[i].y, v[i].z)), v[i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    // exp(x - m) = 2^(x log2e - m log2e): one FFMA + one MUFU.EX2 per element (relative error 2^-22,
    // i.e. 1e-7 on the log-sum; the row maximum m keeps every term in (0, 1])
    const float kLog2e = 1.4426950408889634f;
    const float nm = -m * kLog2e;
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 32 + lane < n4)
        s += (Ex2(fmaf(v[i].x, kLog2e, nm)) + Ex2(fmaf(v[i].y, kLog2e, nm))) +
             (Ex2(fmaf(v[i].z, kLog2e, nm)) + Ex2(fmaf(v[i].w, kLog2e, nm)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    lse = m + logf(s);
  }
  float best = -FLT_MAX;
  int best_i = 0x7fffffff;
  float4 *y4 = loglik ? reinterpret_cast<float4 *>(loglik + orow * ld_out) : nullptr;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 32 + lane;
    if (c < n4) {
      const float4 lp = __ldg(lp4 + c);
      float4 r = v[i];
      if (log_softmax) {                                 // x -= log(sum)            vector.cc:120
        r.x = __fsub_rn(r.x, lse); r.y = __fsub_rn(r.y, lse);
        r.z = __fsub_rn(r.z, lse); r.w = __fsub_rn(r.w, lse);
      }
      r.x = __fsub_rn(r.x, lp.x); r.y = __fsub_rn(r.y, lp.y);   // AddVec(-1, log_prior_)  am.cc:111
      r.z = __fsub_rn(r.z, lp.z); r.w = __fsub_rn(r.w, lp.w);
      if (y4) __stcs(y4 + c, r);
      const float e[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (e[q] > best) {                               // first maximum wins (columns ascend per lane)
          best = e[q];
          best_i = 4 * c + q;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) {
      best = ob;
      best_i = oi;
    }
  }
  if (argmax && lane == 0) argmax[orow] = best_i == 0x7fffffff ? 0 : best_i;
  }
}

// ---- selected output for the decoder feed ----------------------------
// The decoder reads frame_logp(tid2pdf[ilabel]) for its active arcs only (src/decoder.cc:97-102),
// while a dense row is 12 KB a frame over PCIe.  Both kernels below compute the row exactly like
// finalize_rowcache_kernel (same operations, same order) and differ only in what they write.
// (The dense kernel keeps its own fused subtract-and-store loop: built on FinalRow it measured
// 2.30-2.47 ms instead of 2.18-2.24 ms per step, the stores starting later.)

// The finished row of one warp: NV float4 per lane, column 4 (i 32 + lane) + q; padding = -FLT_MAX.
template <int NV>
__device__ __forceinline__ void FinalRow(const float *__restrict__ x, int N, int lane, int log_softmax,
                                         const float *__restrict__ log_prior, float4 (&v)[NV],
                                         const float *__restrict__ lse_row = nullptr) {
  const float4 *x4 = reinterpret_cast<const float4 *>(x);
  const float4 *lp4 = reinterpret_cast<const float4 *>(log_prior);
  const int n4 = N >> 2;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 32 + lane;
    v[i] = (c < n4) ? __ldcs(x4 + c) : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
  }
  float lse = 0.0f;
  if (log_softmax && lse_row) {
    // the row's log-sum-exp as the output layer's own epilogue reduced it (GemmArgs::lsm_single): the same
    // number the fused dense output subtracts, so selected and dense rows agree bit for bit
    lse = __ldg(lse_row);
  } else if (log_softmax) {
    float m = -FLT_MAX;
#pragma unroll
    for (int i = 0; i < NV; ++i) m = fmaxf(fmaxf(fmaxf(m, v[i].x), fmaxf(v[i].y, v[i].z)), v[i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float kLog2e = 1.4426950408889634f;
    const float nm = -m * kLog2e;
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 32 + lane < n4)
        s += (Ex2(fmaf(v[i].x, kLog2e, nm)) + Ex2(fmaf(v[i].y, kLog2e, nm))) +
             (Ex2(fmaf(v[i].z, kLog2e, nm)) + Ex2(fmaf(v[i].w, kLog2e, nm)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    lse = m + logf(s);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 32 + lane;
    if (c < n4) {
      const float4 lp = __ldg(lp4 + c);
      float4 r = v[i];
      if (log_softmax) {                                 // x -= log(sum)            vector.cc:120
        r.x = __fsub_rn(r.x, lse); r.y = __fsub_rn(r.y, lse);
        r.z = __fsub_rn(r.z, lse); r.w = __fsub_rn(r.w, lse);
      }
      r.x = __fsub_rn(r.x, lp.x); r.y = __fsub_rn(r.y, lp.y);   // AddVec(-1, log_prior_)  am.cc:111
      r.z = __fsub_rn(r.z, lp.z); r.w = __fsub_rn(r.w, lp.w);
      v[i] = r;
    }
  }
}

// Valid output row of padded row `row`, or -1.
__device__ __forceinline__ int64_t OutputRowOf(int row, const int32_t *__restrict__ tile_utt,
                                               const UttRows *__restrict__ utts,
                                               const int64_t *__restrict__ out_row_off, int left,
                                               int right) {
  const int utt = tile_utt[row / kRowGran];
  const UttRows ur = utts[utt];
  const int pos = row - ur.row_off;
  if (pos < left || pos >= ur.rows - right) return -1;
  return out_row_off[utt] + (pos - left);
}

// Subset: out[orow][j] = loglik[ids[j]], j < n_ids.  The row goes through shared memory (N floats
// per warp) for the gather; the argmax is still the argmax over all N pdfs.
template <int NV>
__global__ void __launch_bounds__(128, 3)
finalize_subset_kernel(const float *__restrict__ logits, int64_t ld, int N, int M,
                       const int32_t *__restrict__ tile_utt, const UttRows *__restrict__ utts,
                       const int64_t *__restrict__ out_row_off, int left, int right, int log_softmax,
                       const float *__restrict__ log_prior, const int32_t *__restrict__ ids, int n_ids,
                       float *__restrict__ out, int64_t ld_out, int32_t *__restrict__ argmax,
                       const float *__restrict__ lse_in) {
  extern __shared__ float4 s_rows4[];
  const int lane = threadIdx.x & 31;
  const int n4 = N >> 2;
  float4 *srow4 = s_rows4 + (size_t)(threadIdx.x >> 5) * n4;
  const float *srow = reinterpret_cast<const float *>(srow4);
  const int warps = gridDim.x * (blockDim.x >> 5);
  for (int rr = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rr < M; rr += warps) {
    const int row = M - 1 - rr;
    const int64_t orow = OutputRowOf(row, tile_utt, utts, out_row_off, left, right);
    if (orow < 0) continue;
    float4 v[NV];
    FinalRow<NV>(logits + (int64_t)row * ld, N, lane, log_softmax, log_prior, v, lse_in ? lse_in + row : nullptr);
    float best = -FLT_MAX;
    int best_i = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < n4) {
        srow4[c] = v[i];
        const float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (e[q] > best) {
            best = e[q];
            best_i = 4 * c + q;
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ob > best || (ob == best && oi < best_i)) {
        best = ob;
        best_i = oi;
      }
    }
    if (argmax && lane == 0) argmax[orow] = best_i == 0x7fffffff ? 0 : best_i;
    __syncwarp();
    if (out) {
      float *y = out + orow * ld_out;
      for (int j = lane; j < n_ids; j += 32) __stcs(y + j, srow[__ldg(ids + j)]);
    }
    __syncwarp();                                        // the next row overwrites srow
  }
}

// fp32 -> unsigned key with the same order (larger float, larger key), and back.
__device__ __forceinline__ uint32_t OrderedKey(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float KeyToFloat(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// The k-th largest of the warp's 32 x LEN keys: the largest T with count(key >= T) >= k, found bit
// by bit from the top (0 when fewer than k keys are non-zero).
template <int LEN>
__device__ __forceinline__ uint32_t WarpKthLargest(const uint32_t (&key)[LEN], int k) {
  uint32_t T = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < LEN; ++e) cnt += (key[e] >= cand) ? 1 : 0;
    if (__reduce_add_sync(0xffffffffu, cnt) >= k) T = cand;
  }
  return T;
}

// Top-k: out[orow][j] = (loglik, pdf) of the j-th largest entry of the row, j < k; equal values in
// ascending pdf order (so entry 0 is the argmax with "first maximum wins").  One warp per row, the
// row's keys in registers.  The work is finding a lower bound T of the k-th largest key that leaves
// few candidates, compacting the entries >= T into shared memory as 64-bit (inverted key : pdf)
// words and sorting those (bitonic; ties order themselves by pdf):
//   1. small k (MT > 0): every lane keeps its MT largest keys; the k-th largest of those 32 MT
//      keys is such a bound, and a tight one while MT is about twice k / 32;
//   2. otherwise, or when (1) leaves more than `cap` candidates: the k-th largest key bit by bit
//      from the top over all keys, stopping as soon as at most `limit` candidates remain;
//   3. rows where even the exact k-th largest key has more than `cap` entries at or above it (many
//      equal values) compact exactly k entries: those above T and the lowest-numbered ties.
template <int NV, int MT>
__global__ void __launch_bounds__(128, 3)
finalize_topk_kernel(const float *__restrict__ logits, int64_t ld, int N, int M,
                     const int32_t *__restrict__ tile_utt, const UttRows *__restrict__ utts,
                     const int64_t *__restrict__ out_row_off, int left, int right, int log_softmax,
                     const float *__restrict__ log_prior, int k, int limit, int cap,
                     uint2 *__restrict__ out, int64_t ld_out_pairs, int32_t *__restrict__ argmax,
                     const float *__restrict__ lse_in) {
  extern __shared__ unsigned long long s_list_all[];
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int n4 = N >> 2;
  unsigned long long *list = s_list_all + (size_t)(threadIdx.x >> 5) * cap;
  const int warps = gridDim.x * (blockDim.x >> 5);
  for (int rr = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rr < M; rr += warps) {
    const int row = M - 1 - rr;
    const int64_t orow = OutputRowOf(row, tile_utt, utts, out_row_off, left, right);
    if (orow < 0) continue;
    uint32_t key[4 * NV];
    {
      float4 v[NV];
      FinalRow<NV>(logits + (int64_t)row * ld, N, lane, log_softmax, log_prior, v, lse_in ? lse_in + row : nullptr);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const bool in = i * 32 + lane < n4;              // padding sorts below every real entry
        key[4 * i + 0] = in ? OrderedKey(v[i].x) : 0u;
        key[4 * i + 1] = in ? OrderedKey(v[i].y) : 0u;
        key[4 * i + 2] = in ? OrderedKey(v[i].z) : 0u;
        key[4 * i + 3] = in ? OrderedKey(v[i].w) : 0u;
      }
    }
    constexpr int kUnknown = 0x7fffffff;
    uint32_t T = 0;
    int n_cand = kUnknown;                               // count(key >= T) in the row,
    int my_cand = 0;                                     // and among this lane's keys
    if (MT > 0) {
      uint32_t top[MT > 0 ? MT : 1];
#pragma unroll
      for (int j = 0; j < MT; ++j) top[j] = 0u;
#pragma unroll
      for (int e = 0; e < 4 * NV; ++e) {
        uint32_t x = key[e];
#pragma unroll
        for (int j = 0; j < MT; ++j) {                    // top[] stays sorted, x sinks through it
          const uint32_t hi = max(top[j], x);
          x = min(top[j], x);
          top[j] = hi;
        }
      }
      const uint32_t L = WarpKthLargest<(MT > 0 ? MT : 1)>(top, k);
      if (L != 0u) {
        int c = 0;
#pragma unroll
        for (int e = 0; e < 4 * NV; ++e) c += (key[e] >= L) ? 1 : 0;
        const int total = __reduce_add_sync(0xffffffffu, c);
        if (total <= cap) {
          T = L;
          n_cand = total;
          my_cand = c;
        }
      }
    }
    if (n_cand == kUnknown) {
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = T | (1u << bit);
        int c = 0;
#pragma unroll
        for (int e = 0; e < 4 * NV; ++e) c += (key[e] >= cand) ? 1 : 0;
        const int total = __reduce_add_sync(0xffffffffu, c);
        if (total >= k) {
          T = cand;
          n_cand = total;
          my_cand = c;
          if (total <= limit) break;
        }
      }
    }
    int n_out = 0;
    if (n_cand <= cap) {                                 // candidates: everything >= T, in any order
      int incl = my_cand;                                // every lane appends its own after the lower lanes'
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      unsigned long long *mine = list + (incl - my_cand);
      n_out = n_cand;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t kk = key[4 * i + q];
          if (kk >= T) *mine++ = ((unsigned long long)(~kk) << 32) | (uint32_t)(4 * (i * 32 + lane) + q);
        }
      }
    } else {                                             // T is exact here: above T, then the lowest ties
      int n_gt = 0;
#pragma unroll
      for (int e = 0; e < 4 * NV; ++e) n_gt += (key[e] > T) ? 1 : 0;
      const int need = k - __reduce_add_sync(0xffffffffu, n_gt);
      int ties_seen = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const bool mine = key[4 * i] >= T || key[4 * i + 1] >= T || key[4 * i + 2] >= T || key[4 * i + 3] >= T;
        if (!__any_sync(0xffffffffu, mine)) continue;
        // pdf order inside one i is (lane, q): ties in lower lanes come first, then my earlier q
        int before = 0, total_eq = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t b_eq = __ballot_sync(0xffffffffu, key[4 * i + q] == T);
          before += __popc(b_eq & lt_mask);
          total_eq += __popc(b_eq);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t kk = key[4 * i + q];
          bool sel = kk > T;
          if (kk == T) {
            sel = ties_seen + before < need;
            ++before;
          }
          const uint32_t b = __ballot_sync(0xffffffffu, sel);
          if (sel)
            list[n_out + __popc(b & lt_mask)] =
                ((unsigned long long)(~kk) << 32) | (uint32_t)(4 * (i * 32 + lane) + q);
          n_out += __popc(b);
        }
        ties_seen += total_eq;
      }
    }
    int n_sort = 1;                                      // n_out <= cap, a power of two
    while (n_sort < n_out) n_sort <<= 1;
    for (int j = n_out + lane; j < n_sort; j += 32) list[j] = ~0ull;
    __syncwarp();
    for (int size = 2; size <= n_sort; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = lane; t < (n_sort >> 1); t += 32) {
          const int lo = 2 * t - (t & (stride - 1));
          const int hi = lo + stride;
          const unsigned long long a = list[lo], b = list[hi];
          const bool up = (lo & size) == 0;
          if ((a > b) == up) {
            list[lo] = b;
            list[hi] = a;
          }
        }
        __syncwarp();
      }
    }
    if (argmax && lane == 0) argmax[orow] = (int32_t)(uint32_t)list[0];
    if (out) {
      uint2 *y = out + orow * ld_out_pairs;
      for (int j = lane; j < k; j += 32) {
        const unsigned long long e = list[j];
        __stcs(y + j, make_uint2(__float_as_uint(KeyToFloat(~(uint32_t)(e >> 32))), (uint32_t)e));
      }
    }
    __syncwarp();                                        // the next row overwrites the list
  }
}

}  // namespace

int InitMinMaxLaunch(uint32_t *mm, int n, cudaStream_t s) {
  if (n <= 0) return CE_GPU_OK;
  ProfScope prof(kProfQuantize, s);
  init_minmax_kernel<<<(n + 255) / 256, 256, 0, s>>>(mm, n);
  CE_LAUNCHED();
  return CE_GPU_OK;
}

int MinMaxLaunch(const float *x, int64_t ld, int C, int M, const int32_t *tile_utt,
                 const UttRows *utts, const RowUse &use, uint32_t *minmax, cudaStream_t s) {
  if (M <= 0) return CE_GPU_OK;
  ProfScope prof(kProfQuantize, s);
  minmax_kernel<<<std::min((M + 7) / 8, 4 * SmCount()), 256, 0, s>>>(x, ld, C, M, tile_utt, utts, use, minmax);
  CE_LAUNCHED();
  return CE_GPU_OK;
}

int QParamsLaunch(const uint32_t *minmax, QParam *q, int n, cudaStream_t s) {
  if (n <= 0) return CE_GPU_OK;
  ProfScope prof(kProfQuantize, s);
  qparams_kernel<<<(n + 127) / 128, 128, 0, s>>>(minmax, q, n);
  CE_LAUNCHED();
  return CE_GPU_OK;
}

int QuantizeLaunch(const float *x, int64_t ld_in, int C, int M, int c_pad, const int32_t *tile_utt,
                   const uint32_t *minmax, int n_utts, QParam *qp, uint8_t *q, int32_t *rowsum,
                   cudaStream_t s) {
  if (M <= 0) return CE_GPU_OK;
  if (c_pad % 4 != 0) {
    SetError("QuantizeLaunch: c_pad %d is not a multiple of 4", c_pad);
    return CE_GPU_EINVAL;
  }
  const bool rows_path = c_pad % 128 == 0 && c_pad <= 1024 && C % 4 == 0 && ld_in % 4 == 0 &&
                         ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (!rows_path) {                                      // odd shapes: one row per warp, generic loads
    if (minmax) CE_CHECK(QParamsLaunch(minmax, qp, n_utts, s));
    ProfScope prof(kProfQuantize, s);
    const unsigned grid = (unsigned)std::min((M + 7) / 8, 32 * SmCount());
    quantize_kernel<<<grid, 256, 0, s>>>(x, ld_in, C, M, c_pad, tile_utt, qp, q, rowsum);
    CE_LAUNCHED();
    return CE_GPU_OK;
  }
  ProfScope prof(kProfQuantize, s);
  const int n_units = (M + kQuantUnitRows - 1) / kQuantUnitRows;
  static const int per_sm = getenv("CE_GPU_QUANT_GRID") ? atoi(getenv("CE_GPU_QUANT_GRID")) : 4;
  const unsigned grid = (unsigned)std::min(n_units, per_sm * SmCount());
#define CE_QUANT_ROWS(NV)                                                                       \
  quantize_rows_kernel<NV><<<grid, 32 * kQuantWarps, 0, s>>>(x, ld_in, C, M, c_pad, tile_utt,   \
                                                              minmax, qp, q, rowsum)
  const int nv = c_pad / 128;
  if (nv <= 1) { CE_QUANT_ROWS(1); }
  else if (nv <= 2) { CE_QUANT_ROWS(2); }
  else if (nv <= 4) { CE_QUANT_ROWS(4); }
  else { CE_QUANT_ROWS(8); }
#undef CE_QUANT_ROWS
  CE_LAUNCHED();
  return CE_GPU_OK;
}

int QuantSelfTestLaunch(int64_t n, uint64_t seed, unsigned long long *mismatches_dev, cudaStream_t s) {
  quant_selftest_kernel<<<4 * SmCount(), 256, 0, s>>>(n, seed, mismatches_dev);
  CE_LAUNCHED();
  return CE_GPU_OK;
}

int ConvertLaunch(const float *x, int64_t ld_in, int C, int64_t M, int c_pad,
                  __nv_bfloat16 *out_bf16, float *out_hi, float *out_lo, cudaStream_t s,
                  __nv_bfloat16 *out_x3) {
  if (M <= 0) return CE_GPU_OK;
  const int64_t n = M * c_pad;
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
  ProfScope prof(kProfOther, s);
  convert_kernel<<<grid, 256, 0, s>>>(x, ld_in, C, M, c_pad, out_bf16, out_hi, out_lo, out_x3);
  CE_LAUNCHED();
  return CE_GPU_OK;
}

int FinalizeLaunch(const float *logits, int64_t ld, int N, int M, const int32_t *tile_utt,
                   const UttRows *utts, const int64_t *out_row_off, int left, int right,
                   bool log_softmax, const float *log_prior, float *loglik, int64_t ld_out,
                   int32_t *argmax, cudaStream_t s, const OutSel &sel, const float *lse_in) {
  if (M <= 0) return CE_GPU_OK;
  ProfScope prof(kProfFinalize, s);
  const unsigned fgrid = (unsigned)std::min((M + 3) / 4, 32 * SmCount());
  if (sel.mode != kOutDense) {
    // the selecting kernels keep the row in registers: N % 4 == 0, N <= 4096 (checked when the
    // selection is set, ce_gpu_model_set_output)
    if (N % 4 != 0 || ld % 4 != 0 || N > 4096 || sel.n <= 0 ||
        ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(log_prior)) & 15) != 0) {
      SetError("FinalizeLaunch: output selection needs num_pdfs %% 4 == 0 and num_pdfs <= 4096");
      return CE_GPU_EUNSUPPORTED;
    }
    // top-k: the candidates of one row sit in shared memory, `cap` entries per warp; the bit-by-bit
    // search stops at `limit` candidates (about 1.5 k, a power of two: what the sort works on)
    int limit = 32;
    while (limit < sel.n + sel.n / 2) limit <<= 1;
    const int cap = std::max(256, limit);
    const size_t smem = sel.mode == kOutSubset ? sizeof(float) * 4 * (size_t)N
                                               : sizeof(unsigned long long) * 4 * (size_t)cap;
#define CE_FINALIZE_TOPK(NV, MT)                                                                  \
  do {                                                                                            \
    if (smem > 48 * 1024)                                                                         \
      CE_CUDA(cudaFuncSetAttribute(finalize_topk_kernel<NV, MT>,                                  \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    finalize_topk_kernel<NV, MT><<<fgrid, 128, smem, s>>>(                                        \
        logits, ld, N, M, tile_utt, utts, out_row_off, left, right, log_softmax ? 1 : 0,          \
        log_prior, sel.n, limit, cap, reinterpret_cast<uint2 *>(loglik), ld_out / 2, argmax,      \
        lse_in);                                                                                  \
  } while (0)
#define CE_FINALIZE_SEL(NV)                                                                       \
  do {                                                                                            \
    if (sel.mode == kOutSubset) {                                                                 \
      if (smem > 48 * 1024)                                                                       \
        CE_CUDA(cudaFuncSetAttribute(finalize_subset_kernel<NV>,                                  \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
      finalize_subset_kernel<NV><<<fgrid, 128, smem, s>>>(                                        \
          logits, ld, N, M, tile_utt, utts, out_row_off, left, right, log_softmax ? 1 : 0,        \
          log_prior, sel.ids, sel.n, loglik, ld_out, argmax, lse_in);                             \
    } else if (sel.n <= 32) {                                                                     \
      CE_FINALIZE_TOPK(NV, 2);                                                                    \
    } else if (sel.n <= 64) {                                                                     \
      CE_FINALIZE_TOPK(NV, 4);                                                                    \
    } else if (sel.n <= 128) {                                                                    \
      CE_FINALIZE_TOPK(NV, 8);                                                                    \
    } else {                                                                                      \
      CE_FINALIZE_TOPK(NV, 0);                                                                    \
    }                                                                                             \
  } while (0)
    if (N <= 1024) CE_FINALIZE_SEL(8);
    else if (N <= 2048) CE_FINALIZE_SEL(16);
    else if (N <= 3072) CE_FINALIZE_SEL(24);
    else CE_FINALIZE_SEL(32);
#undef CE_FINALIZE_TOPK
#undef CE_FINALIZE_SEL
    CE_LAUNCHED();
    return CE_GPU_OK;
  }
  const bool vec = (N % 4 == 0) && (ld % 4 == 0) && (ld_out % 4 == 0) && N <= 4096 &&
                   ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(loglik) |
                     reinterpret_cast<uintptr_t>(log_prior)) & 15) == 0;
#define CE_FINALIZE(NV)                                                                            \
  finalize_rowcache_kernel<NV><<<fgrid, 128, 0, s>>>(logits, ld, N, M, tile_utt, utts,             \
                                                           out_row_off, left, right,               \
                                                           log_softmax ? 1 : 0, log_prior, loglik, \
                                                           ld_out, argmax)
  if (vec && N <= 1024) {
    CE_FINALIZE(8);
  } else if (vec && N <= 2048) {
    CE_FINALIZE(16);
  } else if (vec && N <= 3072) {
    CE_FINALIZE(24);
  } else if (vec) {
    CE_FINALIZE(32);
  } else {
    finalize_kernel<<<(M + 7) / 8, 256, 0, s>>>(logits, ld, N, M, tile_utt, utts, out_row_off, left,
                                               right, log_softmax ? 1 : 0, log_prior, loglik, ld_out,
                                               argmax);
  }
#undef CE_FINALIZE
  CE_LAUNCHED();
  return CE_GPU_OK;
}

}  // namespace ce
