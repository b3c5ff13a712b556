// fbank.cu -- K1: batched log-mel filterbank for sm_100a.
//
// Replaces, per frame, the reference chain
//   WaveReader::Process   src/pcm_reader.cc:148-190   int16 -> float, unscaled
//   ExtractWindow         src/fbank.cc:74-100         400 samples, zero-pad to 512
//   ProcessWindow         src/fbank.cc:44-69          -mean, pre-emphasis 0.97, Hamming
//   SRFFT::Compute        src/srfft.cc:370-445        512-point real FFT
//   ComputePowerSpectrum  src/fbank.cc:193-211
//   Melbanks::Compute     src/fbank.cc:165-184        triangular filters (tables :103-163)
//   ApplyFloor + ApplyLog src/fbank.cc:243-244
//
// Design (B200): one CTA owns a run of up to kChunkFrames consecutive frames of one utterance.
//   phase 1  the run's PCM (each sample read from HBM exactly once, 16-byte vector loads) is
//            staged in shared memory as int16, turned into the frame-independent part of the
//            pre-emphasis d[s] = x[s] - 0.97 x[s-1] (fp32) and into exact integer sums of
//            80-sample blocks (the per-frame DC offset is a sum of five blocks).
//   phase 2  every half-warp (16 lanes) takes one frame at a time: 16 complex points per lane
//            in registers, a 256-point complex FFT as two register-resident radix-16 passes
//            with ONE transpose through shared memory, the real-FFT split done pairwise with
//            warp shuffles directly into power values, then the sparse mel filters, floor, log.
// The FFT is any-algorithm-admissible (the contract is the un-normalised DFT);
// twiddles come from double-precision tables, which is more accurate than the reference's fp32
// recurrence (src/srfft.cc:392), so parity is by tolerance (1e-4), not bits.

#include <float.h>
#include <math.h>

#include <algorithm>
#include <map>
#include <mutex>

#include "common.h"

namespace ce {
namespace {

constexpr int kChunkFrames = 32;                       // frames per CTA
constexpr int kThreads = 256;                          // 8 warps = 16 half-warps
constexpr int kHalfWarps = kThreads / 16;
constexpr int kChunkSamples = (kChunkFrames - 1) * kFrameShift + kFrameLen;   // 5360
constexpr int kXsLen = kChunkSamples + 16;             // + alignment slack
constexpr int kBlocksPerChunk = kChunkSamples / 80;    // 67 blocks of 80 samples
constexpr int kXchgStride = 17;                        // padded 16x16 float2 transpose
// float2 per half-warp buffer: 16 x 17 + 8, i.e. 560 floats == 16 (mod 32), so the two half-warps of
// a warp (which run the same instruction on their own buffers) land on disjoint shared-memory banks
constexpr int kXchgBuf = 16 * kXchgStride + 8;
constexpr int kMaxSlots = kMaxMel / 16;

struct ChunkDesc {        // one per CTA
  int64_t sample_begin;   // absolute index into pcm of the chunk's first sample
  int64_t out_row;        // first output row
  int32_t n_frames;
  int32_t pad;
};

struct MelSlot {          // work item of one lane in one slot
  int16_t k0, width, mel, pad;
  int32_t woff;
};

struct FbankTablesDev {   // lives in global memory, copied to smem by every CTA
  float hamming[kFrameLen];
  float2 tw256[256];      // tw256[n1*16+k2] = exp(-2 pi i n1 k2 / 256)
  float2 tw512[16];       // exp(-2 pi i t / 512), t = 0..15
  MelSlot slots[kMaxSlots][16];
  int32_t slot_iters[kMaxSlots];
  int32_t n_slots;
  int32_t n_weights;
  float weights[1];       // n_weights floats, each 0.25 * reference weight; slot-major, then
                          // [iteration][lane] so that a half-warp reads 16 consecutive floats
};

// ---------------------------------------------------------------------------
// complex helpers
// ---------------------------------------------------------------------------
// Complex add / subtract as ONE packed fp32x2 instruction (sm_100 add.f32x2 / sub.f32x2): the FFT's
// butterflies are mostly these, so the packed forms halve its issue slots; each half is an ordinary
// IEEE round-to-nearest fp32 operation.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "sub.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Forward 4-point DFT, natural order in and out.
__device__ __forceinline__ void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  const float2 r3 = make_float2(t3.y, -t3.x);   // -i t3
  a1 = cadd(t1, r3);                             // t1 - i t3
  a3 = csub(t1, r3);                             // t1 + i t3
}

// a * exp(-2 pi i M / 16)
template <int M>
__device__ __forceinline__ float2 mul_w16(float2 a) {
  constexpr float kC1 = 0.92387953251128674f, kS1 = 0.38268343236508977f, kH = 0.70710678118654752f;
  if (M == 0) return a;
  if (M == 4) return make_float2(a.y, -a.x);
  if (M == 2) return make_float2(kH * (a.x + a.y), kH * (a.y - a.x));
  if (M == 6) return make_float2(kH * (a.y - a.x), -kH * (a.x + a.y));
  // general: (c, -s)
  constexpr float c = (M == 1) ? kC1 : (M == 3) ? kS1 : (M == 9) ? -kC1 : 0.f;
  constexpr float s = (M == 1) ? kS1 : (M == 3) ? kC1 : (M == 9) ? -kS1 : 0.f;
  return make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
}

// Forward 16-point DFT in registers. Input v[n] natural order; output X[k] is left at
// v[FFT16_POS(k)] (a digit reversal resolved at compile time).
#define FFT16_POS(k) (4 * ((k) & 3) + ((k) >> 2))
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) fft4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
  // y[n1][k2] now at v[n1 + 4 k2]; twiddle by W16^(n1 k2)
  v[5] = mul_w16<1>(v[5]);   v[9] = mul_w16<2>(v[9]);   v[13] = mul_w16<3>(v[13]);
  v[6] = mul_w16<2>(v[6]);   v[10] = mul_w16<4>(v[10]); v[14] = mul_w16<6>(v[14]);
  v[7] = mul_w16<3>(v[7]);   v[11] = mul_w16<6>(v[11]); v[15] = mul_w16<9>(v[15]);
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) fft4(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
}

// 256-point forward complex FFT across a half-warp (16 lanes).
//   in : lane n1 holds z[n1 + 16 n2] in v[n2]
//   out: lane k2 holds Z[k2 + 16 k1] in v[FFT16_POS(k1)]
// tw[k2] = exp(-2 pi i n1 k2 / 256) for this lane; xchg = this half-warp's 16x17 float2 buffer.
template <int TW_STRIDE>
__device__ __forceinline__ void fft256_halfwarp(float2 (&v)[16], const float2 *tw,
                                                float2 *xchg, int t) {
  fft16(v);
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    float2 y = v[FFT16_POS(k2)];
    if (k2 > 0) y = cmul(y, tw[k2 * TW_STRIDE]);
    xchg[k2 * kXchgStride + t] = y;
  }
  __syncwarp();
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) v[n1] = xchg[t * kXchgStride + n1];
  __syncwarp();
  fft16(v);
}

// exp(-2 pi i (16 k1) / 512) = exp(-2 pi i k1 / 32), k1 = 0..7 (compile-time constants)
__device__ __forceinline__ float2 w32_const(int k1) {
  switch (k1) {
    case 0: return make_float2(1.0f, 0.0f);
    case 1: return make_float2(0.98078528040323043f, -0.19509032201612825f);
    case 2: return make_float2(0.92387953251128674f, -0.38268343236508977f);
    case 3: return make_float2(0.83146961230254524f, -0.55557023301960218f);
    case 4: return make_float2(0.70710678118654752f, -0.70710678118654752f);
    case 5: return make_float2(0.55557023301960218f, -0.83146961230254524f);
    case 6: return make_float2(0.38268343236508977f, -0.92387953251128674f);
    default: return make_float2(0.19509032201612825f, -0.98078528040323043f);
  }
}

// Real-FFT split straight to power: from Z (layout of fft256_halfwarp's output) writes
// P4[k] = 4 |X_k|^2 for k = 0..255 (X = 512-point real DFT) into p4[0..256].
// Pairs (k, 256-k) share C and w^k D (src/srfft.cc:391-433):  X_k = C + w^k D,
// X_{256-k} = conj(C - w^k D); the factor 4 (we skip the two 1/2's) is folded into the
// mel weights.
__device__ __forceinline__ void rfft_power(const float2 (&v)[16], float2 wt, float *p4, int t,
                                           int lane) {
  const int partner = (lane & 16) | ((16 - t) & 15);
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    float2 zk = v[FFT16_POS(k1)];
    float2 zp = v[FFT16_POS(15 - k1)];
    float2 zm;
    zm.x = __shfl_sync(0xffffffffu, zp.x, partner);
    zm.y = __shfl_sync(0xffffffffu, zp.y, partner);
    if (t == 0) zm = v[FFT16_POS((16 - k1) & 15)];   // residue 0 mirrors inside lane 0
    float2 c2 = make_float2(zk.x + zm.x, zk.y - zm.y);           // 2C
    float2 d2 = make_float2(zk.y + zm.y, zm.x - zk.x);           // 2D = -i (Zk - conj Zm)
    float2 w = cmul(wt, w32_const(k1));                          // exp(-2 pi i (t + 16 k1)/512)
    float2 e = cmul(w, d2);
    float2 a = cadd(c2, e), b = csub(c2, e);
    int k = t + 16 * k1;
    p4[k] = a.x * a.x + a.y * a.y;
    p4[256 - k] = b.x * b.x + b.y * b.y;
  }
  if (t == 0) {                                                  // k = 128 pairs with itself
    float2 z = v[FFT16_POS(8)];
    p4[128] = 4.0f * (z.x * z.x + z.y * z.y);
  }
}

// ---------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------
struct SmemLayout {
  // byte offsets into dynamic shared memory
  int xs, d, bsum, x0, xchg, hamming, weights, tw, total;
};

__host__ __device__ inline SmemLayout MakeLayout(int n_weights) {
  SmemLayout L;
  int o = 0;
  L.d = o;        o += kXsLen * 4;                         // float d[]
  L.xchg = o;     o += kHalfWarps * kXchgBuf * 8;          // float2 per half-warp (phase 2)
  L.xs = L.xchg;                                           // int16 xs[] (phase 1 only): same bytes
  L.hamming = o;  o += kFrameLen * 4;
  L.weights = o;  o += ((n_weights + 3) & ~3) * 4;
  L.tw = o;       o += 256 * 8;                            // float2 tw[k2][t] (lane-contiguous)
  L.bsum = o;     o += ((kBlocksPerChunk + 1 + 3) & ~3) * 4;
  L.x0 = o;       o += kChunkFrames * 4;                   // first sample of every frame
  L.total = (o + 15) & ~15;
  return L;
}

__global__ void __launch_bounds__(kThreads, 3)
fbank_kernel(const int16_t *__restrict__ pcm, int64_t total_samples,
             const ChunkDesc *__restrict__ chunks, const FbankTablesDev *__restrict__ tab,
             int num_mel, float *__restrict__ out, int64_t out_stride) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = MakeLayout(tab->n_weights);
  float *d = reinterpret_cast<float *>(smem + L.d);
  float2 *xchg_all = reinterpret_cast<float2 *>(smem + L.xchg);
  float *s_ham = reinterpret_cast<float *>(smem + L.hamming);
  float *s_w = reinterpret_cast<float *>(smem + L.weights);
  float2 *s_tw = reinterpret_cast<float2 *>(smem + L.tw);
  int *bsum = reinterpret_cast<int *>(smem + L.bsum);
  int16_t *xs = reinterpret_cast<int16_t *>(smem + L.xs);
  float *x0s = reinterpret_cast<float *>(smem + L.x0);

  const ChunkDesc cd = chunks[blockIdx.x];
  const int tid = threadIdx.x;
  const int n_samples = (cd.n_frames - 1) * kFrameShift + kFrameLen;

  // ---- phase 1a: PCM -> smem (16-byte vector loads from an 8-sample aligned start) ----
  const int shift = static_cast<int>(cd.sample_begin & 7);
  const int64_t g0 = cd.sample_begin - shift;              // multiple of 8 samples
  const int n_vec = (shift + n_samples + 7) >> 3;
  const bool aligned = (reinterpret_cast<uintptr_t>(pcm) & 15) == 0;
  {
    // every 16-byte load of this thread is issued before the first one is stored
    constexpr int kVecPerThread = (kXsLen / 8 + kThreads) / kThreads;
    int4 q[kVecPerThread];
    bool fast[kVecPerThread];
#pragma unroll
    for (int k = 0; k < kVecPerThread; ++k) {
      const int i = tid + k * kThreads;
      const int64_t g = g0 + 8 * (int64_t)i;
      fast[k] = i < n_vec && aligned && g + 8 <= total_samples;
      q[k] = fast[k] ? __ldg(reinterpret_cast<const int4 *>(pcm + g)) : make_int4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < kVecPerThread; ++k) {
      const int i = tid + k * kThreads;
      if (i >= n_vec) continue;
      if (fast[k]) {
        *reinterpret_cast<int4 *>(xs + 8 * i) = q[k];
      } else {
        const int64_t g = g0 + 8 * (int64_t)i;
        for (int j = 0; j < 8; ++j) xs[8 * i + j] = (g + j < total_samples) ? pcm[g + j] : (int16_t)0;
      }
    }
  }
  for (int i = tid; i < kFrameLen; i += kThreads) s_ham[i] = tab->hamming[i];
  for (int i = tid; i < tab->n_weights; i += kThreads) s_w[i] = tab->weights[i];
  // [n1][k2] -> [k2][n1]: the transposing side is the (cached, 2 KB) global read; written the other way round
  // every warp's store hit one bank pair 16 times over (ncu: 4 % of the kernel's shared-memory wavefronts)
  for (int i = tid; i < 256; i += kThreads) s_tw[i] = tab->tw256[(i & 15) * 16 + (i >> 4)];
  __syncthreads();

  // ---- phase 1b: d[s] = x[s] - 0.97 x[s-1]; exact sums of 80-sample blocks ----
  const int16_t *x = xs + shift;                           // x[0] = first sample of the chunk
  for (int i = tid; i < n_samples; i += kThreads) {
    float cur = (float)x[i];
    float prev = (i > 0) ? (float)x[i - 1] : 0.0f;
    d[i] = fmaf(-0.97f, prev, cur);
  }
  {
    // one warp per 80-sample block: consecutive lanes read consecutive samples (no bank conflicts)
    const int n_blocks = n_samples / 80;
    for (int b = tid >> 5; b < n_blocks; b += kThreads / 32) {
      const int j = tid & 31;
      int s = x[80 * b + j] + x[80 * b + 32 + j] + (j < 16 ? x[80 * b + 64 + j] : 0);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (j == 0) bsum[b] = s;
    }
    for (int f = tid; f < cd.n_frames; f += kThreads) x0s[f] = (float)x[f * kFrameShift];
  }
  __syncthreads();                                         // xs is dead from here: xchg reuses its bytes

  // ---- phase 2: one frame per half-warp ----
  const int lane = tid & 31;
  const int t = tid & 15;
  const int hw = tid >> 4;
  float2 *xchg = xchg_all + hw * kXchgBuf;
  float *p4 = reinterpret_cast<float *>(xchg);             // reused after the transpose

  const float2 *tw = s_tw + t;                             // tw[16 k2] = exp(-2 pi i t k2 / 256)
  const float2 wt = tab->tw512[t];
  const int n_slots = tab->n_slots;

  // Both half-warps of a warp must run the same number of iterations (shuffles, syncwarp).
  const int n_iter = (cd.n_frames + kHalfWarps - 1) / kHalfWarps;
  for (int it = 0; it < n_iter; ++it) {
    const int f = it * kHalfWarps + hw;
    const bool live = f < cd.n_frames;
    const int fs = live ? f * kFrameShift : 0;             // frame start within the chunk

    // DC offset: five 80-sample blocks (exact integer sum), then fp32 divide (fbank.cc:48-52)
    const int b0 = fs / 80;
    const int isum = bsum[b0] + bsum[b0 + 1] + bsum[b0 + 2] + bsum[b0 + 3] + bsum[b0 + 4];
    const float mean = (float)isum / (float)kFrameLen;
    const float c = fmaf(-0.97f, mean, mean);              // the mean's share of every p[i], i >= 1

    float2 v[16];
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) {
      const int i = 2 * t + 32 * n2;                       // sample index within the frame
      if (n2 < 12 || (n2 == 12 && t < 8)) {
        float2 dd = *reinterpret_cast<const float2 *>(d + fs + i);
        float2 ww = *reinterpret_cast<const float2 *>(s_ham + i);
        v[n2] = make_float2((dd.x - c) * ww.x, (dd.y - c) * ww.y);
      } else {
        v[n2] = make_float2(0.0f, 0.0f);
      }
    }
    if (t == 0) {                                          // sample 0: y0 - 0.97 y0 (fbank.cc:61)
      float y0 = x0s[live ? f : 0] - mean;
      v[0].x = fmaf(-0.97f, y0, y0) * s_ham[0];
    }

    fft256_halfwarp<16>(v, tw, xchg, t);
    rfft_power(v, wt, p4, t, lane);
    __syncwarp();

    // mel filters: slot s hands lane t one filter; every lane of the warp runs slot_iters[s] steps
    float *orow = out + (cd.out_row + f) * out_stride;
    for (int s = 0; s < n_slots; ++s) {
      const MelSlot ms = tab->slots[s][t];
      const int iters = tab->slot_iters[s];
      float acc = 0.0f;
      for (int i = 0; i < iters; ++i) {
        if (i < ms.width) acc = fmaf(s_w[ms.woff + 16 * i], p4[ms.k0 + i], acc);
      }
      if (live && ms.mel >= 0) orow[ms.mel] = logf(fmaxf(acc, FLT_EPSILON));
    }
    __syncwarp();
  }
}

// Test hook: rows of 512 floats -> packed real FFT (src/srfft.cc:370 layout). One frame per
// half-warp, same fft256_halfwarp as the production kernel; the split is the plain formula.
__global__ void __launch_bounds__(kThreads)
rfft512_kernel(const float *__restrict__ in, int n_frames, const FbankTablesDev *__restrict__ tab,
               float *__restrict__ out) {
  __shared__ float2 xchg_all[kHalfWarps * 16 * kXchgStride];
  const int tid = threadIdx.x, t = tid & 15, hw = tid >> 4;
  float2 *xchg = xchg_all + hw * 16 * kXchgStride;
  float2 *zbuf = xchg;   // 272 >= 256 entries; free again once fft256_halfwarp has returned
  float2 tw[16];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) tw[k2] = tab->tw256[t * 16 + k2];
  const int n_iter = (n_frames + gridDim.x * kHalfWarps - 1) / (gridDim.x * kHalfWarps);
  for (int it = 0; it < n_iter; ++it) {
    const int f = (it * gridDim.x + blockIdx.x) * kHalfWarps + hw;
    const bool live = f < n_frames;
    const float *row = in + (size_t)(live ? f : 0) * 512;
    float2 v[16];
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = *reinterpret_cast<const float2 *>(row + 2 * t + 32 * n2);
    fft256_halfwarp<1>(v, tw, xchg, t);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) zbuf[t + 16 * k1] = v[FFT16_POS(k1)];
    __syncwarp();
    if (live) {
      float *orow = out + (size_t)f * 512;
      for (int k = t; k < 256; k += 16) {
        float2 zk = zbuf[k], zm = zbuf[(256 - k) & 255];
        float2 cc = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        float2 dd = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));
        float sn, cs;
        sincospif(-(float)k / 256.0f, &sn, &cs);
        float2 e = cmul(make_float2(cs, sn), dd);
        if (k == 0) {
          orow[0] = zk.x + zk.y;
          orow[1] = zk.x - zk.y;
        } else {
          orow[2 * k] = cc.x + e.x;
          orow[2 * k + 1] = cc.y + e.y;
        }
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// host: tables (fp32 arithmetic exactly as the reference builds them)
// ---------------------------------------------------------------------------
float MelScale(float freq) { return 1127.0f * logf(1.0f + freq / 700.0f); }   // src/fbank.h:30-32

struct HostTables {
  std::vector<unsigned char> blob;
  int n_weights = 0;
};

int BuildTables(int num_mel, HostTables *ht) {
  // Mel filters, src/fbank.cc:103-163. (The reference asserts every filter spans >= 2 FFT
  // bins; we keep single-bin filters so that 80 bins can run -- an extension)
  const int num_fft_bins = kFftSize / 2;
  const float sample_freq = 16000;
  const float bin_width = sample_freq / kFftSize;
  const float mel_low = MelScale(20), mel_high = MelScale(8000);
  const float delta = (mel_high - mel_low) / (num_mel + 1);
  std::vector<int> off(num_mel), width(num_mel);
  std::vector<std::vector<float>> fw(num_mel);
  for (int b = 0; b < num_mel; ++b) {
    float left = mel_low + b * delta;
    float center = mel_low + (b + 1) * delta;
    float right = mel_low + (b + 2) * delta;
    int first = -1, last = -1;
    std::vector<float> tmp(num_fft_bins, 0.0f);
    for (int i = 0; i < num_fft_bins; ++i) {
      float mel = MelScale(bin_width * i);
      if (mel > left && mel < right) {
        tmp[i] = (mel <= center) ? (mel - left) / (center - left) : (right - mel) / (right - center);
        if (first == -1) first = i;
        last = i;
      }
    }
    if (first == -1) {
      SetError("mel filter %d of %d covers no FFT bin", b, num_mel);
      return CE_GPU_EINVAL;
    }
    off[b] = first;
    width[b] = last + 1 - first;
    for (int i = first; i <= last; ++i) fw[b].push_back(0.25f * tmp[i]);   // exact scaling
  }

  // Slots: 16 filters per slot, one per lane.  Filters are placed widest first into the earliest
  // slot with a free lane in which no filter has the same first bin modulo 16: lane t reads
  // p4[k0_t + i] in iteration i, so distinct k0 mod 16 means distinct banks for the whole warp
  // (the other half-warp's buffer is 16 banks away).
  const int n_slots = (num_mel + 15) / 16;
  std::vector<std::vector<int>> members(n_slots);
  {
    std::vector<int> order(num_mel);
    for (int b = 0; b < num_mel; ++b) order[b] = b;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return width[a] > width[b]; });
    for (int b : order) {
      int pick = -1;
      for (int s2 = 0; s2 < n_slots && pick < 0; ++s2) {
        if ((int)members[s2].size() >= 16) continue;
        bool clash = false;
        for (int o2 : members[s2]) clash = clash || (off[o2] % 16 == off[b] % 16);
        if (!clash) pick = s2;
      }
      for (int s2 = 0; s2 < n_slots && pick < 0; ++s2)
        if ((int)members[s2].size() < 16) pick = s2;       // no clash-free slot left: accept a conflict
      members[pick].push_back(b);
    }
  }
  int total_w = 0;
  std::vector<int> slot_iters(kMaxSlots, 0), slot_off(kMaxSlots, 0);
  for (int s2 = 0; s2 < n_slots; ++s2) {
    for (int b : members[s2]) slot_iters[s2] = std::max(slot_iters[s2], width[b]);
    slot_off[s2] = total_w;
    total_w += 16 * slot_iters[s2];
  }
  std::vector<float> weights((size_t)total_w, 0.0f);
  ht->n_weights = total_w;
  size_t bytes = sizeof(FbankTablesDev) + sizeof(float) * weights.size();
  ht->blob.assign(bytes, 0);
  FbankTablesDev *T = reinterpret_cast<FbankTablesDev *>(ht->blob.data());

  {  // Hamming, src/fbank.cc:248-255 (M_2PI = 6.28318530718 there, :19)
    float a = (float)(6.28318530718 / (kFrameLen - 1));
    for (int i = 0; i < kFrameLen; ++i) {
      float fi = (float)i;
      T->hamming[i] = (float)(0.54 - 0.46 * cosf(a * fi));
    }
  }
  const double kPi = 3.14159265358979323846;
  for (int n1 = 0; n1 < 16; ++n1)
    for (int k2 = 0; k2 < 16; ++k2) {
      double ang = -2.0 * kPi * (double)(n1 * k2) / 256.0;
      T->tw256[n1 * 16 + k2] = make_float2((float)cos(ang), (float)sin(ang));
    }
  for (int t = 0; t < 16; ++t) {
    double ang = -2.0 * kPi * (double)t / 512.0;
    T->tw512[t] = make_float2((float)cos(ang), (float)sin(ang));
  }
  T->n_slots = n_slots;
  for (int s2 = 0; s2 < kMaxSlots; ++s2) {
    T->slot_iters[s2] = slot_iters[s2];
    for (int t = 0; t < 16; ++t) {
      MelSlot ms = {0, 0, -1, 0, 0};
      if (s2 < n_slots && t < (int)members[s2].size()) {
        const int b = members[s2][t];
        ms.k0 = (int16_t)off[b];
        ms.width = (int16_t)width[b];
        ms.mel = (int16_t)b;
        ms.woff = slot_off[s2] + t;
        for (int i = 0; i < width[b]; ++i) weights[(size_t)slot_off[s2] + 16 * i + t] = fw[b][i];
      }
      T->slots[s2][t] = ms;
    }
  }
  T->n_weights = ht->n_weights;
  memcpy(T->weights, weights.data(), sizeof(float) * weights.size());
  return CE_GPU_OK;
}

struct DeviceTables {
  FbankTablesDev *dev = nullptr;
  int n_weights = 0;
};

std::mutex g_tab_mu;
std::map<std::pair<int, int>, DeviceTables> g_tables;   // (device, num_mel)

int GetTables(int num_mel, DeviceTables *out) {
  int dev = 0;
  CE_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_tab_mu);
  auto key = std::make_pair(dev, num_mel);
  auto it = g_tables.find(key);
  if (it == g_tables.end()) {
    HostTables ht;
    CE_CHECK(BuildTables(num_mel, &ht));
    DeviceTables dt;
    dt.n_weights = ht.n_weights;
    CE_CUDA(cudaMalloc(&dt.dev, ht.blob.size()));
    CE_CUDA(cudaMemcpy(dt.dev, ht.blob.data(), ht.blob.size(), cudaMemcpyHostToDevice));
    it = g_tables.emplace(key, dt).first;
  }
  *out = it->second;
  return CE_GPU_OK;
}

}  // namespace

int FbankLaunch(const int16_t *pcm_dev, int64_t total_samples, const int64_t *sample_off,
                const int64_t *frame_off, int n_utts, int num_mel, float *feats_dev,
                int64_t out_row_stride, Table *chunks, cudaStream_t s) {
  if (num_mel < 3 || num_mel > kMaxMel) {
    SetError("num_mel = %d not in [3, %d]", num_mel, kMaxMel);
    return CE_GPU_EINVAL;
  }
  DeviceTables dt;
  CE_CHECK(GetTables(num_mel, &dt));

  // chunk table
  int64_t n_chunks = 0;
  for (int u = 0; u < n_utts; ++u) {
    int64_t T = frame_off[u + 1] - frame_off[u];
    n_chunks += (T + kChunkFrames - 1) / kChunkFrames;
  }
  if (n_chunks == 0) return CE_GPU_OK;
  if (n_chunks > 0x7fffffff) {
    SetError("too many fbank chunks (%lld)", (long long)n_chunks);
    return CE_GPU_EINVAL;
  }
  size_t bytes = sizeof(ChunkDesc) * (size_t)n_chunks;
  CE_CHECK(chunks->Acquire(bytes));
  ChunkDesc *h = chunks->host<ChunkDesc>();
  int64_t c = 0;
  for (int u = 0; u < n_utts; ++u) {
    int64_t T = frame_off[u + 1] - frame_off[u];
    for (int64_t f0 = 0; f0 < T; f0 += kChunkFrames) {
      h[c].sample_begin = sample_off[u] + f0 * kFrameShift;
      h[c].out_row = frame_off[u] + f0;
      h[c].n_frames = (int32_t)std::min<int64_t>(kChunkFrames, T - f0);
      h[c].pad = 0;
      ++c;
    }
  }
  CE_CHECK(chunks->Upload(bytes, s));

  SmemLayout L = MakeLayout(dt.n_weights);
  // the opt-in above the 48 KB default is a per-DEVICE attribute: one host thread may drive several
  static thread_local int configured_smem[64] = {0};
  int dev = 0;
  CE_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || L.total > configured_smem[dev]) {
    CE_CUDA(cudaFuncSetAttribute(fbank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    if (dev < 64) configured_smem[dev] = L.total;
  }
  ProfScope prof(kProfFbank, s);
  fbank_kernel<<<(unsigned)n_chunks, kThreads, L.total, s>>>(
      pcm_dev, total_samples, chunks->dev<ChunkDesc>(), dt.dev, num_mel, feats_dev, out_row_stride);
  CE_LAUNCHED();
  return CE_GPU_OK;
}

int Rfft512Launch(const float *in_dev, int n_frames, float *out_dev, cudaStream_t s) {
  if (n_frames <= 0) return CE_GPU_OK;
  DeviceTables dt;
  CE_CHECK(GetTables(40, &dt));
  int blocks = std::min(1184, (n_frames + kHalfWarps - 1) / kHalfWarps);
  rfft512_kernel<<<blocks, kThreads, 0, s>>>(in_dev, n_frames, dt.dev, out_dev);
  CE_LAUNCHED();
  return CE_GPU_OK;
}

}  // namespace ce
