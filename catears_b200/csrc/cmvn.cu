// cmvn.cu -- K2: online mean-only CMVN (+ the AM's replicate padding) for sm_100a.
//
// Replaces CMVN::GetFrame (src/cmvn.cc:100-110) called for frames 0..T-1 in order:
//   ComputeStats  src/cmvn.cc:35-68   S_t = float(double(S_{t-1}) + x_t - x_{t-600}), count likewise
//   SmoothStats   src/cmvn.cc:70-89   if count < 600: S += (float)(min(600-count,200)/G_count) * G
//   Apply         src/cmvn.cc:91-98   y = x + (-(float)(1/count)) * S
// and AcousticModel's edge replication (src/am.cc:119-124,152-155).
//
// The reference re-rounds the running sum to fp32 after every frame, so the chain is replayed
// in order: one thread per (utterance, mel bin) walks the frames sequentially with
// exactly the reference's operations (double add, float round; un-fused float multiply/add).
// Parallelism comes from utterances x bins; the count / smoothing weight / 1/count depend only
// on the frame index and are taken from a 600-entry table built on the host with the same
// arithmetic.  Loads run one 16-frame batch ahead of the chain; a warp reads 128 contiguous bytes.
//
// HBM traffic: 4*mel bytes read + 4*mel written per frame (x_{t-600} is an L2 hit).

#include <algorithm>
#include <map>
#include <mutex>

#include "common.h"

namespace ce {
namespace {

struct CmvnUtt {
  int64_t in_row;    // first row of the utterance in feats
  int64_t out_row;   // first row of the utterance's block in out (before pad_left)
  int32_t T;
  int32_t pad;
};

struct CmvnStep {    // frame-index-only part of the chain (t < 600; t >= 599 uses entry 599)
  float alpha;       // smoothing weight, 0 when count >= 600
  float nscale;      // -(float)(1 / count_after_smoothing)
};

constexpr int kUnroll = 16;

// One thread per (utt, d), flattened so that warps stay full for any mel.
__global__ void __launch_bounds__(128)
cmvn_kernel(const float *__restrict__ g, const CmvnStep *__restrict__ steps,
            const float *__restrict__ feats, const CmvnUtt *__restrict__ utts, int n_utts,
            int mel, int pad_left, int pad_right, float *__restrict__ out, int64_t out_stride) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int u = (int)(idx / mel);
  const int d = (int)(idx % mel);
  if (u >= n_utts) return;
  const CmvnUtt ut = utts[u];
  const float *x = feats + ut.in_row * mel + d;
  float *y = out + (ut.out_row + pad_left) * out_stride + d;
  const int T = ut.T;
  const bool apply = g != nullptr;
  const float gd = apply ? g[d] : 0.0f;

  // Software pipeline: the loads of batch k+1 are in flight while the (sequential, rounding-
  // exact) chain of batch k runs.
  float cached = 0.0f, y_first = 0.0f, y_last = 0.0f;
  float xv[kUnroll], xo[kUnroll];
  auto load_batch = [&](int t0, float (&a)[kUnroll], float (&b)[kUnroll]) {
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      const int t = t0 + j;
      a[j] = (t < T) ? __ldg(x + (int64_t)t * mel) : 0.0f;
      b[j] = (apply && t < T && t >= kCmvnWindow) ? __ldg(x + (int64_t)(t - kCmvnWindow) * mel) : 0.0f;
    }
  };
  load_batch(0, xv, xo);
  for (int t0 = 0; t0 < T; t0 += kUnroll) {
    float nv[kUnroll], no[kUnroll];
    load_batch(t0 + kUnroll, nv, no);
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      const int t = t0 + j;
      if (t < T) {
        float r = xv[j];
        if (apply) {
          double s = (double)cached;                       // cmvn.cc:42-47 (double accumulate)
          s += (double)xv[j];
          if (t >= kCmvnWindow) s += -1.0 * (double)xo[j];
          cached = (float)s;                               // cmvn.cc:63-67 (stored as float)
          const CmvnStep st = steps[min(t, kCmvnWindow - 1)];
          float stat = cached;
          if (t < kCmvnWindow - 1) stat = __fadd_rn(stat, __fmul_rn(st.alpha, gd));   // AddVec
          r = __fadd_rn(xv[j], __fmul_rn(st.nscale, stat));                          // cmvn.cc:96-97
        }
        y[(int64_t)t * out_stride] = r;
        if (t == 0) y_first = r;
        y_last = r;
      }
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      xv[j] = nv[j];
      xo[j] = no[j];
    }
  }
  if (T > 0) {
    for (int p = 0; p < pad_left; ++p) y[(int64_t)(p - pad_left) * out_stride] = y_first;
    for (int p = 0; p < pad_right; ++p) y[(int64_t)(T + p) * out_stride] = y_last;
  }
}

struct StepTable {
  CmvnStep *dev = nullptr;
};
std::mutex g_mu;
std::map<std::pair<int, uint32_t>, StepTable> g_steps;   // (device, bits of global count)

// The count chain (cmvn.cc:49-60), SmoothStats on the count entry and Apply's scale, for
// t = 0..599, in the reference's arithmetic.
int GetSteps(float global_count, const CmvnStep **out) {
  int dev = 0;
  CE_CUDA(cudaGetDevice(&dev));
  uint32_t bits;
  memcpy(&bits, &global_count, 4);
  std::lock_guard<std::mutex> lock(g_mu);
  auto key = std::make_pair(dev, bits);
  auto it = g_steps.find(key);
  if (it == g_steps.end()) {
    std::vector<CmvnStep> h(kCmvnWindow);
    float cached_count = 0.0f;
    for (int t = 0; t < kCmvnWindow; ++t) {
      double c = (double)cached_count + 1.0;
      cached_count = (float)c;
      float stat = cached_count;
      float alpha = 0.0f;
      if ((double)stat < kCmvnWindow) {
        double from_global = kCmvnWindow - (double)stat;
        if (from_global > kCmvnGlobal) from_global = kCmvnGlobal;
        alpha = (float)(from_global / (double)global_count);
        volatile float prod = alpha * global_count;      // separate multiply, then add
        stat = stat + prod;
      }
      h[t].alpha = alpha;
      h[t].nscale = -(float)(1 / (double)stat);
    }
    StepTable st;
    CE_CUDA(cudaMalloc(&st.dev, sizeof(CmvnStep) * kCmvnWindow));
    CE_CUDA(cudaMemcpy(st.dev, h.data(), sizeof(CmvnStep) * kCmvnWindow, cudaMemcpyHostToDevice));
    it = g_steps.emplace(key, st).first;
  }
  *out = it->second.dev;
  return CE_GPU_OK;
}

}  // namespace

int CmvnLaunch(const float *global_stats_dev, float global_count, const float *feats_dev,
               const int64_t *frame_off, const int64_t *out_row_off, int n_utts, int num_mel,
               int pad_left, int pad_right, float *out_dev, int64_t out_stride, Table *utts,
               cudaStream_t s) {
  if (n_utts <= 0) return CE_GPU_OK;
  const CmvnStep *steps = nullptr;
  if (global_stats_dev) CE_CHECK(GetSteps(global_count, &steps));
  size_t bytes = sizeof(CmvnUtt) * (size_t)n_utts;
  CE_CHECK(utts->Acquire(bytes));
  CmvnUtt *h = utts->host<CmvnUtt>();
  for (int u = 0; u < n_utts; ++u) {
    h[u].in_row = frame_off[u];
    h[u].out_row = out_row_off[u];
    h[u].T = (int32_t)(frame_off[u + 1] - frame_off[u]);
    h[u].pad = 0;
  }
  CE_CHECK(utts->Upload(bytes, s));
  const int64_t n_threads = (int64_t)n_utts * num_mel;
  const unsigned grid = (unsigned)((n_threads + 127) / 128);
  ProfScope prof(kProfCmvn, s);
  cmvn_kernel<<<grid, 128, 0, s>>>(global_stats_dev, steps, feats_dev, utts->dev<CmvnUtt>(),
                                     n_utts, num_mel, pad_left, pad_right, out_dev, out_stride);
  CE_LAUNCHED();
  return CE_GPU_OK;
}

}  // namespace ce
