// cmvn.cu -- K2: online mean-only CMVN (+ the AM's replicate padding) for sm_100a.
//
// Replaces CMVN::GetFrame (src/cmvn.cc:100-110) called for frames 0..T-1 in order:
//   ComputeStats  src/cmvn.cc:35-68   S_t = float(double(S_{t-1}) + x_t - x_{t-600}), count likewise
//   SmoothStats   src/cmvn.cc:70-89   if count < 600: S += (float)(min(600-count,200)/G_count) * G
//   Apply         src/cmvn.cc:91-98   y = x + (-(float)(1/count)) * S
// and AcousticModel's edge replication (src/am.cc:119-124,152-155).
//
// The reference re-rounds the running sum to fp32 after every frame, so the chain is replayed
// in order with exactly the reference's operations (double add, float round; un-fused float
// multiply/add) and stays bit-exact at any utterance length.  Only that chain is
// sequential -- four dependent operations per frame and bin -- so it is all the chain threads do:
// one CTA owns one utterance and works on tiles of kTileFrames frames in a three-stage pipeline,
//   loader/output warps   global -> shared memory for tile j+1 (x_t and x_{t-600}, coalesced),
//                         y_t = x_t + nscale_t (S_t + alpha_t g) and the stores for tile j-1,
//   chain warps           S_t for tile j, one thread per mel bin, operands read from shared memory
//                         (independent of the chain, so their latency is hidden by unrolling).
// The count / smoothing weight / 1/count depend only on the frame index and come from a
// 600-entry table built on the host with the same arithmetic.
//
// HBM traffic: 4*mel bytes read + 4*mel written per frame (x_{t-600} is an L2 hit).

#include <algorithm>
#include <map>
#include <mutex>

#include <float.h>

#include "common.h"
#include "gemm.h"

namespace ce {
namespace {

struct CmvnUtt {
  int64_t in_row;    // first row of the utterance's (new) frames in feats
  int64_t out_row;   // first row of the utterance's block in out (before pad_left)
  int32_t T;         // frames to normalise in this launch
  int32_t t_base;    // frames of the utterance normalised by earlier launches (streaming; else 0):
                     // the raw frames t_base - 600 .. t_base - 1 then precede in_row in feats
};

struct CmvnStep {    // frame-index-only part of the chain (t < 600; t >= 599 uses entry 599)
  float alpha;       // smoothing weight, 0 when count >= 600
  float nscale;      // -(float)(1 / count_after_smoothing)
};

constexpr int kCmvnThreads = 256;

// One CTA = one utterance x one group of `nb_per` consecutive mel bins (all of them when the batch has enough
// utterances to fill the GPU; groups of 8 for a few long utterances, so that an hour-long stream spreads over
// several SMs and its tiles get long: the chain of a bin is sequential, its bins are not).
// Shared memory: x[2][TF][nb], d[2][TF][nb] (x_t - x_{t-600}), S[2][TF][nb] floats, then inexact[2].
// NBT: nb_per as a compile-time constant (the chain's loads and stores then take immediate offsets), 0 = any.
// CHAIN_DIFF: who makes d = x_t - x_{t-600} and its exactness test.  false: the workers, while they stage the
// tile (a few long utterances: the chain is the bound and does nothing but add); true: the chain threads, eight
// frames at a time ahead of the dependent adds (a batch of many utterances: there the workers are the bound --
// the test costs them a fifth more instructions -- and the chain has time to spare).
template <int NBT, bool CHAIN_DIFF>
__global__ void __launch_bounds__(kCmvnThreads, 2)
cmvn_kernel(const float *__restrict__ g, const CmvnStep *__restrict__ steps,
            const float *__restrict__ feats, const CmvnUtt *__restrict__ utts, int n_utts,
            int mel, int nb_per, int tile_frames, int pad_left, int pad_right, float *__restrict__ out,
            int64_t out_stride, float *__restrict__ state, uint32_t *__restrict__ minmax) {
  extern __shared__ float cmvn_smem[];
  const CmvnUtt ut = utts[blockIdx.x];
  const int T = ut.T;
  if (T <= 0) return;
  const int b0 = blockIdx.y * nb_per;            // this CTA's bins [b0, b0 + nb)
  const int nb = min(nb_per, mel - b0);
  const int tb = ut.t_base;                    // absolute index of this launch's first frame
  const int TF = tile_frames;
  const int tile_elems = TF * nb_per;
  float *xs = cmvn_smem;                         // [2][tile_elems]
  float *dds = cmvn_smem + 2 * tile_elems;       // [2][tile_elems]
  float *ss = cmvn_smem + 4 * tile_elems;        // [2][tile_elems]
  int *inexact = reinterpret_cast<int *>(cmvn_smem + 6 * tile_elems);   // [2]: some x - x_old of the tile is not exact
  const bool apply = g != nullptr;
  const int chain_threads = apply ? ((nb_per + 31) / 32) * 32 : 0;   // whole warps
  const int tid = threadIdx.x;
  const bool is_chain = tid < chain_threads;
  const int wtid = tid - chain_threads;          // index among the loader/output threads
  const int n_workers = kCmvnThreads - chain_threads;
  const float *x = feats + ut.in_row * mel + b0;
  float *y = out + (ut.out_row + pad_left) * out_stride + b0;
  const int n_tiles = (T + TF - 1) / TF;

  // Workers index a tile flat, element i = (frame i / nb, bin i % nb); i / nb by multiply-shift (i < 2^13,
  // nb <= 128: exact with a 20-bit reciprocal).  Every load of a tile is issued before any is used.
  const uint32_t inv_nb = ((1u << 20) + nb - 1) / nb;
  constexpr int kPerThread = 16;                 // >= tile_elems / n_workers for every configuration
  // Loading a tile is split in two so that its latency hides behind the stores of the previous tile:
  // fetch_tile issues every load into registers, commit_tile writes them to shared memory once the
  // buffer is free.
  float la[kPerThread], lb[kPerThread];
  auto fetch_tile = [&](int j) {                 // workers: global -> registers, tile j
    const int t0 = j * TF;
    const int n = min(TF, T - t0) * nb;
    const float *src = x + (int64_t)t0 * mel;
    const bool need_old = apply && tb + t0 + TF > kCmvnWindow;   // some frame of the tile has t >= 600
    const int first_old = kCmvnWindow - tb - t0;                 // tile frames before it have t < 600
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) {
      const int i = wtid + k * n_workers;
      const int tl = (int)(((uint32_t)i * inv_nb) >> 20), d = i - tl * nb;
      const int64_t o = (int64_t)tl * mel + d;
      la[k] = (i < n) ? __ldg(src + o) : 0.0f;
      lb[k] = (need_old && i < n && tl >= first_old) ? __ldg(src + o - (int64_t)kCmvnWindow * mel) : 0.0f;
    }
    asm volatile("" ::: "memory");               // the loads are issued here, not where they are used
  };
  // The reference's  S = float(double(S) + x - x_old)  (cmvn.cc:42-47,63-67) without fp64, whose
  // CUDA-core rate on this part makes a dependent chain cost ~280 clocks per frame: the double sums
  // are exact (24-bit operands a few binades apart), so S is the correctly rounded three-term sum.
  // With d = x - x_old: if the subtraction is exact (its TwoSum error term is zero -- always, for
  // log-mel magnitudes) the result is RN(S + d), one fp32 add; otherwise the tile falls back to fp64.
  // For t < 600 it is RN(S + x) (x_old = 0).  The differences and the exactness test are element-wise,
  // so the WORKERS make them while they stage the tile; only the add is left on the dependent chain.
  auto commit_tile = [&](int j) {                // workers: registers -> smem, tile j
    const int t0 = j * TF;
    const int n = min(TF, T - t0) * nb;
    float *dx = xs + (j & 1) * tile_elems;
    float *dd = dds + (j & 1) * tile_elems;
    bool bad = false;
    if (CHAIN_DIFF) {
#pragma unroll
      for (int k = 0; k < kPerThread; ++k) {
        const int i = wtid + k * n_workers;
        if (i < n) {
          dx[i] = la[k];
          dd[i] = lb[k];                         // x_{t-600} itself (0 for t < 600)
        }
      }
      return;
    }
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) {
      const int i = wtid + k * n_workers;
      if (i < n) {
        const float xv = la[k], xo = lb[k];
        const float dh = __fsub_rn(xv, xo);                            // TwoSum(x, -x_old)
        const float bv = __fsub_rn(dh, xv);
        const float dl = __fadd_rn(__fsub_rn(xv, __fsub_rn(dh, bv)), __fsub_rn(-xo, bv));
        bad |= dl != 0.0f;
        dx[i] = xv;
        dd[i] = dh;
      }
    }
    if (apply && bad) inexact[j & 1] = 1;
  };
  float wmin = FLT_MAX, wmax = -FLT_MAX;         // FindMinMax of what this thread writes (matrix.cc:329-345)
  auto store_tile = [&](int j) {                 // workers: y for tile j, and the replicated edges
    const int t0 = j * TF;
    const int n = min(TF, T - t0) * nb;
    const float *dx = xs + (j & 1) * tile_elems;
    const float *ds = ss + (j & 1) * tile_elems;
#pragma unroll 4
    for (int i = wtid; i < n; i += n_workers) {
      const int tl = (int)(((uint32_t)i * inv_nb) >> 20), d = i - tl * nb;
      const int t = t0 + tl;
      float r = dx[i];
      if (apply) {
        const CmvnStep st = steps[min(tb + t, kCmvnWindow - 1)];
        float stat = ds[i];
        if (tb + t < kCmvnWindow - 1) stat = __fadd_rn(stat, __fmul_rn(st.alpha, __ldg(g + b0 + d)));   // AddVec
        r = __fadd_rn(r, __fmul_rn(st.nscale, stat));                                       // cmvn.cc:96-97
      }
      y[(int64_t)t * out_stride + d] = r;
      wmin = (r < wmin) ? r : wmin;
      wmax = (r > wmax) ? r : wmax;
      if (t == 0)
        for (int p = 1; p <= pad_left; ++p) y[-(int64_t)p * out_stride + d] = r;
      if (t == T - 1)
        for (int p = 1; p <= pad_right; ++p) y[(int64_t)(T - 1 + p) * out_stride + d] = r;
    }
  };

  float cached = 0.0f;                           // chain state of this thread's bin
  if (state && is_chain && tid < nb) cached = state[(int64_t)blockIdx.x * mel + b0 + tid];   // streaming: resume
  if (tid < 2) inexact[tid] = 0;
  __syncthreads();
  if (!is_chain) {
    fetch_tile(0);
    commit_tile(0);
  }
  __syncthreads();
  for (int j = 0; j < n_tiles; ++j) {
    if (is_chain) {
      if (tid < nb) {
        const int t0 = j * TF;
        const int nt = min(TF, T - t0);
        const float *dd = dds + (j & 1) * tile_elems + tid;
        float *ds = ss + (j & 1) * tile_elems + tid;
        if (CHAIN_DIFF) {
          const int st = NBT ? NBT : nb;
          const float *px = xs + (j & 1) * tile_elems + tid;
          constexpr int kB = 8;
          for (int tl0 = 0; tl0 < nt; tl0 += kB) {
            float a[kB], o[kB], dh[kB];
            bool bad = false;
#pragma unroll
            for (int u = 0; u < kB; ++u) {
              const bool in = tl0 + u < nt;
              a[u] = in ? px[(tl0 + u) * st] : 0.0f;
              o[u] = in ? dd[(tl0 + u) * st] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < kB; ++u) {                             // TwoSum(x, -x_old), off the chain
              dh[u] = __fsub_rn(a[u], o[u]);
              const float bv = __fsub_rn(dh[u], a[u]);
              const float dl = __fadd_rn(__fsub_rn(a[u], __fsub_rn(dh[u], bv)), __fsub_rn(-o[u], bv));
              bad |= dl != 0.0f;
            }
            if (!bad) {
#pragma unroll
              for (int u = 0; u < kB; ++u) {
                if (tl0 + u < nt) {
                  cached = __fadd_rn(cached, dh[u]);
                  ds[(tl0 + u) * st] = cached;
                }
              }
            } else {                             // an inexact difference among these frames: as written, in fp64
              for (int u = 0; u < kB; ++u) {
                if (tl0 + u < nt) {
                  if (tb + t0 + tl0 + u >= kCmvnWindow) {
                    double s2 = (double)cached;
                    s2 += (double)a[u];
                    s2 += -1.0 * (double)o[u];
                    cached = (float)s2;
                  } else {
                    cached = __fadd_rn(cached, a[u]);
                  }
                  ds[(tl0 + u) * st] = cached;
                }
              }
            }
          }
        } else if (!inexact[j & 1]) {            // (always, for log-mel features)
          // 16 differences are in registers before the first sum is stored, and the next 16 are on their
          // way while these are added: as a plain loop every load waited behind the store before it (the
          // compiler must assume they alias) and a frame cost ~140 clocks; only the add is sequential.
          constexpr int kBatch = 16;
          const int st = NBT ? NBT : nb;         // floats between consecutive frames of a bin
          int tl0 = 0;
          float v[kBatch], w[kBatch];
          if (nt >= kBatch) {
#pragma unroll
            for (int u = 0; u < kBatch; ++u) v[u] = dd[u * st];
          }
          for (; tl0 + kBatch <= nt; tl0 += kBatch) {
            const bool more = tl0 + 2 * kBatch <= nt;
            if (more) {
#pragma unroll
              for (int u = 0; u < kBatch; ++u) w[u] = dd[(tl0 + kBatch + u) * st];
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
              cached = __fadd_rn(cached, v[u]);
              ds[(tl0 + u) * st] = cached;
            }
            if (more) {
#pragma unroll
              for (int u = 0; u < kBatch; ++u) v[u] = w[u];
            }
          }
          for (; tl0 < nt; ++tl0) {              // the last frames of an utterance
            cached = __fadd_rn(cached, dd[tl0 * st]);
            ds[tl0 * st] = cached;
          }
        } else {                                 // some difference of the tile is inexact: as written, in fp64
          const float *xg = x + (int64_t)t0 * mel + tid;
          for (int tl = 0; tl < nt; ++tl) {
            const float xv = xg[(int64_t)tl * mel];
            if (tb + t0 + tl >= kCmvnWindow) {
              const float xo = xg[((int64_t)tl - kCmvnWindow) * mel];
              double s2 = (double)cached;
              s2 += (double)xv;
              s2 += -1.0 * (double)xo;
              cached = (float)s2;
            } else {
              cached = __fadd_rn(cached, xv);
            }
            ds[tl * nb] = cached;
          }
        }
      }
    } else {
      if (j + 1 < n_tiles) fetch_tile(j + 1);    // in flight during the stores below
      if (j > 0) store_tile(j - 1);              // reads buffers (j-1)&1 ...
      if (wtid == 0) inexact[(j + 1) & 1] = 0;   // (the chain read that flag one iteration ago)
      asm volatile("bar.sync 1, %0;" ::"r"(n_workers) : "memory");
      if (j + 1 < n_tiles) commit_tile(j + 1);   // ... which tile j+1 then overwrites
    }
    __syncthreads();
  }
  if (!is_chain) store_tile(n_tiles - 1);
  if (state && is_chain && tid < nb) state[(int64_t)blockIdx.x * mel + b0 + tid] = cached;
  if (minmax && !is_chain) {                     // the utterance's min/max for the first Quantize
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wmin = fminf(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
      wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    }
    if ((tid & 31) == 0 && wmin <= wmax) {
      atomicMin(minmax + 2 * blockIdx.x, OrderedFromFloat(wmin));
      atomicMax(minmax + 2 * blockIdx.x + 1, OrderedFromFloat(wmax));
    }
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
               cudaStream_t s, const CmvnResume *resume, uint32_t *minmax_dev) {
  if (n_utts <= 0) return CE_GPU_OK;
  HostMark(nullptr);
  const CmvnStep *steps = nullptr;
  if (global_stats_dev) CE_CHECK(GetSteps(global_count, &steps));
  size_t bytes = sizeof(CmvnUtt) * (size_t)n_utts;
  CE_CHECK(utts->Acquire(bytes));
  HostMark("cmvn: acquire");
  CmvnUtt *h = utts->host<CmvnUtt>();
  for (int u = 0; u < n_utts; ++u) {
    h[u].in_row = frame_off[u];
    h[u].out_row = out_row_off[u];
    h[u].T = (int32_t)(frame_off[u + 1] - frame_off[u]);
    h[u].t_base = 0;
    if (resume) {                                          // rows [frame_off[u], +n_hist) are history
      h[u].in_row += resume->n_hist[u];
      h[u].T -= resume->n_hist[u];
      h[u].t_base = (int32_t)std::min<int64_t>(resume->t_base[u], 0x7fffffff);
    }
  }
  HostMark("cmvn: table");
  CE_CHECK(utts->Upload(bytes, s));
  HostMark("cmvn: upload");
  if (num_mel > kMaxMel) {
    SetError("CmvnLaunch: num_mel %d > %d", num_mel, kMaxMel);
    return CE_GPU_EINVAL;
  }
  // Bins per CTA: all of them when there are utterances enough to fill the GPU; groups of 8 (one 32-byte
  // sector of every frame) for a few long ones -- the bins' chains are independent, and a CTA with fewer
  // bins affords longer tiles (fewer barriers per frame).
  int64_t total_frames = 0;
  for (int u = 0; u < n_utts; ++u) total_frames += h[u].T;
  const bool split = global_stats_dev && num_mel % 8 == 0 && n_utts < 64 && total_frames / std::max(1, n_utts) >= 4096;
  static const int nb_env = getenv("CE_GPU_CMVN_BINS") ? atoi(getenv("CE_GPU_CMVN_BINS")) : 0;   // A/B: bins per CTA
  const int nb_per = (nb_env > 0 && num_mel % nb_env == 0) ? nb_env : split ? (n_utts < 4 ? 4 : 8) : num_mel;
  const int n_groups = (num_mel + nb_per - 1) / nb_per;
  // frames per tile: the largest power of two whose tile the worker threads cover with 16 elements each
  const int workers = kCmvnThreads - (global_stats_dev ? (nb_per + 31) / 32 * 32 : 0);
  int tile_frames = 512;
  while (tile_frames > 1 && tile_frames * nb_per > 16 * workers) tile_frames >>= 1;
  const size_t smem = sizeof(float) * 6 * (size_t)tile_frames * nb_per + 16;
  static thread_local size_t configured[64] = {0};
  int dev = 0;
  CE_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || smem > configured[dev]) {
    CE_CUDA(cudaFuncSetAttribute(cmvn_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CE_CUDA(cudaFuncSetAttribute(cmvn_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CE_CUDA(cudaFuncSetAttribute(cmvn_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CE_CUDA(cudaFuncSetAttribute(cmvn_kernel<40, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev < 64) configured[dev] = smem;
  }
  ProfScope prof(kProfCmvn, s);
#define CE_CMVN_LAUNCH(NBT, CD)                                                                               \
  cmvn_kernel<NBT, CD><<<dim3((unsigned)n_utts, (unsigned)n_groups), kCmvnThreads, smem, s>>>(                 \
      global_stats_dev, steps, feats_dev, utts->dev<CmvnUtt>(), n_utts, num_mel, nb_per, tile_frames, pad_left, \
      pad_right, out_dev, out_stride, resume ? resume->state_dev : nullptr, minmax_dev)
  const bool whole = num_mel % nb_per == 0;      // every CTA has exactly nb_per bins
  if (whole && nb_per == 40) CE_CMVN_LAUNCH(40, true);
  else if (whole && nb_per == 8) CE_CMVN_LAUNCH(8, false);
  else if (whole && nb_per == 4) CE_CMVN_LAUNCH(4, false);
  else CE_CMVN_LAUNCH(0, true);
#undef CE_CMVN_LAUNCH
  CE_LAUNCHED();
  HostMark("cmvn: launch");
  return CE_GPU_OK;
}

}  // namespace ce
