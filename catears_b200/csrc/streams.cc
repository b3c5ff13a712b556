// streams.cc -- live utterances whose state stays on the device (include/ce_gpu.h,
// ce_gpu_streams_*).
//
// What the reference keeps per utterance in Fbank::Instance (the samples that do not fill a frame
// yet, src/fbank.cc:308-313), in CMVN (the running sums and the last 600 raw frames,
// src/cmvn.cc:35-68) and in AcousticModel::Instance (the frames still waiting for their right
// context, src/am.cc:115-142) lives in per-slot device buffers here.  One process call takes
// whatever PCM has arrived for a set of slots and runs ONE fbank, ONE CMVN and ONE acoustic-model
// pass for all of them; only the new samples go up and only the finished rows come down.  The
// host keeps counters, never data.  The kernels and their inputs are those of the batch entry
// points, so the rows equal ce_host::StreamBatch's (host-resident state) bit for bit.
#include <string.h>

#include <algorithm>
#include <chrono>
#include <memory>
#include <vector>

#include "api_kernels.h"
#include "common.h"
#include "nnet.h"

using namespace ce;

struct ce_gpu_streams {
  ce_gpu_model *model = nullptr;
  int max_streams = 0;
  int mel = 0;
  struct Slot {
    bool in_use = false, started = false;
    bool fresh = true;       // the CMVN running sums have not been zeroed yet (src/cmvn.cc:20-25)
    int rem_n = 0;           // samples waiting for a full frame (< 400)
    int n_hist = 0;          // raw frames kept for the CMVN window (<= 600)
    int64_t t_base = 0;      // frames normalised so far
    int ctx_n = 0;           // frames waiting for their right context (<= L + R)
  };
  std::vector<Slot> slots;
  // per-slot state, slot i at i * stride
  DevBuf rem, hist, cmvn_state, ctx;
  // per-call scratch
  PinnedBuf pcm_host;
  DevBuf pcm_new, wave, raw, cm_in, cm_state, norm, x, ll, rows_stage;
  Table segs, fbank_chunks, cmvn_utts;
  // host time spent inside ce_gpu_streams_process (ce_gpu_streams_call_stats)
  int64_t stat_calls = 0;
  double stat_enqueue_us = 0, stat_total_us = 0;

  size_t rem_stride() const { return kFrameLen; }                          // int16
  size_t hist_stride() const { return (size_t)kCmvnWindow * mel; }         // float
  size_t ctx_stride() const { return (size_t)(model->left + model->right) * mel; }

  ~ce_gpu_streams() {
    if (!model) return;
    cudaSetDevice(model->device);
    cudaDeviceSynchronize();
    rem.Free(); hist.Free(); cmvn_state.Free(); ctx.Free();
    pcm_host.Free();
    pcm_new.Free(); wave.Free(); raw.Free(); cm_in.Free(); cm_state.Free(); norm.Free(); x.Free();
    ll.Free(); rows_stage.Free();
    segs.Free(); fbank_chunks.Free(); cmvn_utts.Free();
  }
};

namespace {

int64_t FramesOf(int64_t n) { return n < kFrameLen ? 0 : 1 + (n - kFrameLen) / kFrameShift; }

// What one process call will do to one slot, from the counters alone.
struct Plan {
  int slot;
  int new_n;                 // new samples
  int64_t total;             // rem + new
  int64_t T;                 // new frames
  int rem_after;
  bool starting;             // left padding goes in now
  bool closing;              // right padding goes in now
  int64_t have;              // frames in the AM input: ctx + [L] + T + [R]
  int64_t n_ready;           // rows that come out
};

int MakePlan(const ce_gpu_streams *S, const int *slots, int n, const int *n_samples,
             const unsigned char *eos, std::vector<Plan> *plan) {
  const int L = S->model->left, R = S->model->right;
  std::vector<char> seen(S->max_streams, 0);
  plan->resize(n);
  for (int i = 0; i < n; ++i) {
    const int id = slots[i];
    if (id < 0 || id >= S->max_streams || !S->slots[id].in_use || seen[id] || n_samples[i] < 0) {
      SetError("ce_gpu_streams: entry %d: slot %d is not open, listed twice, or has a negative sample count", i, id);
      return CE_GPU_EINVAL;
    }
    seen[id] = 1;
    const ce_gpu_streams::Slot &st = S->slots[id];
    Plan &p = (*plan)[i];
    p.slot = id;
    p.new_n = n_samples[i];
    p.total = (int64_t)st.rem_n + p.new_n;
    p.T = FramesOf(p.total);
    p.rem_after = (int)(p.total - kFrameShift * p.T);
    p.starting = p.T > 0 && !st.started;
    int64_t c = st.ctx_n + (p.starting ? L : 0) + p.T;
    p.closing = eos && eos[i] && c > 0;                  // right padding, src/am.cc:152-155
    if (p.closing) c += R;
    p.have = c;
    p.n_ready = std::max<int64_t>(0, c - L - R);
  }
  return CE_GPU_OK;
}

}  // namespace

extern "C" {

ce_gpu_streams_t *ce_gpu_streams_create(ce_gpu_model_t *m, int max_streams) {
  if (!m || max_streams < 1) {
    SetError("ce_gpu_streams_create: bad arguments");
    return nullptr;
  }
  if (UseDevice(m->device) != CE_GPU_OK) return nullptr;
  std::unique_ptr<ce_gpu_streams> S(new ce_gpu_streams());
  S->model = m;
  S->max_streams = max_streams;
  S->mel = m->prog.feat_dim;
  S->slots.resize(max_streams);
  const size_t n = (size_t)max_streams;
  if (S->rem.Reserve(sizeof(int16_t) * n * S->rem_stride()) != CE_GPU_OK ||
      S->hist.Reserve(sizeof(float) * n * S->hist_stride()) != CE_GPU_OK ||
      S->cmvn_state.Reserve(sizeof(float) * n * S->mel) != CE_GPU_OK ||
      S->ctx.Reserve(sizeof(float) * n * std::max<size_t>(S->ctx_stride(), 1)) != CE_GPU_OK)
    return nullptr;
  return S.release();
}

void ce_gpu_streams_free(ce_gpu_streams_t *S) { delete S; }

int ce_gpu_streams_open(ce_gpu_streams_t *S) {
  if (!S) {
    SetError("ce_gpu_streams_open: null set");
    return CE_GPU_EINVAL;
  }
  for (int i = 0; i < S->max_streams; ++i) {
    if (S->slots[i].in_use) continue;
    S->slots[i] = ce_gpu_streams::Slot();              // buffers are length-tracked: nothing to clear
    S->slots[i].in_use = true;
    return i;
  }
  SetError("ce_gpu_streams_open: all %d slots are in use", S->max_streams);
  return CE_GPU_ENOMEM;
}

int ce_gpu_streams_call_stats(ce_gpu_streams_t *S, int64_t *calls, double *enqueue_us, double *total_us, int reset) {
  if (!S) {
    SetError("ce_gpu_streams_call_stats: null set");
    return CE_GPU_EINVAL;
  }
  if (calls) *calls = S->stat_calls;
  if (enqueue_us) *enqueue_us = S->stat_enqueue_us;
  if (total_us) *total_us = S->stat_total_us;
  if (reset) {
    S->stat_calls = 0;
    S->stat_enqueue_us = S->stat_total_us = 0;
  }
  return CE_GPU_OK;
}

int64_t ce_gpu_streams_rows_ready(const ce_gpu_streams_t *S, const int *slots, int n, const int *n_samples,
                                  const unsigned char *end_of_stream) {
  if (!S || n < 0 || (n > 0 && (!slots || !n_samples))) {
    SetError("ce_gpu_streams_rows_ready: bad arguments");
    return CE_GPU_EINVAL;
  }
  std::vector<Plan> plan;
  CE_CHECK(MakePlan(S, slots, n, n_samples, end_of_stream, &plan));
  int64_t rows = 0;
  for (const Plan &p : plan) rows += p.n_ready;
  return rows;
}

int ce_gpu_streams_process(ce_gpu_streams_t *S, const int *slots, int n, const int16_t *const *pcm,
                           const int *n_samples, const unsigned char *end_of_stream, float *rows,
                           int64_t rows_cap, int64_t *row_offsets, void *stream) {
  if (!S || n < 0 || (n > 0 && (!slots || !n_samples || !row_offsets))) {
    SetError("ce_gpu_streams_process: bad arguments");
    return CE_GPU_EINVAL;
  }
  if (row_offsets) row_offsets[0] = 0;
  if (n == 0) return CE_GPU_OK;
  auto now = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_start = now();
  ce_gpu_model *m = S->model;
  const int L = m->left, R = m->right, mel = S->mel, W = m->out_words();
  std::vector<Plan> plan;
  CE_CHECK(MakePlan(S, slots, n, n_samples, end_of_stream, &plan));
  int64_t new_total = 0, wave_total = 0, frames_total = 0, rows_total = 0, x_total = 0;
  for (int i = 0; i < n; ++i) {
    const Plan &p = plan[i];
    if (p.new_n > 0 && (!pcm || !pcm[i])) {
      SetError("ce_gpu_streams_process: entry %d has samples but no pcm pointer", i);
      return CE_GPU_EINVAL;
    }
    new_total += p.new_n;
    wave_total += p.total;
    frames_total += p.T;
    rows_total += p.n_ready;
    x_total += p.have;
    row_offsets[i + 1] = rows_total;
  }
  if (rows_total > rows_cap || (rows_total > 0 && !rows)) {
    SetError("ce_gpu_streams_process: %lld rows are ready but the buffer holds %lld "
             "(ce_gpu_streams_rows_ready tells beforehand); nothing was changed",
             (long long)rows_total, (long long)rows_cap);
    return CE_GPU_EINVAL;
  }
  CE_CHECK(UseDevice(m->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  // Every batched copy of the call is planned first (they are executed in six groups, in stream
  // order around the three compute passes).  One CTA copies one entry, so long copies (a minute of
  // rows, a full CMVN history) are cut into 256 KB pieces.
  std::vector<SegCopy> seg;
  auto add = [&seg](const void *src, void *dst, size_t bytes, uint32_t repeat = 1) {
    if (bytes == 0 || repeat == 0) return;
    constexpr size_t kPiece = 256 * 1024;
    if (repeat > 1 || bytes <= kPiece) {
      seg.push_back(SegCopy{src, dst, (uint32_t)bytes, repeat});
      return;
    }
    for (size_t o = 0; o < bytes; o += kPiece)
      seg.push_back(SegCopy{static_cast<const char *>(src) + o, static_cast<char *>(dst) + o,
                            (uint32_t)std::min(kPiece, bytes - o), 1});
  };
  CE_CHECK(S->pcm_new.Reserve(sizeof(int16_t) * (size_t)std::max<int64_t>(new_total, 1)));
  CE_CHECK(S->wave.Reserve(sizeof(int16_t) * (size_t)std::max<int64_t>(wave_total, 1)));
  CE_CHECK(S->raw.Reserve(sizeof(float) * (size_t)std::max<int64_t>(frames_total, 1) * mel));
  CE_CHECK(S->x.Reserve(sizeof(float) * (size_t)std::max<int64_t>(x_total, 1) * mel));

  for (int i = 0; i < n; ++i) {                           // new utterances: sums start at zero
    ce_gpu_streams::Slot &st = S->slots[plan[i].slot];
    if (!st.fresh) continue;
    CE_CUDA(cudaMemsetAsync(S->cmvn_state.as<float>() + (size_t)plan[i].slot * mel, 0, sizeof(float) * mel, s));
    st.fresh = false;
  }
  // ---- 1. new samples up; wave_i = [remainder | new samples] ----
  if (new_total > 0) {
    CE_CHECK(S->pcm_host.Acquire(sizeof(int16_t) * (size_t)new_total));
    int64_t o = 0;
    for (int i = 0; i < n; ++i) {
      if (plan[i].new_n > 0) memcpy(S->pcm_host.as<int16_t>() + o, pcm[i], sizeof(int16_t) * (size_t)plan[i].new_n);
      o += plan[i].new_n;
    }
    CE_CUDA(cudaMemcpyAsync(S->pcm_new.ptr, S->pcm_host.ptr, sizeof(int16_t) * (size_t)new_total,
                            cudaMemcpyHostToDevice, s));
    CE_CHECK(S->pcm_host.Release(s));
  }
  std::vector<int64_t> soff(n + 1, 0), foff(n + 1, 0);
  const size_t g0 = seg.size();
  {
    int64_t o_new = 0;
    for (int i = 0; i < n; ++i) {
      const Plan &p = plan[i];
      const ce_gpu_streams::Slot &st = S->slots[p.slot];
      int16_t *w = S->wave.as<int16_t>() + soff[i];
      add(S->rem.as<int16_t>() + (size_t)p.slot * S->rem_stride(), w, sizeof(int16_t) * (size_t)st.rem_n);
      add(S->pcm_new.as<int16_t>() + o_new, w + st.rem_n, sizeof(int16_t) * (size_t)p.new_n);
      o_new += p.new_n;
      soff[i + 1] = soff[i] + p.total;
      foff[i + 1] = foff[i] + p.T;
    }
  }
  // ---- 2. after the fbank: the new remainder back into the slot ----
  const size_t g1 = seg.size();
  for (int i = 0; i < n; ++i) {
    const Plan &p = plan[i];
    add(S->wave.as<int16_t>() + soff[i] + kFrameShift * p.T,
        S->rem.as<int16_t>() + (size_t)p.slot * S->rem_stride(), sizeof(int16_t) * (size_t)p.rem_after);
  }
  // ---- 3. CMVN input [history | new raw frames] and running sums gathered per call ----
  const bool cmvn = m->has_cmvn && frames_total > 0;
  std::vector<int64_t> coff(n + 1, 0), tbase(n, 0), cm_out_off(n + 1, 0);
  std::vector<int32_t> nhist(n, 0);
  const size_t g2 = seg.size();
  if (cmvn) {
    int64_t cm_rows = 0;
    for (int i = 0; i < n; ++i) cm_rows += S->slots[plan[i].slot].n_hist + plan[i].T;
    CE_CHECK(S->cm_in.Reserve(sizeof(float) * (size_t)cm_rows * mel));
    CE_CHECK(S->cm_state.Reserve(sizeof(float) * (size_t)n * mel));
    CE_CHECK(S->norm.Reserve(sizeof(float) * (size_t)frames_total * mel));
    for (int i = 0; i < n; ++i) {
      const Plan &p = plan[i];
      const ce_gpu_streams::Slot &st = S->slots[p.slot];
      float *in = S->cm_in.as<float>() + coff[i] * mel;
      add(S->hist.as<float>() + (size_t)p.slot * S->hist_stride(), in, sizeof(float) * (size_t)st.n_hist * mel);
      add(S->raw.as<float>() + foff[i] * mel, in + (size_t)st.n_hist * mel, sizeof(float) * (size_t)p.T * mel);
      add(S->cmvn_state.as<float>() + (size_t)p.slot * mel, S->cm_state.as<float>() + (size_t)i * mel,
          sizeof(float) * mel);
      nhist[i] = st.n_hist;
      tbase[i] = st.t_base;
      coff[i + 1] = coff[i] + st.n_hist + p.T;
      cm_out_off[i + 1] = cm_out_off[i] + p.T;
    }
  }
  // ---- 4. after the CMVN: sums and the last 600 raw frames back into the slot ----
  const size_t g3 = seg.size();
  if (cmvn) {
    for (int i = 0; i < n; ++i) {
      const Plan &p = plan[i];
      const ce_gpu_streams::Slot &st = S->slots[p.slot];
      const int64_t have = st.n_hist + p.T, keep = std::min<int64_t>(have, kCmvnWindow);
      add(S->cm_state.as<float>() + (size_t)i * mel, S->cmvn_state.as<float>() + (size_t)p.slot * mel,
          sizeof(float) * mel);
      if (p.T > 0)
        add(S->cm_in.as<float>() + (coff[i] + have - keep) * mel,
            S->hist.as<float>() + (size_t)p.slot * S->hist_stride(), sizeof(float) * (size_t)keep * mel);
    }
  }
  // ---- 5. AM input x_i = [waiting frames | L x first | new frames | R x last]; the utterances
  //         that produce rows come first so that they are contiguous for the forward pass ----
  const float *feat = cmvn ? S->norm.as<float>() : S->raw.as<float>();
  std::vector<int64_t> xpos(n, 0), xoff(1, 0);
  std::vector<int> ready_idx;
  {
    int64_t o = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int i = 0; i < n; ++i) {
        if ((plan[i].n_ready > 0) != (pass == 0)) continue;
        xpos[i] = o;
        o += plan[i].have;
        if (pass == 0) {
          xoff.push_back(o);
          ready_idx.push_back(i);
        }
      }
    }
  }
  const size_t g4 = seg.size();
  for (int i = 0; i < n; ++i) {
    const Plan &p = plan[i];
    const ce_gpu_streams::Slot &st = S->slots[p.slot];
    float *x = S->x.as<float>() + xpos[i] * mel;
    const float *ctx = S->ctx.as<float>() + (size_t)p.slot * S->ctx_stride();
    const float *nf = feat + foff[i] * mel;
    const size_t fb = sizeof(float) * mel;
    add(ctx, x, fb * st.ctx_n);
    x += (size_t)st.ctx_n * mel;
    if (p.starting) {                                    // left padding, src/am.cc:119-124
      add(nf, x, fb, L);
      x += (size_t)L * mel;
    }
    add(nf, x, fb * p.T);
    x += (size_t)p.T * mel;
    if (p.closing) add(p.T > 0 ? nf + (p.T - 1) * mel : ctx + (size_t)(st.ctx_n - 1) * mel, x, fb, R);
  }
  // ---- 6. after the forward pass: rows out, the frames still waiting back into the slot ----
  const size_t g5 = seg.size();
  const bool rows_host = rows && !IsDevicePtr(rows);
  float *rows_dev = rows;
  if (rows_total > 0) {
    CE_CHECK(S->ll.Reserve(sizeof(float) * (size_t)rows_total * W));
    if (rows_host) {
      CE_CHECK(S->rows_stage.Reserve(sizeof(float) * (size_t)rows_total * W));
      rows_dev = S->rows_stage.as<float>();
    }
  }
  {                                                      // the forward pass packs block k's rows behind block k-1's
    int64_t o = 0;
    for (size_t k = 0; k < ready_idx.size(); ++k) {
      const int i = ready_idx[k];
      add(S->ll.as<float>() + o * W, rows_dev + row_offsets[i] * W, sizeof(float) * (size_t)plan[i].n_ready * W);
      o += plan[i].n_ready;
    }
  }
  for (int i = 0; i < n; ++i) {
    const Plan &p = plan[i];
    if (end_of_stream && end_of_stream[i]) continue;     // the slot is closed below
    add(S->x.as<float>() + (xpos[i] + p.n_ready) * mel, S->ctx.as<float>() + (size_t)p.slot * S->ctx_stride(),
        sizeof(float) * (size_t)(p.have - p.n_ready) * mel);
  }
  const size_t g6 = seg.size();

  CE_CHECK(S->segs.Acquire(sizeof(SegCopy) * std::max<size_t>(seg.size(), 1)));
  if (!seg.empty()) memcpy(S->segs.host<SegCopy>(), seg.data(), sizeof(SegCopy) * seg.size());
  CE_CHECK(S->segs.Upload(sizeof(SegCopy) * std::max<size_t>(seg.size(), 1), s));
  const SegCopy *d = S->segs.dev<SegCopy>();
  auto run = [&](size_t a, size_t b) { return SegCopyLaunch(d + a, (int)(b - a), s); };

  CE_CHECK(run(g0, g1));
  if (frames_total > 0)
    CE_CHECK(FbankLaunch(S->wave.as<int16_t>(), wave_total, soff.data(), foff.data(), n, mel,
                         S->raw.as<float>(), mel, &S->fbank_chunks, s));
  CE_CHECK(run(g1, g2));
  if (cmvn) {
    CE_CHECK(run(g2, g3));
    CmvnResume resume = {nhist.data(), tbase.data(), S->cm_state.as<float>()};
    CE_CHECK(CmvnLaunch(m->cmvn_dev.as<float>(), m->cmvn_host[mel], S->cm_in.as<float>(), coff.data(),
                        cm_out_off.data(), n, mel, 0, 0, S->norm.as<float>(), mel, &S->cmvn_utts, s, &resume));
    CE_CHECK(run(g3, g4));
  }
  CE_CHECK(run(g4, g5));
  if (rows_total > 0)
    CE_CHECK(NnetForward(m, S->x.as<float>(), xoff.data(), (int)ready_idx.size(), /*apply_cmvn=*/false,
                         S->ll.as<float>(), nullptr, s, /*contexted=*/true));   // blocks carry their context
  CE_CHECK(run(g5, g6));
  const double t_enqueued = now();                         // everything but the rows' way home is queued
  if (rows_host && rows_total > 0)
    CE_CUDA(cudaMemcpyAsync(rows, rows_dev, sizeof(float) * (size_t)rows_total * W, cudaMemcpyDeviceToHost, s));

  // ---- counters ----
  for (int i = 0; i < n; ++i) {
    const Plan &p = plan[i];
    ce_gpu_streams::Slot &st = S->slots[p.slot];
    st.rem_n = p.rem_after;
    if (m->has_cmvn) {
      st.n_hist = (int)std::min<int64_t>(st.n_hist + p.T, kCmvnWindow);
      st.t_base += p.T;
    }
    st.started = st.started || p.starting;
    st.ctx_n = (int)(p.have - p.n_ready);
    if (end_of_stream && end_of_stream[i]) st = ce_gpu_streams::Slot();   // free again
  }
  if (rows_host) CE_CUDA(cudaStreamSynchronize(s));        // host rows are complete on return
  S->stat_calls += 1;
  S->stat_enqueue_us += t_enqueued - t_start;
  S->stat_total_us += now() - t_start;
  return CE_GPU_OK;
}

}  // extern "C"
