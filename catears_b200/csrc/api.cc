// api.cc -- the C ABI of include/ce_gpu.h.
//
// Every entry point validates its arguments, makes the device current, stages host buffers
// through device workspace and calls the CUDA launchers.  There is no CPU implementation
// behind any of them: without a usable sm_100 device they fail with CE_GPU_ENODEVICE.
#include <math.h>
#include <string.h>

#include <exception>
#include <map>
#include <memory>
#include <string>

#include "api_kernels.h"
#include "common.h"
#include "gemm.h"
#include "model.h"
#include "nnet.h"
#include "nnet_kernels.h"

using namespace ce;

namespace {

// Workspace of the handle-less stage entry points (one per host thread and device).
struct StageWs {
  DevBuf in, in2, out, out2, tmp[8];
  Table t0, t1;
};

StageWs *GetWs(int device) {
  static thread_local std::map<int, std::unique_ptr<StageWs>> ws;
  auto &p = ws[device];
  if (!p) p.reset(new StageWs());
  return p.get();
}

inline int RoundUp(int v, int m) { return (v + m - 1) / m * m; }

int CheckOffsets(const int64_t *off, int n, const char *what) {
  if (n < 0 || (n > 0 && off == nullptr)) {
    SetError("%s: bad utterance count / offsets", what);
    return CE_GPU_EINVAL;
  }
  for (int u = 0; u < n; ++u) {
    if (off[u + 1] < off[u] || off[u] < 0) {
      SetError("%s: offsets must be non-negative and non-decreasing (utterance %d)", what, u);
      return CE_GPU_EINVAL;
    }
  }
  return CE_GPU_OK;
}

// Copies a device result into the caller's buffer when that is host memory.
int Deliver(void *user, const void *dev, size_t bytes, cudaStream_t s) {
  if (user == dev || bytes == 0) return CE_GPU_OK;
  return StageOut(user, dev, bytes, s);
}

}  // namespace

extern "C" {

const char *ce_gpu_last_error(void) { return LastError(); }

int ce_gpu_device_count(void) { return DeviceCount(); }

int ce_gpu_version(void) { return 200; }   // 0.2.0

int64_t ce_gpu_launch_count(int reset) {
  int64_t v = LaunchCounter();
  if (reset) LaunchCounter() = 0;
  return v;
}

int ce_gpu_profile_enable(int on) {
  ProfEnable(on != 0);
  return CE_GPU_OK;
}

int ce_gpu_profile_read(double *ms, int64_t *launches) {
  if (!ms || !launches) {
    SetError("ce_gpu_profile_read: null output");
    return CE_GPU_EINVAL;
  }
  return ProfRead(ms, launches);
}

int ce_gpu_profile_trace(int cap, int32_t *cat, double *t0_ms, double *t1_ms) {
  if (cap < 0 || !cat || !t0_ms || !t1_ms) {
    SetError("ce_gpu_profile_trace: bad arguments");
    return CE_GPU_EINVAL;
  }
  return ProfTrace(cap, cat, t0_ms, t1_ms);
}

// ---- model ----------------------------------------------------------------------------

namespace {
ce_gpu_model_t *ModelLoad(const char *nnet_path, const char *prior_path, const char *cmvn_stats_path,
                          int left_context, int right_context, int precision, int device);
}  // namespace

// No C++ exception may cross the C ABI: an allocation failure while reading or packing a model
// (std::bad_alloc, std::length_error) becomes NULL + a message like every other load error.
ce_gpu_model_t *ce_gpu_model_load(const char *nnet_path, const char *prior_path,
                                  const char *cmvn_stats_path, int left_context,
                                  int right_context, int precision, int device) {
  try {
    return ModelLoad(nnet_path, prior_path, cmvn_stats_path, left_context, right_context, precision, device);
  } catch (const std::exception &e) {
    SetError("ce_gpu_model_load: %s", e.what());
    return nullptr;
  }
}

namespace {
ce_gpu_model_t *ModelLoad(const char *nnet_path, const char *prior_path, const char *cmvn_stats_path,
                          int left_context, int right_context, int precision, int device) {
  if (!nnet_path || !prior_path || left_context < 0 || right_context < 0) {
    SetError("ce_gpu_model_load: bad arguments");
    return nullptr;
  }
  HostNnet nn;
  std::vector<float> prior, cmvn;
  if (ReadNnetFile(nnet_path, &nn) != CE_GPU_OK) return nullptr;
  if (ReadVectorFile(prior_path, &prior) != CE_GPU_OK) return nullptr;
  if (cmvn_stats_path && ReadVectorFile(cmvn_stats_path, &cmvn) != CE_GPU_OK) return nullptr;
  if (nn.left_context != left_context || nn.right_context != right_context) {
    // Q10: the reference ignores the header values (src/am.cc:48-50); so do we, loudly.
    fprintf(stderr, "ce_gpu: note: NN02 header contexts %d/%d differ from the configured %d/%d\n",
            nn.left_context, nn.right_context, left_context, right_context);
  }
  if (UseDevice(device) != CE_GPU_OK) return nullptr;
  std::unique_ptr<ce_gpu_model> m(new ce_gpu_model());
  if (ModelBuild(nn, prior, cmvn_stats_path ? &cmvn : nullptr, left_context, right_context,
                 precision, device, m.get()) != CE_GPU_OK)
    return nullptr;
  return m.release();
}
}  // namespace

ce_gpu_model_t *ce_gpu_model_load_config(const char *config_path, int precision, int device) {
  if (!config_path) {
    SetError("ce_gpu_model_load_config: null path");
    return nullptr;
  }
  std::map<std::string, std::string> kv;
  std::string dir;
  if (ReadConfigFile(config_path, &kv, &dir) != CE_GPU_OK) return nullptr;
  auto path_of = [&](const char *key, std::string *out) -> bool {
    auto it = kv.find(key);
    if (it == kv.end()) return false;
    *out = it->second[0] == '/' ? it->second : dir + it->second;   // configuration.cc:56-70
    return true;
  };
  std::string nnet, prior, cmvn;
  if (!path_of("nnet", &nnet) || !path_of("prior", &prior)) {
    SetError("%s: keys 'nnet' and 'prior' are required", config_path);   // am.cc:31,41
    return nullptr;
  }
  int lr[2];
  const char *keys[2] = {"left_context", "right_context"};
  for (int i = 0; i < 2; ++i) {
    auto it = kv.find(keys[i]);
    if (it == kv.end()) {
      SetError("%s: key '%s' is required", config_path, keys[i]);          // am.cc:48-49
      return nullptr;
    }
    lr[i] = atoi(it->second.c_str());
  }
  const bool has_cmvn = path_of("cmvn_stats", &cmvn);
  ce_gpu_model_t *m = ce_gpu_model_load(nnet.c_str(), prior.c_str(), has_cmvn ? cmvn.c_str() : nullptr,
                                        lr[0], lr[1], precision, device);
  if (m) {
    auto it = kv.find("num_pdfs");
    if (it != kv.end() && atoi(it->second.c_str()) != m->prog.num_pdfs) {
      SetError("%s: num_pdfs = %s but the nnet has %d outputs", config_path, it->second.c_str(),
               m->prog.num_pdfs);
      delete m;
      return nullptr;
    }
  }
  return m;
}

void ce_gpu_model_free(ce_gpu_model_t *m) { delete m; }

int ce_gpu_model_info(const ce_gpu_model_t *m, int *num_pdfs, int *left_context,
                      int *right_context, int *feat_dim, int *precision, int *device) {
  if (!m) {
    SetError("ce_gpu_model_info: null model");
    return CE_GPU_EINVAL;
  }
  if (num_pdfs) *num_pdfs = m->prog.num_pdfs;
  if (left_context) *left_context = m->left;
  if (right_context) *right_context = m->right;
  if (feat_dim) *feat_dim = m->prog.feat_dim;
  if (precision) *precision = m->precision;
  if (device) *device = m->device;
  return CE_GPU_OK;
}

// ---- stages ---------------------------------------------------------------------------

int64_t ce_gpu_frame_offsets(const int64_t *utt_sample_offsets, int n_utts,
                             int64_t *utt_frame_offsets) {
  int rc = CheckOffsets(utt_sample_offsets, n_utts, "ce_gpu_frame_offsets");
  if (rc != CE_GPU_OK) return rc;
  if (!utt_frame_offsets) {
    SetError("ce_gpu_frame_offsets: null output");
    return CE_GPU_EINVAL;
  }
  int64_t total = 0;
  utt_frame_offsets[0] = 0;
  for (int u = 0; u < n_utts; ++u) {
    total += NumFrames(utt_sample_offsets[u + 1] - utt_sample_offsets[u]);
    utt_frame_offsets[u + 1] = total;
  }
  return total;
}

int ce_gpu_fbank(const int16_t *pcm, const int64_t *utt_sample_offsets, int n_utts, int num_mel,
                 float *feats, int device, void *stream) {
  CE_CHECK(CheckOffsets(utt_sample_offsets, n_utts, "ce_gpu_fbank"));
  if (n_utts == 0) return CE_GPU_OK;
  if (!pcm || !feats) {
    SetError("ce_gpu_fbank: null buffer");
    return CE_GPU_EINVAL;
  }
  CE_CHECK(UseDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  StageWs *ws = GetWs(device);
  std::vector<int64_t> foff(n_utts + 1);
  const int64_t total_frames = ce_gpu_frame_offsets(utt_sample_offsets, n_utts, foff.data());
  if (total_frames <= 0) return (int)std::min<int64_t>(total_frames, 0);
  const int64_t base = utt_sample_offsets[0], total_samples = utt_sample_offsets[n_utts];
  const void *pcm_dev = nullptr;
  CE_CHECK(StageIn(pcm + base, sizeof(int16_t) * (size_t)(total_samples - base), &ws->in, s, &pcm_dev));
  std::vector<int64_t> soff(n_utts + 1);
  for (int u = 0; u <= n_utts; ++u) soff[u] = utt_sample_offsets[u] - base;
  float *out_dev = feats;
  const size_t out_bytes = sizeof(float) * (size_t)total_frames * num_mel;
  if (!IsDevicePtr(feats)) {
    CE_CHECK(ws->out.Reserve(out_bytes));
    out_dev = ws->out.as<float>();
  }
  CE_CHECK(FbankLaunch(static_cast<const int16_t *>(pcm_dev), total_samples - base, soff.data(),
                       foff.data(), n_utts, num_mel, out_dev, num_mel, &ws->t0, s));
  return Deliver(feats, out_dev, out_bytes, s);
}

int ce_gpu_cmvn(const float *global_stats, const float *feats, const int64_t *utt_frame_offsets,
                int n_utts, int num_mel, float *out, int device, void *stream) {
  CE_CHECK(CheckOffsets(utt_frame_offsets, n_utts, "ce_gpu_cmvn"));
  if (n_utts == 0) return CE_GPU_OK;
  if (!global_stats || !feats || !out || num_mel < 1) {
    SetError("ce_gpu_cmvn: null buffer / bad num_mel");
    return CE_GPU_EINVAL;
  }
  CE_CHECK(UseDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  StageWs *ws = GetWs(device);
  const int64_t total = utt_frame_offsets[n_utts];
  if (total == 0) return CE_GPU_OK;
  const size_t bytes = sizeof(float) * (size_t)total * num_mel;
  const void *in_dev = nullptr;
  CE_CHECK(StageIn(feats, bytes, &ws->in, s, &in_dev));
  CE_CHECK(ws->in2.Reserve(sizeof(float) * (num_mel + 1)));
  CE_CUDA(cudaMemcpyAsync(ws->in2.ptr, global_stats, sizeof(float) * (num_mel + 1),
                          cudaMemcpyHostToDevice, s));
  CE_CUDA(cudaStreamSynchronize(s));                     // global_stats may be a stack array
  float *out_dev = out;
  if (!IsDevicePtr(out)) {
    CE_CHECK(ws->out.Reserve(bytes));
    out_dev = ws->out.as<float>();
  } else if (out == feats) {
    // in place on the device: the window term x_{t-600} must be the ORIGINAL value -> copy first
    CE_CHECK(ws->out2.Reserve(bytes));
    CE_CUDA(cudaMemcpyAsync(ws->out2.ptr, feats, bytes, cudaMemcpyDeviceToDevice, s));
    in_dev = ws->out2.ptr;
  }
  CE_CHECK(CmvnLaunch(ws->in2.as<float>(), global_stats[num_mel], static_cast<const float *>(in_dev),
                      utt_frame_offsets, utt_frame_offsets, n_utts, num_mel, 0, 0, out_dev, num_mel,
                      &ws->t0, s));
  return Deliver(out, out_dev, bytes, s);
}

int ce_gpu_cmvn_stream(const float *global_stats, const float *feats, const int64_t *utt_frame_offsets,
                       const int32_t *n_hist, const int64_t *t_base, float *state, int n_utts,
                       int num_mel, float *out, int device, void *stream) {
  CE_CHECK(CheckOffsets(utt_frame_offsets, n_utts, "ce_gpu_cmvn_stream"));
  if (n_utts == 0) return CE_GPU_OK;
  if (!global_stats || !feats || !out || !n_hist || !t_base || !state || num_mel < 1) {
    SetError("ce_gpu_cmvn_stream: null buffer / bad num_mel");
    return CE_GPU_EINVAL;
  }
  std::vector<int64_t> out_off(n_utts + 1, 0);
  for (int u = 0; u < n_utts; ++u) {
    const int64_t rows = utt_frame_offsets[u + 1] - utt_frame_offsets[u];
    if (n_hist[u] < 0 || n_hist[u] > rows || t_base[u] < n_hist[u] ||
        n_hist[u] != std::min<int64_t>(t_base[u], kCmvnWindow)) {
      SetError("ce_gpu_cmvn_stream: utterance %d must bring min(t_base, %d) history frames (has %d, t_base %lld)",
               u, kCmvnWindow, n_hist[u], (long long)t_base[u]);
      return CE_GPU_EINVAL;
    }
    out_off[u + 1] = out_off[u] + (rows - n_hist[u]);
  }
  const int64_t total_in = utt_frame_offsets[n_utts], total_out = out_off[n_utts];
  if (total_out == 0) return CE_GPU_OK;
  CE_CHECK(UseDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  StageWs *ws = GetWs(device);
  const void *in_dev = nullptr;
  CE_CHECK(StageIn(feats, sizeof(float) * (size_t)total_in * num_mel, &ws->in, s, &in_dev));
  CE_CHECK(ws->in2.Reserve(sizeof(float) * (num_mel + 1)));
  CE_CUDA(cudaMemcpyAsync(ws->in2.ptr, global_stats, sizeof(float) * (num_mel + 1), cudaMemcpyHostToDevice, s));
  const size_t state_bytes = sizeof(float) * (size_t)n_utts * num_mel;
  CE_CHECK(ws->tmp[0].Reserve(state_bytes));
  CE_CUDA(cudaMemcpyAsync(ws->tmp[0].ptr, state, state_bytes, cudaMemcpyHostToDevice, s));
  CE_CUDA(cudaStreamSynchronize(s));                     // global_stats / state may be reused by the caller
  const size_t out_bytes = sizeof(float) * (size_t)total_out * num_mel;
  float *out_dev = out;
  if (!IsDevicePtr(out)) {
    CE_CHECK(ws->out.Reserve(out_bytes));
    out_dev = ws->out.as<float>();
  }
  CmvnResume resume = {n_hist, t_base, ws->tmp[0].as<float>()};
  CE_CHECK(CmvnLaunch(ws->in2.as<float>(), global_stats[num_mel], static_cast<const float *>(in_dev),
                      utt_frame_offsets, out_off.data(), n_utts, num_mel, 0, 0, out_dev, num_mel, &ws->t0, s,
                      &resume));
  CE_CUDA(cudaMemcpyAsync(state, ws->tmp[0].ptr, state_bytes, cudaMemcpyDeviceToHost, s));
  CE_CHECK(Deliver(out, out_dev, out_bytes, s));
  CE_CUDA(cudaStreamSynchronize(s));                     // the state is complete on return
  return CE_GPU_OK;
}

int ce_gpu_rfft512(const float *in, int n_frames, float *out, int device, void *stream) {
  if (n_frames < 0 || (n_frames > 0 && (!in || !out))) {
    SetError("ce_gpu_rfft512: bad arguments");
    return CE_GPU_EINVAL;
  }
  if (n_frames == 0) return CE_GPU_OK;
  CE_CHECK(UseDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  StageWs *ws = GetWs(device);
  const size_t bytes = sizeof(float) * 512 * (size_t)n_frames;
  const void *in_dev = nullptr;
  CE_CHECK(StageIn(in, bytes, &ws->in, s, &in_dev));
  float *out_dev = out;
  if (!IsDevicePtr(out)) {
    CE_CHECK(ws->out.Reserve(bytes));
    out_dev = ws->out.as<float>();
  }
  CE_CHECK(Rfft512Launch(static_cast<const float *>(in_dev), n_frames, out_dev, s));
  return Deliver(out, out_dev, bytes, s);
}

int ce_gpu_nnet(ce_gpu_model_t *m, const float *feats, const int64_t *utt_frame_offsets,
                int n_utts, float *loglik, int32_t *argmax, void *stream) {
  if (!m) {
    SetError("ce_gpu_nnet: null model");
    return CE_GPU_EINVAL;
  }
  CE_CHECK(CheckOffsets(utt_frame_offsets, n_utts, "ce_gpu_nnet"));
  if (n_utts == 0 || utt_frame_offsets[n_utts] == utt_frame_offsets[0]) return CE_GPU_OK;
  if (!feats) {
    SetError("ce_gpu_nnet: null features");
    return CE_GPU_EINVAL;
  }
  CE_CHECK(UseDevice(m->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = utt_frame_offsets[n_utts];
  const void *f_dev = nullptr;
  CE_CHECK(StageIn(feats, sizeof(float) * (size_t)total * m->prog.feat_dim, &m->stage_feats, s, &f_dev));
  return NnetForward(m, static_cast<const float *>(f_dev), utt_frame_offsets, n_utts, false, loglik,
                     argmax, s);
}

int ce_gpu_nnet_chunks(ce_gpu_model_t *m, const float *feats, const int64_t *block_offsets, int n_blocks,
                       float *loglik, int32_t *argmax, void *stream) {
  if (!m) {
    SetError("ce_gpu_nnet_chunks: null model");
    return CE_GPU_EINVAL;
  }
  CE_CHECK(CheckOffsets(block_offsets, n_blocks, "ce_gpu_nnet_chunks"));
  if (n_blocks == 0 || block_offsets[n_blocks] == block_offsets[0]) return CE_GPU_OK;
  if (!feats) {
    SetError("ce_gpu_nnet_chunks: null features");
    return CE_GPU_EINVAL;
  }
  CE_CHECK(UseDevice(m->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = block_offsets[n_blocks];
  const void *f_dev = nullptr;
  CE_CHECK(StageIn(feats, sizeof(float) * (size_t)total * m->prog.feat_dim, &m->stage_feats, s, &f_dev));
  return NnetForward(m, static_cast<const float *>(f_dev), block_offsets, n_blocks, false, loglik, argmax, s,
                     /*contexted=*/true);
}

int ce_gpu_forward(ce_gpu_model_t *m, const int16_t *pcm, const int64_t *utt_sample_offsets,
                   int n_utts, float *loglik, int32_t *argmax, int64_t *utt_frame_offsets_out,
                   void *stream) {
  if (!m) {
    SetError("ce_gpu_forward: null model");
    return CE_GPU_EINVAL;
  }
  CE_CHECK(CheckOffsets(utt_sample_offsets, n_utts, "ce_gpu_forward"));
  std::vector<int64_t> foff(n_utts + 1, 0);
  const int64_t total_frames = ce_gpu_frame_offsets(utt_sample_offsets, n_utts, foff.data());
  if (utt_frame_offsets_out) memcpy(utt_frame_offsets_out, foff.data(), sizeof(int64_t) * (n_utts + 1));
  if (n_utts == 0 || total_frames == 0) return CE_GPU_OK;
  if (!pcm) {
    SetError("ce_gpu_forward: null pcm");
    return CE_GPU_EINVAL;
  }
  CE_CHECK(UseDevice(m->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t base = utt_sample_offsets[0], total_samples = utt_sample_offsets[n_utts];
  std::vector<int64_t> soff(n_utts + 1);
  for (int u = 0; u <= n_utts; ++u) soff[u] = utt_sample_offsets[u] - base;
  return PcmForward(m, pcm + base, total_samples - base, soff.data(), foff.data(), n_utts, loglik,
                    argmax, s);
}

int ce_gpu_model_set_output(ce_gpu_model_t *m, int mode, const int32_t *pdf_ids, int n) {
  if (!m) {
    SetError("ce_gpu_model_set_output: null model");
    return CE_GPU_EINVAL;
  }
  const int NP = m->prog.num_pdfs;
  if (mode == CE_GPU_OUTPUT_DENSE) {
    m->out_sel = ce::OutSel();
    return CE_GPU_OK;
  }
  if (mode != CE_GPU_OUTPUT_SUBSET && mode != CE_GPU_OUTPUT_TOPK) {
    SetError("ce_gpu_model_set_output: unknown mode %d", mode);
    return CE_GPU_EINVAL;
  }
  if (NP % 4 != 0 || NP > 4096) {
    SetError("ce_gpu_model_set_output: a selected output needs num_pdfs %% 4 == 0 and <= 4096 (have %d)", NP);
    return CE_GPU_EUNSUPPORTED;
  }
  if (mode == CE_GPU_OUTPUT_TOPK) {
    if (n < 1 || n > NP || n > ce::kMaxTopK) {
      SetError("ce_gpu_model_set_output: top-k needs 1 <= k <= min(num_pdfs, %d), got %d", ce::kMaxTopK, n);
      return CE_GPU_EINVAL;
    }
    m->out_sel.mode = ce::kOutTopK;
    m->out_sel.n = n;
    m->out_sel.ids = nullptr;
    return CE_GPU_OK;
  }
  if (!pdf_ids || n < 1) {
    SetError("ce_gpu_model_set_output: empty pdf subset");
    return CE_GPU_EINVAL;
  }
  for (int j = 0; j < n; ++j) {
    if (pdf_ids[j] < 0 || pdf_ids[j] >= NP) {
      SetError("ce_gpu_model_set_output: pdf_ids[%d] = %d is outside [0, %d)", j, pdf_ids[j], NP);
      return CE_GPU_EINVAL;
    }
  }
  CE_CHECK(UseDevice(m->device));
  CE_CUDA(cudaDeviceSynchronize());                      // no earlier call still reads the old list
  CE_CHECK(m->out_ids.Reserve(sizeof(int32_t) * (size_t)n));
  CE_CUDA(cudaMemcpy(m->out_ids.ptr, pdf_ids, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
  m->out_sel.mode = ce::kOutSubset;
  m->out_sel.n = n;
  m->out_sel.ids = m->out_ids.as<int32_t>();
  return CE_GPU_OK;
}

int ce_gpu_model_set_rows_callback(ce_gpu_model_t *m, ce_gpu_rows_ready_fn fn, void *user) {
  if (!m) {
    SetError("ce_gpu_model_set_rows_callback: null model");
    return CE_GPU_EINVAL;
  }
  m->rows_cb = fn;
  m->rows_cb_user = fn ? user : nullptr;
  return CE_GPU_OK;
}

int ce_gpu_model_output_width(const ce_gpu_model_t *m) {
  if (!m) {
    SetError("ce_gpu_model_output_width: null model");
    return CE_GPU_EINVAL;
  }
  return m->out_words();
}

int ce_gpu_nnet_keep_acc(ce_gpu_model_t *m, int linear_ordinal) {
  if (!m || linear_ordinal >= (int)m->blocks.size()) {
    SetError("ce_gpu_nnet_keep_acc: bad model / layer ordinal");
    return CE_GPU_EINVAL;
  }
  if (linear_ordinal >= 0 && m->kind != kKindI8) {
    SetError("ce_gpu_nnet_keep_acc: only int8 models have integer accumulators");
    return CE_GPU_EINVAL;
  }
  m->keep_acc = linear_ordinal;
  m->kept_valid = false;
  return CE_GPU_OK;
}

int ce_gpu_nnet_get_acc(ce_gpu_model_t *m, int utt, int32_t *acc, int64_t cap, int *rows, int *cols) {
  if (!m || !m->kept_valid || utt < 0 || utt >= (int)m->kept_rows.size()) {
    SetError("ce_gpu_nnet_get_acc: nothing kept (call ce_gpu_nnet_keep_acc, then a forward pass "
             "that fits one chunk)");
    return CE_GPU_EINVAL;
  }
  const int r = std::max(0, m->kept_rows[utt] - m->kept_lo - m->kept_hi);
  if (rows) *rows = r;
  if (cols) *cols = m->kept_cols;
  if (!acc) return CE_GPU_OK;
  if ((int64_t)r * m->kept_cols > cap) {
    SetError("ce_gpu_nnet_get_acc: buffer too small (%lld needed)", (long long)r * m->kept_cols);
    return CE_GPU_EINVAL;
  }
  CE_CHECK(UseDevice(m->device));
  const int32_t *src = m->acc_dump.as<int32_t>() + (int64_t)(m->kept_row_off[utt] + m->kept_lo) * m->kept_ld;
  CE_CUDA(cudaDeviceSynchronize());
  CE_CUDA(cudaMemcpy2D(acc, sizeof(int32_t) * m->kept_cols, src, sizeof(int32_t) * m->kept_ld,
                       sizeof(int32_t) * m->kept_cols, r, cudaMemcpyDeviceToHost));
  return CE_GPU_OK;
}

// The activation QuantizationParams (src/matrix.h:231-234) every Linear layer's Quantize produced
// for utterance `utt` of the last (one-chunk) forward call on an int8 model.
int ce_gpu_nnet_get_qparams(ce_gpu_model_t *m, int utt, float *scale, int32_t *zero_point, int cap) {
  if (!m || m->kind != ce::kKindI8 || utt < 0 || utt >= m->last_n_utts) {
    SetError("ce_gpu_nnet_get_qparams: needs an int8 model and an utterance of the last forward call");
    return CE_GPU_EINVAL;
  }
  const int nb = (int)m->blocks.size();
  if (cap < nb) {
    SetError("ce_gpu_nnet_get_qparams: %d layers, room for %d", nb, cap);
    return CE_GPU_EINVAL;
  }
  CE_CHECK(UseDevice(m->device));
  CE_CUDA(cudaDeviceSynchronize());
  for (int b = 0; b < nb; ++b) {
    ce::QParam q;
    CE_CUDA(cudaMemcpy(&q, m->ws[0].qparams.as<ce::QParam>() + (size_t)b * m->last_n_utts + utt, sizeof(q),
                       cudaMemcpyDeviceToHost));
    if (scale) scale[b] = q.scale;
    if (zero_point) zero_point[b] = q.zero_point;
  }
  return nb;
}

// ---- pinned host memory for callers that do not link the CUDA runtime -------------------------

void *ce_gpu_host_alloc(size_t bytes) {
  if (DeviceCount() < 1) {
    SetError("ce_gpu_host_alloc: no CUDA device");
    return nullptr;
  }
  void *p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    SetError("cudaHostAlloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}

void ce_gpu_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// ---- multi-GPU planning ---------------------------------------------------------------------

int ce_gpu_partition(const int64_t *utt_frame_offsets, int n_utts, int n_parts, int32_t *part_begin) {
  CE_CHECK(CheckOffsets(utt_frame_offsets, n_utts, "ce_gpu_partition"));
  if (n_parts < 1 || !part_begin) {
    SetError("ce_gpu_partition: bad arguments");
    return CE_GPU_EINVAL;
  }
  const int64_t base = n_utts > 0 ? utt_frame_offsets[0] : 0;
  const int64_t total = n_utts > 0 ? utt_frame_offsets[n_utts] - base : 0;
  part_begin[0] = 0;
  int u = 0;
  for (int p = 1; p < n_parts; ++p) {
    // first utterance boundary at or after the ideal cut p/n of the frames (ties: earlier)
    const int64_t cut = base + (total * p + n_parts / 2) / n_parts;
    while (u < n_utts && utt_frame_offsets[u] < cut) {
      // step to the boundary nearest to the cut
      if (utt_frame_offsets[u + 1] - cut > cut - utt_frame_offsets[u]) break;
      ++u;
    }
    part_begin[p] = u;
  }
  part_begin[n_parts] = n_utts;
  return CE_GPU_OK;
}

int ce_gpu_time_shards(int64_t total_frames, int n_parts, int left_context, int right_context,
                       int cmvn_history, int64_t *keep_begin, int64_t *keep_end,
                       int64_t *feed_begin, int64_t *feed_end) {
  if (total_frames < 0 || n_parts < 1 || left_context < 0 || right_context < 0 || cmvn_history < 0 ||
      !keep_begin || !keep_end || !feed_begin || !feed_end) {
    SetError("ce_gpu_time_shards: bad arguments");
    return CE_GPU_EINVAL;
  }
  for (int p = 0; p < n_parts; ++p) {
    keep_begin[p] = total_frames * p / n_parts;
    keep_end[p] = total_frames * (p + 1) / n_parts;
    feed_begin[p] = std::max<int64_t>(0, keep_begin[p] - left_context - cmvn_history);
    feed_end[p] = std::min<int64_t>(total_frames, keep_end[p] + right_context);
    if (keep_end[p] == keep_begin[p]) feed_begin[p] = feed_end[p] = keep_begin[p];
  }
  return CE_GPU_OK;
}

// ---- matrix level -----------------------------------------------------------------------

int ce_gpu_quantize(const float *src, int64_t rows, int cols, uint8_t *dst, float *scale,
                    int32_t *zero_point, int device, void *stream) {
  if (!src || !dst || rows <= 0 || cols <= 0 || rows > 0x7fffffff) {
    SetError("ce_gpu_quantize: bad arguments");           // matrix.cc:369 asserts non-empty
    return CE_GPU_EINVAL;
  }
  CE_CHECK(UseDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  StageWs *ws = GetWs(device);
  const void *x = nullptr;
  CE_CHECK(StageIn(src, sizeof(float) * (size_t)rows * cols, &ws->in, s, &x));
  // the production kernel wants 128-column padding; odd widths take the generic one
  const int c_pad = (cols % 4 == 0 && cols <= 1024) ? RoundUp(cols, 128) : RoundUp(cols, 4);
  CE_CHECK(ws->tmp[0].Reserve(sizeof(uint32_t) * 2));
  CE_CHECK(ws->tmp[1].Reserve(sizeof(QParam)));
  CE_CHECK(ws->tmp[2].Reserve((size_t)rows * c_pad));
  RowUse use;
  memset(&use, 0, sizeof(use));
  CE_CHECK(InitMinMaxLaunch(ws->tmp[0].as<uint32_t>(), 1, s));
  CE_CHECK(MinMaxLaunch(static_cast<const float *>(x), cols, cols, (int)rows, nullptr, nullptr, use,
                        ws->tmp[0].as<uint32_t>(), s));
  CE_CHECK(QuantizeLaunch(static_cast<const float *>(x), cols, cols, (int)rows, c_pad, nullptr,
                          ws->tmp[0].as<uint32_t>(), 1, ws->tmp[1].as<QParam>(),
                          ws->tmp[2].as<uint8_t>(), nullptr, s));
  QParam q;
  CE_CUDA(cudaMemcpyAsync(&q, ws->tmp[1].ptr, sizeof(q), cudaMemcpyDeviceToHost, s));
  CE_CUDA(cudaMemcpy2DAsync(dst, cols, ws->tmp[2].ptr, c_pad, cols, rows,
                            IsDevicePtr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  CE_CUDA(cudaStreamSynchronize(s));
  if (scale) *scale = q.scale;
  if (zero_point) *zero_point = q.zero_point;
  return CE_GPU_OK;
}

int64_t ce_gpu_selftest_quantizer(int64_t n, uint64_t seed, int device) {
  if (n <= 0) {
    SetError("ce_gpu_selftest_quantizer: n must be positive");
    return CE_GPU_EINVAL;
  }
  CE_CHECK(UseDevice(device));
  StageWs *ws = GetWs(device);
  CE_CHECK(ws->tmp[0].Reserve(sizeof(unsigned long long)));
  CE_CUDA(cudaMemset(ws->tmp[0].ptr, 0, sizeof(unsigned long long)));
  CE_CHECK(QuantSelfTestLaunch(n, seed, ws->tmp[0].as<unsigned long long>(), nullptr));
  unsigned long long bad = 0;
  CE_CUDA(cudaMemcpy(&bad, ws->tmp[0].ptr, sizeof(bad), cudaMemcpyDeviceToHost));
  return (int64_t)bad;
}

int ce_gpu_gemm_u8(const uint8_t *a, float scale_a, int32_t zp_a, const uint8_t *b, float scale_b,
                   int32_t zp_b, int m, int n, int k, float *c, int32_t *acc, int device,
                   void *stream) {
  if (!a || !b || !c || m <= 0 || n <= 0 || k <= 0) {
    SetError("ce_gpu_gemm_u8: bad arguments");
    return CE_GPU_EINVAL;
  }
  CE_CHECK(UseDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  StageWs *ws = GetWs(device);
  const void *a_dev = nullptr, *b_dev = nullptr;
  CE_CHECK(StageIn(a, (size_t)m * k, &ws->in, s, &a_dev));
  CE_CHECK(StageIn(b, (size_t)k * n, &ws->in2, s, &b_dev));
  const int k_pad = RoundUp(k, KindTileK(kKindI8));
  const int ldc = RoundUp(n, 4), n_par = RoundUp(n, kTileN);
  CE_CHECK(ws->tmp[0].Reserve((size_t)m * k_pad));                    // A padded
  CE_CHECK(ws->tmp[1].Reserve((size_t)n * k_pad));                    // B^T padded
  CE_CHECK(ws->tmp[2].Reserve(sizeof(int32_t) * (size_t)m));          // row sums of A
  CE_CHECK(ws->tmp[3].Reserve(sizeof(int32_t) * (size_t)n_par));      // column sums of B
  CE_CHECK(ws->tmp[4].Reserve(sizeof(QParam)));
  CE_CHECK(ws->tmp[5].Reserve(sizeof(float) * (size_t)m * ldc));
  CE_CHECK(ws->tmp[6].Reserve(sizeof(int32_t) * (size_t)m * ldc));
  CE_CHECK(PadRowsLaunch<uint8_t>(static_cast<const uint8_t *>(a_dev), m, k, ws->tmp[0].as<uint8_t>(), k_pad, s));
  CE_CHECK(TransposePadLaunch<uint8_t>(static_cast<const uint8_t *>(b_dev), k, n, ws->tmp[1].as<uint8_t>(), k_pad, s));
  CE_CHECK(RowSumU8Launch(ws->tmp[0].as<uint8_t>(), m, k_pad, k, ws->tmp[2].as<int32_t>(), s));
  CE_CUDA(cudaMemsetAsync(ws->tmp[3].ptr, 0, sizeof(int32_t) * (size_t)n_par, s));
  CE_CHECK(RowSumU8Launch(ws->tmp[1].as<uint8_t>(), n, k_pad, k, ws->tmp[3].as<int32_t>(), s));
  QParam qa = {scale_a, zp_a};
  CE_CUDA(cudaMemcpyAsync(ws->tmp[4].ptr, &qa, sizeof(qa), cudaMemcpyHostToDevice, s));
  CE_CUDA(cudaStreamSynchronize(s));                     // qa lives on this stack frame

  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = m; g.N = n; g.c_pad = k_pad; g.n_taps = 1; g.n_pass = 1;
  g.a_rowsum = ws->tmp[2].as<int32_t>();
  g.b_colsum = ws->tmp[3].as<int32_t>();
  g.zp_b = zp_b; g.scale_b = scale_b; g.k_true = k;
  g.qa = ws->tmp[4].as<QParam>();
  g.out_f32 = ws->tmp[5].as<float>();
  g.out_acc = acc ? ws->tmp[6].as<int32_t>() : nullptr;
  g.ld_out = ldc; g.n_store = n;
  GemmOperands ops;
  memset(&ops, 0, sizeof(ops));
  ops.a[0] = ws->tmp[0].ptr; ops.rows_a = m;
  ops.b[0] = ws->tmp[1].ptr; ops.rows_b = n; ops.k_total = k_pad;
  CE_CHECK(GemmLaunch(kKindI8, ops, g, s));
  CE_CUDA(cudaMemcpy2DAsync(c, sizeof(float) * n, ws->tmp[5].ptr, sizeof(float) * ldc, sizeof(float) * n, m,
                            IsDevicePtr(c) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  if (acc) {
    CE_CUDA(cudaMemcpy2DAsync(acc, sizeof(int32_t) * n, ws->tmp[6].ptr, sizeof(int32_t) * ldc,
                              sizeof(int32_t) * n, m,
                              IsDevicePtr(acc) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  }
  CE_CUDA(cudaStreamSynchronize(s));
  return CE_GPU_OK;
}

int ce_gpu_gemm_f32(const float *a, const float *b, int m, int n, int k, float *c, int precision,
                    int device, void *stream) {
  if (!a || !b || !c || m <= 0 || n <= 0 || k <= 0) {
    SetError("ce_gpu_gemm_f32: bad arguments");
    return CE_GPU_EINVAL;
  }
  int kind, n_pass = 1;
  switch (precision) {
    case CE_GPU_PRECISION_BF16: kind = kKindBF16; break;
    case CE_GPU_PRECISION_FP32: kind = kKindTF32; n_pass = 3; break;
    case CE_GPU_PRECISION_TF32: kind = kKindTF32; break;
    case CE_GPU_PRECISION_BF16X3: kind = kKindBF16X3; break;
    default:
      SetError("ce_gpu_gemm_f32: precision must be BF16, FP32, TF32 or BF16X3");
      return CE_GPU_EINVAL;
  }
  CE_CHECK(UseDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  StageWs *ws = GetWs(device);
  const void *a_dev = nullptr, *b_dev = nullptr;
  CE_CHECK(StageIn(a, sizeof(float) * (size_t)m * k, &ws->in, s, &a_dev));
  CE_CHECK(StageIn(b, sizeof(float) * (size_t)k * n, &ws->in2, s, &b_dev));
  const int k_pad = RoundUp(k, KindTileK(kind));
  const int ldc = RoundUp(n, 4);
  const size_t elt = (size_t)KindEltBytes(kind) * (kind == kKindBF16X3 ? 2 : 1);   // stored bytes per channel
  CE_CHECK(ws->tmp[0].Reserve(sizeof(float) * (size_t)n * k));        // B^T fp32
  CE_CHECK(ws->tmp[1].Reserve(elt * (size_t)m * k_pad));              // A operand (hi)
  CE_CHECK(ws->tmp[2].Reserve(elt * (size_t)m * k_pad));              // A lo
  CE_CHECK(ws->tmp[3].Reserve(elt * (size_t)n * k_pad));              // B operand (hi)
  CE_CHECK(ws->tmp[4].Reserve(elt * (size_t)n * k_pad));              // B lo
  CE_CHECK(ws->tmp[5].Reserve(sizeof(float) * (size_t)m * ldc));
  CE_CHECK(TransposePadLaunch<float>(static_cast<const float *>(b_dev), k, n, ws->tmp[0].as<float>(), k, s));
  if (kind == kKindBF16X3) {
    CE_CHECK(ConvertLaunch(static_cast<const float *>(a_dev), k, k, m, k_pad, nullptr, nullptr, nullptr, s,
                           ws->tmp[1].as<__nv_bfloat16>()));
    CE_CHECK(ConvertLaunch(ws->tmp[0].as<float>(), k, k, n, k_pad, nullptr, nullptr, nullptr, s,
                           ws->tmp[3].as<__nv_bfloat16>()));
  } else if (kind == kKindBF16) {
    CE_CHECK(ConvertLaunch(static_cast<const float *>(a_dev), k, k, m, k_pad,
                           ws->tmp[1].as<__nv_bfloat16>(), nullptr, nullptr, s));
    CE_CHECK(ConvertLaunch(ws->tmp[0].as<float>(), k, k, n, k_pad, ws->tmp[3].as<__nv_bfloat16>(),
                           nullptr, nullptr, s));
  } else {
    CE_CHECK(ConvertLaunch(static_cast<const float *>(a_dev), k, k, m, k_pad, nullptr,
                           ws->tmp[1].as<float>(), n_pass == 3 ? ws->tmp[2].as<float>() : nullptr, s));
    CE_CHECK(ConvertLaunch(ws->tmp[0].as<float>(), k, k, n, k_pad, nullptr, ws->tmp[3].as<float>(),
                           n_pass == 3 ? ws->tmp[4].as<float>() : nullptr, s));
  }
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = m; g.N = n; g.c_pad = KindPhysCols(kind, k_pad); g.n_taps = 1; g.n_pass = n_pass;
  if (n_pass == 3) {
    g.pass_a[0] = 1; g.pass_b[0] = 0;
    g.pass_a[1] = 0; g.pass_b[1] = 1;
    g.pass_a[2] = 0; g.pass_b[2] = 0;
  }
  g.out_f32 = ws->tmp[5].as<float>();
  g.ld_out = ldc; g.n_store = n;
  GemmOperands ops;
  memset(&ops, 0, sizeof(ops));
  ops.a[0] = ws->tmp[1].ptr; ops.a[1] = n_pass == 3 ? ws->tmp[2].ptr : nullptr; ops.rows_a = m;
  ops.b[0] = ws->tmp[3].ptr; ops.b[1] = n_pass == 3 ? ws->tmp[4].ptr : nullptr; ops.rows_b = n;
  ops.k_total = KindPhysCols(kind, k_pad);
  CE_CHECK(GemmLaunch(kind, ops, g, s));
  CE_CUDA(cudaMemcpy2DAsync(c, sizeof(float) * n, ws->tmp[5].ptr, sizeof(float) * ldc, sizeof(float) * n, m,
                            IsDevicePtr(c) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  CE_CUDA(cudaStreamSynchronize(s));
  return CE_GPU_OK;
}

}  // extern "C"
