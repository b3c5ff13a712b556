// ce_host.hpp -- the C++ host side of the GPU path: the reference's operator surface for
// fbank / CMVN / acoustic model, implemented on the C ABI of ce_gpu.h (header-only, C++11).
//
// Same names, argument meaning and streaming behaviour as the reference classes, so that code
// written against them (src/ce_stt.cc:295-362) reads the same:
//
//   ce_host::Fbank::Process          Fbank::Process            src/fbank.h:57-59, src/fbank.cc:265-314
//   ce_host::CMVN::GetFrame          CMVN::GetFrame            src/cmvn.h:26,   src/cmvn.cc:100-110
//   ce_host::AcousticModel::Read     AcousticModel::Read       src/am.h:34,     src/am.cc:26-64
//   ce_host::AcousticModel::Process  AcousticModel::Process    src/am.h:42-44,  src/am.cc:115-142
//   ce_host::AcousticModel::EndOfStream                        src/am.h:47,     src/am.cc:144-164
//   TransitionPdfIdMap(), num_pdfs()                           src/am.h:37-39,50
//
// Beyond the reference's surface (serving forms of the same operators):
//   ce_host::AcousticModel::SelectPdfs / SelectTopK   narrower log_prob rows for the decoder
//   ce_host::StreamBatch         many live utterances per call, per-stream state on the host
//   ce_host::DeviceStreamBatch   the same with the state in device buffers (ce_gpu_streams_*)
//   ce_host::ShardedModel        a list of utterances over ALL GPUs of the box: one model handle and
//                                one host thread per GPU, rows handed to a pool of consumer threads
//                                (CPU decoders) utterance by utterance while the GPUs keep working
//
// Differences, all deliberate: the types are plain (std::vector-backed Matrix instead of
// pocketkaldi::Matrix<float>); every call that can fail on the device returns a Status instead of
// asserting; there is no CPU implementation behind any of it (no device -> Status with the
// CE_GPU_ENODEVICE message).  For throughput use the batch entry points of ce_gpu.h directly
// (ce_gpu_forward over many utterances); this header is the drop-in for per-utterance callers.
#ifndef CE_HOST_HPP_
#define CE_HOST_HPP_

#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ce_gpu.h"

namespace ce_host {

// Mirrors pocketkaldi::Status (src/status.h): ok() / what().
class Status {
 public:
  Status() : code_(0) {}
  static Status OK() { return Status(); }
  static Status IOError(const std::string &m) { return Status(1, "IOError: " + m); }
  static Status Corruption(const std::string &m) { return Status(2, "Corruption: " + m); }
  static Status RuntimeError(const std::string &m) { return Status(3, "RuntimeError: " + m); }
  // The last error of the C library on this thread, as a Status.
  static Status FromGpu(int rc) {
    if (rc >= 0) return OK();
    return rc == CE_GPU_EIO ? IOError(ce_gpu_last_error()) : RuntimeError(ce_gpu_last_error());
  }
  bool ok() const { return code_ == 0; }
  const std::string &what() const { return msg_; }

 private:
  Status(int c, const std::string &m) : code_(c), msg_(m) {}
  int code_;
  std::string msg_;
};

// Row-major float matrix, stride == cols (src/matrix.h:162).  "Nothing ready" is a 0-row matrix.
struct Matrix {
  int rows, cols;
  std::vector<float> data;
  Matrix() : rows(0), cols(0) {}
  Matrix(int r, int c) : rows(r), cols(c), data((size_t)r * c, 0.0f) {}
  void Resize(int r, int c) {
    rows = r;
    cols = c;
    data.assign((size_t)r * c, 0.0f);
  }
  int NumRows() const { return rows; }
  int NumCols() const { return cols; }
  float *Row(int r) { return data.data() + (size_t)r * cols; }
  const float *Row(int r) const { return data.data() + (size_t)r * cols; }
};

namespace detail {

// "key = value" file, values that are paths are relative to the file (src/configuration.cc:14-88).
inline Status ReadConfig(const std::string &path, std::map<std::string, std::string> *kv, std::string *dir) {
  FILE *f = fopen(path.c_str(), "r");
  if (!f) return Status::IOError("unable to open " + path);
  size_t slash = path.find_last_of('/');
  *dir = slash == std::string::npos ? std::string("") : path.substr(0, slash + 1);
  char line[4096];
  while (fgets(line, sizeof(line), f)) {
    std::string s(line);
    size_t hash = s.find('#');
    if (hash != std::string::npos) s = s.substr(0, hash);
    size_t eq = s.find('=');
    if (eq == std::string::npos) continue;
    auto trim = [](std::string v) {
      size_t a = v.find_first_not_of(" \t\r\n"), b = v.find_last_not_of(" \t\r\n");
      return a == std::string::npos ? std::string("") : v.substr(a, b - a + 1);
    };
    std::string k = trim(s.substr(0, eq)), v = trim(s.substr(eq + 1));
    if (!k.empty()) (*kv)[k] = v;
  }
  fclose(f);
  return Status::OK();
}

// VEC0 | int32 bytes (= 4 dim + 4) | int32 dim | dim x 4-byte items (src/vector.cc:267-300).
template <typename T>
inline Status ReadVec0(const std::string &path, std::vector<T> *out) {
  static_assert(sizeof(T) == 4, "VEC0 holds 4-byte items");
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return Status::IOError("unable to open " + path);
  char magic[4];
  int32_t bytes = 0, dim = 0;
  bool good = fread(magic, 1, 4, f) == 4 && memcmp(magic, "VEC0", 4) == 0 &&
              fread(&bytes, 4, 1, f) == 1 && fread(&dim, 4, 1, f) == 1 && dim >= 0 && bytes == 4 * dim + 4;
  if (good) {
    out->resize((size_t)dim);
    good = dim == 0 || fread(out->data(), 4, (size_t)dim, f) == (size_t)dim;
  }
  fclose(f);
  return good ? Status::OK() : Status::Corruption(path + ": not a VEC0 section");
}

}  // namespace detail

// ---------------------------------------------------------------------------------------------
// Fbank: streaming log-mel filterbank.  Process() appends the samples to the utterance's buffer,
// returns every complete frame (T = 1 + (n - 400) / 160, snip-edges) and keeps the remainder
// n - 160 T for the next call (src/fbank.cc:275-313).  Samples are the UNSCALED int16 values
// (src/pcm_reader.cc:168-182).
// ---------------------------------------------------------------------------------------------
class Fbank {
 public:
  class Instance {
   public:
    std::vector<int16_t> wave_buffer;
  };

  explicit Fbank(int num_mel = 40, int device = 0) : num_mel_(num_mel), device_(device) {}

  Status Process(Instance *inst, const int16_t *wave, int n, Matrix *feats) const {
    inst->wave_buffer.insert(inst->wave_buffer.end(), wave, wave + n);
    const int64_t total = (int64_t)inst->wave_buffer.size();
    const int64_t soff[2] = {0, total};
    int64_t foff[2] = {0, 0};
    const int64_t frames = ce_gpu_frame_offsets(soff, 1, foff);
    if (frames < 0) return Status::FromGpu((int)frames);
    feats->Resize((int)frames, num_mel_);
    if (frames == 0) return Status::OK();                  // fewer than 400 samples: stay buffered
    int rc = ce_gpu_fbank(inst->wave_buffer.data(), soff, 1, num_mel_, feats->data.data(), device_, nullptr);
    if (rc != CE_GPU_OK) return Status::FromGpu(rc);
    inst->wave_buffer.erase(inst->wave_buffer.begin(), inst->wave_buffer.begin() + frames * 160);
    return Status::OK();
  }

  // The reference's signature takes float samples (WaveReader's output); they must be int16 values.
  Status Process(Instance *inst, const std::vector<float> &wave, Matrix *feats) const {
    std::vector<int16_t> pcm(wave.size());
    for (size_t i = 0; i < wave.size(); ++i) {
      const float v = wave[i];
      if (v < -32768.0f || v > 32767.0f || v != (float)(int)v)
        return Status::RuntimeError("Fbank::Process: samples must be unscaled int16 values");
      pcm[i] = (int16_t)v;
    }
    return Process(inst, pcm.data(), (int)pcm.size(), feats);
  }

 private:
  int num_mel_, device_;
};

// ---------------------------------------------------------------------------------------------
// CMVN: online mean-only normalisation of one utterance's features.  Like the reference, the
// constructor takes the whole raw matrix (src/cmvn.cc:113-119 deep-copies it) and GetFrame must be
// called for frames 0, 1, 2, ... in order (src/cmvn.cc:38).
// ---------------------------------------------------------------------------------------------
class CMVN {
 public:
  CMVN(const std::vector<float> &global_stats, const Matrix &raw_feats, int device = 0)
      : next_(0), norm_(raw_feats.rows, raw_feats.cols) {
    if ((int)global_stats.size() != raw_feats.cols + 1) {
      status_ = Status::RuntimeError("CMVN: global stats must hold dim + 1 values");
      return;
    }
    if (raw_feats.rows == 0) return;
    const int64_t foff[2] = {0, raw_feats.rows};
    status_ = Status::FromGpu(ce_gpu_cmvn(global_stats.data(), raw_feats.data.data(), foff, 1, raw_feats.cols,
                                          norm_.data.data(), device, nullptr));
  }

  Status GetFrame(int frame, float *feats /* [dim] */) {
    if (!status_.ok()) return status_;
    if (frame != next_ || frame >= norm_.rows)
      return Status::RuntimeError("CMVN::GetFrame: frames must be requested in order");
    memcpy(feats, norm_.Row(frame), sizeof(float) * (size_t)norm_.cols);
    ++next_;
    return Status::OK();
  }

 private:
  int next_;
  Matrix norm_;
  Status status_;
};

// ---------------------------------------------------------------------------------------------
// AcousticModel: the streaming operator of src/am.h.  Frames are buffered with the reference's
// edge replication (left_context copies of the first frame, right_context copies of the last);
// a batch of chunk_size rows is computed as soon as chunk_size + left + right frames are buffered
// (src/am.cc:73-80,126-141), EndOfStream computes what is left (src/am.cc:144-164).  Rows are
// log-likelihoods: network output minus log prior.
// ---------------------------------------------------------------------------------------------
class AcousticModel {
 public:
  class Instance {
   public:
    Instance() : started(false), dim(0) {}
    bool started;
    int dim;
    std::vector<float> feats_buffer;     // [frames x dim], oldest first
    int frames() const { return dim ? (int)(feats_buffer.size() / dim) : 0; }
  };

  AcousticModel() : model_(nullptr), left_context_(0), right_context_(0), chunk_size_(0), num_pdfs_(0), feat_dim_(0) {}
  ~AcousticModel() { ce_gpu_model_free(model_); }
  AcousticModel(const AcousticModel &) = delete;
  AcousticModel &operator=(const AcousticModel &) = delete;

  // Reads the model named by the reference's configuration file (keys nnet, prior, left_context,
  // right_context, chunk_size, num_pdfs, tid2pdf [, cmvn_stats]).
  Status Read(const std::string &config_file, int precision = CE_GPU_PRECISION_FP32, int device = 0) {
    std::map<std::string, std::string> kv;
    std::string dir;
    Status st = detail::ReadConfig(config_file, &kv, &dir);
    if (!st.ok()) return st;
    for (const char *key : {"nnet", "prior", "left_context", "right_context", "chunk_size", "num_pdfs", "tid2pdf"})
      if (!kv.count(key)) return Status::Corruption(config_file + ": key '" + key + "' is missing");
    chunk_size_ = atoi(kv["chunk_size"].c_str());
    if (chunk_size_ <= 0) return Status::Corruption(config_file + ": chunk_size must be positive");
    const std::string tid = kv["tid2pdf"][0] == '/' ? kv["tid2pdf"] : dir + kv["tid2pdf"];
    st = detail::ReadVec0<int32_t>(tid, &tid2pdf_);
    if (!st.ok()) return st;
    ce_gpu_model_free(model_);
    model_ = ce_gpu_model_load_config(config_file.c_str(), precision, device);
    if (!model_) return Status::FromGpu(CE_GPU_EIO);
    st = Status::FromGpu(ce_gpu_model_info(model_, &num_pdfs_, &left_context_, &right_context_, &feat_dim_,
                                           nullptr, nullptr));
    out_width_ = num_pdfs_;
    return st;
  }

  // What a row of log_prob is from now on (ce_gpu_model_set_output): every pdf
  // (the reference's row), the listed pdfs only -- hand the decoder the matching remapped
  // TransitionPdfIdMap and it computes the same costs --, or the k best as ce_gpu_scored_pdf_t
  // pairs (two 4-byte words each; see ScoredRow).  Not while an Instance / Stream is mid-utterance.
  Status SelectAllPdfs() { return SetOutput(CE_GPU_OUTPUT_DENSE, nullptr, 0); }
  Status SelectPdfs(const std::vector<int32_t> &pdf_ids) {
    return SetOutput(CE_GPU_OUTPUT_SUBSET, pdf_ids.data(), (int)pdf_ids.size());
  }
  Status SelectTopK(int k) { return SetOutput(CE_GPU_OUTPUT_TOPK, nullptr, k); }
  // Columns (4-byte words) of every log_prob row under the current selection.
  int output_width() const { return out_width_; }
  static const ce_gpu_scored_pdf_t *ScoredRow(const Matrix &log_prob, int row) {
    return reinterpret_cast<const ce_gpu_scored_pdf_t *>(log_prob.Row(row));
  }

  const std::vector<int32_t> &TransitionPdfIdMap() const { return tid2pdf_; }
  int num_pdfs() const { return num_pdfs_; }
  int left_context() const { return left_context_; }
  int right_context() const { return right_context_; }
  int chunk_size() const { return chunk_size_; }
  int feat_dim() const { return feat_dim_; }
  ce_gpu_model_t *handle() const { return model_; }

  Status Process(Instance *inst, const float *frame_feat, Matrix *log_prob) const {
    if (!model_) return Status::RuntimeError("AcousticModel::Process before Read");
    if (!inst->started) {                                  // left padding, src/am.cc:119-124
      inst->dim = feat_dim_;
      for (int i = 0; i < left_context_; ++i) Append(inst, frame_feat);
      inst->started = true;
    }
    Append(inst, frame_feat);
    if (inst->frames() < left_context_ + right_context_ + chunk_size_) {   // BatchAvailable
      log_prob->Resize(0, 0);
      return Status::OK();
    }
    Status st = ComputeBatch(inst, chunk_size_, log_prob);
    if (!st.ok()) return st;
    inst->feats_buffer.erase(inst->feats_buffer.begin(),
                             inst->feats_buffer.begin() + (size_t)chunk_size_ * inst->dim);
    return Status::OK();
  }

  Status EndOfStream(Instance *inst, Matrix *log_prob) const {
    if (!model_) return Status::RuntimeError("AcousticModel::EndOfStream before Read");
    if (inst->feats_buffer.empty()) {
      log_prob->Resize(0, 0);
      return Status::OK();
    }
    const std::vector<float> last(inst->feats_buffer.end() - inst->dim, inst->feats_buffer.end());
    for (int i = 0; i < right_context_; ++i) Append(inst, last.data());   // src/am.cc:152-155
    const int n = inst->frames() - left_context_ - right_context_;
    if (n <= 0) {
      log_prob->Resize(0, 0);
      return Status::OK();
    }
    Status st = ComputeBatch(inst, n, log_prob);
    inst->feats_buffer.clear();
    return st;
  }

 private:
  static void Append(Instance *inst, const float *frame) {
    inst->feats_buffer.insert(inst->feats_buffer.end(), frame, frame + inst->dim);
  }

  // batch_size output rows from the first batch_size + left + right buffered frames, which carry
  // their own context (src/am.cc:82-113): one block of ce_gpu_nnet_chunks.
  Status ComputeBatch(Instance *inst, int batch_size, Matrix *log_prob) const {
    const int in_rows = batch_size + left_context_ + right_context_;
    const int64_t foff[2] = {0, in_rows};
    log_prob->Resize(batch_size, out_width_);
    int rc = ce_gpu_nnet_chunks(model_, inst->feats_buffer.data(), foff, 1, log_prob->data.data(), nullptr, nullptr);
    return rc != CE_GPU_OK ? Status::FromGpu(rc) : Status::OK();
  }

  Status SetOutput(int mode, const int32_t *ids, int n) {
    if (!model_) return Status::RuntimeError("AcousticModel: output selection before Read");
    int rc = ce_gpu_model_set_output(model_, mode, ids, n);
    if (rc != CE_GPU_OK) return Status::FromGpu(rc);
    out_width_ = ce_gpu_model_output_width(model_);
    return Status::OK();
  }

  ce_gpu_model_t *model_;
  int left_context_, right_context_, chunk_size_, num_pdfs_, feat_dim_;
  int out_width_ = 0;
  std::vector<int32_t> tid2pdf_;
  mutable Matrix scratch_;
};

// ---------------------------------------------------------------------------------------------
// StreamBatch: many live utterances evaluated together.  Every Process() call
// takes whatever PCM has arrived for each stream and runs ONE fbank, ONE CMVN and ONE acoustic-model
// pass for all of them; what the reference keeps per utterance in Fbank::Instance, CMVN and
// AcousticModel::Instance is carried here between calls:
//   * the samples that do not fill a frame yet (< 400 + 160, src/fbank.cc:308-313),
//   * the last <= 600 raw feature frames and the CMVN running sums (src/cmvn.cc:35-68),
//   * the normalised frames whose right context has not arrived, behind `left_context` frames of
//     left context (first frame replicated at the start, last frame at end of stream,
//     src/am.cc:119-124,152-155).
// Rows come out as soon as their right context exists; concatenated over the calls they are the rows
// of a whole-utterance evaluation (bit-identical CMVN; float log-likelihoods identical row by row --
// for int8 models the Quantize granularity is the micro-batch, as it is the chunk in the reference).
// ---------------------------------------------------------------------------------------------
class StreamBatch {
 public:
  struct Stream {
    std::vector<int16_t> wave;       // samples not yet framed
    std::vector<float> raw_hist;     // last <= 600 raw fbank frames [n x mel]
    std::vector<float> cmvn_state;   // running sums [mel]
    int64_t frames_normalised = 0;
    std::vector<float> ctx;          // normalised frames buffered for the AM [n x mel]
    bool started = false, ended = false;
    int64_t rows_emitted = 0;
  };

  // global_cmvn_stats: mel sums + count, or empty for no CMVN (what src/ce_stt.cc does).
  StreamBatch(const AcousticModel *am, const std::vector<float> &global_cmvn_stats, int device = 0)
      : am_(am), stats_(global_cmvn_stats), device_(device) {}

  // pcm[i] / n_samples[i]: new audio of streams[i] (may be 0); end_of_stream[i]: no more audio will
  // come.  rows[i] receives the log-likelihood rows that became ready (0 rows if none).
  Status Process(const std::vector<Stream *> &streams, const std::vector<const int16_t *> &pcm,
                 const std::vector<int> &n_samples, const std::vector<bool> &end_of_stream,
                 std::vector<Matrix> *rows) const {
    const int n = (int)streams.size(), mel = am_->feat_dim(), P = am_->output_width();
    const int L = am_->left_context(), R = am_->right_context();
    rows->assign(n, Matrix());
    // ---- 1. fbank of every stream's buffered samples, one call ----
    std::vector<int16_t> all_pcm;
    std::vector<int64_t> soff(n + 1, 0), foff(n + 1, 0);
    for (int i = 0; i < n; ++i) {
      Stream *st = streams[i];
      if (st->ended) return Status::RuntimeError("StreamBatch: stream already ended");
      if (n_samples[i] > 0) st->wave.insert(st->wave.end(), pcm[i], pcm[i] + n_samples[i]);
      all_pcm.insert(all_pcm.end(), st->wave.begin(), st->wave.end());
      soff[i + 1] = (int64_t)all_pcm.size();
    }
    const int64_t new_frames = ce_gpu_frame_offsets(soff.data(), n, foff.data());
    if (new_frames < 0) return Status::FromGpu((int)new_frames);
    std::vector<float> raw((size_t)new_frames * mel);
    if (new_frames > 0) {
      int rc = ce_gpu_fbank(all_pcm.data(), soff.data(), n, mel, raw.data(), device_, nullptr);
      if (rc != CE_GPU_OK) return Status::FromGpu(rc);
    }
    for (int i = 0; i < n; ++i) {
      const int64_t f = foff[i + 1] - foff[i];
      Stream *st = streams[i];
      st->wave.erase(st->wave.begin(), st->wave.begin() + f * 160);
    }
    // ---- 2. CMVN continued from every stream's state, one call ----
    std::vector<float> norm = raw;
    if (!stats_.empty() && new_frames > 0) {
      std::vector<float> in, state((size_t)n * mel, 0.0f);
      std::vector<int64_t> coff(n + 1, 0), tbase(n);
      std::vector<int32_t> nhist(n);
      for (int i = 0; i < n; ++i) {
        Stream *st = streams[i];
        if (st->cmvn_state.empty()) st->cmvn_state.assign(mel, 0.0f);
        nhist[i] = (int32_t)(st->raw_hist.size() / mel);
        tbase[i] = st->frames_normalised;
        in.insert(in.end(), st->raw_hist.begin(), st->raw_hist.end());
        in.insert(in.end(), raw.begin() + foff[i] * mel, raw.begin() + foff[i + 1] * mel);
        coff[i + 1] = (int64_t)(in.size() / mel);
        std::copy(st->cmvn_state.begin(), st->cmvn_state.end(), state.begin() + (size_t)i * mel);
      }
      int rc = ce_gpu_cmvn_stream(stats_.data(), in.data(), coff.data(), nhist.data(), tbase.data(), state.data(),
                                  n, mel, norm.data(), device_, nullptr);
      if (rc != CE_GPU_OK) return Status::FromGpu(rc);
      for (int i = 0; i < n; ++i) {
        Stream *st = streams[i];
        std::copy(state.begin() + (size_t)i * mel, state.begin() + (size_t)(i + 1) * mel, st->cmvn_state.begin());
        st->raw_hist.insert(st->raw_hist.end(), raw.begin() + foff[i] * mel, raw.begin() + foff[i + 1] * mel);
        const size_t keep = (size_t)600 * mel;
        if (st->raw_hist.size() > keep) st->raw_hist.erase(st->raw_hist.begin(), st->raw_hist.end() - keep);
        st->frames_normalised += foff[i + 1] - foff[i];
      }
    }
    // ---- 3. acoustic model on every stream's rows whose context is complete, one call ----
    std::vector<float> x;
    std::vector<int64_t> xoff(1, 0);
    std::vector<int> who, ready;
    for (int i = 0; i < n; ++i) {
      Stream *st = streams[i];
      const float *f0 = norm.data() + foff[i] * mel;
      const int64_t f = foff[i + 1] - foff[i];
      if (f > 0 && !st->started) {                           // left padding, src/am.cc:119-124
        for (int k = 0; k < L; ++k) st->ctx.insert(st->ctx.end(), f0, f0 + mel);
        st->started = true;
      }
      st->ctx.insert(st->ctx.end(), f0, f0 + f * mel);
      if (end_of_stream[i]) {
        st->ended = true;
        if (!st->ctx.empty()) {                              // right padding, src/am.cc:152-155
          const std::vector<float> last(st->ctx.end() - mel, st->ctx.end());
          for (int k = 0; k < R; ++k) st->ctx.insert(st->ctx.end(), last.begin(), last.end());
        }
      }
      const int have = (int)(st->ctx.size() / mel);
      const int n_ready = have - L - R;
      if (n_ready <= 0) continue;
      x.insert(x.end(), st->ctx.begin(), st->ctx.begin() + (size_t)(n_ready + L + R) * mel);
      xoff.push_back((int64_t)(x.size() / mel));
      who.push_back(i);
      ready.push_back(n_ready);
    }
    if (who.empty()) return Status::OK();
    int64_t n_rows = 0;
    for (int r : ready) n_rows += r;
    std::vector<float> ll((size_t)n_rows * P);
    int rc = ce_gpu_nnet_chunks(am_->handle(), x.data(), xoff.data(), (int)who.size(), ll.data(), nullptr, nullptr);
    if (rc != CE_GPU_OK) return Status::FromGpu(rc);
    int64_t o = 0;
    for (size_t k = 0; k < who.size(); ++k) {
      Stream *st = streams[who[k]];
      Matrix &out = (*rows)[who[k]];
      out.Resize(ready[k], P);
      memcpy(out.data.data(), ll.data() + (size_t)o * P, sizeof(float) * (size_t)ready[k] * P);
      o += ready[k];
      st->ctx.erase(st->ctx.begin(), st->ctx.begin() + (size_t)ready[k] * mel);
      st->rows_emitted += ready[k];
      if (st->ended) st->ctx.clear();
    }
    return Status::OK();
  }

 private:
  const AcousticModel *am_;
  std::vector<float> stats_;
  int device_;
};

// ---------------------------------------------------------------------------------------------
// DeviceStreamBatch: the same micro-batching with the per-stream state kept ON THE DEVICE
// (ce_gpu_streams_*): the host holds a slot number per live utterance, only the new samples go up
// and only the finished rows come down.  Same kernels on the same inputs as StreamBatch, so the
// rows are identical bit for bit.  The CMVN statistics are the model's (`cmvn_stats` in its config).
// ---------------------------------------------------------------------------------------------
class DeviceStreamBatch {
 public:
  struct Stream {
    int slot = -1;                   // -1: not opened yet (Process opens it)
    bool ended = false;
    int64_t rows_emitted = 0;
  };

  DeviceStreamBatch(const AcousticModel *am, int max_streams)
      : am_(am), set_(ce_gpu_streams_create(am->handle(), max_streams)) {}
  ~DeviceStreamBatch() { ce_gpu_streams_free(set_); }
  DeviceStreamBatch(const DeviceStreamBatch &) = delete;
  DeviceStreamBatch &operator=(const DeviceStreamBatch &) = delete;

  Status Process(const std::vector<Stream *> &streams, const std::vector<const int16_t *> &pcm,
                 const std::vector<int> &n_samples, const std::vector<bool> &end_of_stream,
                 std::vector<Matrix> *rows) const {
    if (!set_) return Status::FromGpu(CE_GPU_ENOMEM);
    const int n = (int)streams.size(), W = am_->output_width();
    rows->assign(n, Matrix());
    std::vector<int> slots(n);
    std::vector<unsigned char> eos(n);
    for (int i = 0; i < n; ++i) {
      if (streams[i]->ended) return Status::RuntimeError("DeviceStreamBatch: stream already ended");
      if (streams[i]->slot < 0) {
        const int id = ce_gpu_streams_open(set_);
        if (id < 0) return Status::FromGpu(id);
        streams[i]->slot = id;
      }
      slots[i] = streams[i]->slot;
      eos[i] = end_of_stream[i] ? 1 : 0;
    }
    const int64_t ready = ce_gpu_streams_rows_ready(set_, slots.data(), n, n_samples.data(), eos.data());
    if (ready < 0) return Status::FromGpu((int)ready);
    std::vector<float> flat((size_t)ready * W);
    std::vector<int64_t> off(n + 1, 0);
    int rc = ce_gpu_streams_process(set_, slots.data(), n, pcm.data(), n_samples.data(), eos.data(), flat.data(),
                                    ready, off.data(), nullptr);
    if (rc != CE_GPU_OK) return Status::FromGpu(rc);
    for (int i = 0; i < n; ++i) {
      const int r = (int)(off[i + 1] - off[i]);
      (*rows)[i].Resize(r, W);
      if (r > 0) memcpy((*rows)[i].data.data(), flat.data() + (size_t)off[i] * W, sizeof(float) * (size_t)r * W);
      streams[i]->rows_emitted += r;
      if (end_of_stream[i]) {
        streams[i]->ended = true;
        streams[i]->slot = -1;
      }
    }
    return Status::OK();
  }

 private:
  const AcousticModel *am_;
  ce_gpu_streams_t *set_;
};

// ---------------------------------------------------------------------------------------------
// ShardedModel: the batch form over every GPU of the box (utterances are independent,
// so there is no collective).  The utterance list is cut into contiguous, frame-balanced groups
// (ce_gpu_partition), one per GPU; every GPU has its own model handle and its own host thread,
// which runs ce_gpu_forward on its group chunk by chunk into a pinned host buffer; as soon as a
// chunk's rows have arrived (ce_gpu_model_set_rows_callback) its utterances are queued for a pool
// of consumer threads -- the CPU decoders -- so decoding of chunk c overlaps the GPU work on c + 1
// and everything the other GPUs do.  Use a narrow row (SelectPdfs / SelectTopK): a dense row is
// 4 num_pdfs bytes a frame of pinned memory and PCIe traffic.
// ---------------------------------------------------------------------------------------------
class ShardedModel {
 public:
  // One finished utterance: rows [n_frames x width] (valid during the call only), called once per
  // utterance with >= 1 frame, from one of the consumer threads, in no particular order.
  typedef std::function<void(int utt, const float *rows, int64_t n_frames, int width)> UtteranceFn;

  ShardedModel() {}
  ~ShardedModel() {
    for (Shard &sh : shards_) {
      ce_gpu_host_free(sh.rows);
      ce_gpu_model_free(sh.model);
    }
  }
  ShardedModel(const ShardedModel &) = delete;
  ShardedModel &operator=(const ShardedModel &) = delete;

  // One handle per device; `devices` empty = every visible device.
  Status Read(const std::string &config_file, int precision = CE_GPU_PRECISION_FP32,
              std::vector<int> devices = std::vector<int>()) {
    if (devices.empty())
      for (int d = 0; d < ce_gpu_device_count(); ++d) devices.push_back(d);
    if (devices.empty()) return Status::RuntimeError("ShardedModel: no CUDA device");
    for (int d : devices) {
      Shard sh;
      sh.device = d;
      sh.model = ce_gpu_model_load_config(config_file.c_str(), precision, d);
      if (!sh.model) return Status::FromGpu(CE_GPU_EIO);
      shards_.push_back(sh);
    }
    return Status::FromGpu(ce_gpu_model_info(shards_[0].model, &num_pdfs_, &left_, &right_, nullptr, nullptr, nullptr));
  }

  Status SelectPdfs(const std::vector<int32_t> &pdf_ids) {
    return SetOutput(CE_GPU_OUTPUT_SUBSET, pdf_ids.data(), (int)pdf_ids.size());
  }
  Status SelectTopK(int k) { return SetOutput(CE_GPU_OUTPUT_TOPK, nullptr, k); }
  Status SelectAllPdfs() { return SetOutput(CE_GPU_OUTPUT_DENSE, nullptr, 0); }
  int n_devices() const { return (int)shards_.size(); }
  int num_pdfs() const { return num_pdfs_; }
  int output_width() const { return shards_.empty() ? 0 : ce_gpu_model_output_width(shards_[0].model); }

  // pcm: HOST samples of all utterances, utterance u = [sample_off[u], sample_off[u+1]).
  // n_consumers threads call `fn`.  frame_off_out (nullable): frames of every utterance.
  Status Forward(const int16_t *pcm, const int64_t *sample_off, int n_utts, const UtteranceFn &fn,
                 int n_consumers = 1, std::vector<int64_t> *frame_off_out = nullptr) {
    if (shards_.empty()) return Status::RuntimeError("ShardedModel::Forward before Read");
    std::vector<int64_t> foff((size_t)n_utts + 1, 0);
    const int64_t total = ce_gpu_frame_offsets(sample_off, n_utts, foff.data());
    if (total < 0) return Status::FromGpu((int)total);
    if (frame_off_out) *frame_off_out = foff;
    const int n_parts = (int)shards_.size();
    std::vector<int32_t> part((size_t)n_parts + 1, 0);
    int rc = ce_gpu_partition(foff.data(), n_utts, n_parts, part.data());
    if (rc != CE_GPU_OK) return Status::FromGpu(rc);

    Queue q;
    std::vector<Status> status((size_t)n_parts);
    std::vector<std::thread> gpus, consumers;
    for (int c = 0; c < std::max(1, n_consumers); ++c)
      consumers.emplace_back([&]() {
        Item it;
        while (q.Pop(&it)) fn(it.utt, it.rows, it.n_frames, it.width);
      });
    for (int p = 0; p < n_parts; ++p)
      gpus.emplace_back([&, p]() { status[p] = RunShard(&shards_[p], pcm, sample_off, foff.data(), part[p], part[p + 1], &q); });
    for (std::thread &t : gpus) t.join();
    q.Close();
    for (std::thread &t : consumers) t.join();
    for (const Status &st : status)
      if (!st.ok()) return st;
    return Status::OK();
  }

 private:
  struct Shard {
    int device = 0;
    ce_gpu_model_t *model = nullptr;
    float *rows = nullptr;              // pinned host rows of the group being evaluated
    size_t rows_bytes = 0;
  };
  struct Item {
    int utt;
    const float *rows;
    int64_t n_frames;
    int width;
  };
  class Queue {
   public:
    void Push(const Item &it) {
      std::lock_guard<std::mutex> lock(mu_);
      items_.push_back(it);
      cv_.notify_one();
    }
    void Close() {
      std::lock_guard<std::mutex> lock(mu_);
      closed_ = true;
      cv_.notify_all();
    }
    bool Pop(Item *it) {
      std::unique_lock<std::mutex> lock(mu_);
      cv_.wait(lock, [&]() { return closed_ || !items_.empty(); });
      if (items_.empty()) return false;
      *it = items_.front();
      items_.pop_front();
      return true;
    }

   private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Item> items_;
    bool closed_ = false;
  };
  struct ChunkCtx {                     // what the rows-ready callback of one Forward call needs
    Queue *q;
    const int64_t *foff;                // frame offsets of ALL utterances
    int first_utt;                      // of this shard's group
    int64_t first_frame;
    const float *rows;
    int width;
  };
  static void RowsReady(void *user, int first_utt, int n_utts, int64_t, int64_t) {
    const ChunkCtx *c = static_cast<const ChunkCtx *>(user);
    for (int u = first_utt; u < first_utt + n_utts; ++u) {          // indices within the group
      const int g = c->first_utt + u;
      const int64_t n = c->foff[g + 1] - c->foff[g];
      if (n > 0) c->q->Push(Item{g, c->rows + (size_t)(c->foff[g] - c->first_frame) * c->width, n, c->width});
    }
  }

  Status RunShard(Shard *sh, const int16_t *pcm, const int64_t *sample_off, const int64_t *foff, int u0, int u1,
                  Queue *q) {
    if (u1 <= u0) return Status::OK();
    const int width = ce_gpu_model_output_width(sh->model);
    const int64_t frames = foff[u1] - foff[u0];
    if (frames == 0) return Status::OK();
    const size_t bytes = sizeof(float) * (size_t)frames * width;
    if (bytes > sh->rows_bytes) {
      ce_gpu_host_free(sh->rows);
      sh->rows = static_cast<float *>(ce_gpu_host_alloc(bytes));
      sh->rows_bytes = sh->rows ? bytes : 0;
      if (!sh->rows) return Status::FromGpu(CE_GPU_ENOMEM);
    }
    // the group's own offset arrays start at 0 (the sample data is addressed from its first sample)
    std::vector<int64_t> soff((size_t)(u1 - u0) + 1);
    for (int u = u0; u <= u1; ++u) soff[u - u0] = sample_off[u] - sample_off[u0];
    ChunkCtx ctx = {q, foff, u0, foff[u0], sh->rows, width};
    int rc = ce_gpu_model_set_rows_callback(sh->model, &ShardedModel::RowsReady, &ctx);
    if (rc == CE_GPU_OK)
      rc = ce_gpu_forward(sh->model, pcm + sample_off[u0], soff.data(), u1 - u0, sh->rows, nullptr, nullptr, nullptr);
    Status st = Status::FromGpu(rc);                     // (before the next call overwrites the message)
    ce_gpu_model_set_rows_callback(sh->model, nullptr, nullptr);
    return st;
  }

  Status SetOutput(int mode, const int32_t *ids, int n) {
    for (Shard &sh : shards_) {
      int rc = ce_gpu_model_set_output(sh.model, mode, ids, n);
      if (rc != CE_GPU_OK) return Status::FromGpu(rc);
    }
    return Status::OK();
  }

  std::vector<Shard> shards_;
  int num_pdfs_ = 0, left_ = 0, right_ = 0;
};

}  // namespace ce_host
#endif  // CE_HOST_HPP_
