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
#include <map>
#include <string>
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
    return Status::FromGpu(ce_gpu_model_info(model_, &num_pdfs_, &left_context_, &right_context_, &feat_dim_,
                                             nullptr, nullptr));
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
  // their own context (src/am.cc:82-113).  ce_gpu_nnet treats its input as a whole utterance and
  // replicates the edges once more; the rows whose context is entirely real are [left, left + batch).
  Status ComputeBatch(Instance *inst, int batch_size, Matrix *log_prob) const {
    const int in_rows = batch_size + left_context_ + right_context_;
    const int64_t foff[2] = {0, in_rows};
    scratch_.Resize(in_rows, num_pdfs_);
    int rc = ce_gpu_nnet(model_, inst->feats_buffer.data(), foff, 1, scratch_.data.data(), nullptr, nullptr);
    if (rc != CE_GPU_OK) return Status::FromGpu(rc);
    log_prob->Resize(batch_size, num_pdfs_);
    memcpy(log_prob->data.data(), scratch_.Row(left_context_), sizeof(float) * (size_t)batch_size * num_pdfs_);
    return Status::OK();
  }

  ce_gpu_model_t *model_;
  int left_context_, right_context_, chunk_size_, num_pdfs_, feat_dim_;
  std::vector<int32_t> tid2pdf_;
  mutable Matrix scratch_;
};

}  // namespace ce_host
#endif  // CE_HOST_HPP_
