// ce_stt_gpu.cc -- the reference's C API (src/ce_stt.h:41-76, the eight ce_stt_* / ce_utt_* symbols)
// with the acoustic front end and the acoustic model evaluated by libce_gpu.so.
//
// This file REPLACES src/ce_stt.cc at link time; every
// other reference source is compiled unchanged: decoder.cc, fst.cc, hashtable.cc, symbol_table.cc,
// pcm_reader.cc, configuration.cc, am.cc (only for the transition-id -> pdf-id map it loads) and the
// vendored OpenFst.  fbank.cc, srfft.cc, cmvn.cc are not linked at all, and cblas_sgemm is a stub
// that aborts (integration/cblas_forbidden.cc), so the reference's CPU feature extraction and GEMM
// provably never run.
//
// Streaming, like src/ce_stt.cc:295-362: every utterance owns a slot of a ce_gpu_streams set (sample
// remainder, CMVN state and AM context live on the device).  ce_stt_process hands the GPU whatever
// audio has arrived as soon as the reference itself would compute a batch -- AcousticModel::Process
// emits chunk_size rows once left + right + chunk_size frames are buffered (src/am.cc:73-142) -- the
// rows come back into a pinned host buffer, and the unchanged Decoder::Process consumes them in the
// reference's groups; the hypothesis text is refreshed every 20 decoded frames (src/ce_stt.cc:326-327)
// and at ce_stt_end_of_stream, which flushes the right context (src/am.cc:144-164).  After every call
// utt->hyp is therefore what the all-CPU reference holds after the same call.  The delta-LM rescoring
// option (large_lm / original_lm, src/ce_stt.cc:84-113) is wired to the unchanged DeltaLmFst.
// PCM: 16-bit samples go up as they are; 8-bit (signed, src/pcm_reader.cc:36-40) are widened exactly;
// 32-bit samples that fit 16 bits likewise, others make ce_stt_process fail with a message -- the
// reference would pass them on as floats (src/pcm_reader.cc:168-182), nothing is ever clamped.
//
// CE_STT_GPU_OUTPUT selects what crosses PCIe per frame (ce_gpu_model_set_output):
//   unset / "dense"  all num_pdfs log-likelihoods (12 KB a frame at 3072 pdfs);
//   "subset"         only the pdfs the graph's input labels can reach; the decoder gets the
//                    matching remapped transition-id map, so it computes exactly what it computed
//                    from the dense row;
//   "topk:<k>"       the k best (loglik, pdf) pairs; the row handed to the decoder holds those and
//                    the k-th value (an upper bound) for every other pdf -- an approximation.
// CE_STT_GPU_MAX_UTTS: utterances alive at a time (slots of the stream set), default 64.
#include "ce_stt.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "am.h"
#include "ce_gpu.h"
#include "configuration.h"
#include "decoder.h"
#include "fst.h"
#include "pcm_reader.h"
#include "symbol_table.h"
#include "util.h"

using pocketkaldi::AcousticModel;
using pocketkaldi::Configuration;
using pocketkaldi::Decoder;
using pocketkaldi::DeltaLmFst;
using pocketkaldi::LmFst;
using pocketkaldi::Status;
using pocketkaldi::SymbolTable;
using pocketkaldi::Vector;
using pocketkaldi::WaveReader;

struct ce_stt_t {
  fst::ConstFst<fst::StdArc> *graph = nullptr;
  AcousticModel *am = nullptr;            // host copy: transition-id map and num_pdfs only
  SymbolTable *symbols = nullptr;
  Vector<float> *original_lm = nullptr;   // delta-LM rescoring (optional), src/ce_stt.cc:84-113
  LmFst *large_lm = nullptr;
  DeltaLmFst *delta_lm = nullptr;
  ce_gpu_model_t *gpu = nullptr;          // the model that actually runs
  ce_gpu_streams_t *streams = nullptr;    // per-utterance state on the device
  std::mutex gpu_mu;                      // calls on one stream set must not overlap (ce_gpu.h)
  int left = 0, right = 0, chunk = 1;
  // what a row of the GPU output is, and the transition-id map that goes with it
  int out_mode = CE_GPU_OUTPUT_DENSE;
  int top_k = 0;
  Vector<int32_t> tid2col;                // subset mode: transition-id -> column of the gathered row
  const Vector<int32_t> &DecoderMap() const {
    return out_mode == CE_GPU_OUTPUT_SUBSET ? tid2col : am->TransitionPdfIdMap();
  }
};

struct ce_utt_internal_t {
  ce_stt_t *recognizer = nullptr;
  WaveReader wave_reader;
  std::unique_ptr<Decoder> decoder;
  int slot = -1;                          // ce_gpu_streams slot; -1 once the stream has ended
  std::vector<int16_t> pending;           // samples the GPU has not seen yet
  int64_t samples_total = 0;              // samples received so far
  int64_t rows_received = 0;              // rows the GPU has produced for this utterance
  int64_t rows_decoded = 0;               // rows Decoder::Process has consumed
  float *rows = nullptr;                  // pinned: the rows received and not decoded yet, oldest first
  int64_t rows_cap = 0;
  ~ce_utt_internal_t() { ce_gpu_host_free(rows); }
};

namespace {

thread_local char g_error[2048] = "";     // ce_stt_last_error(): thread-local here

void SetError(const std::string &msg) {
  strncpy(g_error, msg.c_str(), sizeof(g_error) - 1);
  g_error[sizeof(g_error) - 1] = '\0';
}

// Restricts the GPU output to the pdfs some arc of the graph can ask for (Decoder::LogLikelihood
// reads frame_logp(tid2pdf[arc.ilabel]) and nothing else, src/decoder.cc:97-102,325-350).
bool SelectGraphPdfs(ce_stt_t *r) {
  const Vector<int32_t> &tid2pdf = r->am->TransitionPdfIdMap();
  std::vector<int32_t> col_of(r->am->num_pdfs(), -1), ids;
  r->tid2col.Resize(tid2pdf.Dim());
  for (fst::StateIterator<fst::Fst<fst::StdArc>> si(*r->graph); !si.Done(); si.Next()) {
    for (fst::ArcIterator<fst::Fst<fst::StdArc>> ai(*r->graph, si.Value()); !ai.Done(); ai.Next()) {
      const int tid = ai.Value().ilabel;
      if (tid == 0) continue;
      if (tid < 0 || tid >= tid2pdf.Dim()) return false;
      const int pdf = tid2pdf(tid);
      if (col_of[pdf] < 0) {
        col_of[pdf] = (int32_t)ids.size();
        ids.push_back(pdf);
      }
      r->tid2col(tid) = col_of[pdf];
    }
  }
  if (ids.empty()) return false;
  return ce_gpu_model_set_output(r->gpu, CE_GPU_OUTPUT_SUBSET, ids.data(), (int)ids.size()) == CE_GPU_OK;
}

// The optional G^-1 o G' rescoring FST (keys large_lm + original_lm of the config file).
Status ReadDeltaLm(ce_stt_t *r, const Configuration &conf) {
  const std::string large = conf.GetPathOrElse("large_lm", "");
  if (large.empty()) return Status::OK();
  const std::string original = conf.GetPathOrElse("original_lm", "");
  if (original.empty()) return Status::Corruption("Unable to find key 'original_lm'");
  pocketkaldi::util::ReadableFile fd_original, fd_large;
  PK_CHECK_STATUS(fd_original.Open(original));
  r->original_lm = new Vector<float>();
  PK_CHECK_STATUS(r->original_lm->Read(&fd_original));
  PK_CHECK_STATUS(fd_large.Open(large));
  r->large_lm = new LmFst();
  PK_CHECK_STATUS(r->large_lm->Read(&fd_large));
  r->large_lm->InitBucket0();
  r->delta_lm = new DeltaLmFst(r->original_lm, r->large_lm, r->symbols);
  return Status::OK();
}

void StoreHyp(ce_utt_t *utt) {
  Decoder *dec = utt->internal->decoder.get();
  Decoder::Hypothesis hyp = dec->BestPath();
  std::vector<int> words = hyp.words();
  std::reverse(words.begin(), words.end());
  std::string text;
  for (size_t i = 0; i < words.size(); ++i) {
    if (i) text += ' ';
    text += utt->internal->recognizer->symbols->Get(words[i]);
  }
  delete[] utt->hyp;
  utt->hyp = new char[text.size() + 1];
  memcpy(utt->hyp, text.c_str(), text.size() + 1);
  if (!words.empty() && dec->NumFramesDecoded() > 0)
    utt->loglikelihood_per_frame = hyp.weight() / dec->NumFramesDecoded();
}

int64_t NumFrames(int64_t samples) { return samples < 400 ? 0 : 1 + (samples - 400) / 160; }   // src/fbank.cc:35-42

// Rows AcousticModel::Process has handed out after n frames: a batch of chunk rows whenever
// left + right + chunk frames are buffered, at most one per frame (src/am.cc:73-142).
int64_t RowsDue(const ce_stt_t *r, int64_t n_frames) {
  return n_frames < r->right + r->chunk ? 0 : (n_frames - r->right) / r->chunk * r->chunk;
}

// One row of GPU output to the unchanged decoder.
void DecodeRow(ce_utt_t *utt, const float *row, int width, Vector<float> *dense) {
  ce_utt_internal_t *in = utt->internal;
  const ce_stt_t *rec = in->recognizer;
  if (rec->out_mode == CE_GPU_OUTPUT_TOPK) {
    const ce_gpu_scored_pdf_t *best = reinterpret_cast<const ce_gpu_scored_pdf_t *>(row);
    for (int j = 0; j < dense->Dim(); ++j) (*dense)(j) = best[rec->top_k - 1].loglik;
    for (int j = 0; j < rec->top_k; ++j) (*dense)(best[j].pdf) = best[j].loglik;
    in->decoder->Process(*dense);
  } else {                                               // dense or gathered rows, as they are
    in->decoder->Process(pocketkaldi::SubVector<float>(const_cast<float *>(row), width));
  }
  ++in->rows_decoded;
}

// Sends the pending samples to the GPU (with the end-of-stream flag: the right context is replicated
// and the slot is freed) and appends the rows that come back to the utterance's pinned buffer.
bool RunGpu(ce_utt_t *utt, bool end_of_stream) {
  ce_utt_internal_t *in = utt->internal;
  ce_stt_t *rec = in->recognizer;
  if (in->slot < 0) return true;
  std::lock_guard<std::mutex> lock(rec->gpu_mu);
  const int width = ce_gpu_model_output_width(rec->gpu);
  const int n_samples = (int)in->pending.size();
  const unsigned char eos = end_of_stream ? 1 : 0;
  const int64_t n_new = ce_gpu_streams_rows_ready(rec->streams, &in->slot, 1, &n_samples, &eos);
  if (n_new < 0) {
    SetError(ce_gpu_last_error());
    return false;
  }
  const int64_t held = in->rows_received - in->rows_decoded;
  if (held + n_new > in->rows_cap) {                     // grow the pinned buffer, keep what is held
    const int64_t cap = std::max<int64_t>(2 * in->rows_cap, held + n_new + rec->chunk);
    float *grown = static_cast<float *>(ce_gpu_host_alloc(sizeof(float) * (size_t)cap * width));
    if (!grown) {
      SetError(ce_gpu_last_error());
      return false;
    }
    if (held > 0) memcpy(grown, in->rows, sizeof(float) * (size_t)held * width);
    ce_gpu_host_free(in->rows);
    in->rows = grown;
    in->rows_cap = cap;
  }
  const int16_t *pcm = in->pending.data();
  int64_t row_off[2] = {0, 0};
  if (ce_gpu_streams_process(rec->streams, &in->slot, 1, &pcm, &n_samples, &eos, in->rows + held * width,
                             in->rows_cap - held, row_off, nullptr) != CE_GPU_OK) {
    SetError(ce_gpu_last_error());
    return false;
  }
  in->pending.clear();
  in->rows_received += row_off[1] - row_off[0];
  if (end_of_stream) in->slot = -1;
  return true;
}

// Decoder::Process over the held rows until `due` rows have been decoded in all; refreshes the
// hypothesis every 20 frames when `partial` (src/ce_stt.cc:322-329).
void Decode(ce_utt_t *utt, int64_t due, bool partial) {
  ce_utt_internal_t *in = utt->internal;
  const ce_stt_t *rec = in->recognizer;
  const int width = ce_gpu_model_output_width(rec->gpu);
  Vector<float> dense(rec->out_mode == CE_GPU_OUTPUT_TOPK ? rec->am->num_pdfs() : 0);
  const int64_t held = in->rows_received - in->rows_decoded;
  const int64_t n = std::min(held, due - in->rows_decoded);
  for (int64_t i = 0; i < n; ++i) {
    DecodeRow(utt, in->rows + i * width, width, &dense);
    if (partial && in->decoder->NumFramesDecoded() % 20 == 0) StoreHyp(utt);
  }
  if (n > 0 && n < held)                                 // the rest moves to the front of the buffer
    memmove(in->rows, in->rows + n * width, sizeof(float) * (size_t)(held - n) * width);
}

}  // namespace

extern "C" {

ce_stt_t *ce_stt_init(const char *config_file) {
  std::unique_ptr<ce_stt_t> r(new ce_stt_t());
  Configuration conf;
  Status st = conf.Read(config_file);
  std::string graph_file;
  if (st.ok()) st = conf.GetPath("fst", &graph_file);
  if (st.ok()) {
    r->graph = fst::ConstFst<fst::StdArc>::Read(graph_file);
    if (!r->graph) st = Status::IOError(graph_file);
  }
  if (st.ok()) {
    r->am = new AcousticModel();
    st = r->am->Read(conf);
  }
  if (st.ok()) st = conf.GetInteger("left_context", &r->left);
  if (st.ok()) st = conf.GetInteger("right_context", &r->right);
  if (st.ok()) st = conf.GetInteger("chunk_size", &r->chunk);
  if (st.ok() && r->chunk < 1) st = Status::Corruption("chunk_size must be positive");
  std::string symbol_file;
  if (st.ok()) st = conf.GetPath("symbol_table", &symbol_file);
  if (st.ok()) {
    r->symbols = new SymbolTable();
    st = r->symbols->Read(symbol_file);
  }
  if (st.ok()) st = ReadDeltaLm(r.get(), conf);
  if (st.ok()) {
    const char *prec = getenv("CE_GPU_PRECISION");
    r->gpu = ce_gpu_model_load_config(config_file, prec ? atoi(prec) : CE_GPU_PRECISION_FP32, 0);
    if (!r->gpu) st = Status::IOError(ce_gpu_last_error());
  }
  if (st.ok()) {
    const char *o = getenv("CE_STT_GPU_OUTPUT");
    if (o && strcmp(o, "subset") == 0) {
      r->out_mode = CE_GPU_OUTPUT_SUBSET;
      if (!SelectGraphPdfs(r.get())) st = Status::Corruption(std::string("pdf subset: ") + ce_gpu_last_error());
    } else if (o && strncmp(o, "topk:", 5) == 0) {
      r->out_mode = CE_GPU_OUTPUT_TOPK;
      r->top_k = atoi(o + 5);
      if (ce_gpu_model_set_output(r->gpu, CE_GPU_OUTPUT_TOPK, nullptr, r->top_k) != CE_GPU_OK)
        st = Status::Corruption(ce_gpu_last_error());
    } else if (o && strcmp(o, "dense") != 0) {
      st = Status::Corruption(std::string("CE_STT_GPU_OUTPUT: ") + o);
    }
  }
  if (st.ok()) {
    const char *mx = getenv("CE_STT_GPU_MAX_UTTS");
    r->streams = ce_gpu_streams_create(r->gpu, mx ? std::max(1, atoi(mx)) : 64);
    if (!r->streams) st = Status::RuntimeError(ce_gpu_last_error());
  }
  if (!st.ok()) {
    SetError(st.what());
    ce_stt_destroy(r.release());
    return nullptr;
  }
  return r.release();
}

void ce_stt_destroy(ce_stt_t *r) {
  if (!r) return;
  ce_gpu_streams_free(r->streams);
  ce_gpu_model_free(r->gpu);
  delete r->delta_lm;
  delete r->large_lm;
  delete r->original_lm;
  delete r->symbols;
  delete r->am;
  delete r->graph;
  delete r;
}

ce_utt_t *ce_utt_init(ce_stt_t *r, const ce_wave_format_t *format) {
  std::unique_ptr<ce_utt_internal_t> in(new ce_utt_internal_t());
  in->recognizer = r;
  in->decoder.reset(new Decoder(r->graph, r->DecoderMap(), 0.1f, r->delta_lm));   // am_scale, src/ce_stt.cc:263
  in->decoder->Initialize();
  Status st = in->wave_reader.SetFormat(*format);
  if (!st.ok()) {
    SetError(st.what());
    return nullptr;
  }
  {
    std::lock_guard<std::mutex> lock(r->gpu_mu);
    in->slot = ce_gpu_streams_open(r->streams);
  }
  if (in->slot < 0) {
    SetError(ce_gpu_last_error());
    return nullptr;
  }
  ce_utt_t *utt = new ce_utt_t;
  utt->hyp = new char[1];
  utt->hyp[0] = '\0';
  utt->loglikelihood_per_frame = 0.0f;
  utt->internal = in.release();
  return utt;
}

void ce_utt_destroy(ce_utt_t *utt) {
  if (!utt) return;
  if (utt->internal && utt->internal->slot >= 0) {       // abandoned before end of stream: free the slot
    utt->internal->pending.clear();
    RunGpu(utt, /*end_of_stream=*/true);
  }
  delete[] utt->hyp;
  delete utt->internal;
  delete utt;
}

int32_t ce_stt_process(ce_utt_t *utt, const char *data, int32_t size) {
  if (!utt || !utt->internal) {
    SetError("utt is NULL");
    return CE_STT_FAILED;
  }
  ce_utt_internal_t *in = utt->internal;
  Vector<float> samples;                  // WaveReader keeps odd trailing bytes and handles 8/16/32 bit
  Status st = in->wave_reader.Process(data, size, &samples);
  if (!st.ok()) {
    SetError(st.what());
    return CE_STT_FAILED;
  }
  if (samples.Dim() == 0) return 0;
  for (int i = 0; i < samples.Dim(); ++i) {
    const float v = samples(i);           // the unscaled sample value (src/pcm_reader.cc:168-182)
    if (!(v >= -32768.0f && v <= 32767.0f)) {
      SetError(pocketkaldi::util::Format(
          "sample value {} does not fit 16 bits: the GPU front end takes 8-bit, 16-bit and 16-bit-range 32-bit PCM", v));
      return CE_STT_FAILED;
    }
    in->pending.push_back((int16_t)v);    // exact: v is an integer in range
  }
  in->samples_total += samples.Dim();
  // the reference computes a batch as soon as left + right + chunk_size frames are buffered; so does the GPU
  const int64_t due = RowsDue(in->recognizer, NumFrames(in->samples_total));
  if (due > in->rows_decoded) {
    if (due > in->rows_received && !RunGpu(utt, false)) return CE_STT_FAILED;
    Decode(utt, due, /*partial=*/true);
  }
  return samples.Dim();
}

void ce_stt_end_of_stream(ce_utt_t *utt) {
  if (!utt || !utt->internal) {
    SetError("utt is NULL");
    return;
  }
  ce_utt_internal_t *in = utt->internal;
  if (!RunGpu(utt, /*end_of_stream=*/true)) return;      // AcousticModel::EndOfStream, src/am.cc:144-164
  Decode(utt, in->rows_received, /*partial=*/false);
  in->decoder->EndOfStream();
  StoreHyp(utt);
}

ce_wave_format_t *ce_read_pcm_header(FILE *fp, ce_wave_format_t *format) {
  pocketkaldi::util::ReadableFile fd(fp);
  Status st = pocketkaldi::ReadPcmHeader(&fd, format);
  if (!st.ok()) {
    SetError(st.what());
    return nullptr;
  }
  return format;
}

const char *ce_stt_last_error() { return g_error; }

}  // extern "C"
