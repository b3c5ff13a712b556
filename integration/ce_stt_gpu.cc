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
// Difference in behaviour, by design: ce_stt_process only buffers PCM; the whole utterance goes
// through ce_gpu_forward at ce_stt_end_of_stream (chunked and whole-utterance evaluation are the same
// function), then the rows are fed to the unchanged Decoder::Process one by one, exactly
// like src/ce_stt.cc:349-357.  The delta-LM rescoring option (src/ce_stt.cc:84-113) is not wired.
#include "ce_stt.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>
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
using pocketkaldi::Matrix;
using pocketkaldi::Status;
using pocketkaldi::SymbolTable;
using pocketkaldi::Vector;
using pocketkaldi::WaveReader;

struct ce_stt_t {
  fst::ConstFst<fst::StdArc> *graph = nullptr;
  AcousticModel *am = nullptr;            // host copy: transition-id map and num_pdfs only
  SymbolTable *symbols = nullptr;
  ce_gpu_model_t *gpu = nullptr;          // the model that actually runs
};

struct ce_utt_internal_t {
  const ce_stt_t *recognizer = nullptr;
  WaveReader wave_reader;
  std::vector<int16_t> pcm;               // the utterance so far (unscaled 16-bit samples)
  std::unique_ptr<Decoder> decoder;
};

namespace {

thread_local char g_error[2048] = "";     // ce_stt_last_error(): thread-local here

void SetError(const std::string &msg) {
  strncpy(g_error, msg.c_str(), sizeof(g_error) - 1);
  g_error[sizeof(g_error) - 1] = '\0';
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
  std::string symbol_file;
  if (st.ok()) st = conf.GetPath("symbol_table", &symbol_file);
  if (st.ok()) {
    r->symbols = new SymbolTable();
    st = r->symbols->Read(symbol_file);
  }
  if (st.ok()) {
    const char *prec = getenv("CE_GPU_PRECISION");
    r->gpu = ce_gpu_model_load_config(config_file, prec ? atoi(prec) : CE_GPU_PRECISION_FP32, 0);
    if (!r->gpu) st = Status::IOError(ce_gpu_last_error());
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
  ce_gpu_model_free(r->gpu);
  delete r->symbols;
  delete r->am;
  delete r->graph;
  delete r;
}

ce_utt_t *ce_utt_init(ce_stt_t *r, const ce_wave_format_t *format) {
  std::unique_ptr<ce_utt_internal_t> in(new ce_utt_internal_t());
  in->recognizer = r;
  in->decoder.reset(new Decoder(r->graph, r->am->TransitionPdfIdMap(), 0.1f, nullptr));   // am_scale, src/ce_stt.cc:263
  in->decoder->Initialize();
  Status st = in->wave_reader.SetFormat(*format);
  if (!st.ok()) {
    SetError(st.what());
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
  delete[] utt->hyp;
  delete utt->internal;
  delete utt;
}

int32_t ce_stt_process(ce_utt_t *utt, const char *data, int32_t size) {
  if (!utt || !utt->internal) {
    SetError("utt is NULL");
    return CE_STT_FAILED;
  }
  Vector<float> samples;                  // WaveReader keeps odd trailing bytes and handles 8/16/32 bit
  Status st = utt->internal->wave_reader.Process(data, size, &samples);
  if (!st.ok()) {
    SetError(st.what());
    return CE_STT_FAILED;
  }
  std::vector<int16_t> &pcm = utt->internal->pcm;
  for (int i = 0; i < samples.Dim(); ++i) {
    const float v = samples(i);           // unscaled sample value (src/pcm_reader.cc:168-182)
    pcm.push_back((int16_t)std::max(-32768.0f, std::min(32767.0f, v)));
  }
  return samples.Dim();
}

void ce_stt_end_of_stream(ce_utt_t *utt) {
  if (!utt || !utt->internal) {
    SetError("utt is NULL");
    return;
  }
  ce_utt_internal_t *in = utt->internal;
  const int64_t soff[2] = {0, (int64_t)in->pcm.size()};
  int64_t foff[2] = {0, 0};
  const int64_t frames = ce_gpu_frame_offsets(soff, 1, foff);
  if (frames > 0) {
    Matrix<float> log_prob((int)frames, in->recognizer->am->num_pdfs());
    if (log_prob.Stride() != log_prob.NumCols() ||
        ce_gpu_forward(in->recognizer->gpu, in->pcm.data(), soff, 1, log_prob.Data(), nullptr, nullptr,
                       nullptr) != CE_GPU_OK) {
      SetError(ce_gpu_last_error());
      return;
    }
    for (int r = 0; r < log_prob.NumRows(); ++r) in->decoder->Process(log_prob.Row(r));
  }
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
