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
//
// CE_STT_GPU_OUTPUT selects what crosses PCIe per frame (ce_gpu_model_set_output):
//   unset / "dense"  all num_pdfs log-likelihoods (12 KB a frame at 3072 pdfs);
//   "subset"         only the pdfs the graph's input labels can reach; the decoder gets the
//                    matching remapped transition-id map, so it computes exactly what it computed
//                    from the dense row;
//   "topk:<k>"       the k best (loglik, pdf) pairs; the row handed to the decoder holds those and
//                    the k-th value (an upper bound) for every other pdf -- an approximation.
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
  // what a row of ce_gpu_forward's output is, and the transition-id map that goes with it
  int out_mode = CE_GPU_OUTPUT_DENSE;
  int top_k = 0;
  Vector<int32_t> tid2col;                // subset mode: transition-id -> column of the gathered row
  const Vector<int32_t> &DecoderMap() const {
    return out_mode == CE_GPU_OUTPUT_SUBSET ? tid2col : am->TransitionPdfIdMap();
  }
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
  in->decoder.reset(new Decoder(r->graph, r->DecoderMap(), 0.1f, nullptr));   // am_scale, src/ce_stt.cc:263
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
    const ce_stt_t *rec = in->recognizer;
    const int width = ce_gpu_model_output_width(rec->gpu);          // 4-byte words per row
    std::vector<float> rows((size_t)frames * width);
    if (ce_gpu_forward(rec->gpu, in->pcm.data(), soff, 1, rows.data(), nullptr, nullptr, nullptr) != CE_GPU_OK) {
      SetError(ce_gpu_last_error());
      return;
    }
    if (rec->out_mode == CE_GPU_OUTPUT_TOPK) {
      Vector<float> dense(rec->am->num_pdfs());
      for (int64_t r = 0; r < frames; ++r) {
        const ce_gpu_scored_pdf_t *best = reinterpret_cast<const ce_gpu_scored_pdf_t *>(rows.data() + r * width);
        for (int j = 0; j < dense.Dim(); ++j) dense(j) = best[rec->top_k - 1].loglik;
        for (int j = 0; j < rec->top_k; ++j) dense(best[j].pdf) = best[j].loglik;
        in->decoder->Process(dense);
      }
    } else {                                             // dense or gathered rows, as they are
      for (int64_t r = 0; r < frames; ++r)
        in->decoder->Process(pocketkaldi::SubVector<float>(rows.data() + r * width, width));
    }
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
