// batch_decode.cc -- `pocketkaldi_batch <config> <scp> [decoder threads]`: the scp mode of the
// reference's CLI (src/main.cc:55-84) evaluated as ONE batch over every GPU of the box.
//
// src/main.cc decodes a list one utterance at a time through the streaming C API.  Utterances are
// independent, so here the whole list goes through ce_host::ShardedModel: contiguous
// frame-balanced groups, one model handle + host thread per GPU, log-likelihood rows narrowed to the
// pdfs the graph can reach (exact, src/decoder.cc:97-102) into pinned host memory, and a pool of CPU
// threads running the reference's UNCHANGED Decoder on every utterance as soon as its rows have
// arrived -- chunk c is decoded while the GPUs work on chunk c + 1.  Output: the reference's own
// "<name> <hypothesis>" lines in list order; tests/test_dropin_decoder.py demands that they equal
// `pocketkaldi <config> <scp>` of the all-CPU reference.
//
// Linked with the reference's decoder.cc, fst.cc, hashtable.cc, symbol_table.cc, pcm_reader.cc,
// configuration.cc, am.cc (transition-id map only) and OpenFst, all unchanged; fbank.cc / srfft.cc /
// cmvn.cc are not linked and cblas_sgemm aborts (integration/cblas_forbidden.cc).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <memory>
#include <string>
#include <vector>

#include "am.h"
#include "ce_host.hpp"
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

namespace {

void Die(const std::string &what) {
  fprintf(stderr, "pocketkaldi_batch: %s\n", what.c_str());
  exit(1);
}
void Check(const Status &st) {
  if (!st.ok()) Die(st.what());
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <config> <scp> [decoder threads]\n", argv[0]);
    return 2;
  }
  const int n_threads = argc > 3 ? std::max(1, atoi(argv[3])) : 4;
  Configuration conf;
  Check(conf.Read(argv[1]));
  std::string graph_file, symbol_file;
  Check(conf.GetPath("fst", &graph_file));
  std::unique_ptr<fst::ConstFst<fst::StdArc>> graph(fst::ConstFst<fst::StdArc>::Read(graph_file));
  if (!graph) Die("unable to read " + graph_file);
  AcousticModel am;                                       // host copy: transition-id map, num_pdfs
  Check(am.Read(conf));
  Check(conf.GetPath("symbol_table", &symbol_file));
  SymbolTable symbols;
  Check(symbols.Read(symbol_file));
  std::unique_ptr<Vector<float>> original_lm;
  std::unique_ptr<LmFst> large_lm;
  std::unique_ptr<DeltaLmFst> delta_lm;
  const std::string large = conf.GetPathOrElse("large_lm", "");
  if (!large.empty()) {                                   // src/ce_stt.cc:84-113
    pocketkaldi::util::ReadableFile fd_o, fd_l;
    Check(fd_o.Open(conf.GetPathOrElse("original_lm", "")));
    original_lm.reset(new Vector<float>());
    Check(original_lm->Read(&fd_o));
    Check(fd_l.Open(large));
    large_lm.reset(new LmFst());
    Check(large_lm->Read(&fd_l));
    large_lm->InitBucket0();
    delta_lm.reset(new DeltaLmFst(original_lm.get(), large_lm.get(), &symbols));
  }

  // ---- the utterance list (src/main.cc:55-84) ----
  std::vector<std::string> names;
  std::vector<int16_t> pcm;
  std::vector<int64_t> sample_off(1, 0);
  {
    pocketkaldi::util::ReadableFile fd;
    Check(fd.Open(argv[2]));
    std::string line;
    Status st;
    while (fd.ReadLine(&line, &st) && st.ok()) {
      std::vector<std::string> fields = pocketkaldi::util::Split(line, " ");
      if (fields.size() != 2) Die("scp: unexpected line: " + line);
      Vector<float> wave;
      Check(pocketkaldi::Read16kPcm(fields[1].c_str(), &wave));
      for (int i = 0; i < wave.Dim(); ++i) {
        const float v = wave(i);                          // unscaled sample values, src/pcm_reader.cc:168-182
        if (!(v >= -32768.0f && v <= 32767.0f)) Die(fields[1] + ": sample value does not fit 16 bits");
        pcm.push_back((int16_t)v);
      }
      names.push_back(fields[0]);
      sample_off.push_back((int64_t)pcm.size());
    }
    Check(st);
  }
  const int n_utts = (int)names.size();

  // ---- the GPUs: rows narrowed to the pdfs some arc of the graph can ask for ----
  ce_host::ShardedModel gpus;
  const char *prec = getenv("CE_GPU_PRECISION");
  ce_host::Status gs = gpus.Read(argv[1], prec ? atoi(prec) : CE_GPU_PRECISION_FP32);
  if (!gs.ok()) Die(gs.what());
  const Vector<int32_t> &tid2pdf = am.TransitionPdfIdMap();
  Vector<int32_t> tid2col(tid2pdf.Dim());
  std::vector<int32_t> col_of(am.num_pdfs(), -1), ids;
  for (fst::StateIterator<fst::Fst<fst::StdArc>> si(*graph); !si.Done(); si.Next())
    for (fst::ArcIterator<fst::Fst<fst::StdArc>> ai(*graph, si.Value()); !ai.Done(); ai.Next()) {
      const int tid = ai.Value().ilabel;
      if (tid == 0) continue;
      if (tid < 0 || tid >= tid2pdf.Dim()) Die("graph: transition-id out of range");
      const int pdf = tid2pdf(tid);
      if (col_of[pdf] < 0) {
        col_of[pdf] = (int32_t)ids.size();
        ids.push_back(pdf);
      }
      tid2col(tid) = col_of[pdf];
    }
  gs = gpus.SelectPdfs(ids);
  if (!gs.ok()) Die(gs.what());

  // ---- forward over all GPUs; every utterance is decoded as soon as its rows are on the host ----
  std::vector<std::string> hyps((size_t)n_utts);
  auto decode = [&](int utt, const float *rows, int64_t n_frames, int width) {
    Decoder dec(graph.get(), tid2col, 0.1f, delta_lm.get());   // am_scale, src/ce_stt.cc:263
    dec.Initialize();
    for (int64_t r = 0; r < n_frames; ++r)
      dec.Process(pocketkaldi::SubVector<float>(const_cast<float *>(rows) + r * width, width));
    dec.EndOfStream();
    Decoder::Hypothesis hyp = dec.BestPath();
    std::vector<int> words = hyp.words();
    std::reverse(words.begin(), words.end());
    std::string text;
    for (size_t i = 0; i < words.size(); ++i) {
      if (i) text += ' ';
      text += symbols.Get(words[i]);
    }
    hyps[utt] = text;
  };
  const auto t0 = std::chrono::steady_clock::now();
  // the delta-LM FST caches internally and is shared: one decoder thread then
  std::vector<int64_t> frame_off;
  gs = gpus.Forward(pcm.data(), sample_off.data(), n_utts, decode, delta_lm ? 1 : n_threads, &frame_off);
  if (!gs.ok()) Die(gs.what());
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (int u = 0; u < n_utts; ++u) {
    if (frame_off[u + 1] == frame_off[u]) {               // shorter than one frame: the empty hypothesis
      Decoder dec(graph.get(), tid2col, 0.1f, delta_lm.get());
      dec.Initialize();
      dec.EndOfStream();
    }
    printf("%s %s\n", names[u].c_str(), hyps[u].c_str());
  }
  fprintf(stderr, "pocketkaldi_batch: %d utterances, %.1f s of audio, %d GPU(s), %d decoder thread(s): %.3f s (%.0f x real time)\n",
          n_utts, pcm.size() / 16000.0, gpus.n_devices(), delta_lm ? 1 : n_threads, sec, pcm.size() / 16000.0 / sec);
  return 0;
}
