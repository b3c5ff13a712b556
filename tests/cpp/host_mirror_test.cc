// Drives include/ce_host.hpp the way src/ce_stt.cc:295-362 drives the reference's stage objects:
// PCM in 1 KB pieces -> Fbank::Process -> [CMVN] -> AcousticModel::Process per frame ->
// EndOfStream; the log-likelihood rows a decoder would consume are written to a file.
//
//   host_mirror_test errors
//   host_mirror_test stream <conf> <pcm.s16le> <out.bin> <precision 0..3> [cmvn_stats.vec0]
//   host_mirror_test streams <conf> <precision> <cmvn_stats.vec0 | -> <seed> <out_prefix> <pcm>...
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <string>
#include <vector>

#include "ce_host.hpp"

using namespace ce_host;

static int Fail(const char *what, const Status &st) {
  fprintf(stderr, "FAIL %s: %s\n", what, st.what().c_str());
  return 1;
}

static int Errors() {
  AcousticModel am;
  Status st = am.Read("/nonexistent/dir/model.conf");
  if (st.ok() || st.what().empty()) return Fail("Read of a missing config must fail", st);
  Matrix lp;
  AcousticModel::Instance inst;
  float frame[40] = {0};
  st = am.Process(&inst, frame, &lp);
  if (st.ok()) return Fail("Process before Read must fail", st);
  std::vector<int32_t> v;
  st = detail::ReadVec0<int32_t>("/nonexistent/file", &v);
  if (st.ok()) return Fail("ReadVec0 of a missing file must fail", st);
  if (ce_gpu_device_count() == 0) {                        // no CPU fallback behind the mirror
    Fbank fb;
    Fbank::Instance fi;
    std::vector<int16_t> pcm(1600, 100);
    Matrix feats;
    st = fb.Process(&fi, pcm.data(), (int)pcm.size(), &feats);
    if (st.ok() || st.what().find("no CUDA device") == std::string::npos)
      return Fail("Fbank::Process without a device must report it", st);
  }
  Fbank fb;
  Fbank::Instance fi;
  Matrix feats;
  std::vector<float> bad(500, 0.5f);                        // not int16-valued
  st = fb.Process(&fi, bad, &feats);
  if (st.ok()) return Fail("non-integral samples must be rejected", st);
  printf("OK\n");
  return 0;
}

static int Stream(int argc, char **argv) {
  const std::string conf = argv[2], pcm_path = argv[3], out_path = argv[4];
  const int precision = atoi(argv[5]);
  AcousticModel am;
  Status st = am.Read(conf, precision, 0);
  if (!st.ok()) return Fail("AcousticModel::Read", st);
  std::vector<float> stats;
  if (argc > 6) {
    st = detail::ReadVec0<float>(argv[6], &stats);
    if (!st.ok()) return Fail("cmvn stats", st);
  }
  FILE *f = fopen(pcm_path.c_str(), "rb");
  if (!f) return Fail("open pcm", Status::IOError(pcm_path));
  Fbank fbank(am.feat_dim(), 0);
  Fbank::Instance fbank_inst;
  AcousticModel::Instance am_inst;
  std::vector<float> rows;                                  // what Decoder::Process would receive
  Matrix all_feats(0, am.feat_dim());
  char buf[1024];                                           // src/main.cc:38-44 reads 1024-byte pieces
  size_t got;
  while ((got = fread(buf, 1, sizeof(buf), f)) > 0) {
    Matrix feats;
    st = fbank.Process(&fbank_inst, reinterpret_cast<const int16_t *>(buf), (int)(got / 2), &feats);
    if (!st.ok()) return Fail("Fbank::Process", st);
    all_feats.data.insert(all_feats.data.end(), feats.data.begin(), feats.data.end());
    all_feats.rows += feats.rows;
  }
  fclose(f);
  // The reference runs no CMVN between fbank and AM (src/ce_stt.cc:307-331); with stats given the
  // test inserts it, whole utterance then frame by frame in order.
  Matrix norm = all_feats;
  if (!stats.empty()) {
    CMVN cmvn(stats, all_feats, 0);
    for (int t = 0; t < all_feats.rows; ++t) {
      st = cmvn.GetFrame(t, norm.Row(t));
      if (!st.ok()) return Fail("CMVN::GetFrame", st);
    }
  }
  int batches = 0;
  for (int t = 0; t < norm.rows; ++t) {
    Matrix lp;
    st = am.Process(&am_inst, norm.Row(t), &lp);
    if (!st.ok()) return Fail("AcousticModel::Process", st);
    if (lp.NumRows()) ++batches;
    rows.insert(rows.end(), lp.data.begin(), lp.data.end());
  }
  Matrix lp;
  st = am.EndOfStream(&am_inst, &lp);
  if (!st.ok()) return Fail("AcousticModel::EndOfStream", st);
  rows.insert(rows.end(), lp.data.begin(), lp.data.end());
  const int32_t hdr[3] = {(int32_t)(rows.size() / am.num_pdfs()), am.num_pdfs(), batches};
  FILE *o = fopen(out_path.c_str(), "wb");
  if (!o) return Fail("open out", Status::IOError(out_path));
  fwrite(hdr, 4, 3, o);
  fwrite(rows.data(), 4, rows.size(), o);
  fclose(o);
  printf("OK rows=%d cols=%d batches=%d tid2pdf=%zu\n", hdr[0], hdr[1], batches, am.TransitionPdfIdMap().size());
  return 0;
}

// HOST_MIRROR_SELECT = "subset:<id>,<id>,..." | "topk:<k>": narrower rows for the decoder
static Status SelectFromEnv(AcousticModel *am) {
  const char *sel = getenv("HOST_MIRROR_SELECT");
  if (!sel) return Status::OK();
  if (strncmp(sel, "topk:", 5) == 0) return am->SelectTopK(atoi(sel + 5));
  if (strncmp(sel, "subset:", 7) != 0) return Status::RuntimeError("HOST_MIRROR_SELECT");
  std::vector<int32_t> ids;
  for (const char *p = sel + 7; *p;) {
    char *end;
    ids.push_back((int32_t)strtol(p, &end, 10));
    p = (*end == ',') ? end + 1 : end;
  }
  return am->SelectPdfs(ids);
}

// Feeds the utterances in random-sized pieces through one batcher until all have ended; writes
// <prefix>.<i>.bin with the rows of stream i in the order they came out.
template <typename Batch, typename Rng>
static int RunStreams(const Batch &batch, const AcousticModel &am, const std::vector<std::vector<int16_t>> &audio,
                      Rng &next, const std::string &prefix) {
  const int n = (int)audio.size();
  Status st;
  std::vector<typename Batch::Stream> state(n);
  std::vector<size_t> pos(n, 0);
  std::vector<std::vector<float>> rows(n);
  int calls = 0, live = n;
  while (live > 0) {
    std::vector<typename Batch::Stream *> streams;
    std::vector<const int16_t *> pcm;
    std::vector<int> cnt;
    std::vector<bool> eos;
    std::vector<int> idx;
    for (int i = 0; i < n; ++i) {
      if (state[i].ended) continue;
      size_t take = next() % 3 == 0 ? 0 : next() % 9000;     // sometimes nothing arrives
      take = std::min(take, audio[i].size() - pos[i]);
      streams.push_back(&state[i]);
      pcm.push_back(audio[i].data() + pos[i]);
      cnt.push_back((int)take);
      pos[i] += take;
      eos.push_back(pos[i] == audio[i].size());
      idx.push_back(i);
    }
    std::vector<Matrix> out;
    st = batch.Process(streams, pcm, cnt, eos, &out);
    if (!st.ok()) return Fail("StreamBatch::Process", st);
    for (size_t k = 0; k < idx.size(); ++k) {
      rows[idx[k]].insert(rows[idx[k]].end(), out[k].data.begin(), out[k].data.end());
      if (state[idx[k]].ended) --live;
    }
    ++calls;
  }
  for (int i = 0; i < n; ++i) {
    const int32_t hdr[3] = {(int32_t)(rows[i].size() / am.output_width()), am.output_width(), calls};
    FILE *o = fopen((prefix + "." + std::to_string(i) + ".bin").c_str(), "wb");
    if (!o) return Fail("open out", Status::IOError(prefix));
    fwrite(hdr, 4, 3, o);
    fwrite(rows[i].data(), 4, rows[i].size(), o);
    fclose(o);
  }
  printf("OK streams=%d calls=%d\n", n, calls);
  return 0;
}


// streams <conf> <precision> <cmvn_stats.vec0 | -> <seed> <out_prefix> <pcm.s16le>...
// Several live utterances fed in random-sized pieces through ONE StreamBatch; writes
// <out_prefix>.<i>.bin with the rows of stream i in the order they came out.
static int Streams(int argc, char **argv) {
  const std::string conf = argv[2], stats_path = argv[4], prefix = argv[6];
  const int precision = atoi(argv[3]);
  uint64_t rng = strtoull(argv[5], nullptr, 10) * 2654435761ull + 12345;
  auto next = [&rng]() {
    rng = rng * 6364136223846793005ull + 1442695040888963407ull;
    return (uint32_t)(rng >> 33);
  };
  AcousticModel am;
  Status st = am.Read(conf, precision, 0);
  if (!st.ok()) return Fail("AcousticModel::Read", st);
  std::vector<float> stats;
  if (stats_path != "-") {
    st = detail::ReadVec0<float>(stats_path, &stats);
    if (!st.ok()) return Fail("cmvn stats", st);
  }
  st = SelectFromEnv(&am);
  if (!st.ok()) return Fail("output selection", st);
  const int n = argc - 7;
  std::vector<std::vector<int16_t>> audio(n);
  for (int i = 0; i < n; ++i) {
    FILE *f = fopen(argv[7 + i], "rb");
    if (!f) return Fail("open pcm", Status::IOError(argv[7 + i]));
    int16_t buf[4096];
    size_t got;
    while ((got = fread(buf, 2, 4096, f)) > 0) audio[i].insert(audio[i].end(), buf, buf + got);
    fclose(f);
  }
  if (getenv("HOST_MIRROR_DEVICE_STATE"))              // same schedule, state on the device
    return RunStreams<DeviceStreamBatch>(DeviceStreamBatch(&am, n + 2), am, audio, next, prefix);
  return RunStreams<StreamBatch>(StreamBatch(&am, stats, 0), am, audio, next, prefix);
}

// streambench <conf> <precision> <cmvn_stats.vec0 | -> <n_streams> <samples_per_call> <calls>
// Serving-shaped timing: n live streams each receive samples_per_call new samples per call; wall
// milliseconds per call (median) for host-resident (StreamBatch) and device-resident
// (DeviceStreamBatch) state.  Not a test: tools/stream_probe.py runs it.
template <typename Batch>
static double TimeCalls(const Batch &batch, int n, int per_call, int calls) {
  std::vector<typename Batch::Stream> state(n);
  std::vector<typename Batch::Stream *> streams;
  std::vector<std::vector<int16_t>> audio(n, std::vector<int16_t>(per_call));
  std::vector<const int16_t *> pcm;
  uint32_t x = 12345;
  for (int i = 0; i < n; ++i) {
    for (int16_t &v : audio[i]) {
      x = x * 1664525u + 1013904223u;
      v = (int16_t)((int)(x >> 18) - 8192);
    }
    streams.push_back(&state[i]);
    pcm.push_back(audio[i].data());
  }
  const std::vector<int> cnt(n, per_call);
  const std::vector<bool> eos(n, false);
  std::vector<Matrix> out;
  std::vector<double> ms;
  for (int c = 0; c < calls + 3; ++c) {                     // 3 warm-up calls
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    Status st = batch.Process(streams, pcm, cnt, eos, &out);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (!st.ok()) return -1.0;
    if (c >= 3) ms.push_back(1e3 * (t1.tv_sec - t0.tv_sec) + 1e-6 * (t1.tv_nsec - t0.tv_nsec));
  }
  std::sort(ms.begin(), ms.end());                          // the median: the shared boxes hiccup
  return ms[ms.size() / 2];
}

static int StreamBench(char **argv) {
  const std::string conf = argv[2], stats_path = argv[4];
  const int n = atoi(argv[5]), per_call = atoi(argv[6]), calls = atoi(argv[7]);
  AcousticModel am;
  Status st = am.Read(conf, atoi(argv[3]), 0);
  if (!st.ok()) return Fail("AcousticModel::Read", st);
  st = SelectFromEnv(&am);
  if (!st.ok()) return Fail("output selection", st);
  std::vector<float> stats;
  if (stats_path != "-") {
    st = detail::ReadVec0<float>(stats_path, &stats);
    if (!st.ok()) return Fail("cmvn stats", st);
  }
  const double host_ms = TimeCalls(StreamBatch(&am, stats, 0), n, per_call, calls);
  ce_gpu_profile_enable(1);                                // kernel time of the device-state calls
  double kernel_ms[CE_GPU_PROFILE_CATEGORIES];
  int64_t launches[CE_GPU_PROFILE_CATEGORIES];
  const double dev_ms = TimeCalls(DeviceStreamBatch(&am, n), n, per_call, calls);
  ce_gpu_profile_read(kernel_ms, launches);
  ce_gpu_profile_enable(0);
  if (host_ms < 0 || dev_ms < 0) return Fail("Process", Status::RuntimeError("streambench"));
  double gpu_ms = 0.0;
  for (double v : kernel_ms) gpu_ms += v;
  const double per = 1.0 / (calls + 3);
  printf("streams=%d samples_per_call=%d row_words=%d host_state_ms=%.3f device_state_ms=%.3f "
         "(kernels %.3f ms per call: fbank %.3f cmvn %.3f gemm %.3f quantize %.3f finalize %.3f copies %.3f)\n",
         n, per_call, am.output_width(), host_ms, dev_ms, gpu_ms * per, kernel_ms[0] * per, kernel_ms[1] * per,
         kernel_ms[2] * per, kernel_ms[3] * per, kernel_ms[4] * per, kernel_ms[5] * per);
  return 0;
}

int main(int argc, char **argv) {
  if (argc >= 8 && strcmp(argv[1], "streambench") == 0) return StreamBench(argv);
  if (argc >= 2 && strcmp(argv[1], "errors") == 0) return Errors();
  if (argc >= 6 && strcmp(argv[1], "stream") == 0) return Stream(argc, argv);
  if (argc >= 8 && strcmp(argv[1], "streams") == 0) return Streams(argc, argv);
  fprintf(stderr, "usage: %s errors | stream <conf> <pcm> <out> <precision> [cmvn_stats]\n", argv[0]);
  return 2;
}
