// oracle/ref_shim.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin C ABI over the UNMODIFIED reference sources, which are compiled where
// they lie under /root/reference by oracle/Makefile into oracle/_ref/libce_ref.so.
// Nothing here restates an algorithm: every entry point just drives the
// reference's own classes so tests can compare the CUDA path with them.
//
//   ref_srfft            -> SRFFT::Compute                 (src/srfft.cc:370)
//   ref_fbank_*          -> WaveReader::Process + Fbank::Process
//                                                          (src/pcm_reader.cc:148, src/fbank.cc:265)
//   ref_cmvn             -> CMVN::GetFrame                 (src/cmvn.cc:100)
//   ref_nnet_*           -> Nnet::Read / Nnet::Propagate   (src/nnet.cc:273,295)
//   ref_am_*             -> AcousticModel::Read/Process/EndOfStream (src/am.cc:26,115,144)
//   ref_quantize         -> Quantize                       (src/matrix.cc:366)
//   ref_gemm_u8          -> MatMat_U8U8F32                 (src/matrix.cc:389)
//   ref_u8_*             -> the int8 layer composition of SURVEY.md D3 / §8c:
//                           Quantize(in) + MatMat_U8U8F32 + AddVec(b) per Linear
//                           layer, every other layer the reference's own class.
//
// cblas_sgemm (the only external arithmetic, src/matrix.cc:308) is defined here
// and is either a deterministic in-order fp32 loop (the summation order of the
// reference's SimpleMatMat, src/matrix.cc:275-292) or a dlopen()ed OpenBLAS.

#include <assert.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <memory>
#include <string>
#include <vector>

#include <cblas.h>

#include "am.h"
#include "cmvn.h"
#include "configuration.h"
#include "fbank.h"
#include "matrix.h"
#include "nnet.h"
#include "pcm_reader.h"
#include "srfft.h"
#include "util.h"
#include "vector.h"
#include "gemmlowp/public/gemmlowp.h"

using namespace pocketkaldi;

// ---------------------------------------------------------------------------
// cblas_sgemm back-ends
// ---------------------------------------------------------------------------
namespace {

typedef void (*sgemm_fn)(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE,
                         enum CBLAS_TRANSPOSE, int, int, int, float,
                         const float *, int, const float *, int, float, float *,
                         int);

int g_sgemm_backend = 0;  // 0 = in-order fp32 loop, 1 = OpenBLAS
sgemm_fn g_openblas_sgemm = nullptr;
void *g_openblas_handle = nullptr;

// C[i][j] = sum_k A[i][k]*B[k][j], k ascending, every partial sum rounded to
// fp32: per element this is the summation order of SimpleMatMat.
void InOrderSgemm(int m, int n, int k, const float *a, int lda, const float *b,
                  int ldb, float *c, int ldc) {
  for (int i = 0; i < m; ++i) {
    float *crow = c + (size_t)i * ldc;
    for (int j = 0; j < n; ++j) crow[j] = 0.0f;
    const float *arow = a + (size_t)i * lda;
    for (int kk = 0; kk < k; ++kk) {
      const float av = arow[kk];
      const float *brow = b + (size_t)kk * ldb;
      for (int j = 0; j < n; ++j) crow[j] += av * brow[j];
    }
  }
}

}  // namespace

extern "C" void cblas_sgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta,
                            enum CBLAS_TRANSPOSE tb, int m, int n, int k,
                            float alpha, const float *a, int lda,
                            const float *b, int ldb, float beta, float *c,
                            int ldc) {
  if (g_sgemm_backend == 1 && g_openblas_sgemm != nullptr) {
    g_openblas_sgemm(order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
    return;
  }
  // The reference only ever calls RowMajor/NoTrans/NoTrans, alpha=1, beta=0.
  assert(order == CblasRowMajor && ta == CblasNoTrans && tb == CblasNoTrans);
  assert(alpha == 1.0f && beta == 0.0f);
  InOrderSgemm(m, n, k, a, lda, b, ldb, c, ldc);
}

extern "C" {

// backend 0: in-order loop; 1: OpenBLAS loaded from `path`. Returns 0 on success.
int ref_set_sgemm_backend(int backend, const char *path) {
  if (backend == 0) {
    g_sgemm_backend = 0;
    return 0;
  }
  if (g_openblas_sgemm == nullptr) {
    if (path == nullptr) return -1;
    g_openblas_handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (g_openblas_handle == nullptr) {
      fprintf(stderr, "ref_set_sgemm_backend: %s\n", dlerror());
      return -2;
    }
    g_openblas_sgemm = (sgemm_fn)dlsym(g_openblas_handle, "cblas_sgemm");
    if (g_openblas_sgemm == nullptr) return -3;
  }
  g_sgemm_backend = 1;
  return 0;
}

int ref_fbank_dim() { return PK_FBANK_DIM; }

// ---------------------------------------------------------------------------
// FFT / fbank / CMVN
// ---------------------------------------------------------------------------

// In-place forward real FFT of n floats (n a power of two), packed output.
int ref_srfft(float *data, int n) {
  SRFFT fft(n);
  std::vector<float> buffer(n);
  fft.Compute(data, n, true, buffer.data(), n);
  return 0;
}

// Whole-buffer fbank of 16-bit PCM. Returns the number of frames written
// (rows of PK_FBANK_DIM floats), or -1 if `cap_frames` is too small.
int ref_fbank_pcm16(const int16_t *pcm, int n_samples, float *out,
                    int cap_frames) {
  if (n_samples <= 0) return 0;
  ce_wave_format_t fmt = {1, 16000, 16};
  WaveReader reader;
  Status s = reader.SetFormat(fmt);
  if (!s.ok()) return -2;
  Vector<float> samples;
  s = reader.Process(reinterpret_cast<const char *>(pcm), n_samples * 2, &samples);
  if (!s.ok()) return -3;

  Fbank fbank;
  Fbank::Instance inst;
  Matrix<float> feat;
  fbank.Process(&inst, samples, &feat);
  if (feat.NumRows() > cap_frames) return -1;
  for (int r = 0; r < feat.NumRows(); ++r)
    memcpy(out + (size_t)r * PK_FBANK_DIM, feat.Row(r).Data(),
           sizeof(float) * PK_FBANK_DIM);
  return feat.NumRows();
}

// Streaming fbank: the byte stream is fed in `chunk_bytes` pieces through
// WaveReader + Fbank::Instance (the shape of test/fbank_test.cc:85-136).
int ref_fbank_stream(const char *bytes, int n_bytes, int chunk_bytes,
                     float *out, int cap_frames) {
  ce_wave_format_t fmt = {1, 16000, 16};
  WaveReader reader;
  if (!reader.SetFormat(fmt).ok()) return -2;
  Fbank fbank;
  Fbank::Instance inst;
  Vector<float> samples;
  Matrix<float> feat;
  int total = 0;
  for (int pos = 0; pos < n_bytes; pos += chunk_bytes) {
    int len = n_bytes - pos < chunk_bytes ? n_bytes - pos : chunk_bytes;
    if (!reader.Process(bytes + pos, len, &samples).ok()) return -3;
    if (samples.Dim() == 0) continue;
    fbank.Process(&inst, samples, &feat);
    for (int r = 0; r < feat.NumRows(); ++r) {
      if (total >= cap_frames) return -1;
      memcpy(out + (size_t)total * PK_FBANK_DIM, feat.Row(r).Data(),
             sizeof(float) * PK_FBANK_DIM);
      ++total;
    }
  }
  return total;
}

// Online CMVN over a whole [T x PK_FBANK_DIM] matrix, frames in order.
int ref_cmvn(const float *global_stats, const float *feats, int T, float *out) {
  if (T <= 0) return 0;
  Vector<float> g(PK_FBANK_DIM + 1);
  memcpy(g.Data(), global_stats, sizeof(float) * (PK_FBANK_DIM + 1));
  Matrix<float> raw(T, PK_FBANK_DIM);
  memcpy(raw.Data(), feats, sizeof(float) * (size_t)T * PK_FBANK_DIM);
  CMVN cmvn(g, raw);
  Vector<float> row(PK_FBANK_DIM);
  for (int t = 0; t < T; ++t) {
    cmvn.GetFrame(t, &row);
    memcpy(out + (size_t)t * PK_FBANK_DIM, row.Data(), sizeof(float) * PK_FBANK_DIM);
  }
  return T;
}

// ---------------------------------------------------------------------------
// Nnet (float path)
// ---------------------------------------------------------------------------

void *ref_nnet_open(const char *path) {
  util::ReadableFile fd;
  if (!fd.Open(path).ok()) return nullptr;
  Nnet *nnet = new Nnet();
  Status s = nnet->Read(&fd);
  if (!s.ok()) {
    fprintf(stderr, "ref_nnet_open: %s\n", s.what().c_str());
    delete nnet;
    return nullptr;
  }
  return nnet;
}

void ref_nnet_close(void *h) { delete static_cast<Nnet *>(h); }

// Nnet::Propagate on a [rows x cols] matrix. Writes at most cap floats.
int ref_nnet_propagate(void *h, const float *in, int rows, int cols, float *out,
                       long cap, int *out_rows, int *out_cols) {
  Nnet *nnet = static_cast<Nnet *>(h);
  SubMatrix<float> x(const_cast<float *>(in), rows, cols, cols);
  Matrix<float> y;
  nnet->Propagate(x, &y);
  *out_rows = y.NumRows();
  *out_cols = y.NumCols();
  if ((long)y.NumRows() * y.NumCols() > cap) return -1;
  for (int r = 0; r < y.NumRows(); ++r)
    memcpy(out + (size_t)r * y.NumCols(), y.Row(r).Data(),
           sizeof(float) * y.NumCols());
  return 0;
}

// ---------------------------------------------------------------------------
// AcousticModel (float path, streaming exactly as ce_stt.cc:317-331,351-357)
// ---------------------------------------------------------------------------

void *ref_am_open(const char *conf_path) {
  Configuration conf;
  Status s = conf.Read(conf_path);
  if (!s.ok()) {
    fprintf(stderr, "ref_am_open: %s\n", s.what().c_str());
    return nullptr;
  }
  AcousticModel *am = new AcousticModel();
  s = am->Read(conf);
  if (!s.ok()) {
    fprintf(stderr, "ref_am_open: %s\n", s.what().c_str());
    delete am;
    return nullptr;
  }
  return am;
}

void ref_am_close(void *h) { delete static_cast<AcousticModel *>(h); }

int ref_am_num_pdfs(void *h) { return static_cast<AcousticModel *>(h)->num_pdfs(); }

// Feeds T feature rows one by one through Process, then EndOfStream; appends
// every returned log_prob row to `out`. Returns rows written (or -1 on overflow).
int ref_am_forward(void *h, const float *feats, int T, int dim, float *out,
                   int cap_rows, int *out_cols) {
  AcousticModel *am = static_cast<AcousticModel *>(h);
  AcousticModel::Instance inst;
  Matrix<float> log_prob;
  int total = 0;
  *out_cols = 0;
  auto append = [&](const Matrix<float> &m) -> bool {
    if (m.NumRows() == 0) return true;
    *out_cols = m.NumCols();
    for (int r = 0; r < m.NumRows(); ++r) {
      if (total >= cap_rows) return false;
      memcpy(out + (size_t)total * m.NumCols(), m.Row(r).Data(),
             sizeof(float) * m.NumCols());
      ++total;
    }
    return true;
  };
  for (int t = 0; t < T; ++t) {
    SubVector<float> row(const_cast<float *>(feats) + (size_t)t * dim, dim);
    am->Process(&inst, row, &log_prob);
    if (!append(log_prob)) return -1;
  }
  am->EndOfStream(&inst, &log_prob);
  if (!append(log_prob)) return -1;
  return total;
}

// ---------------------------------------------------------------------------
// Quantisation + u8 GEMM
// ---------------------------------------------------------------------------

void ref_quantize(const float *src, int rows, int cols, uint8_t *dst,
                  float *scale, int32_t *zero_point) {
  SubMatrix<float> m(const_cast<float *>(src), rows, cols, cols);
  Matrix<uint8_t> q;
  QuantizationParams p;
  Quantize(m, &q, &p);
  memcpy(dst, q.Data(), (size_t)rows * cols);
  *scale = p.scale;
  *zero_point = p.zero_point;
}

// The int32 accumulators gemmlowp produces for MatMat_U8U8F32's arguments:
// same template, offsets and (empty) output pipeline as
// eight_bit_int_gemm.cc:107-133.
static void GemmlowpInt32(const uint8_t *a, int32_t zp_a, const uint8_t *b,
                          int32_t zp_b, int m, int n, int k, int32_t *acc) {
  gemmlowp::GemmContext context;
  gemmlowp::MatrixMap<const std::uint8_t, gemmlowp::MapOrder::RowMajor> lhs(a, m, k, k);
  gemmlowp::MatrixMap<const std::uint8_t, gemmlowp::MapOrder::RowMajor> rhs(b, k, n, n);
  gemmlowp::MatrixMap<std::int32_t, gemmlowp::MapOrder::RowMajor> result(acc, m, n, n);
  auto empty_pipeline = std::make_tuple();
  gemmlowp::GemmWithOutputPipeline<std::uint8_t, std::int32_t,
                                   gemmlowp::DefaultL8R8BitDepthParams>(
      &context, lhs, rhs, &result, -zp_a, -zp_b, empty_pipeline);
}

// C = MatMat_U8U8F32(A, B); optionally also the int32 accumulators.
void ref_gemm_u8(const uint8_t *a, float scale_a, int32_t zp_a, const uint8_t *b,
                 float scale_b, int32_t zp_b, int m, int n, int k, float *c,
                 int32_t *acc /*nullable*/) {
  // (SubMatrix<uint8_t> is not instantiated by the reference: copy into Matrix.)
  Matrix<uint8_t> A(m, k, Matrix<uint8_t>::kUndefined);
  Matrix<uint8_t> B(k, n, Matrix<uint8_t>::kUndefined);
  memcpy(A.Data(), a, (size_t)m * k);
  memcpy(B.Data(), b, (size_t)k * n);
  SubMatrix<float> C(c, m, n, n);
  QuantizationParams pa = {scale_a, zp_a}, pb = {scale_b, zp_b};
  MatMat_U8U8F32(A, pa, B, pb, &C);
  if (acc != nullptr) GemmlowpInt32(a, zp_a, b, zp_b, m, n, k, acc);
}

// ---------------------------------------------------------------------------
// int8 AM composition
// ---------------------------------------------------------------------------

namespace {

struct U8Layer {
  int type;
  std::unique_ptr<Layer> layer;     // every non-Linear layer: the reference's class
  Matrix<float> W;                  // Linear: [in x out] as on disk
  Vector<float> b;
  Matrix<uint8_t> W8;
  QuantizationParams qW;
};

struct U8Model {
  std::vector<U8Layer> layers;
  Vector<float> log_prior;
  int left, right;
};

Status ReadU8Layer(util::ReadableFile *fd, U8Layer *out) {
  PK_CHECK_STATUS(fd->ReadAndVerifyString(PK_NNET_LAYER_SECTION));
  int32_t type;
  PK_CHECK_STATUS(fd->ReadValue<int32_t>(&type));
  out->type = type;
  switch (type) {
    case Layer::kLinear:
      PK_CHECK_STATUS(out->W.Read(fd));
      PK_CHECK_STATUS(out->b.Read(fd));
      Quantize(out->W, &out->W8, &out->qW);
      return Status::OK();
    case Layer::kReLU: out->layer.reset(new ReLULayer()); break;
    case Layer::kNormalize: out->layer.reset(new NormalizeLayer()); break;
    case Layer::kSoftmax: out->layer.reset(new SoftmaxLayer()); break;
    case Layer::kSplice: out->layer.reset(new SpliceLayer()); break;
    case Layer::kBatchNorm: out->layer.reset(new BatchNormLayer()); break;
    case Layer::kLogSoftmax: out->layer.reset(new LogSoftmaxLayer()); break;
    case Layer::kNarrow: out->layer.reset(new NarrowLayer()); break;
    default: return Status::Corruption("unexpected layer type");
  }
  return out->layer->Read(fd);
}

}  // namespace

void *ref_u8_open(const char *nnet_path, const char *prior_path, int left,
                  int right) {
  std::unique_ptr<U8Model> model(new U8Model());
  model->left = left;
  model->right = right;
  util::ReadableFile fd;
  if (!fd.Open(nnet_path).ok()) return nullptr;
  if (!fd.ReadAndVerifyString(PK_NNET_SECTION).ok()) return nullptr;
  int32_t l, r, n;
  if (!fd.ReadValue<int32_t>(&l).ok() || !fd.ReadValue<int32_t>(&r).ok() ||
      !fd.ReadValue<int32_t>(&n).ok())
    return nullptr;
  model->layers.resize(n);
  for (int i = 0; i < n; ++i) {
    Status s = ReadU8Layer(&fd, &model->layers[i]);
    if (!s.ok()) {
      fprintf(stderr, "ref_u8_open: %s\n", s.what().c_str());
      return nullptr;
    }
  }
  util::ReadableFile fp;
  if (!fp.Open(prior_path).ok()) return nullptr;
  if (!model->log_prior.Read(&fp).ok()) return nullptr;
  model->log_prior.ApplyLog();  // am.cc:43-44
  return model.release();
}

void ref_u8_close(void *h) { delete static_cast<U8Model *>(h); }

// One utterance as ONE batch: replicate-pad left/right rows as
// am.cc:119-124,152-155, propagate, subtract the log prior (am.cc:109-112).
// `dump_layer` >= 0: also return that Linear layer's (ordinal among Linear
// layers) int32 accumulators [rows x out] in acc_out and its dims; dump_layer == -2 with a trace:
// the accumulators of EVERY Linear layer, concatenated (acc_off[i] = first element of layer i), and
// the activation QuantizationParams each Linear layer's Quantize produced.
struct U8Trace {
  int cap_layers;
  int n_layers;
  float *q_scale;
  int32_t *q_zp;
  long *acc_off;      // [cap_layers + 1]
  int *acc_rows, *acc_cols;
};

static int U8Forward(U8Model *model, const float *feats, int T, int dim, float *out,
                     long cap, int *out_rows, int *out_cols, int dump_layer,
                     int32_t *acc_out, long acc_cap, int *acc_rows, int *acc_cols,
                     U8Trace *trace) {
  int rows = T + model->left + model->right;
  Matrix<float> cur(rows, dim), next;
  for (int r = 0; r < rows; ++r) {
    int src = r - model->left;
    if (src < 0) src = 0;
    if (src > T - 1) src = T - 1;
    memcpy(cur.Row(r).Data(), feats + (size_t)src * dim, sizeof(float) * dim);
  }
  int linear_ordinal = 0;
  long acc_used = 0;
  if (acc_rows) *acc_rows = 0;
  if (acc_cols) *acc_cols = 0;
  for (U8Layer &L : model->layers) {
    if (L.type == Layer::kLinear) {
      Matrix<uint8_t> in8;
      QuantizationParams qa;
      Quantize(cur, &in8, &qa);
      next.Resize(cur.NumRows(), L.W.NumCols());
      MatMat_U8U8F32(in8, qa, L.W8, L.qW, &next);
      if (linear_ordinal == dump_layer && acc_out != nullptr) {
        if ((long)next.NumRows() * next.NumCols() > acc_cap) return -2;
        GemmlowpInt32(in8.Data(), qa.zero_point, L.W8.Data(), L.qW.zero_point,
                      next.NumRows(), next.NumCols(), in8.NumCols(), acc_out);
        *acc_rows = next.NumRows();
        *acc_cols = next.NumCols();
      }
      if (trace != nullptr) {
        if (linear_ordinal >= trace->cap_layers) return -3;
        trace->q_scale[linear_ordinal] = qa.scale;
        trace->q_zp[linear_ordinal] = qa.zero_point;
        trace->acc_off[linear_ordinal] = acc_used;
        trace->acc_rows[linear_ordinal] = next.NumRows();
        trace->acc_cols[linear_ordinal] = next.NumCols();
        if (acc_out != nullptr) {
          const long n = (long)next.NumRows() * next.NumCols();
          if (acc_used + n > acc_cap) return -2;
          GemmlowpInt32(in8.Data(), qa.zero_point, L.W8.Data(), L.qW.zero_point,
                        next.NumRows(), next.NumCols(), in8.NumCols(), acc_out + acc_used);
          acc_used += n;
        }
        trace->acc_off[linear_ordinal + 1] = acc_used;
        trace->n_layers = linear_ordinal + 1;
      }
      for (int r = 0; r < next.NumRows(); ++r) {
        SubVector<float> row = next.Row(r);
        row.AddVec(1.0f, L.b);  // nnet.cc:32-35
      }
      ++linear_ordinal;
    } else {
      L.layer->Propagate(cur, &next);
    }
    cur.Swap(&next);
  }
  for (int r = 0; r < cur.NumRows(); ++r) {
    SubVector<float> row = cur.Row(r);
    row.AddVec(-1.0f, model->log_prior);
  }
  *out_rows = cur.NumRows();
  *out_cols = cur.NumCols();
  if ((long)cur.NumRows() * cur.NumCols() > cap) return -1;
  for (int r = 0; r < cur.NumRows(); ++r)
    memcpy(out + (size_t)r * cur.NumCols(), cur.Row(r).Data(),
           sizeof(float) * cur.NumCols());
  return 0;
}

int ref_u8_forward(void *h, const float *feats, int T, int dim, float *out,
                   long cap, int *out_rows, int *out_cols, int dump_layer,
                   int32_t *acc_out, long acc_cap, int *acc_rows,
                   int *acc_cols) {
  return U8Forward(static_cast<U8Model *>(h), feats, T, dim, out, cap, out_rows, out_cols,
                   dump_layer, acc_out, acc_cap, acc_rows, acc_cols, nullptr);
}

// Every Linear layer in one pass: activation quantisation parameters and (acc_all nullable) the
// int32 accumulators of all of them.  Returns the number of Linear layers (>= 0) or an error (< 0).
int ref_u8_forward_trace(void *h, const float *feats, int T, int dim, float *out, long cap,
                         int *out_rows, int *out_cols, int cap_layers, float *q_scale,
                         int32_t *q_zp, int32_t *acc_all, long acc_cap, long *acc_off,
                         int *acc_rows, int *acc_cols) {
  U8Trace t = {cap_layers, 0, q_scale, q_zp, acc_off, acc_rows, acc_cols};
  int rc = U8Forward(static_cast<U8Model *>(h), feats, T, dim, out, cap, out_rows, out_cols, -2,
                     acc_all, acc_cap, nullptr, nullptr, &t);
  return rc < 0 ? rc : t.n_layers;
}

}  // extern "C"
