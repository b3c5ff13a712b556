// nnet.cc -- AcousticModel on the device: weight packing at load, and the batched forward pass.
//
// Replaces AcousticModel::Read (src/am.cc:26-64), ComputeBatch (src/am.cc:82-113) and
// Nnet::Propagate (src/nnet.cc:295-307) for whole utterances evaluated as one batch each
// (chunked and whole-utterance evaluation are the same function).
//
// Row space: utterance u of a chunk owns rows [row_off[u], row_off[u] + P_u) with
// P_u = T_u + left + right and row_off[u] a multiple of 128, in EVERY layer's activation
// matrix (rows are never compacted).  After block b the rows [cum_left_b, P_u - cum_right_b)
// are the reference's rows; the others hold values that no valid row ever reads.
#include "nnet.h"

#include "layers.h"

#include <math.h>
#include <string.h>

#include <algorithm>

void ce_gpu_model::ChunkWs::Free() {
  x0.Free(); feats.Free(); fbank_chunks.Free();
  for (int i = 0; i < 2; ++i) { act_f32[i].Free(); act_lo[i].Free(); act_bf16[i].Free(); }
  act_u8.Free(); rowsum.Free(); logits.Free(); row_lse.Free(); minmax.Free(); qparams.Free();
  stage_loglik.Free();
  cmvn_utts.Free(); utt_table.Free(); tile_table.Free(); outrow_table.Free();
  if (stream) cudaStreamDestroy(stream);
  if (stream_hi) cudaStreamDestroy(stream_hi);
  if (done) cudaEventDestroy(done);
  if (to_hi) cudaEventDestroy(to_hi);
  if (to_lo) cudaEventDestroy(to_lo);
  stream = stream_hi = nullptr;
  done = to_hi = to_lo = nullptr;
}

ce_gpu_model::~ce_gpu_model() {
  cudaSetDevice(device);
  cudaDeviceSynchronize();
  for (auto &b : blocks) {
    b.w[0].Free(); b.w[1].Free(); b.bias.Free(); b.bn_scale.Free(); b.bn_offset.Free(); b.colsum.Free();
  }
  for (auto &g : gen) { g.idx.Free(); g.scale.Free(); g.offset.Free(); }
  log_prior.Free(); zero_prior.Free(); cmvn_dev.Free(); out_ids.Free();
  stage_pcm.Free(); stage_feats.Free(); feats.Free(); fbank_chunks.Free(); acc_dump.Free();
  stage_argmax_all.Free();
  ws[0].Free(); ws[1].Free();
  if (inputs_ready) cudaEventDestroy(inputs_ready);
  if (call_start) cudaEventDestroy(call_start);
  for (cudaEvent_t e : copy_done) cudaEventDestroy(e);
  if (copy_stream) cudaStreamDestroy(copy_stream);
  if (d2h_stream) cudaStreamDestroy(d2h_stream);
  for (int i = 0; i < 2; ++i) {
    if (ll_ready[i]) cudaEventDestroy(ll_ready[i]);
    if (ll_copied[i]) cudaEventDestroy(ll_copied[i]);
  }
}

namespace ce {
namespace {

inline int RoundUp(int v, int m) { return (v + m - 1) / m * m; }

uint16_t Bf16Bits(float f) {                              // round to nearest even
  uint32_t b;
  memcpy(&b, &f, 4);
  if ((b & 0x7f800000u) == 0x7f800000u) return (uint16_t)(b >> 16);
  b += 0x7fffu + ((b >> 16) & 1u);
  return (uint16_t)(b >> 16);
}

float Tf32Round(float f) {                                // cvt.rna.tf32.f32
  uint32_t b;
  memcpy(&b, &f, 4);
  if ((b & 0x7f800000u) == 0x7f800000u) return f;
  b = (b + 0x1000u) & ~0x1fffu;
  float r;
  memcpy(&r, &b, 4);
  return r;
}

int Upload(DevBuf *buf, const void *src, size_t bytes) {
  CE_CHECK(buf->Reserve(bytes));
  CE_CUDA(cudaMemcpy(buf->ptr, src, bytes, cudaMemcpyHostToDevice));
  return CE_GPU_OK;
}

int UploadPadded(DevBuf *buf, const std::vector<float> &v, int n_pad, float fill) {
  std::vector<float> h((size_t)n_pad, fill);
  std::copy(v.begin(), v.end(), h.begin());
  return Upload(buf, h.data(), sizeof(float) * h.size());
}

}  // namespace

int ModelBuild(const HostNnet &nn, const std::vector<float> &prior,
               const std::vector<float> *cmvn_stats, int left, int right, int precision,
               int device, ce_gpu_model *m) {
  m->device = device;
  m->precision = precision;
  m->left = left;
  m->right = right;
  switch (precision) {
    case CE_GPU_PRECISION_INT8: m->kind = kKindI8; m->n_pass = 1; break;
    case CE_GPU_PRECISION_BF16: m->kind = kKindBF16; m->n_pass = 1; break;
    case CE_GPU_PRECISION_FP32: m->kind = kKindTF32; m->n_pass = 3; break;
    case CE_GPU_PRECISION_TF32: m->kind = kKindTF32; m->n_pass = 1; break;
    case CE_GPU_PRECISION_BF16X3: m->kind = kKindBF16X3; m->n_pass = 1; break;
    default:
      SetError("unknown precision %d", precision);
      return CE_GPU_EINVAL;
  }
  CE_CHECK(CompileProgram(nn, left, right, &m->prog, (int)prior.size()));
  if ((int)prior.size() != m->prog.num_pdfs) {
    SetError("prior has %zu entries, the nnet has %d outputs", prior.size(), m->prog.num_pdfs);
    return CE_GPU_EINVAL;
  }
  if (cmvn_stats) {
    if ((int)cmvn_stats->size() != m->prog.feat_dim + 1) {
      SetError("cmvn stats have %zu entries, expected feat_dim + 1 = %d", cmvn_stats->size(),
               m->prog.feat_dim + 1);
      return CE_GPU_EINVAL;
    }
    m->has_cmvn = true;
    m->cmvn_host = *cmvn_stats;
    CE_CHECK(Upload(&m->cmvn_dev, cmvn_stats->data(), sizeof(float) * cmvn_stats->size()));
  }
  {  // log_prior_.ApplyLog(), src/am.cc:43-44 (logf of every entry, src/vector.cc:167-174)
    std::vector<float> lp(prior.size());
    for (size_t i = 0; i < prior.size(); ++i) lp[i] = logf(prior[i]);
    CE_CHECK(Upload(&m->log_prior, lp.data(), sizeof(float) * lp.size()));
    std::fill(lp.begin(), lp.end(), 0.0f);               // rows that are already final (fused output layer)
    CE_CHECK(Upload(&m->zero_prior, lp.data(), sizeof(float) * lp.size()));
  }

  const int tile_k = KindTileK(m->kind);
  m->blocks.resize(m->prog.blocks.size());
  for (size_t bi = 0; bi < m->prog.blocks.size(); ++bi) {
    const Block &B = m->prog.blocks[bi];
    DeviceBlock &D = m->blocks[bi];
    const HostLayer &L = nn.layers[B.linear];
    D.meta = B;
    D.c_pad = RoundUp(B.in_dim, tile_k);
    const int n_taps = (int)B.taps.size();
    D.k_total = (int64_t)n_taps * KindPhysCols(m->kind, D.c_pad);
    const int N = B.out_dim, C = B.in_dim, K = n_taps * C;
    // per-column arrays cover every column the epilogue may touch (the next block's K padding)
    int next_pad = N;
    if (bi + 1 < m->prog.blocks.size()) next_pad = RoundUp(m->prog.blocks[bi + 1].in_dim, tile_k);
    D.n_pad = RoundUp(std::max(N, next_pad), kTileN);
    CE_CHECK(UploadPadded(&D.bias, L.b, D.n_pad, 0.0f));
    if (B.batchnorm >= 0) {
      CE_CHECK(UploadPadded(&D.bn_scale, nn.layers[B.batchnorm].scale, D.n_pad, 0.0f));
      CE_CHECK(UploadPadded(&D.bn_offset, nn.layers[B.batchnorm].offset, D.n_pad, 0.0f));
    }
    const size_t elems = (size_t)N * (size_t)D.k_total;
    if (m->kind == kKindI8) {
      std::vector<uint8_t> w8((size_t)K * N);
      QuantizeHost(L.W.data(), (int64_t)K * N, w8.data(), &D.scale_b, &D.zp_b);
      std::vector<uint8_t> packed(elems, 0);
      std::vector<int32_t> colsum((size_t)D.n_pad, 0);
      for (int t = 0; t < n_taps; ++t)
        for (int c = 0; c < C; ++c) {
          const uint8_t *src = w8.data() + (size_t)(t * C + c) * N;
          const size_t kk = (size_t)t * D.c_pad + c;
          for (int n = 0; n < N; ++n) {
            packed[(size_t)n * D.k_total + kk] = src[n];
            colsum[n] += src[n];
          }
        }
      CE_CHECK(Upload(&D.w[0], packed.data(), packed.size()));
      CE_CHECK(Upload(&D.colsum, colsum.data(), sizeof(int32_t) * colsum.size()));
    } else if (m->kind == kKindBF16) {
      std::vector<uint16_t> packed(elems, 0);
      for (int t = 0; t < n_taps; ++t)
        for (int c = 0; c < C; ++c) {
          const float *src = L.W.data() + (size_t)(t * C + c) * N;
          const size_t kk = (size_t)t * D.c_pad + c;
          for (int n = 0; n < N; ++n) packed[(size_t)n * D.k_total + kk] = Bf16Bits(src[n]);
        }
      CE_CHECK(Upload(&D.w[0], packed.data(), packed.size() * 2));
    } else if (m->kind == kKindBF16X3) {                 // per 32 channels: [32 hi | 32 lo]
      std::vector<uint16_t> packed(elems, 0);
      for (int t = 0; t < n_taps; ++t)
        for (int c = 0; c < C; ++c) {
          const float *src = L.W.data() + (size_t)(t * C + c) * N;
          const size_t kk = (size_t)t * 2 * D.c_pad + (size_t)(c >> 5) * 64 + (c & 31);
          for (int n = 0; n < N; ++n) {
            const uint16_t hb = Bf16Bits(src[n]);
            uint32_t hw = (uint32_t)hb << 16;
            float hf;
            memcpy(&hf, &hw, 4);
            packed[(size_t)n * D.k_total + kk] = hb;
            packed[(size_t)n * D.k_total + kk + 32] = Bf16Bits(src[n] - hf);
          }
        }
      CE_CHECK(Upload(&D.w[0], packed.data(), packed.size() * 2));
    } else {
      std::vector<float> hi(elems, 0.0f), lo;
      if (m->n_pass == 3) lo.assign(elems, 0.0f);
      for (int t = 0; t < n_taps; ++t)
        for (int c = 0; c < C; ++c) {
          const float *src = L.W.data() + (size_t)(t * C + c) * N;
          const size_t kk = (size_t)t * D.c_pad + c;
          for (int n = 0; n < N; ++n) {
            const float h = Tf32Round(src[n]);
            hi[(size_t)n * D.k_total + kk] = h;
            if (m->n_pass == 3) lo[(size_t)n * D.k_total + kk] = src[n] - h;
          }
        }
      CE_CHECK(Upload(&D.w[0], hi.data(), hi.size() * 4));
      if (m->n_pass == 3) CE_CHECK(Upload(&D.w[1], lo.data(), lo.size() * 4));
    }
  }
  m->gen.resize(m->prog.steps.size());
  for (size_t i = 0; i < m->prog.steps.size(); ++i) {
    const Step &st = m->prog.steps[i];
    const HostLayer &L = nn.layers[st.layer];
    if (st.type == kSplice) CE_CHECK(Upload(&m->gen[i].idx, L.indices.data(), sizeof(int32_t) * L.indices.size()));
    if (st.type == kBatchNorm) {
      CE_CHECK(Upload(&m->gen[i].scale, L.scale.data(), sizeof(float) * L.scale.size()));
      CE_CHECK(Upload(&m->gen[i].offset, L.offset.data(), sizeof(float) * L.offset.size()));
    }
  }
  if (const char *e = getenv("CE_GPU_CHUNK_ROWS")) {
    long v = atol(e);
    if (v >= kTileM) m->max_chunk_rows = v;
  }
  if (const char *e = getenv("CE_GPU_OVERLAP")) m->overlap = atoi(e) != 0;
  if (const char *e = getenv("CE_GPU_FUSED_OUTPUT")) m->fused_output = atoi(e);   // 0: A/B against the separate kernel
  int prio_lo = 0, prio_hi = 0;
  CE_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  for (int i = 0; i < 2; ++i) {
    CE_CUDA(cudaStreamCreateWithPriority(&m->ws[i].stream, cudaStreamNonBlocking, prio_lo));
    CE_CUDA(cudaStreamCreateWithPriority(&m->ws[i].stream_hi, cudaStreamNonBlocking, prio_hi));
    CE_CUDA(cudaEventCreateWithFlags(&m->ws[i].done, cudaEventDisableTiming));
    CE_CUDA(cudaEventCreateWithFlags(&m->ws[i].to_hi, cudaEventDisableTiming));
    CE_CUDA(cudaEventCreateWithFlags(&m->ws[i].to_lo, cudaEventDisableTiming));
  }
  CE_CUDA(cudaEventCreateWithFlags(&m->inputs_ready, cudaEventDisableTiming));
  CE_CUDA(cudaEventCreateWithFlags(&m->call_start, cudaEventDisableTiming));
  CE_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  CE_CUDA(cudaStreamCreateWithFlags(&m->d2h_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CE_CUDA(cudaEventCreateWithFlags(&m->ll_ready[i], cudaEventDisableTiming));
    CE_CUDA(cudaEventCreateWithFlags(&m->ll_copied[i], cudaEventDisableTiming));
  }
  return CE_GPU_OK;
}

namespace {

RowUse MakeRowUse(const ce_gpu_model *m, int producer /* -1 = network input */) {
  RowUse u;
  memset(&u, 0, sizeof(u));
  if (producer >= 0) {
    u.lo = m->blocks[producer].meta.cum_left;
    u.hi = m->blocks[producer].meta.cum_right;
  }
  const Block &nx = m->blocks[producer + 1].meta;
  u.next_n_taps = (int)nx.taps.size();
  for (int t = 0; t < u.next_n_taps; ++t) u.next_tap_off[t] = nx.taps[t];
  u.next_lo = nx.cum_left;
  u.next_hi = nx.cum_right;
  return u;
}

// Output rows are addressed by absolute frame index (frame_off), so loglik_dev / argmax_dev are
// the bases of the whole batch's outputs.
struct PcmSource {
  const int16_t *pcm_dev = nullptr;     // nullptr: features are given
  const int16_t *pcm_host = nullptr;    // set: pcm_dev is a staging buffer filled chunk by chunk
  int64_t total_samples = 0;
  const int64_t *sample_off = nullptr;  // [n_utts + 1] of this chunk
};

// The row space of one chunk (see the header comment) and its tables on the device.
struct RowSpace {
  int M = 0;                            // rows, a multiple of kTileM
  bool gran = false;                    // blocks at multiples of kRowGran instead of kTileM
  std::vector<int64_t> row_off64;       // first row of every utterance block
  const UttRows *d_utts = nullptr;
  const int32_t *d_tile = nullptr;
  const UttRows *h_utts = nullptr;
};

// frame_off: input rows of every utterance; out_off: where its output rows go.  `contexted`: the input
// rows of a block already carry their left / right context (ComputeBatch of a chunk, src/am.cc:82-113),
// so nothing is replicated and a block of P rows yields P - L - R output rows.
int BuildRowSpace(ce_gpu_model *m, ce_gpu_model::ChunkWs *w, const int64_t *frame_off, const int64_t *out_off,
                  bool contexted, int n_utts, cudaStream_t s, RowSpace *rs) {
  const int L = contexted ? 0 : m->left, R = contexted ? 0 : m->right;
  // ---- row space ----
  // Every utterance block starts at a multiple of kTileM rows (one utterance per GEMM tile).  When
  // the blocks are short -- micro-batches of live streams, a few frames plus context each -- that
  // padding is most of the work, so they are packed at multiples of kRowGran (32) rows instead and
  // the int8 GEMM runs its granule-mode epilogue (per-warp instead of per-tile parameters).
  int64_t m_tile = 0, m_gran = 0;
  for (int u = 0; u < n_utts; ++u) {
    const int64_t T = frame_off[u + 1] - frame_off[u];
    if (T <= 0) continue;
    m_tile += (T + L + R + kTileM - 1) / kTileM * kTileM;
    m_gran += (T + L + R + kRowGran - 1) / kRowGran * kRowGran;
  }
  const char *force_env = getenv("CE_GPU_ROW_GRAN");     // 32 / 128 force one layout (tests)
  const int force_gran = force_env ? atoi(force_env) : 0;
  const bool gran = force_gran == kRowGran || (force_gran == 0 && m_gran * 5 <= m_tile * 4);   // saves >= 20 %
  const int align = gran ? kRowGran : kTileM;
  std::vector<int64_t> row_off64(n_utts);
  int64_t M64 = 0;
  for (int u = 0; u < n_utts; ++u) {
    const int64_t T = frame_off[u + 1] - frame_off[u];
    row_off64[u] = M64;
    if (T > 0) M64 += (T + L + R + align - 1) / align * align;
  }
  const int64_t m_used = M64;
  M64 = (M64 + kTileM - 1) / kTileM * kTileM;            // whole GEMM tiles; the tail belongs to nobody
  if (M64 == 0) return CE_GPU_OK;                      // rs->M stays 0
  if (M64 > 0x7fffff00LL) {
    SetError("a chunk of %lld rows exceeds the 2^31 row limit", (long long)M64);
    return CE_GPU_EINVAL;
  }
  const int M = (int)M64;
  const int m_tiles = M / kRowGran;                       // entries of the row -> utterance table
  CE_CHECK(w->utt_table.Acquire(sizeof(UttRows) * n_utts));
  CE_CHECK(w->tile_table.Acquire(sizeof(int32_t) * m_tiles));
  CE_CHECK(w->outrow_table.Acquire(sizeof(int64_t) * n_utts));
  UttRows *hu = w->utt_table.host<UttRows>();
  int32_t *ht = w->tile_table.host<int32_t>();
  int64_t *ho = w->outrow_table.host<int64_t>();
  for (int u = 0; u < n_utts; ++u) {
    const int64_t T = frame_off[u + 1] - frame_off[u];
    hu[u].row_off = (int32_t)row_off64[u];
    hu[u].rows = T > 0 ? (int32_t)(T + L + R) : 0;
    ho[u] = out_off[u];
    const int64_t end = (u + 1 < n_utts) ? row_off64[u + 1] : m_used;
    for (int64_t t = row_off64[u] / kRowGran; t < end / kRowGran; ++t) ht[t] = u;
  }
  // rows past the last block: the last utterance's, beyond its `rows` (padding like any other)
  for (int64_t t = m_used / kRowGran; t < m_tiles; ++t) ht[t] = n_utts - 1;
  CE_CHECK(w->utt_table.Upload(sizeof(UttRows) * n_utts, s));
  CE_CHECK(w->tile_table.Upload(sizeof(int32_t) * m_tiles, s));
  CE_CHECK(w->outrow_table.Upload(sizeof(int64_t) * n_utts, s));
  const UttRows *d_utts = w->utt_table.dev<UttRows>();
  const int32_t *d_tile = w->tile_table.dev<int32_t>();

  rs->M = M;
  rs->gran = gran;
  rs->row_off64.swap(row_off64);
  rs->d_utts = d_utts;
  rs->d_tile = d_tile;
  rs->h_utts = hu;
  return CE_GPU_OK;
}

// fbank of the chunk's utterances (when the input is PCM) + replicate padding (+ CMVN) into x0
// (src/am.cc:119-124,152-155).  For the u8 path the same kernel reduces each utterance's min/max (the
// padding rows are copies, so the frames' min/max is the matrix's) for the first Quantize.
int ChunkInput(ce_gpu_model *m, ce_gpu_model::ChunkWs *w, const PcmSource &src, const float *feats_dev,
               const int64_t *frame_off, int n_utts, bool apply_cmvn, const RowSpace &rs, uint32_t *mm_first,
               cudaStream_t s, bool contexted) {
  const int L = contexted ? 0 : m->left, R = contexted ? 0 : m->right, F = m->prog.feat_dim;
  std::vector<int64_t> local_off;
  const int64_t *feat_off = frame_off;
  if (src.pcm_dev) {
    local_off.resize(n_utts + 1);
    for (int u = 0; u <= n_utts; ++u) local_off[u] = frame_off[u] - frame_off[0];
    CE_CHECK(w->feats.Reserve(sizeof(float) * (size_t)local_off[n_utts] * F));
    CE_CHECK(FbankLaunch(src.pcm_dev, src.total_samples, src.sample_off, local_off.data(), n_utts, F,
                         w->feats.as<float>(), F, &w->fbank_chunks, s));
    feats_dev = w->feats.as<float>();
    feat_off = local_off.data();
  }
  return CmvnLaunch(apply_cmvn ? m->cmvn_dev.as<float>() : nullptr, apply_cmvn ? m->cmvn_host[F] : 0.0f,
                    feats_dev, feat_off, rs.row_off64.data(), n_utts, F, L, R, w->x0.as<float>(), F,
                    &w->cmvn_utts, s, nullptr, mm_first);
}

// The general layer program (Program::general): any layer list of src/nnet.h:21-30, one kernel per
// layer over fp32 activations in the padded row space.  Linear layers run on the tensor cores in the
// model's precision -- for int8 that is Quantize(in) + MatMat_U8U8F32 + AddVec(b) per Linear layer,
// the composition of SURVEY D3 -- everything else as the row-wise kernels of layers.cu.
int ForwardChunkGeneral(ce_gpu_model *m, ce_gpu_model::ChunkWs *w, const PcmSource &src,
                        const float *feats_dev, const int64_t *frame_off, const int64_t *out_off,
                        bool contexted, int n_utts, bool apply_cmvn, float *loglik_dev, int32_t *argmax_dev,
                        cudaStream_t s) {
  const int L = m->left, R = m->right, F = m->prog.feat_dim, NP = m->prog.num_pdfs;
  const int nb = (int)m->blocks.size();
  RowSpace rs;
  CE_CHECK(BuildRowSpace(m, w, frame_off, out_off, contexted, n_utts, s, &rs));
  if (rs.M == 0) return CE_GPU_OK;
  const int M = rs.M;
  const int wmax = RoundUp(std::max(m->prog.max_dim, 4), 4);
  int cmax = 4;
  for (const DeviceBlock &D : m->blocks) cmax = std::max(cmax, D.c_pad);
  CE_CHECK(w->x0.Reserve(sizeof(float) * (size_t)M * F));
  for (int i = 0; i < 2; ++i) CE_CHECK(w->act_f32[i].Reserve(sizeof(float) * (size_t)M * wmax));
  if (nb > 0) {
    if (m->kind == kKindI8) {
      CE_CHECK(w->act_u8.Reserve((size_t)M * cmax));
      CE_CHECK(w->rowsum.Reserve(sizeof(int32_t) * (size_t)M));
      CE_CHECK(w->minmax.Reserve(sizeof(uint32_t) * 2 * (size_t)nb * n_utts));
      CE_CHECK(w->qparams.Reserve(sizeof(QParam) * (size_t)nb * n_utts));
      CE_CHECK(InitMinMaxLaunch(w->minmax.as<uint32_t>(), nb * n_utts, s));
    } else if (m->kind == kKindBF16 || m->kind == kKindBF16X3) {
      CE_CHECK(w->act_bf16[0].Reserve(2 * (size_t)M * KindPhysCols(m->kind, cmax)));
    } else {
      CE_CHECK(w->act_lo[0].Reserve(sizeof(float) * (size_t)M * cmax));   // operand hi
      if (m->n_pass == 3) CE_CHECK(w->act_lo[1].Reserve(sizeof(float) * (size_t)M * cmax));
    }
  }
  CE_CHECK(ChunkInput(m, w, src, feats_dev, frame_off, n_utts, apply_cmvn, rs, nullptr, s, contexted));

  m->kept_valid = false;
  m->last_n_utts = n_utts;
  float *cur = w->x0.as<float>();
  int64_t ld = F;
  int dim = F, which = 0;                                // act_f32[which] is the next free buffer
  for (const Step &st : m->prog.steps) {
    const ce::GenStepDev &G = m->gen[&st - m->prog.steps.data()];
    switch (st.type) {
      case kNarrow:                                      // only the valid range shrinks (st.lo / st.hi)
        break;
      case kSplice: {
        float *out = w->act_f32[which].as<float>();
        CE_CHECK(SpliceLaunch(cur, ld, dim, M, rs.d_tile, rs.d_utts, st.lo, st.hi, G.idx.as<int32_t>(),
                              st.out_dim / std::max(1, st.in_dim), out, RoundUp(st.out_dim, 4), s));
        cur = out;
        ld = RoundUp(st.out_dim, 4);
        which ^= 1;
        break;
      }
      case kReLU:
        CE_CHECK(RowwiseLaunch(kRowReLU, cur, ld, dim, M, rs.d_tile, rs.d_utts, st.lo, st.hi, nullptr, nullptr, s));
        break;
      case kBatchNorm:
        CE_CHECK(RowwiseLaunch(kRowBatchNorm, cur, ld, dim, M, rs.d_tile, rs.d_utts, st.lo, st.hi,
                               G.scale.as<float>(), G.offset.as<float>(), s));
        break;
      case kNormalize:
        CE_CHECK(RowwiseLaunch(kRowNormalize, cur, ld, dim, M, rs.d_tile, rs.d_utts, st.lo, st.hi, nullptr, nullptr, s));
        break;
      case kSoftmax:
        CE_CHECK(RowwiseLaunch(kRowSoftmax, cur, ld, dim, M, rs.d_tile, rs.d_utts, st.lo, st.hi, nullptr, nullptr, s));
        break;
      case kLogSoftmax:
        CE_CHECK(RowwiseLaunch(kRowLogSoftmax, cur, ld, dim, M, rs.d_tile, rs.d_utts, st.lo, st.hi, nullptr, nullptr, s));
        break;
      case kLinear: {
        const int b = st.block;
        const DeviceBlock &D = m->blocks[b];
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        GemmOperands ops;
        memset(&ops, 0, sizeof(ops));
        a.M = M;
        a.N = D.meta.out_dim;
        a.c_pad = KindPhysCols(m->kind, D.c_pad);
        a.n_taps = 1;
        a.n_pass = m->n_pass;
        if (m->n_pass == 3) {
          a.pass_a[0] = 1; a.pass_b[0] = 0;
          a.pass_a[1] = 0; a.pass_b[1] = 1;
          a.pass_a[2] = 0; a.pass_b[2] = 0;
        }
        a.bias = D.bias.as<float>();
        a.tile_utt = rs.d_tile;
        a.gran = rs.gran ? 1 : 0;
        a.utts = rs.d_utts;
        ops.rows_a = M;
        ops.rows_b = D.meta.out_dim;
        ops.k_total = D.k_total;
        ops.b[0] = D.w[0].ptr;
        ops.b[1] = D.w[1].ptr;
        if (m->kind == kKindI8) {
          uint32_t *mm = w->minmax.as<uint32_t>() + 2 * (size_t)b * n_utts;
          QParam *qp = w->qparams.as<QParam>() + (size_t)b * n_utts;
          RowUse use;
          memset(&use, 0, sizeof(use));
          use.lo = st.lo;
          use.hi = st.hi;
          CE_CHECK(MinMaxLaunch(cur, ld, dim, M, rs.d_tile, rs.d_utts, use, mm, s));
          CE_CHECK(QuantizeLaunch(cur, ld, dim, M, D.c_pad, rs.d_tile, mm, n_utts, qp, w->act_u8.as<uint8_t>(),
                                  w->rowsum.as<int32_t>(), s));
          ops.a[0] = w->act_u8.ptr;
          a.a_rowsum = w->rowsum.as<int32_t>();
          a.b_colsum = D.colsum.as<int32_t>();
          a.zp_b = D.zp_b;
          a.scale_b = D.scale_b;
          a.k_true = D.meta.in_dim;
          a.qa = qp;
        } else if (m->kind == kKindBF16) {
          CE_CHECK(ConvertLaunch(cur, ld, dim, M, D.c_pad, w->act_bf16[0].as<__nv_bfloat16>(), nullptr, nullptr, s));
          ops.a[0] = w->act_bf16[0].ptr;
        } else if (m->kind == kKindBF16X3) {
          CE_CHECK(ConvertLaunch(cur, ld, dim, M, D.c_pad, nullptr, nullptr, nullptr, s,
                                 w->act_bf16[0].as<__nv_bfloat16>()));
          ops.a[0] = w->act_bf16[0].ptr;
        } else {
          CE_CHECK(ConvertLaunch(cur, ld, dim, M, D.c_pad, nullptr, w->act_lo[0].as<float>(),
                                 m->n_pass == 3 ? w->act_lo[1].as<float>() : nullptr, s));
          ops.a[0] = w->act_lo[0].ptr;
          ops.a[1] = m->n_pass == 3 ? w->act_lo[1].ptr : nullptr;
        }
        float *out = w->act_f32[which].as<float>();
        a.out_f32 = out;
        a.ld_out = RoundUp(a.N, 4);
        a.n_store = a.N;
        if (m->kind == kKindI8 && m->keep_acc == b) {
          CE_CHECK(m->acc_dump.Reserve(sizeof(int32_t) * (size_t)M * a.ld_out));
          a.out_acc = m->acc_dump.as<int32_t>();
          m->kept_row_off.resize(n_utts);
          m->kept_rows.resize(n_utts);
          for (int u = 0; u < n_utts; ++u) {
            m->kept_row_off[u] = rs.h_utts[u].row_off;
            m->kept_rows[u] = rs.h_utts[u].rows;
          }
          m->kept_lo = st.lo;
          m->kept_hi = st.hi;
          m->kept_cols = a.N;
          m->kept_ld = a.ld_out;
          m->kept_valid = true;
        }
        CE_CHECK(GemmLaunch(m->kind, ops, a, s));
        cur = out;
        ld = a.ld_out;
        which ^= 1;
        break;
      }
      default:
        SetError("general program: unexpected layer type %d", st.type);
        return CE_GPU_EUNSUPPORTED;
    }
    dim = st.out_dim;
  }
  return FinalizeLaunch(cur, ld, NP, M, rs.d_tile, rs.d_utts, w->outrow_table.dev<int64_t>(), L, R,
                        m->prog.log_softmax, m->log_prior.as<float>(), loglik_dev, m->out_words(), argmax_dev,
                        s, m->out_sel);
}

// s: stream of the memory-bound kernels; s_gemm: stream of the GEMMs (== s when chunks are not
// overlapped).
int ForwardChunk(ce_gpu_model *m, ce_gpu_model::ChunkWs *w, const PcmSource &src,
                 const float *feats_dev, const int64_t *frame_off, const int64_t *out_off, bool contexted,
                 int n_utts, bool apply_cmvn, float *loglik_dev, int32_t *argmax_dev, cudaStream_t s,
                 cudaStream_t s_gemm) {
  const int L = m->left, R = m->right, F = m->prog.feat_dim, NP = m->prog.num_pdfs;
  const int nb = (int)m->blocks.size();

  if (m->prog.general)
    return ForwardChunkGeneral(m, w, src, feats_dev, frame_off, out_off, contexted, n_utts, apply_cmvn, loglik_dev,
                               argmax_dev, s);
  HostMark(nullptr);
  RowSpace rs;
  CE_CHECK(BuildRowSpace(m, w, frame_off, out_off, contexted, n_utts, s, &rs));
  HostMark("chunk: row space");
  if (rs.M == 0) return CE_GPU_OK;
  const int M = rs.M;
  const bool gran = rs.gran;
  const UttRows *d_utts = rs.d_utts;
  const int32_t *d_tile = rs.d_tile;
  const UttRows *hu = rs.h_utts;

  // ---- workspace ----
  int wmax = 4;
  for (const DeviceBlock &D : m->blocks) wmax = std::max(wmax, std::max(D.c_pad, RoundUp(D.meta.out_dim, 4)));
  const int ldp = RoundUp(NP, 4);
  CE_CHECK(w->x0.Reserve(sizeof(float) * (size_t)M * F));
  CE_CHECK(w->logits.Reserve(sizeof(float) * (size_t)M * ldp));
  CE_CHECK(w->row_lse.Reserve(sizeof(float) * (size_t)M));
  if (m->kind == kKindI8) {
    CE_CHECK(w->act_f32[0].Reserve(sizeof(float) * (size_t)M * wmax));
    CE_CHECK(w->act_u8.Reserve((size_t)M * wmax));
    CE_CHECK(w->rowsum.Reserve(sizeof(int32_t) * (size_t)M));
    CE_CHECK(w->minmax.Reserve(sizeof(uint32_t) * 2 * (size_t)nb * n_utts));
    CE_CHECK(w->qparams.Reserve(sizeof(QParam) * (size_t)nb * n_utts));
  } else if (m->kind == kKindBF16 || m->kind == kKindBF16X3) {
    for (int i = 0; i < 2; ++i)
      CE_CHECK(w->act_bf16[i].Reserve(2 * (size_t)M * KindPhysCols(m->kind, wmax)));
  } else {
    for (int i = 0; i < 2; ++i) {
      CE_CHECK(w->act_f32[i].Reserve(sizeof(float) * (size_t)M * wmax));
      if (m->n_pass == 3) CE_CHECK(w->act_lo[i].Reserve(sizeof(float) * (size_t)M * wmax));
    }
  }

  // ---- fbank (when the input is PCM), replicate padding (+ CMVN) into x0, first min/max ----
  uint32_t *mm = w->minmax.as<uint32_t>();
  QParam *qp = w->qparams.as<QParam>();
  HostMark("chunk: reserve");
  if (m->kind == kKindI8) CE_CHECK(InitMinMaxLaunch(mm, nb * n_utts, s));
  HostMark("chunk: init minmax");
  CE_CHECK(ChunkInput(m, w, src, feats_dev, frame_off, n_utts, apply_cmvn, rs, m->kind == kKindI8 ? mm : nullptr, s,
                      contexted));

  HostMark("chunk: workspace + input");
  // ---- network input in the operand format of the data path ----
  const int c0 = m->blocks[0].c_pad;
  if (m->kind == kKindI8) {
    CE_CHECK(QuantizeLaunch(w->x0.as<float>(), F, F, M, c0, d_tile, mm, n_utts, qp,
                            w->act_u8.as<uint8_t>(), w->rowsum.as<int32_t>(), s));
  } else if (m->kind == kKindBF16) {
    CE_CHECK(ConvertLaunch(w->x0.as<float>(), F, F, M, c0, w->act_bf16[0].as<__nv_bfloat16>(),
                           nullptr, nullptr, s));
  } else if (m->kind == kKindBF16X3) {
    CE_CHECK(ConvertLaunch(w->x0.as<float>(), F, F, M, c0, nullptr, nullptr, nullptr, s,
                           w->act_bf16[0].as<__nv_bfloat16>()));
  } else {
    CE_CHECK(ConvertLaunch(w->x0.as<float>(), F, F, M, c0, nullptr, w->act_f32[0].as<float>(),
                           m->n_pass == 3 ? w->act_lo[0].as<float>() : nullptr, s));
  }

  m->kept_valid = false;
  m->last_n_utts = n_utts;
  // Fused output layer (GemmArgs::lsm): every row's arithmetic is the same whatever the output mode, so
  // dense rows, selected rows, batches and micro-batches stay bit-identical to each other.
  const bool lsm_dense = m->out_sel.mode == kOutDense;
  const DeviceBlock &Dl = m->blocks[nb - 1];
  // int8 only: its output layer is epilogue-bound either way.  The float kinds' is bound by the multiplications,
  // and multiplying twice costs them more than the log-softmax kernel did (fp32 = 3 x TF32: 64 k -> 50 k x real
  // time, bf16x3 125 k -> 107 k, measured) -- they keep the separate kernel (CE_GPU_FUSED_OUTPUT=2 forces it on).
  const bool lsm = (m->kind == kKindI8 ? m->fused_output != 0 : m->fused_output == 2) && NP % 4 == 0 &&
                   (lsm_dense || m->prog.log_softmax) &&   // (selection without a LogSoftmax layer: nothing to fuse)
                   !Dl.meta.relu && Dl.meta.batchnorm < 0 && m->keep_acc != nb - 1 &&
                   (!lsm_dense || loglik_dev == nullptr || (reinterpret_cast<uintptr_t>(loglik_dev) & 15) == 0);
  HostMark("chunk: first quantize");
  for (int b = 0; b < nb; ++b) {
    const DeviceBlock &D = m->blocks[b];
    const bool last = (b == nb - 1);
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    GemmOperands ops;
    memset(&ops, 0, sizeof(ops));
    a.M = M;
    a.N = D.meta.out_dim;
    a.c_pad = KindPhysCols(m->kind, D.c_pad);
    a.n_taps = (int)D.meta.taps.size();
    for (int t = 0; t < a.n_taps; ++t) a.tap_off[t] = D.meta.taps[t];
    a.n_pass = m->n_pass;
    if (m->n_pass == 3) {                                // small terms first: lo*hi, hi*lo, hi*hi
      a.pass_a[0] = 1; a.pass_b[0] = 0;
      a.pass_a[1] = 0; a.pass_b[1] = 1;
      a.pass_a[2] = 0; a.pass_b[2] = 0;
    }
    a.bias = D.bias.as<float>();
    a.relu = D.meta.relu ? 1 : 0;
    if (D.meta.batchnorm >= 0) {
      a.bn_scale = D.bn_scale.as<float>();
      a.bn_offset = D.bn_offset.as<float>();
    }
    a.tile_utt = d_tile;
    a.gran = gran ? 1 : 0;
    a.utts = d_utts;
    ops.rows_a = M;
    ops.rows_b = D.meta.out_dim;
    ops.k_total = D.k_total;
    ops.b[0] = D.w[0].ptr;
    ops.b[1] = D.w[1].ptr;
    const int next_c = last ? 0 : m->blocks[b + 1].c_pad;

    if (m->kind == kKindI8) {
      ops.a[0] = w->act_u8.ptr;
      a.a_rowsum = w->rowsum.as<int32_t>();
      a.b_colsum = D.colsum.as<int32_t>();
      a.zp_b = D.zp_b;
      a.scale_b = D.scale_b;
      a.k_true = a.n_taps * D.meta.in_dim;
      a.qa = qp + (size_t)b * n_utts;
      a.out_f32 = last ? w->logits.as<float>() : w->act_f32[0].as<float>();
      a.ld_out = last ? ldp : RoundUp(a.N, 4);
      a.n_store = a.N;
      if (!last) {
        a.minmax = mm + 2 * (size_t)(b + 1) * n_utts;
        const RowUse ru = MakeRowUse(m, b);
        a.mm_lo = ru.lo; a.mm_hi = ru.hi;
        a.next_n_taps = ru.next_n_taps;
        memcpy(a.next_tap_off, ru.next_tap_off, sizeof(a.next_tap_off));
        a.next_lo = ru.next_lo; a.next_hi = ru.next_hi;
      }
      if (m->keep_acc == b) {
        CE_CHECK(m->acc_dump.Reserve(sizeof(int32_t) * (size_t)M * a.ld_out));
        a.out_acc = m->acc_dump.as<int32_t>();
        m->kept_row_off.resize(n_utts);
        m->kept_rows.resize(n_utts);
        for (int u = 0; u < n_utts; ++u) {
          m->kept_row_off[u] = hu[u].row_off;
          m->kept_rows[u] = hu[u].rows;
        }
        m->kept_lo = D.meta.cum_left;
        m->kept_hi = D.meta.cum_right;
        m->kept_cols = a.N;
        m->kept_ld = a.ld_out;
        m->kept_valid = true;
      }
    } else if (m->kind == kKindBF16 || m->kind == kKindBF16X3) {
      ops.a[0] = w->act_bf16[b & 1].ptr;
      if (last) {
        a.out_f32 = w->logits.as<float>();
        a.ld_out = ldp;
        a.n_store = a.N;
      } else {
        a.out_bf16 = w->act_bf16[(b + 1) & 1].as<__nv_bfloat16>();
        a.ld_out = KindPhysCols(m->kind, next_c);        // stored elements per row
        a.n_store = next_c;                              // logical columns
      }
    } else {
      ops.a[0] = w->act_f32[b & 1].ptr;
      ops.a[1] = m->n_pass == 3 ? w->act_lo[b & 1].ptr : nullptr;
      if (last) {
        a.out_f32 = w->logits.as<float>();
        a.ld_out = ldp;
        a.n_store = a.N;
      } else {
        a.out_f32 = w->act_f32[(b + 1) & 1].as<float>();
        a.out_lo = m->n_pass == 3 ? w->act_lo[(b + 1) & 1].as<float>() : nullptr;
        a.round_tf32 = 1;
        a.ld_out = next_c;
        a.n_store = next_c;
      }
    }
    if (last && lsm) {
      // the output layer writes the finished rows itself: LogSoftmax, prior and argmax in its epilogue
      a.lsm = 1;
      a.lsm_softmax = m->prog.log_softmax ? 1 : 0;
      a.lsm_left = L;
      a.lsm_right = R;
      a.lsm_prior = m->log_prior.as<float>();
      a.out_lo = nullptr;
      a.round_tf32 = 0;
      if (lsm_dense) {
        a.out_f32 = loglik_dev;                          // may be nullptr: argmax only
        a.ld_out = NP;
        a.lsm_out_row_off = w->outrow_table.dev<int64_t>();     // absolute frame index over the batch
        a.lsm_out_rows = out_off[n_utts];
        a.lsm_argmax = argmax_dev;
      } else {
        // selecting outputs: ONE sweep writes the plain logits in row space plus every row's log-sum-exp
        // -- reduced exactly as the dense fused output reduces it, so that selected and dense rows agree
        // bit for bit -- and the selecting kernel below subtracts it
        a.out_f32 = w->logits.as<float>();
        a.ld_out = ldp;
        a.lsm_rowspace = 1;
        a.lsm_out_rows = M;
        a.lsm_single = 1;
        a.lsm_lse_out = w->row_lse.as<float>();
      }
    }
    if (s_gemm != s) {
      CE_CUDA(cudaEventRecord(w->to_hi, s));
      CE_CUDA(cudaStreamWaitEvent(s_gemm, w->to_hi, 0));
    }
    HostMark("chunk: gemm args");
    CE_CHECK(GemmLaunch(m->kind, ops, a, s_gemm));
    HostMark("chunk: GemmLaunch");
    if (s_gemm != s && (m->kind == kKindI8 || last)) {   // float paths chain GEMM -> GEMM
      CE_CUDA(cudaEventRecord(w->to_lo, s_gemm));
      CE_CUDA(cudaStreamWaitEvent(s, w->to_lo, 0));
    }

    if (m->kind == kKindI8 && !last) {
      QParam *q_next = qp + (size_t)(b + 1) * n_utts;
      CE_CHECK(QuantizeLaunch(w->act_f32[0].as<float>(), a.ld_out, a.N, M, next_c, d_tile,
                              mm + 2 * (size_t)(b + 1) * n_utts, n_utts, q_next,
                              w->act_u8.as<uint8_t>(), w->rowsum.as<int32_t>(), s));
    }
  }

  HostMark("chunk: quantize launches");
  if (lsm && lsm_dense) return CE_GPU_OK;
  CE_CHECK(FinalizeLaunch(w->logits.as<float>(), ldp, NP, M, d_tile, d_utts,
                          w->outrow_table.dev<int64_t>(), L, R, m->prog.log_softmax, m->log_prior.as<float>(),
                          loglik_dev, m->out_words(), argmax_dev, s, m->out_sel,
                          lsm ? w->row_lse.as<float>() : nullptr));
  HostMark("chunk: finalize");
  return CE_GPU_OK;
}

}  // namespace

namespace {
// One chunk's "rows ready" notification, run by the CUDA runtime's host-function thread.
struct RowsReady {
  void (*fn)(void *, int, int, int64_t, int64_t);
  void *user;
  int first_utt, n_utts;
  int64_t first_frame, n_frames;
};
void CUDART_CB RowsReadyTrampoline(void *p) {
  RowsReady *r = static_cast<RowsReady *>(p);
  r->fn(r->user, r->first_utt, r->n_utts, r->first_frame, r->n_frames);
  delete r;
}

int ForwardAllImpl(ce_gpu_model *m, const PcmSource &all, const float *feats_dev, const int64_t *frame_off,
                   int n_utts, bool apply_cmvn, float *loglik, int32_t *argmax, cudaStream_t s, bool contexted);

// A call that fails half way (e.g. the workspace cannot grow for a later chunk) may already have queued
// copies into the caller's host buffers and rows-ready callbacks for earlier chunks: nothing of that
// may still be in flight when the error is returned, or the caller would free buffers under it.
int ForwardAll(ce_gpu_model *m, const PcmSource &all, const float *feats_dev, const int64_t *frame_off,
               int n_utts, bool apply_cmvn, float *loglik, int32_t *argmax, cudaStream_t s,
               bool contexted = false) {
  const int rc = ForwardAllImpl(m, all, feats_dev, frame_off, n_utts, apply_cmvn, loglik, argmax, s, contexted);
  if (rc != CE_GPU_OK) {
    const std::string msg = LastError();                 // keep the first error's message
    cudaStreamSynchronize(s);
    cudaStreamSynchronize(m->copy_stream);
    cudaStreamSynchronize(m->d2h_stream);
    for (int i = 0; i < 2; ++i) {
      if (m->ws[i].stream) cudaStreamSynchronize(m->ws[i].stream);
      if (m->ws[i].stream_hi) cudaStreamSynchronize(m->ws[i].stream_hi);
    }
    cudaGetLastError();
    SetError("%s", msg.c_str());
  }
  return rc;
}

int ForwardAllImpl(ce_gpu_model *m, const PcmSource &all, const float *feats_dev, const int64_t *frame_off,
                   int n_utts, bool apply_cmvn, float *loglik, int32_t *argmax, cudaStream_t s, bool contexted) {
  if (apply_cmvn && !m->has_cmvn) {
    SetError("the model was loaded without CMVN statistics");
    return CE_GPU_EINVAL;
  }
  const int L = m->left, R = m->right;
  // output rows: one per input frame, or (contexted blocks) one per frame whose context is in the block
  std::vector<int64_t> out_off_v;
  const int64_t *out_off = frame_off;
  if (contexted) {
    out_off_v.assign(n_utts + 1, 0);
    for (int u = 0; u < n_utts; ++u)
      out_off_v[u + 1] = out_off_v[u] + std::max<int64_t>(0, frame_off[u + 1] - frame_off[u] - L - R);
    out_off = out_off_v.data();
  }
  const int pad = contexted ? 0 : L + R;
  const int W = m->out_words();                          // 4-byte words per output row
  const bool ll_host = loglik && !IsDevicePtr(loglik);
  const bool am_host = argmax && !IsDevicePtr(argmax);
  const bool overlap = m->overlap && m->keep_acc < 0;
  if (overlap) CE_CUDA(cudaEventRecord(m->inputs_ready, s));
  if (all.pcm_host) {                                    // the staging buffer is free once the work
    CE_CUDA(cudaEventRecord(m->call_start, s));          // already queued on s has run
    CE_CUDA(cudaStreamWaitEvent(m->copy_stream, m->call_start, 0));
  }
  // The per-frame argmax is 4 bytes a frame: it is staged for the whole batch and copied out once,
  // so that no device-to-host copy sits between two chunks on the compute stream.
  const int64_t total_frames = out_off[n_utts] - out_off[0];
  if (am_host) CE_CHECK(m->stage_argmax_all.Reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(total_frames, 1)));
  bool used[2] = {false, false};
  bool ll_pending[2] = {false, false};
  int u0 = 0, chunk = 0;
  while (u0 < n_utts) {
    int u1 = u0;
    int64_t rows = 0;
    // Host PCM arrives over PCIe behind the compute: a short first chunk (a quarter of the cap)
    // keeps the copy nobody can hide short; every later copy runs under the previous chunk.
    // Inputs that are already in HBM have no copy to hide, so they run in chunks twice as large
    // (fewer launch ramps and tails: 15.4 -> 15.0 ms per step on the bench batch; for host input
    // larger later chunks were measured to bring nothing, and chunks growing 1/4, 1, 4 x the cap -- 32 +
    // 128 + 352 utterances -- measured WORSE, 330 k against 345-351 k x end to end: the 113 MB copy of the
    // last chunk does not fit under the 3.3 ms the chunk before it computes).
    const int64_t cap = !all.pcm_host ? 2 * m->max_chunk_rows
                        : chunk == 0  ? std::max<int64_t>(kTileM, m->max_chunk_rows / 4)
                                      : m->max_chunk_rows;
    while (u1 < n_utts) {
      const int64_t T = frame_off[u1 + 1] - frame_off[u1];
      const int64_t r = T > 0 ? (T + pad + kTileM - 1) / kTileM * kTileM : 0;
      if (u1 > u0 && rows + r > cap) break;
      rows += r;
      ++u1;
    }
    ce_gpu_model::ChunkWs *w = &m->ws[overlap ? (chunk & 1) : 0];
    cudaStream_t cs = overlap ? w->stream : s;
    if (overlap && !used[chunk & 1]) {
      CE_CUDA(cudaStreamWaitEvent(cs, m->inputs_ready, 0));
      used[chunk & 1] = true;
    }
    const int64_t f0 = out_off[u0], nf = out_off[u1] - f0;
    float *ll_dev = loglik;
    int32_t *am_dev = argmax;
    ce::DevBuf *ll_stage = &m->ws[chunk & 1].stage_loglik;   // two staging buffers, alternating
    if (ll_host) {
      if (ll_pending[chunk & 1]) CE_CUDA(cudaStreamWaitEvent(cs, m->ll_copied[chunk & 1], 0));   // buffer free again
      CE_CHECK(ll_stage->Reserve(sizeof(float) * (size_t)nf * W));
      ll_dev = ll_stage->as<float>() - f0 * W;           // row f0 lands on the staging buffer's row 0
    }
    if (am_host) am_dev = m->stage_argmax_all.as<int32_t>() - out_off[0];   // whole batch, one copy at the end
    PcmSource src = all;
    if (all.pcm_dev) src.sample_off = all.sample_off + u0;
    if (all.pcm_host && all.sample_off[u1] > all.sample_off[u0]) {
      if ((int)m->copy_done.size() <= chunk) {
        m->copy_done.resize(chunk + 1, nullptr);
      }
      if (!m->copy_done[chunk]) CE_CUDA(cudaEventCreateWithFlags(&m->copy_done[chunk], cudaEventDisableTiming));
      const int64_t a = all.sample_off[u0], b = all.sample_off[u1];
      CE_CUDA(cudaMemcpyAsync(const_cast<int16_t *>(all.pcm_dev) + a, all.pcm_host + a,
                              sizeof(int16_t) * (size_t)(b - a), cudaMemcpyHostToDevice, m->copy_stream));
      CE_CUDA(cudaEventRecord(m->copy_done[chunk], m->copy_stream));
      CE_CUDA(cudaStreamWaitEvent(cs, m->copy_done[chunk], 0));
    }
    CE_CHECK(ForwardChunk(m, w, src, feats_dev, frame_off + u0, out_off + u0, contexted, u1 - u0, apply_cmvn, ll_dev, am_dev, cs,
                          overlap ? w->stream_hi : cs));
    if ((ll_host || m->rows_cb) && nf > 0) {               // off the compute stream: the next chunk starts now
      CE_CUDA(cudaEventRecord(m->ll_ready[chunk & 1], cs));
      CE_CUDA(cudaStreamWaitEvent(m->d2h_stream, m->ll_ready[chunk & 1], 0));
      if (ll_host)
        CE_CUDA(cudaMemcpyAsync(loglik + f0 * W, ll_stage->ptr, sizeof(float) * (size_t)nf * W,
                                cudaMemcpyDeviceToHost, m->d2h_stream));
      if (m->rows_cb) {                                    // the consumer may start on this chunk now
        if (am_host)
          CE_CUDA(cudaMemcpyAsync(argmax + f0, m->stage_argmax_all.as<int32_t>() + (f0 - out_off[0]),
                                  sizeof(int32_t) * (size_t)nf, cudaMemcpyDeviceToHost, m->d2h_stream));
        RowsReady *note = new RowsReady{m->rows_cb, m->rows_cb_user, u0, u1 - u0, f0, nf};
        cudaError_t e = cudaLaunchHostFunc(m->d2h_stream, RowsReadyTrampoline, note);
        if (e != cudaSuccess) {
          delete note;
          CE_CUDA(e);
        }
      }
      CE_CUDA(cudaEventRecord(m->ll_copied[chunk & 1], m->d2h_stream));
      ll_pending[chunk & 1] = true;
    }
    u0 = u1;
    ++chunk;
  }
  if (overlap) {
    for (int i = 0; i < 2; ++i) {
      if (!used[i]) continue;
      CE_CUDA(cudaEventRecord(m->ws[i].done, m->ws[i].stream));
      CE_CUDA(cudaStreamWaitEvent(s, m->ws[i].done, 0));
    }
  }
  for (int i = 0; i < 2; ++i)
    if (ll_pending[i]) CE_CUDA(cudaStreamWaitEvent(s, m->ll_copied[i], 0));
  if (am_host && total_frames > 0 && !m->rows_cb) {      // (with a callback it went out chunk by chunk)
    CE_CUDA(cudaMemcpyAsync(argmax + out_off[0], m->stage_argmax_all.ptr, sizeof(int32_t) * (size_t)total_frames,
                            cudaMemcpyDeviceToHost, s));
  }
  // host outputs are complete, and every callback has returned, on return
  if (ll_host || am_host || m->rows_cb) CE_CUDA(cudaStreamSynchronize(s));
  return CE_GPU_OK;
}
}  // namespace

int NnetForward(ce_gpu_model *m, const float *feats_dev, const int64_t *frame_off, int n_utts,
                bool apply_cmvn, float *loglik, int32_t *argmax, cudaStream_t s, bool contexted) {
  return ForwardAll(m, PcmSource(), feats_dev, frame_off, n_utts, apply_cmvn, loglik, argmax, s, contexted);
}

int PcmForward(ce_gpu_model *m, const int16_t *pcm, int64_t total_samples,
               const int64_t *sample_off, const int64_t *frame_off, int n_utts, float *loglik,
               int32_t *argmax, cudaStream_t s) {
  PcmSource src;
  if (IsDevicePtr(pcm)) {
    src.pcm_dev = pcm;
  } else {
    CE_CHECK(m->stage_pcm.Reserve(sizeof(int16_t) * (size_t)total_samples));
    src.pcm_dev = m->stage_pcm.as<int16_t>();
    src.pcm_host = pcm;
  }
  src.total_samples = total_samples;
  src.sample_off = sample_off;
  return ForwardAll(m, src, nullptr, frame_off, n_utts, m->has_cmvn, loglik, argmax, s);
}

}  // namespace ce
