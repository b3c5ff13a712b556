// nnet.h -- the acoustic model on the device: packed weights, workspace, forward pass.
#ifndef CE_GPU_NNET_H_
#define CE_GPU_NNET_H_

#include <vector>

#include "common.h"
#include "gemm.h"
#include "model.h"
#include "nnet_kernels.h"

namespace ce {

struct DeviceBlock {
  Block meta;
  int c_pad = 0;             // input channels per tap, padded to the K tile of the data path
  int n_pad = 0;             // length of the per-column parameter arrays (multiple of kTileN)
  int64_t k_total = 0;       // taps * c_pad
  DevBuf w[2];               // packed weights [out_dim x k_total]: (hi, lo) / only
  DevBuf bias, bn_scale, bn_offset, colsum;
  float scale_b = 0.0f;      // Quantize(W_) of the whole [in x out] matrix
  int32_t zp_b = 0;
};

// Device-side parameters of one step of the general program (Program::steps).
struct GenStepDev {
  DevBuf idx;                // Splice indices
  DevBuf scale, offset;      // BatchNorm
};

}  // namespace ce

// The opaque handle of include/ce_gpu.h.
struct ce_gpu_model {
  int device = 0;
  int precision = 0;
  int kind = 0;              // GemmKind
  int n_pass = 1;
  int left = 0, right = 0;
  ce::Program prog;
  std::vector<ce::DeviceBlock> blocks;
  std::vector<ce::GenStepDev> gen;     // general program only: one entry per Program::steps entry
  ce::DevBuf log_prior;      // log(prior), src/am.cc:43-44
  ce::DevBuf zero_prior;     // zeros of the same length
  int fused_output = 1;      // LogSoftmax + prior + argmax in the output layer's epilogue: 1 = int8 models,
                             // 0 = never, 2 = every precision (CE_GPU_FUSED_OUTPUT)
  bool has_cmvn = false;
  std::vector<float> cmvn_host;   // num_mel sums + count
  ce::DevBuf cmvn_dev;
  // Rows of activations evaluated per pass of the layer stack (CE_GPU_CHUNK_ROWS).  Large chunks
  // amortise the per-launch ramp and tail of every kernel; a batch above the cap runs as several
  // chunks, which is also the granularity at which host PCM is copied in behind the compute
  // (device-resident inputs use twice this cap, see ForwardAll).
  int64_t max_chunk_rows = 131072;

  // ---- workspace (one forward call at a time per handle) ----
  ce::DevBuf stage_pcm, stage_feats, stage_argmax_all;
  ce::DevBuf feats;                    // fbank output / staged features [frames x feat_dim]
  ce::Table fbank_chunks;
  // Chunks of utterances alternate between two workspaces on two internal streams, so that the
  // memory-bound kernels of one chunk (CMVN, quantise, log-softmax) overlap the tensor-core
  // GEMMs of the other.
  struct ChunkWs {
    ce::DevBuf feats;                  // this chunk's fbank output [frames x feat_dim]
    ce::Table fbank_chunks;
    ce::DevBuf x0;                     // padded fp32 input [M x feat_dim]
    ce::DevBuf act_f32[2], act_lo[2], act_bf16[2], act_u8, rowsum, logits, row_lse;
    ce::DevBuf minmax, qparams;
    ce::DevBuf stage_loglik;
    ce::Table cmvn_utts, utt_table, tile_table, outrow_table;
    // `stream` (low priority) carries the memory-bound kernels, `stream_hi` (high priority) the
    // GEMMs: whenever a GEMM is ready its CTAs are placed first and the other chunk's
    // quantise / log-softmax blocks fill the registers and threads it leaves free.
    cudaStream_t stream = nullptr, stream_hi = nullptr;
    cudaEvent_t done = nullptr, to_hi = nullptr, to_lo = nullptr;
    void Free();
  };
  ChunkWs ws[2];
  cudaEvent_t inputs_ready = nullptr;
  // CE_GPU_OVERLAP=1 alternates chunks between the two workspaces on two internal streams.  Off by
  // default: the GEMMs and the memory-bound kernels contend for the same L2 bandwidth, so running
  // them side by side was measured to be no faster than back to back (overlap study).
  bool overlap = false;
  // Host PCM is copied chunk by chunk on its own stream so that chunk k+1 arrives while chunk k
  // is being computed.
  cudaStream_t copy_stream = nullptr;
  // Host log-likelihood output: chunk k's rows leave over PCIe on d2h_stream (from one of two
  // staging buffers) while chunk k+1 is computed -- the row ring a CPU decoder is fed from.
  cudaStream_t d2h_stream = nullptr;
  cudaEvent_t ll_ready[2] = {nullptr, nullptr}, ll_copied[2] = {nullptr, nullptr};
  cudaEvent_t call_start = nullptr;
  std::vector<cudaEvent_t> copy_done;
  ce::DevBuf acc_dump;

  // ---- what a log-likelihood row is written as (ce_gpu_model_set_output) ----
  ce::OutSel out_sel;
  ce::DevBuf out_ids;                  // the pdf subset on the device
  // 4-byte words per output row: num_pdfs, the subset size, or 2 k
  int out_words() const {
    return out_sel.mode == ce::kOutDense ? prog.num_pdfs
           : out_sel.mode == ce::kOutSubset ? out_sel.n : 2 * out_sel.n;
  }

  // ---- per-chunk completion callback (ce_gpu_model_set_rows_callback) ----
  void (*rows_cb)(void *, int, int, int64_t, int64_t) = nullptr;
  void *rows_cb_user = nullptr;

  // ---- debug: kept accumulators ----
  int keep_acc = -1;
  std::vector<int32_t> kept_row_off, kept_rows;   // per utterance of the last call
  int kept_lo = 0, kept_hi = 0, kept_cols = 0;
  int64_t kept_ld = 0;
  bool kept_valid = false;
  int last_n_utts = 0;                 // utterances of the last chunk evaluated (ce_gpu_nnet_get_qparams)

  ~ce_gpu_model();
};

namespace ce {

// Loads and packs the model onto the current device.
int ModelBuild(const HostNnet &nn, const std::vector<float> &prior,
               const std::vector<float> *cmvn_stats, int left, int right, int precision,
               int device, ce_gpu_model *m);

// feats_dev: [total_frames x feat_dim] fp32 on the device, utterance u = rows
// [frame_off[u], frame_off[u+1]).  apply_cmvn: run the online CMVN with the model's global
// stats while building the padded network input.  loglik [total_frames x m->out_words()] and
// argmax [total_frames] may each be nullptr, a device pointer, or a host pointer (host outputs
// are copied back chunk by chunk and are complete on return).
// contexted: every block of rows already carries its own left / right context (one ComputeBatch of a
// chunk, src/am.cc:82-113): nothing is replicated, a block of P rows yields P - L - R output rows, packed
// block after block in loglik / argmax.
int NnetForward(ce_gpu_model *m, const float *feats_dev, const int64_t *frame_off, int n_utts,
                bool apply_cmvn, float *loglik, int32_t *argmax, cudaStream_t s, bool contexted = false);

// The whole path from PCM: like NnetForward, but every chunk first runs the fbank kernel on its
// own utterances (sample_off as in ce_gpu_fbank).  `pcm` is a device pointer or a host pointer;
// host PCM is copied in chunk by chunk on the model's copy stream, behind the previous chunk's
// compute.
int PcmForward(ce_gpu_model *m, const int16_t *pcm, int64_t total_samples,
               const int64_t *sample_off, const int64_t *frame_off, int n_utts, float *loglik,
               int32_t *argmax, cudaStream_t s);

}  // namespace ce
#endif  // CE_GPU_NNET_H_
