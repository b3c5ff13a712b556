/* ce_gpu.h -- C ABI of libce_gpu.so: the B200 (sm_100a) replacement for the
 * CatEars acoustic front-end + acoustic-model forward pass.
 *
 * What this boundary replaces in the reference (ishine/CatEars):
 *
 *   ce_gpu_fbank      WaveReader::Process   src/pcm_reader.cc:148-190  (int16 -> float, unscaled)
 *                     Fbank::Process        src/fbank.cc:265-314       (all frames of an utterance)
 *   ce_gpu_cmvn       CMVN::GetFrame        src/cmvn.cc:100-110        (frames 0..T-1 in order)
 *   ce_gpu_nnet       AcousticModel::Process / EndOfStream  src/am.cc:115-164, i.e.
 *                     ComputeBatch src/am.cc:82-113 = replicate padding + Nnet::Propagate
 *                     (src/nnet.cc:295-307) + "row -= log_prior"
 *   ce_gpu_forward    the body of ce_stt_process/ce_stt_end_of_stream between WaveReader and
 *                     Decoder::Process      src/ce_stt.cc:307-331,349-357
 *   ce_gpu_model_load AcousticModel::Read   src/am.cc:26-64 (+ Nnet::Read src/nnet.cc:273-293)
 *   ce_gpu_quantize   Quantize              src/matrix.cc:366-387
 *   ce_gpu_gemm_u8    MatMat_U8U8F32        src/matrix.cc:389-420
 *   ce_gpu_last_error ce_stt_last_error     src/ce_stt.cc:375-377 (thread-local here)
 *   ce_gpu_streams_*  the per-utterance state of Fbank::Instance (src/fbank.cc:275-313), CMVN
 *                     (src/cmvn.cc:35-68) and AcousticModel::Instance (src/am.cc:115-142), kept in
 *                     device buffers for many live utterances at once
 *   ce_gpu_model_set_output / ce_gpu_model_set_rows_callback
 *                     what Decoder::Process reads of a row (src/decoder.cc:97-102) and when it
 *                     may start reading (src/ce_stt.cc:349-357): narrower rows, per-chunk hand-over
 *
 * Conventions
 *   - Plain pointers and sizes only.  Every DATA pointer (pcm, feats, loglik, argmax, A, B, C...)
 *     may be a device pointer on the handle's device or a host pointer (pinned or pageable);
 *     host buffers are staged through the handle's own pinned/device workspace inside the call.
 *     OFFSET arrays (utt_sample_offsets, utt_frame_offsets) are always HOST memory.
 *   - Batches are ragged: utterance u owns samples [off[u], off[u+1]) of `pcm` and rows
 *     [foff[u], foff[u+1]) of every per-frame output.  T_u = off<400 ? 0 : 1+(n_u-400)/160
 *     (snip-edges, src/fbank.cc:35-42).  Utterances with T_u == 0 produce no rows.
 *   - `stream` is a cudaStream_t (NULL = the legacy default stream).  Calls return after the
 *     work is enqueued, except that host OUTPUT buffers are complete on return.
 *   - Return value: 0 on success, a negative CE_GPU_E* code otherwise; the message is in
 *     ce_gpu_last_error().  There is no CPU fallback: without a CUDA device every compute entry
 *     point fails with CE_GPU_ENODEVICE.
 *   - One handle per device; calls on one handle must not overlap in time (they share its
 *     workspace); distinct handles are independent.
 */
#ifndef CE_GPU_H_
#define CE_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CE_GPU_OK 0
#define CE_GPU_EINVAL (-1)     /* bad argument                                  */
#define CE_GPU_EIO (-2)        /* file missing / corrupt (Status::IOError/Corruption) */
#define CE_GPU_ECUDA (-3)      /* CUDA runtime / driver error                   */
#define CE_GPU_ENODEVICE (-4)  /* no usable sm_100 device                       */
#define CE_GPU_ENOMEM (-5)
#define CE_GPU_EUNSUPPORTED (-6) /* layer stack the GPU program cannot express  */

/* Arithmetic of the Linear layers. */
#define CE_GPU_PRECISION_INT8 0  /* u8 x u8 -> s32, gemmlowp-compatible (bit-exact accumulators) */
#define CE_GPU_PRECISION_BF16 1  /* bf16 x bf16 -> fp32 (fast mode, OUTSIDE the 1e-3 log-likelihood bar) */
#define CE_GPU_PRECISION_FP32 2  /* 3xTF32 error-compensated, fp32-class accuracy */
#define CE_GPU_PRECISION_TF32 3  /* single-pass TF32 (fast mode, OUTSIDE the 1e-3 log-likelihood bar) */
#define CE_GPU_PRECISION_BF16X3 4 /* bf16 hi/lo operands, 3 products: 16-bit mantissa at the bf16 tensor
                                     rate; meets the 1e-3 bar (the fast in-tolerance float path)    */

typedef struct ce_gpu_model ce_gpu_model_t;

/* Thread-local message of the last failing call on this thread ("" if none). */
const char *ce_gpu_last_error(void);

/* Number of CUDA devices visible (0 if none / no driver). Never fails. */
int ce_gpu_device_count(void);

/* Library/ABI version (for the loader check in tests): 10000*major + 100*minor + patch. */
int ce_gpu_version(void);

/* ---- model ---------------------------------------------------------------- */

/* Loads an NN02 nnet + VEC0 prior (+ optional VEC0 CMVN global stats, NULL = no CMVN, which is
 * what src/ce_stt.cc actually does) onto `device`, packing/quantising the weights once.
 * left/right context are the AM's (config-file) values, src/am.cc:48-50. */
ce_gpu_model_t *ce_gpu_model_load(const char *nnet_path, const char *prior_path,
                                  const char *cmvn_stats_path, int left_context,
                                  int right_context, int precision, int device);

/* Same, from the reference's "key = value" config file (keys nnet, prior, left_context,
 * right_context, num_pdfs [, cmvn_stats]), src/am.cc:26-64, src/configuration.cc:14-88. */
ce_gpu_model_t *ce_gpu_model_load_config(const char *config_path, int precision, int device);

void ce_gpu_model_free(ce_gpu_model_t *m);

/* Any out pointer may be NULL. */
int ce_gpu_model_info(const ce_gpu_model_t *m, int *num_pdfs, int *left_context,
                      int *right_context, int *feat_dim, int *precision, int *device);

/* ---- stage-level entry points (also what the parity tests call) ----------- */

/* Frame bookkeeping only (host): fills utt_frame_offsets[n_utts+1]; returns total frames
 * (>= 0) or a negative error. */
int64_t ce_gpu_frame_offsets(const int64_t *utt_sample_offsets, int n_utts,
                             int64_t *utt_frame_offsets);

/* Log-mel filterbank of int16 PCM: feats[total_frames x num_mel] fp32, packed by utterance.
 * num_mel 40 is the reference's PK_FBANK_DIM; other sizes (<= 128) are an extension. */
int ce_gpu_fbank(const int16_t *pcm, const int64_t *utt_sample_offsets, int n_utts,
                 int num_mel, float *feats, int device, void *stream);

/* Online mean-only CMVN with the reference's 600-frame window / 200-frame global smoothing.
 * global_stats: num_mel sums then the count (HOST pointer, num_mel+1 floats).  In place
 * (out == feats) is allowed. */
int ce_gpu_cmvn(const float *global_stats, const float *feats,
                const int64_t *utt_frame_offsets, int n_utts, int num_mel, float *out,
                int device, void *stream);

/* The same CMVN continued across calls (streaming; the reference's CMVN::GetFrame is causal and is
 * driven frame by frame, src/cmvn.cc:35-68).  For utterance u, rows [off[u], off[u+1]) of `feats` are
 * n_hist[u] RAW frames that earlier calls already normalised -- exactly the last
 * min(t_base[u], 600) of them, oldest first; they are only read as x_{t-600} -- followed by the new
 * raw frames.  t_base[u] = frames normalised before this call.  state[u * num_mel ...] (HOST, in and
 * out) holds the running sums after t_base[u] frames: all zeros for a new utterance.  `out` receives
 * only the NEW frames, packed by utterance.  Bit-identical to one ce_gpu_cmvn call on the whole
 * utterance, for any split into calls. */
int ce_gpu_cmvn_stream(const float *global_stats, const float *feats,
                       const int64_t *utt_frame_offsets, const int32_t *n_hist,
                       const int64_t *t_base, float *state, int n_utts, int num_mel, float *out,
                       int device, void *stream);

/* 512-point forward real FFT of n_frames rows of 512 floats, packed output
 * [Re0, Re256, Re1, Im1, ...] as SRFFT::Compute (src/srfft.cc:370).  Test hook for the FFT
 * inside ce_gpu_fbank (the same device code). */
int ce_gpu_rfft512(const float *in, int n_frames, float *out, int device, void *stream);

/* Acoustic model on ready features [total_frames x feat_dim]: per utterance replicate-pad,
 * propagate, subtract log prior.  loglik[total_frames x num_pdfs] (may be NULL if only argmax
 * is wanted; see ce_gpu_model_set_output for narrower rows); argmax[total_frames] int32 (may be
 * NULL). */
int ce_gpu_nnet(ce_gpu_model_t *m, const float *feats, const int64_t *utt_frame_offsets,
                int n_utts, float *loglik, int32_t *argmax, void *stream);

/* AcousticModel::ComputeBatch (src/am.cc:82-113) for many chunks at once: block b = rows
 * [block_offsets[b], block_offsets[b+1]) of `feats` is one chunk of frames that ALREADY carries its own
 * context -- left_context rows in front of and right_context rows behind the frames it is evaluated
 * for, the way AcousticModel::Process / EndOfStream stack their buffer (src/am.cc:115-164).  Nothing
 * is replicated; a block of P rows yields P - left_context - right_context output rows (none if
 * P <= left + right), packed block after block in `loglik` / `argmax`.  For int8 models each block is
 * one Quantize matrix per layer, exactly like one ComputeBatch call of the reference. */
int ce_gpu_nnet_chunks(ce_gpu_model_t *m, const float *feats, const int64_t *block_offsets, int n_blocks,
                       float *loglik, int32_t *argmax, void *stream);

/* The whole path: PCM -> fbank -> [CMVN if the model has stats] -> AM.
 * utt_frame_offsets_out (HOST, n_utts+1) may be NULL. */
int ce_gpu_forward(ce_gpu_model_t *m, const int16_t *pcm, const int64_t *utt_sample_offsets,
                   int n_utts, float *loglik, int32_t *argmax,
                   int64_t *utt_frame_offsets_out, void *stream);

/* ---- decoder feed: fewer bytes per frame (hazard H6) ----------------------
 * A dense row is 4 num_pdfs bytes a frame (12 KB at 3072 pdfs), which caps a host decoder at what
 * PCIe carries, while Decoder::Process only reads frame_logp(tid2pdf[ilabel]) of its active arcs
 * (src/decoder.cc:97-102).  ce_gpu_model_set_output selects, for all later ce_gpu_nnet /
 * ce_gpu_forward calls on the handle, what a row of `loglik` is:
 *   CE_GPU_OUTPUT_DENSE   float[num_pdfs]                                   (the default)
 *   CE_GPU_OUTPUT_SUBSET  float[n]: column j = log-likelihood of pdf pdf_ids[j] (HOST array, any
 *                         order, repeats allowed).  Exact: give the decoder the matching remapped
 *                         tid2pdf and it computes what it computed from the dense row.
 *   CE_GPU_OUTPUT_TOPK    ce_gpu_scored_pdf_t[n]: the n largest log-likelihoods of the frame in
 *                         descending order, equal values in ascending pdf order (pdf_ids unused);
 *                         1 <= n <= min(num_pdfs, 1024).  What the host assumes for the pdfs not
 *                         listed is its policy; entry n-1 bounds them from above.
 * The values are those of the dense row, bit for bit; `argmax` is unaffected (always over all pdfs).
 * Selected outputs need num_pdfs % 4 == 0 and num_pdfs <= 4096 (CE_GPU_EUNSUPPORTED otherwise).
 * Not to be called while a forward call on the handle is in flight. */
#define CE_GPU_OUTPUT_DENSE 0
#define CE_GPU_OUTPUT_SUBSET 1
#define CE_GPU_OUTPUT_TOPK 2
typedef struct ce_gpu_scored_pdf {
  float loglik;
  int32_t pdf;
} ce_gpu_scored_pdf_t;
int ce_gpu_model_set_output(ce_gpu_model_t *m, int mode, const int32_t *pdf_ids, int n);
/* 4-byte words per row of `loglik` under the current selection: num_pdfs, n, or 2 n. */
int ce_gpu_model_output_width(const ce_gpu_model_t *m);

/* Row ring: a batch above the chunk size (CE_GPU_CHUNK_ROWS activation rows,
 * 128 ten-second utterances by default) is evaluated chunk by chunk, and with a HOST `loglik`
 * buffer chunk c's rows leave over PCIe while chunk c+1 is computed.  With a callback set, `fn` is
 * called once per chunk, in order, as soon as the rows [first_frame, first_frame + n_frames) of
 * `loglik` (and of a host `argmax`) of utterances [first_utt, first_utt + n_utts) are complete in
 * the caller's buffers (utterance indices and frame numbers as in the call's offset arrays) -- so a
 * CPU decoder can consume chunk c while the GPU works on c+1.  It runs on a thread owned by the
 * CUDA runtime: it must not call CUDA or ce_gpu_* functions and should hand the chunk to a worker
 * and return (the GPU runs at most two chunks ahead of a callback that has not returned -- the
 * ring has two slots); the forward call returns only after the last callback has returned.  With
 * device output buffers the callback only reports that the chunk's kernels have finished.
 * fn == NULL removes it. */
typedef void (*ce_gpu_rows_ready_fn)(void *user, int first_utt, int n_utts, int64_t first_frame,
                                     int64_t n_frames);
int ce_gpu_model_set_rows_callback(ce_gpu_model_t *m, ce_gpu_rows_ready_fn fn, void *user);

/* ---- live utterances with their state on the device -----------------------
 * The streaming form of ce_gpu_forward: what the reference keeps per utterance in Fbank::Instance
 * (samples that do not fill a frame yet, src/fbank.cc:308-313), CMVN (running sums + the last 600
 * raw frames, src/cmvn.cc:35-68) and AcousticModel::Instance (frames waiting for their right
 * context, src/am.cc:115-142) lives in per-slot device buffers of a stream set.  One process call
 * takes whatever PCM has arrived for any subset of the open slots and runs ONE fbank, ONE CMVN
 * (if the model has statistics) and ONE acoustic-model pass for all of them; only the new samples
 * go up, only the finished rows come down.  Rows come out as soon as their right context exists
 * and equal the whole-utterance rows (the CMVN chain continues exactly).
 *   create   a set of max_streams slots on model m (m must outlive the set)
 *   open     a free slot for a new utterance: its id (>= 0) or a negative error
 *   process  slots[i]: distinct open slots; pcm[i]: HOST pointer to n_samples[i] new samples (may
 *            be 0 / NULL); end_of_stream[i] != 0 (array may be NULL): no more audio -- the right
 *            context is replicated, the remaining rows come out and the slot is free again.
 *            rows: HOST or DEVICE buffer of rows_cap rows of ce_gpu_model_output_width(m) words;
 *            rows [row_offsets[i], row_offsets[i+1]) belong to slots[i] (row_offsets: HOST, n+1).
 *            If more than rows_cap rows are ready the call fails before changing anything.
 *   rows_ready  how many rows that process call would produce (from the counters alone).
 * Calls on one set (and on its model) must not overlap in time and are ordered by `stream`. */
typedef struct ce_gpu_streams ce_gpu_streams_t;
ce_gpu_streams_t *ce_gpu_streams_create(ce_gpu_model_t *m, int max_streams);
void ce_gpu_streams_free(ce_gpu_streams_t *s);
int ce_gpu_streams_open(ce_gpu_streams_t *s);
int64_t ce_gpu_streams_rows_ready(const ce_gpu_streams_t *s, const int *slots, int n,
                                  const int *n_samples, const unsigned char *end_of_stream);
int ce_gpu_streams_process(ce_gpu_streams_t *s, const int *slots, int n, const int16_t *const *pcm,
                           const int *n_samples, const unsigned char *end_of_stream, float *rows,
                           int64_t rows_cap, int64_t *row_offsets, void *stream);

/* Host time spent inside ce_gpu_streams_process since the set was created (or the last reset), in
 * microseconds: until all work of a call was queued (`enqueue_us`: planning, the PCM upload, ~40 kernel
 * launches) and until the call returned (`total_us`: + waiting for the rows to arrive in a host buffer).
 * Any out pointer may be NULL. */
int ce_gpu_streams_call_stats(ce_gpu_streams_t *s, int64_t *calls, double *enqueue_us, double *total_us,
                              int reset);

/* Debug/parity hook (int8 models): after the next ce_gpu_nnet/ce_gpu_forward call the int32
 * accumulators of the `linear_ordinal`-th Linear layer are kept; fetch them with
 * ce_gpu_nnet_get_acc.  Pass -1 to disable. */
int ce_gpu_nnet_keep_acc(ce_gpu_model_t *m, int linear_ordinal);
/* Copies rows of utterance `utt` of the kept accumulators: acc[rows x cols] (HOST). */
int ce_gpu_nnet_get_acc(ce_gpu_model_t *m, int utt, int32_t *acc, int64_t cap, int *rows,
                        int *cols);

/* Debug/parity hook (int8 models): the activation (scale, zero_point) that the Quantize in front of
 * every Linear layer (src/matrix.cc:348-387) produced for utterance `utt` of the last forward call
 * (which must have fitted one chunk).  scale/zero_point: HOST arrays of `cap` entries.  Returns the
 * number of Linear layers or a negative error. */
int ce_gpu_nnet_get_qparams(ce_gpu_model_t *m, int utt, float *scale, int32_t *zero_point, int cap);

/* ---- matrix-level entry points --------------------------------------------- */

/* Quantize (src/matrix.cc:366-387) of a contiguous [rows x cols] fp32 matrix to u8 with one
 * (scale, zero_point) for the whole matrix; scale/zero_point are HOST out-pointers. */
int ce_gpu_quantize(const float *src, int64_t rows, int cols, uint8_t *dst, float *scale,
                    int32_t *zero_point, int device, void *stream);

/* Device self-test of the quantiser's arithmetic: the production kernel divides by the scale with
 * a correctly rounded reciprocal and two fused corrections and rounds by truncating
 * q + 0.49999997f; this compares it with `roundf(min(max(v / scale + zp, 0), 255))` in plain IEEE
 * operations over n pseudo-random triples concentrated on the rounding ties.  Returns the number
 * of disagreements (0 = bit-exact) or a negative error. */
int64_t ce_gpu_selftest_quantizer(int64_t n, uint64_t seed, int device);

/* MatMat_U8U8F32 (src/matrix.cc:389-420): A[m x k] u8, B[k x n] u8, both row-major;
 * C[m x n] = (float)acc * (scale_a*scale_b), acc = sum_k (A-zp_a)(B-zp_b) in int32.
 * acc (nullable) receives the int32 accumulators. */
int ce_gpu_gemm_u8(const uint8_t *a, float scale_a, int32_t zp_a, const uint8_t *b,
                   float scale_b, int32_t zp_b, int m, int n, int k, float *c, int32_t *acc,
                   int device, void *stream);

/* C[m x n] = A[m x k] * B[k x n], fp32 in/out, `precision` one of BF16/FP32/TF32/BF16X3 (tensor-core
 * replacement for MatMat / cblas_sgemm, src/matrix.cc:300-323). */
int ce_gpu_gemm_f32(const float *a, const float *b, int m, int n, int k, float *c,
                    int precision, int device, void *stream);

/* ---- pinned host buffers --------------------------------------------------------------------
 * Page-locked host memory (the row ring a CPU decoder reads, PCM staging): device <-> host copies
 * to and from it run at the full PCIe rate and overlap compute.  For callers that do not link the
 * CUDA runtime themselves (the ce_stt shim).  NULL (+ ce_gpu_last_error) on failure. */
void *ce_gpu_host_alloc(size_t bytes);
void ce_gpu_host_free(void *p);

/* ---- multi-GPU planning (host only; utterances are independent, so there is no collective) ---- */

/* Splits n_utts utterances (frame offsets as above) into n_parts CONTIGUOUS groups of nearly
 * equal frame counts, one per GPU: part p owns utterances [part_begin[p], part_begin[p+1]).
 * part_begin has n_parts+1 entries.  Deterministic: every rank computes the same plan. */
int ce_gpu_partition(const int64_t *utt_frame_offsets, int n_utts, int n_parts, int32_t *part_begin);

/* Long-form audio (one stream of total_frames frames) as n_parts contiguous time shards with
 * recomputed halos: shard p keeps output frames [keep_begin[p], keep_end[p])
 * and must be fed frames [feed_begin[p], feed_end[p]) = the kept range extended by
 * left_halo = left_context + cmvn_history frames before and right_context frames after (clipped
 * to the stream).  Sample range of a frame range [a, b): [160 a, 160 (b-1) + 400). */
int ce_gpu_time_shards(int64_t total_frames, int n_parts, int left_context, int right_context,
                       int cmvn_history, int64_t *keep_begin, int64_t *keep_end,
                       int64_t *feed_begin, int64_t *feed_end);

/* ---- instrumentation -------------------------------------------------------- */

/* Number of kernels this library has launched on this thread since the last reset. */
int64_t ce_gpu_launch_count(int reset);

/* Per-category device timing of the kernels launched by this thread (CUDA events on the
 * launching stream; off by default).  Categories: 0 fbank, 1 cmvn(+pad), 2 GEMM, 3 min/max +
 * quantise, 4 log-softmax/prior/argmax (int8 models: the output layer's GEMM fused with them), 5 other.  ce_gpu_profile_read waits for the recorded
 * launches, returns the summed milliseconds and launch counts (arrays of 6) and clears them. */
#define CE_GPU_PROFILE_CATEGORIES 6
int ce_gpu_profile_enable(int on);
int ce_gpu_profile_read(double *ms, int64_t *launches);
/* Instead of the sums: one record per timed launch scope, in launch order -- category and
 * begin/end in milliseconds relative to the first record (events of different streams share one
 * clock, so this is a timeline of the overlapped chunk streams).  Returns the number of records
 * written (<= cap) or a negative error; clears the records like ce_gpu_profile_read. */
int ce_gpu_profile_trace(int cap, int32_t *cat, double *t0_ms, double *t1_ms);

#ifdef __cplusplus
}
#endif
#endif /* CE_GPU_H_ */
