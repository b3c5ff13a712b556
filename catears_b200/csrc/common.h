// Internal helpers shared by the translation units of libce_gpu.so.
#ifndef CE_GPU_COMMON_H_
#define CE_GPU_COMMON_H_

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "ce_gpu.h"

namespace ce {

// -- errors -------------------------------------------------------------------
void SetError(const char *fmt, ...);   // thread-local message (ce_gpu_last_error)
const char *LastError();
void ClearError();
int64_t &LaunchCounter();              // thread-local count of kernel launches
int DeviceCount();

#define CE_CUDA(expr)                                                              \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      ::ce::SetError("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                     __FILE__, __LINE__);                                          \
      return CE_GPU_ECUDA;                                                         \
    }                                                                              \
  } while (0)

#define CE_CHECK(expr)                  \
  do {                                  \
    int _rc = (expr);                   \
    if (_rc != CE_GPU_OK) return _rc;   \
  } while (0)

// Checks the launch and counts it.
#define CE_LAUNCHED()                                                              \
  do {                                                                             \
    ++::ce::LaunchCounter();                                                       \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) {                                                       \
      ::ce::SetError("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),   \
                     __FILE__, __LINE__);                                          \
      return CE_GPU_ECUDA;                                                         \
    }                                                                              \
  } while (0)

// -- optional per-category kernel timing (ce_gpu_profile_*) ------------------------------------
enum ProfCat { kProfFbank = 0, kProfCmvn, kProfGemm, kProfQuantize, kProfFinalize, kProfOther, kProfNum };
void ProfBegin(int cat, cudaStream_t s);
void ProfEnd(cudaStream_t s);
struct ProfScope {   // CUDA events around the launches made while it is alive (when enabled)
  cudaStream_t s;
  ProfScope(int cat, cudaStream_t stream) : s(stream) { ProfBegin(cat, s); }
  ~ProfScope() { ProfEnd(s); }
};
void ProfEnable(bool on);
// Waits for the recorded launches, adds their times (ms) and counts per category, clears them.
int ProfRead(double *ms, int64_t *launches);
// Per-scope begin/end times (ms, relative to the first recorded scope) in launch order; returns the
// number of records written (<= cap) or a negative error, and clears the records.
int ProfTrace(int cap, int32_t *cat, double *t0_ms, double *t1_ms);

// -- host-time probe (CE_GPU_HOST_PROF=1): HostMark("label") adds the host time since the previous mark
// of this thread to the label's total; totals are printed at process exit.  A debugging aid.
void HostMark(const char *label);

// -- device selection -----------------------------------------------------------
// Makes `device` current; fails with CE_GPU_ENODEVICE when there is none / not sm_100.
int UseDevice(int device);

// Multiprocessor count of the current device (cached).
int SmCount();

// true if `p` is device memory (cudaMalloc / torch), false for host memory.
bool IsDevicePtr(const void *p);

// -- growable device / pinned buffers --------------------------------------------
struct DevBuf {
  void *ptr = nullptr;
  size_t cap = 0;
  int Reserve(size_t bytes);   // grows (never shrinks); contents are not preserved
  void Free();
  template <typename T> T *as() const { return static_cast<T *>(ptr); }
};

// Pinned host staging for small tables that are uploaded with cudaMemcpyAsync.  Acquire()
// waits until the previous upload from this buffer has been consumed (so the host may
// overwrite it) and grows it; Release() marks the upload that was just enqueued on `s`.
struct PinnedBuf {
  void *ptr = nullptr;
  size_t cap = 0;
  cudaEvent_t inflight = nullptr;
  bool pending = false;
  int Acquire(size_t bytes);
  int Release(cudaStream_t s);
  void Free();
  template <typename T> T *as() const { return static_cast<T *>(ptr); }
};

// A small host table mirrored on the device: Acquire -> fill host() -> Upload.
struct Table {
  PinnedBuf host_buf;
  DevBuf dev_buf;
  int Acquire(size_t bytes) {
    int rc = host_buf.Acquire(bytes);
    return rc != CE_GPU_OK ? rc : dev_buf.Reserve(bytes);
  }
  int Upload(size_t bytes, cudaStream_t s);
  void Free() { host_buf.Free(); dev_buf.Free(); }
  template <typename T> T *host() const { return host_buf.as<T>(); }
  template <typename T> T *dev() const { return dev_buf.as<T>(); }
};

// Stages a (host or device) input so the kernels see device memory. If `src` is already a
// device pointer it is returned as is; otherwise it is copied into `stage`.
int StageIn(const void *src, size_t bytes, DevBuf *stage, cudaStream_t s, const void **dev_out);
// Device -> host copy of a result, complete on return.
int StageOut(void *dst, const void *dev_src, size_t bytes, cudaStream_t s);

// -- frame bookkeeping (src/fbank.cc:35-42) ----------------------------------------
constexpr int kFrameLen = 400;     // PK_FRAMELENGTH_MS * 16   src/fbank.h:9
constexpr int kFrameShift = 160;   // PK_FRAMESHIFT_MS * 16    src/fbank.h:8
constexpr int kFftSize = 512;      // RoundUpToNearestPowerOfTwo(400)  src/fbank.cc:258
constexpr int kMaxMel = 128;
constexpr int kCmvnWindow = 600;   // src/cmvn.h:10
constexpr int kCmvnGlobal = 200;   // src/cmvn.h:11

inline int64_t NumFrames(int64_t n_samples) {
  return n_samples < kFrameLen ? 0 : 1 + (n_samples - kFrameLen) / kFrameShift;
}

// -- stage launchers (device pointers only) -------------------------------------------
// fbank.cu
int FbankLaunch(const int16_t *pcm_dev, int64_t total_samples, const int64_t *sample_off_host,
                const int64_t *frame_off_host, int n_utts, int num_mel, float *feats_dev,
                int64_t out_row_stride, Table *chunks, cudaStream_t s);
int Rfft512Launch(const float *in_dev, int n_frames, float *out_dev, cudaStream_t s);

// cmvn.cu.  Utterance u's frames are rows [frame_off[u], frame_off[u+1]) of feats (row stride
// num_mel).  Frame t is written to out row (out_row_off[u] + pad_left + t) (row stride
// out_stride floats, only the first num_mel columns are written) and, when pad_left /
// pad_right > 0, the first / last frame is replicated into the padding rows
// (src/am.cc:119-124,152-155).  global_stats_dev == nullptr: copy + pad only (no CMVN, which
// is what src/ce_stt.cc does).
// Streaming continuation: utterance u's rows start with n_hist[u] raw frames it has already
// normalised (the last min(t_base[u], 600) of them, needed as x_{t-600}); the chain resumes from
// state_dev[u * num_mel ..] (running sums, updated in place) at absolute frame index t_base[u].
// Only the new frames are written, to out rows out_row_off[u] ...
struct CmvnResume {
  const int32_t *n_hist;     // host [n_utts]
  const int64_t *t_base;     // host [n_utts]
  float *state_dev;          // device [n_utts x num_mel]
};
int CmvnLaunch(const float *global_stats_dev, float global_count, const float *feats_dev,
               const int64_t *frame_off_host, const int64_t *out_row_off_host, int n_utts,
               int num_mel, int pad_left, int pad_right, float *out_dev, int64_t out_stride,
               Table *utts, cudaStream_t s, const CmvnResume *resume = nullptr,
               uint32_t *minmax_dev = nullptr);   // [n_utts][2] ordered-int min/max of the rows written
                                                  // (initialised by the caller), nullptr = off

}  // namespace ce

#endif  // CE_GPU_COMMON_H_
