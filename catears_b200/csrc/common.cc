// common.cc -- error state, device selection and workspace buffers of libce_gpu.so.
#include "common.h"

#include <algorithm>

#include <string>

#include <mutex>

#include <map>

#include <chrono>

#include <string.h>

namespace ce {

namespace {
thread_local char g_error[2048] = "";
thread_local int64_t g_launches = 0;
}  // namespace

void SetError(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

const char *LastError() { return g_error; }
void ClearError() { g_error[0] = '\0'; }
int64_t &LaunchCounter() { return g_launches; }

namespace {
struct ProfRec {
  int cat;
  cudaEvent_t e0, e1;
  int64_t launches_before;
  int64_t launches;
};
thread_local bool g_prof_on = false;
thread_local std::vector<ProfRec> g_prof_recs;
thread_local std::vector<cudaEvent_t> g_prof_pool;
thread_local int g_prof_depth = 0;

cudaEvent_t ProfEvent() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void ProfEnable(bool on) { g_prof_on = on; }

void ProfBegin(int cat, cudaStream_t s) {
  if (!g_prof_on) return;
  if (g_prof_depth++ > 0) return;                        // nested scopes: the outermost one counts
  ProfRec r;
  r.cat = cat;
  r.e0 = ProfEvent();
  r.e1 = ProfEvent();
  r.launches_before = g_launches;
  r.launches = 0;
  cudaEventRecord(r.e0, s);
  g_prof_recs.push_back(r);
}

void ProfEnd(cudaStream_t s) {
  if (!g_prof_on || g_prof_depth == 0) return;
  if (--g_prof_depth > 0) return;
  ProfRec &r = g_prof_recs.back();
  r.launches = g_launches - r.launches_before;
  cudaEventRecord(r.e1, s);
}

int ProfRead(double *ms, int64_t *launches) {
  for (int i = 0; i < kProfNum; ++i) {
    ms[i] = 0.0;
    launches[i] = 0;
  }
  for (ProfRec &r : g_prof_recs) {
    CE_CUDA(cudaEventSynchronize(r.e1));
    float t = 0.0f;
    CE_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms[r.cat] += t;
    launches[r.cat] += r.launches;
    g_prof_pool.push_back(r.e0);
    g_prof_pool.push_back(r.e1);
  }
  g_prof_recs.clear();
  return CE_GPU_OK;
}

int ProfTrace(int cap, int32_t *cat, double *t0_ms, double *t1_ms) {
  int n = 0;
  cudaEvent_t base = g_prof_recs.empty() ? nullptr : g_prof_recs.front().e0;
  for (ProfRec &r : g_prof_recs) {
    CE_CUDA(cudaEventSynchronize(r.e1));
    if (n < cap) {
      float a = 0.0f, b = 0.0f;
      CE_CUDA(cudaEventElapsedTime(&a, base, r.e0));
      CE_CUDA(cudaEventElapsedTime(&b, base, r.e1));
      cat[n] = r.cat;
      t0_ms[n] = a;
      t1_ms[n] = b;
      ++n;
    }
  }
  for (ProfRec &r : g_prof_recs) {
    g_prof_pool.push_back(r.e0);
    g_prof_pool.push_back(r.e1);
  }
  g_prof_recs.clear();
  return n;
}

int DeviceCount() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int UseDevice(int device) {
  int n = DeviceCount();
  if (n <= 0) {
    SetError("no CUDA device is visible: libce_gpu has no CPU fallback");
    return CE_GPU_ENODEVICE;
  }
  if (device < 0 || device >= n) {
    SetError("device %d out of range (%d visible)", device, n);
    return CE_GPU_ENODEVICE;
  }
  static thread_local int checked_major[64] = {0};
  CE_CUDA(cudaSetDevice(device));
  if (device < 64 && checked_major[device] == 0) {
    int major = 0;
    CE_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    checked_major[device] = major;
  }
  if (device < 64 && checked_major[device] != 10) {
    SetError("device %d has compute capability %d.x; libce_gpu is built for sm_100a only", device,
             checked_major[device]);
    return CE_GPU_ENODEVICE;
  }
  return CE_GPU_OK;
}

int SmCount() {
  static thread_local int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool IsDevicePtr(const void *p) {
  if (p == nullptr) return false;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

int DevBuf::Reserve(size_t bytes) {
  if (bytes <= cap) return CE_GPU_OK;
  Free();
  size_t want = (bytes + 255) & ~(size_t)255;
  cudaError_t e = cudaMalloc(&ptr, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    ptr = nullptr;
    SetError("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    return CE_GPU_ENOMEM;
  }
  cap = want;
  return CE_GPU_OK;
}

void DevBuf::Free() {
  if (ptr) cudaFree(ptr);
  ptr = nullptr;
  cap = 0;
}

int PinnedBuf::Acquire(size_t bytes) {
  if (pending) {
    CE_CUDA(cudaEventSynchronize(inflight));
    pending = false;
  }
  if (bytes <= cap) return CE_GPU_OK;
  if (ptr) cudaFreeHost(ptr);
  ptr = nullptr;
  cap = 0;
  size_t want = (bytes + 4095) & ~(size_t)4095;
  cudaError_t e = cudaMallocHost(&ptr, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    ptr = nullptr;
    SetError("cudaMallocHost(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    return CE_GPU_ENOMEM;
  }
  cap = want;
  return CE_GPU_OK;
}

int PinnedBuf::Release(cudaStream_t s) {
  if (!inflight) CE_CUDA(cudaEventCreateWithFlags(&inflight, cudaEventDisableTiming));
  CE_CUDA(cudaEventRecord(inflight, s));
  pending = true;
  return CE_GPU_OK;
}

void PinnedBuf::Free() {
  if (pending) cudaEventSynchronize(inflight);
  pending = false;
  if (inflight) cudaEventDestroy(inflight);
  inflight = nullptr;
  if (ptr) cudaFreeHost(ptr);
  ptr = nullptr;
  cap = 0;
}

int Table::Upload(size_t bytes, cudaStream_t s) {
  if (bytes == 0) return CE_GPU_OK;
  CE_CUDA(cudaMemcpyAsync(dev_buf.ptr, host_buf.ptr, bytes, cudaMemcpyHostToDevice, s));
  return host_buf.Release(s);
}

int StageIn(const void *src, size_t bytes, DevBuf *stage, cudaStream_t s, const void **dev_out) {
  if (bytes == 0 || IsDevicePtr(src)) {
    *dev_out = src;
    return CE_GPU_OK;
  }
  CE_CHECK(stage->Reserve(bytes));
  CE_CUDA(cudaMemcpyAsync(stage->ptr, src, bytes, cudaMemcpyHostToDevice, s));
  *dev_out = stage->ptr;
  return CE_GPU_OK;
}

int StageOut(void *dst, const void *dev_src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return CE_GPU_OK;
  CE_CUDA(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToHost, s));
  CE_CUDA(cudaStreamSynchronize(s));
  return CE_GPU_OK;
}

}  // namespace ce

namespace ce {
namespace {
struct HostProfState {
  std::map<std::string, std::pair<double, long>> acc;
  std::mutex mu;
  ~HostProfState() {
    for (auto &kv : acc)
      fprintf(stderr, "host-prof %-28s %10.1f us total %8ld calls %8.2f us/call\n", kv.first.c_str(), kv.second.first,
              kv.second.second, kv.second.first / std::max<long>(1, kv.second.second));
  }
};
HostProfState g_host_prof;
}  // namespace

void HostMark(const char *label) {
  static const bool on = getenv("CE_GPU_HOST_PROF") != nullptr;
  if (!on) return;
  static thread_local double last = 0;
  const double t = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
  if (label && last > 0) {
    std::lock_guard<std::mutex> lock(g_host_prof.mu);
    auto &e = g_host_prof.acc[label];
    e.first += t - last;
    e.second += 1;
  }
  last = t;
}
}  // namespace ce
