#ifndef CE_GPU_API_KERNELS_H_
#define CE_GPU_API_KERNELS_H_
#include <algorithm>

#include "common.h"
namespace ce {
template <typename T>
int TransposePadLaunch(const T *src, int rows, int cols, T *dst, int ld_dst, cudaStream_t s);
template <typename T>
int PadRowsLaunch(const T *src, int64_t rows, int cols, T *dst, int ld_dst, cudaStream_t s);
int RowSumU8Launch(const uint8_t *x, int rows, int ld, int cols, int32_t *out, cudaStream_t s);

// One entry of a batched device-to-device copy: dst receives `repeat` consecutive copies of the
// `bytes` bytes at src (both 2-byte aligned, bytes even; buffers must not overlap).
struct SegCopy {
  const void *src;
  void *dst;
  uint32_t bytes;
  uint32_t repeat;
};
// segs: device array of n entries; one CTA per entry.
int SegCopyLaunch(const SegCopy *segs_dev, int n, cudaStream_t s);
}  // namespace ce
#endif
