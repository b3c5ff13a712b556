// api_kernels.cu -- small layout helpers behind the matrix-level test hooks of ce_gpu.h
// (ce_gpu_gemm_u8 / ce_gpu_gemm_f32): operands arrive row-major with arbitrary sizes and are
// repacked into the zero-padded K-major layouts the tensor-core kernel reads.
#include "api_kernels.h"

namespace ce {
namespace {

// dst[c][r_pad...] : dst is [cols x ld_dst], dst[c][r] = src[r][c] for r < rows, 0 for r >= rows.
template <typename T>
__global__ void transpose_pad_kernel(const T *__restrict__ src, int rows, int cols, T *__restrict__ dst,
                                     int ld_dst) {
  __shared__ T tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * cols + c] : T(0);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < ld_dst) dst[(int64_t)c * ld_dst + r] = tile[threadIdx.x][j];
  }
}

// dst [rows x ld_dst] = src [rows x cols] zero padded.
template <typename T>
__global__ void pad_rows_kernel(const T *__restrict__ src, int64_t rows, int cols, T *__restrict__ dst,
                                int ld_dst) {
  const int64_t n = rows * ld_dst;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld_dst;
    const int c = (int)(i % ld_dst);
    dst[i] = c < cols ? src[r * cols + c] : T(0);
  }
}

__global__ void rowsum_u8_kernel(const uint8_t *__restrict__ x, int rows, int ld, int cols,
                                 int32_t *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int32_t s = 0;
  for (int c = lane; c < cols; c += 32) s += x[(int64_t)row * ld + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}

// Batched small copies (the per-stream state shuffles of streams.cc): CTA b copies entry b.
__global__ void __launch_bounds__(256) segcopy_kernel(const SegCopy *__restrict__ segs) {
  const SegCopy sc = segs[blockIdx.x];
  const uintptr_t a = reinterpret_cast<uintptr_t>(sc.src) | reinterpret_cast<uintptr_t>(sc.dst) | sc.bytes;
  for (uint32_t r = 0; r < sc.repeat; ++r) {
    char *dst = static_cast<char *>(sc.dst) + (size_t)r * sc.bytes;
    if ((a & 15) == 0) {
      const uint4 *s4 = static_cast<const uint4 *>(sc.src);
      uint4 *d4 = reinterpret_cast<uint4 *>(dst);
      for (uint32_t i = threadIdx.x; i < sc.bytes / 16; i += blockDim.x) d4[i] = s4[i];
    } else if ((a & 3) == 0) {
      const uint32_t *s1 = static_cast<const uint32_t *>(sc.src);
      uint32_t *d1 = reinterpret_cast<uint32_t *>(dst);
      for (uint32_t i = threadIdx.x; i < sc.bytes / 4; i += blockDim.x) d1[i] = s1[i];
    } else {
      const uint16_t *s2 = static_cast<const uint16_t *>(sc.src);
      uint16_t *d2 = reinterpret_cast<uint16_t *>(dst);
      for (uint32_t i = threadIdx.x; i < sc.bytes / 2; i += blockDim.x) d2[i] = s2[i];
    }
  }
}

}  // namespace

int SegCopyLaunch(const SegCopy *segs_dev, int n, cudaStream_t s) {
  if (n <= 0) return CE_GPU_OK;
  ProfScope prof(kProfOther, s);
  segcopy_kernel<<<n, 256, 0, s>>>(segs_dev);
  CE_LAUNCHED();
  return CE_GPU_OK;
}

template <typename T>
int TransposePadLaunch(const T *src, int rows, int cols, T *dst, int ld_dst, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return CE_GPU_OK;
  dim3 block(32, 8), grid((cols + 31) / 32, (ld_dst + 31) / 32);
  transpose_pad_kernel<T><<<grid, block, 0, s>>>(src, rows, cols, dst, ld_dst);
  CE_LAUNCHED();
  return CE_GPU_OK;
}
template int TransposePadLaunch<uint8_t>(const uint8_t *, int, int, uint8_t *, int, cudaStream_t);
template int TransposePadLaunch<float>(const float *, int, int, float *, int, cudaStream_t);

template <typename T>
int PadRowsLaunch(const T *src, int64_t rows, int cols, T *dst, int ld_dst, cudaStream_t s) {
  if (rows <= 0) return CE_GPU_OK;
  const int64_t n = rows * ld_dst;
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
  pad_rows_kernel<T><<<grid, 256, 0, s>>>(src, rows, cols, dst, ld_dst);
  CE_LAUNCHED();
  return CE_GPU_OK;
}
template int PadRowsLaunch<uint8_t>(const uint8_t *, int64_t, int, uint8_t *, int, cudaStream_t);
template int PadRowsLaunch<float>(const float *, int64_t, int, float *, int, cudaStream_t);

int RowSumU8Launch(const uint8_t *x, int rows, int ld, int cols, int32_t *out, cudaStream_t s) {
  if (rows <= 0) return CE_GPU_OK;
  rowsum_u8_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, rows, ld, cols, out);
  CE_LAUNCHED();
  return CE_GPU_OK;
}

}  // namespace ce
