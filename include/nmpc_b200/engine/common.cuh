/* nmpc_b200 -- shared host/device helpers: error plumbing, device buffers, layout kernels. */
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include <nmpc_b200/c_api.h>

namespace nmpc_b200
{
/** Error carrying the C-ABI status code; caught at the extern "C" boundary. */
struct Error : public std::runtime_error
{
  Error(int c, const std::string & msg) : std::runtime_error(msg), code(c) {}
  int code;
};

void setLastError(const std::string & msg);

#define NMPC_CUDA_CHECK(expr)                                                                              \
  do                                                                                                       \
  {                                                                                                        \
    cudaError_t err__ = (expr);                                                                            \
    if(err__ != cudaSuccess)                                                                               \
    {                                                                                                      \
      throw ::nmpc_b200::Error(NMPC_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); \
    }                                                                                                      \
  } while(0)

/** Owning device allocation. */
template<class T>
struct DeviceBuffer
{
  T * ptr = nullptr;
  size_t count = 0;

  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer &) = delete;
  DeviceBuffer & operator=(const DeviceBuffer &) = delete;
  ~DeviceBuffer()
  {
    release();
  }
  void allocate(size_t n)
  {
    release();
    if(n == 0) return;
    NMPC_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&ptr), n * sizeof(T)));
    count = n;
  }
  /** Grow-only variant of allocate(): keeps the buffer when it is already large enough. */
  void reserve(size_t n)
  {
    if(count < n) allocate(n);
  }
  void release()
  {
    if(ptr) cudaFree(ptr);
    ptr = nullptr;
    count = 0;
  }
  size_t bytes() const
  {
    return count * sizeof(T);
  }
};

/** RAII device guard. */
struct DeviceGuard
{
  int prev = -1;
  explicit DeviceGuard(int dev)
  {
    cudaGetDevice(&prev);
    if(prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard()
  {
    if(prev >= 0) cudaSetDevice(prev);
  }
};

/* --------------------------------------------------------------------------- layout kernels ---- */
/* Instance-major [B][R] (the C ABI) <-> batch-innermost [R][Bp] (the engine).  32x32 tiles through
   shared memory so that both the global read and the global write are coalesced. */

/** dst[r * Bp + b] = TD(src[b * R + r]) */
template<class TS, class TD>
__global__ void scatter_rows_kernel(const TS * __restrict__ src, TD * __restrict__ dst, int B, int R, int Bp)
{
  __shared__ TD tile[32][33];
  const int b0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for(int j = threadIdx.y; j < 32; j += blockDim.y)
  {
    const int b = b0 + j, r = r0 + threadIdx.x;
    if(b < B && r < R) tile[j][threadIdx.x] = TD(src[(size_t)b * R + r]);
  }
  __syncthreads();
  for(int j = threadIdx.y; j < 32; j += blockDim.y)
  {
    const int r = r0 + j, b = b0 + threadIdx.x;
    if(b < B && r < R) dst[(size_t)r * Bp + b] = tile[threadIdx.x][j];
  }
}

/** dst[b * R + r] = TD(src_sel(b)[r * Bp + b]); src_sel picks src0 or src1 per instance when sel != nullptr;
    rows >= (row_limit[b] + row_limit_add) * row_group read as zero when row_limit != nullptr (unwritten trace rows). */
template<class TS, class TD>
__global__ void gather_rows_kernel(const TS * __restrict__ src0,
                                   const TS * __restrict__ src1,
                                   const int * __restrict__ sel,
                                   const int * __restrict__ row_limit,
                                   int row_limit_add,
                                   int row_group,
                                   TD * __restrict__ dst,
                                   int B,
                                   int R,
                                   int Bp)
{
  __shared__ TD tile[32][33];
  const int b0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for(int j = threadIdx.y; j < 32; j += blockDim.y)
  {
    const int r = r0 + j, b = b0 + threadIdx.x;
    if(b < B && r < R)
    {
      const TS * s = (sel != nullptr && sel[b] != 0) ? src1 : src0;
      TD v = TD(s[(size_t)r * Bp + b]);
      if(row_limit != nullptr && r >= (row_limit[b] + row_limit_add) * row_group) v = TD(0);
      tile[j][threadIdx.x] = v;
    }
  }
  __syncthreads();
  for(int j = threadIdx.y; j < 32; j += blockDim.y)
  {
    const int b = b0 + j, r = r0 + threadIdx.x;
    if(b < B && r < R) dst[(size_t)b * R + r] = tile[threadIdx.x][j];
  }
}

template<class TS, class TD>
inline void launchScatterRows(const TS * src, TD * dst, int B, int R, int Bp, cudaStream_t stream)
{
  if(B <= 0 || R <= 0) return;
  dim3 grid((B + 31) / 32, (R + 31) / 32), block(32, 8);
  scatter_rows_kernel<TS, TD><<<grid, block, 0, stream>>>(src, dst, B, R, Bp);
}

template<class TS, class TD>
inline void launchGatherRows(const TS * src0,
                             const TS * src1,
                             const int * sel,
                             const int * row_limit,
                             int row_limit_add,
                             int row_group,
                             TD * dst,
                             int B,
                             int R,
                             int Bp,
                             cudaStream_t stream)
{
  if(B <= 0 || R <= 0) return;
  dim3 grid((B + 31) / 32, (R + 31) / 32), block(32, 8);
  gather_rows_kernel<TS, TD><<<grid, block, 0, stream>>>(src0, src1, sel, row_limit, row_limit_add, row_group, dst, B, R, Bp);
}
} // namespace nmpc_b200
