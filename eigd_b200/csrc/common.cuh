// Shared helpers for the CUDA translation units of libeigd_b200.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

void eigd_set_error(const char* fmt, ...);

extern cudaStream_t g_eigd_stream;
extern int64_t g_eigd_launches;

#define EIGD_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      eigd_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return 100 + (int)e_;                                                               \
    }                                                                                     \
  } while (0)

// launch bookkeeping: every kernel launch of the library goes through this macro
#define EIGD_LAUNCH(kernel, grid, block, smem, ...)                                       \
  do {                                                                                    \
    kernel<<<(grid), (block), (smem), g_eigd_stream>>>(__VA_ARGS__);                      \
    ++g_eigd_launches;                                                                    \
  } while (0)

#define EIGD_CHECK_LAUNCH()                                                               \
  do {                                                                                    \
    cudaError_t e_ = cudaGetLastError();                                                  \
    if (e_ != cudaSuccess) {                                                              \
      eigd_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return 100 + (int)e_;                                                               \
    }                                                                                     \
  } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
