// common.cuh -- shared helpers for the sm_100a SGA kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sga_b200.h"

namespace sga {

void set_error(const char* fmt, ...);

#define SGA_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      sga::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);   \
      return SGA_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)

#define SGA_CHECK(cond, ...)                    \
  do {                                          \
    if (!(cond)) {                              \
      sga::set_error(__VA_ARGS__);              \
      return SGA_ERR_INVALID;                   \
    }                                           \
  } while (0)

#define SGA_TRY(expr)          \
  do {                         \
    int _s = (expr);           \
    if (_s != SGA_OK) return _s; \
  } while (0)

extern thread_local long long g_launches;   // kernels launched by this library (sga_launch_count)
#define SGA_LAUNCH_CHECK()        \
  do {                            \
    ++sga::g_launches;            \
    SGA_CUDA(cudaGetLastError()); \
  } while (0)

// A strided NHWC view (channel stride 1). Phase views of a [N,2H,2W,C] tensor use sH = 2*(2W)*C, sW = 2*C.
struct View {
  const void* base;
  long long sN, sH, sW;  // element strides
  int H, W;              // extent of the view's pixel grid
};

struct Tap {
  int view;  // which input view
  int dy, dx;
  int widx;  // index of the (kh*kw) weight slice
};

constexpr int kMaxTaps = 25;
constexpr int kMaxViews = 4;

// kinds of (de)conv, matching sga_op_conv
enum ConvKind { kUp2 = 0, kUp2Dgrad = 1, kS1True = 2, kS1TrueDgrad = 3, kDown2Corr = 4, kS1Corr = 5 };

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

}  // namespace sga
