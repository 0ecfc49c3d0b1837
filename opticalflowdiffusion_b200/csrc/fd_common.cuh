// Shared helpers for libflowdiff.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/flowdiff.h"

#define FD_NUM_SMS 148

void fd_set_error(const char* fmt, ...);

#define FD_REQUIRE(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      fd_set_error(__VA_ARGS__);              \
      return FD_EINVAL;                       \
    }                                         \
  } while (0)

#define FD_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      fd_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return FD_ECUDA;                                                             \
    }                                                                              \
  } while (0)

#define FD_LAUNCH_CHECK() FD_CUDA(cudaGetLastError())

static inline int fd_ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float fd_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float fd_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum of up to NV values; result valid in thread 0. blockDim.x multiple of 32, <= 1024.
template <int NV>
__device__ __forceinline__ void fd_block_sum(float (&v)[NV], float* smem /* >= NV*32 floats */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = fd_warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[i * 32 + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float x = lane < nwarp ? smem[i * 32 + lane] : 0.f;
      v[i] = fd_warp_sum(x);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float fd_silu(float x) { return x / (1.f + __expf(-x)); }

__device__ __forceinline__ uint32_t fd_pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 fd_unpack_bf16(uint32_t u) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(h);
}
