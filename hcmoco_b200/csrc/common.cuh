// Shared helpers for the hcmoco sm_100a kernels: error reporting for the C-ABI, launch checks,
// warp reductions.  The library never allocates, frees or retains pointers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define HCM_OK 0
#define HCM_ERR_ARG (-1)
#define HCM_ERR_CUDA (-2)
#define HCM_ERR_UNSUPPORTED (-3)

void hcm_set_error(const char* fmt, ...);

#define HCM_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      hcm_set_error(__VA_ARGS__);       \
      return HCM_ERR_ARG;               \
    }                                   \
  } while (0)

#define HCM_LAUNCH_CHECK(name)                                                   \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      hcm_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
      return HCM_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

static inline int hcm_cdiv(long a, long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum for blockDim.x <= 1024; `red` is a __shared__ float[32]
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) r = warp_max(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  return red[0];
}
