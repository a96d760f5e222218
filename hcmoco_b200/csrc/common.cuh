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

// ---- Programmatic dependent launch (PDL).  A kernel launched through hcm_launch_pdl may be scheduled while its stream predecessor
// is still running: its CTAs become resident as SM resources free up, run their prologue (barrier init, shared-memory zeroing,
// TMEM allocation: nothing that touches global data of the step) and block in pdl_wait() until the predecessor has completed and
// flushed.  Protocol of every kernel launched this way: [shared-memory-only prologue] -> pdl_wait() -> pdl_trigger() -> body.  The
// trigger comes AFTER the wait, so a dependent can only start once this kernel's own predecessor is complete: completion stays
// transitive along the stream, exactly as plain stream order.  Multi-wave elementwise kernels trigger at their END instead (their
// dependents' resident-but-blocked CTAs would otherwise take slots from their own later waves).  Captured into the step's CUDA
// graph as programmatic edges.  MEASURED (B200, round 2, tc_conv / tc_wgrad2 / the six BatchNorm kernels = 95 % of the launches): parity
// and graph capture are fine, the step does not move (79.03 ms with, 78.95 ms without): the launch gaps and prologues of one stream
// are already covered by the kernels of the other streams of the step graph.  OFF by default; HCM_PDL=1 switches it on.
#include <stdlib.h>
#include <utility>
static inline bool hcm_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("HCM_PDL"); on = (e && e[0] == '1') ? 1 : 0; }
  return on == 1;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t hcm_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                         Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = hcm_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// Division of 0 <= n < 2^31 by a launch constant without the ~40-instruction emulated integer divide (the position decode of
// every staged row and, worse, the per-tile bookkeeping of the single MMA-issuing warp): m = ceil(2^(31+l) / d), l = ceil(log2 d),
// n / d = umulhi(n, m) >> (l - 1)  (error term n / 2^(31+l) < 2^-l <= 1/d, so the floor is exact).
struct FastDiv { uint32_t d, m, s; };
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f; f.d = d; f.m = 0; f.s = 0;
  if (d <= 1) return f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  const unsigned long long k = 31 + l, p2 = 1ull << k;
  f.m = (uint32_t)((p2 + d - 1) / d);
  f.s = l - 1;
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) { return f.d <= 1 ? n : (__umulhi(n, f.m) >> f.s); }
#endif

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
