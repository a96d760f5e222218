// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS mode) as a function of N, operand layout,
// accumulator reuse.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hcmoco_b200/csrc scripts/bench_umma.cu -o /tmp/bench_umma
#include "tc_common.cuh"
#include <cstdio>

__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// mode bits: 1 = swizzle128 (else no-swizzle interleave); 2 = rotate over 3 accumulators; 4 = vary A start address per MMA
__global__ void __launch_bounds__(128) bench(int N, int reps, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tptr), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  if (threadIdx.x < 32) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    const uint32_t idesc = instr_desc(N);
    long long t0 = clock64();
    if (mode & 8) {
      // lean issue: descriptors precomputed, 8 MMAs per loop iteration, no per-MMA integer math
      uint64_t ad[4], bd = sw128_desc(b0);
      for (int k = 0; k < 4; ++k) ad[k] = sw128_desc(a0 + (uint32_t)k * 3072u);
      if (elect_one()) {
        for (int i = 0; i < reps; i += 8) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_bf16(tmem + (uint32_t)((k % 3) * N % 256), ad[k & 3], bd, idesc, 1u);
        }
      }
    } else
    for (int i = 0; i < reps; ++i) {
      const uint32_t ashift = (mode & 4) ? (uint32_t)(i % 9) * 1024u * 3u : 0u;
      uint64_t ad, bd;
      if (mode & 1) { ad = sw128_desc(a0 + ashift); bd = sw128_desc(b0); }
      else { ad = smem_desc(a0 + ashift, 4096, 128); bd = smem_desc(b0, (uint32_t)N * 16, 128); }
      const uint32_t d = tmem + ((mode & 2) ? (uint32_t)((i % 3) * N) : 0u);
      if (elect_one()) umma_bf16(d, ad, bd, idesc, i >= 3 ? 1u : 0u);
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 3000;
  for (int grid : {148})
    for (int mode : {1, 9})
      for (int N : {16, 32, 64, 128, 256}) {
        if ((mode & 2) && !(mode & 8) && 3 * N > 512) continue;
        bench<<<grid, 128, 200 * 1024>>>(N, reps, mode, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("grid %3d mode %d (swz %d, 3acc %d, vary %d) N %3d : %7.1f cycles/MMA  (%s)\n", grid, mode, mode & 1, (mode >> 1) & 1,
               (mode >> 2) & 1, N, (double)mx / reps, cudaGetErrorString(e));
      }
  return 0;
}
