// Micro-benchmark 2: what bounds a stream of small-N tcgen05.mma (kind::f16, K=16, SS operands)?
//   variant A: one issuing warp per CTA, 1 CTA/SM              (baseline: ~83-105 cycles/MMA for N <= 128)
//   variant B: TWO issuing warps per CTA (disjoint accumulators, own operand tiles)
//   variant C: two CTAs per SM, one issuing warp each
//   variant D: M = 64
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hcmoco_b200/csrc scripts/bench_umma2.cu -o scripts/bench_umma2.bin
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
__device__ __forceinline__ uint32_t idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// nissue = issuing warps per CTA (1 or 2); each issues `reps` MMAs into its own accumulator columns from its own smem tiles
__global__ void __launch_bounds__(128) bench(int M, int N, int reps, int nissue, int tmem_cols, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tptr), tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  if (warp < nissue) {
    // A tiles: 4 x 16 KB at [warp*48K, ...), B tile 32 KB behind them
    const uint32_t a0 = smem_u32(smem) + (uint32_t)warp * 48u * 1024u, b0 = a0 + 16u * 1024u;
    const uint32_t idesc = idesc_mn(M, N);
    uint64_t ad[4], bd = sw128_desc(b0);
    for (int k = 0; k < 4; ++k) ad[k] = sw128_desc(a0 + (uint32_t)k * 2048u);
    const uint32_t d = tmem + (uint32_t)warp * (uint32_t)(tmem_cols / 2);
    __syncwarp();
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < reps; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(d, ad[k & 3], bd, idesc, 1u);
      }
      umma_commit(smem_u32(&bar[warp]));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar[warp]), 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 2 + warp] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

int main() {
  long long* out;
  cudaMalloc(&out, 2 * 296 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 4000;
  struct V { const char* name; int grid, nissue, M, tcols; } vs[] = {
      {"A: 1 warp, 1 CTA/SM, M=128", 148, 1, 128, 512}, {"B: 2 warps/CTA, M=128", 148, 2, 128, 512},
      {"C: 2 CTAs/SM, M=128", 296, 1, 128, 256},        {"D: 1 warp, M=64", 148, 1, 64, 512},
      {"E: 2 warps/CTA, M=64", 148, 2, 64, 512}};
  for (auto& v : vs)
    for (int N : {32, 64, 96, 128, 192, 256}) {
      if (v.tcols / (v.nissue == 2 ? 2 : 1) < N) continue;
      cudaMemset(out, 0, 2 * 296 * sizeof(long long));
      bench<<<v.grid, 128, 100 * 1024>>>(v.M, N, reps, v.nissue, v.tcols, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2 * 296];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < 2 * v.grid; ++i) mx = h[i] > mx ? h[i] : mx;
      const double per = (double)mx / reps;                       // cycles per MMA of ONE issuer
      const double streams = (v.grid / 148) * v.nissue;           // concurrent issuers per SM
      printf("%-28s N %3d : %6.1f cycles/MMA per issuer, %6.1f cycles/MMA per SM, %5.0f MAC/clk/SM (%s)\n", v.name, N, per,
             per / streams, (double)v.M * N * 16 * streams / per, cudaGetErrorString(e));
    }
  return 0;
}
