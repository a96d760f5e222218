// Micro-benchmark 4: the in-kernel issue pattern of tc_conv (per step: LDS.64 of the schedule entry, one add per operand, two MMAs
// N=64 / N=32 through umma_bf16_w / umma_bf16_acc) standalone, and with NOISE warps (ALU + shared-memory traffic on every SMSP)
// beside it: is the ~84 cycles per MMA seen inside tc_conv the issue loop, SMSP competition, or the tensor pipe?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hcmoco_b200/csrc scripts/bench_umma4.cu -o scripts/bench_umma4.bin
#include "tc_common.cuh"
#include <cstdio>

__device__ __forceinline__ uint64_t sw_desc(uint32_t saddr, uint32_t SW) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(((8 * SW) >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(SW == 128 ? 2 : 4) << 61;
  return d;
}
// mode 0: table-driven lean loop (as tc_conv); noise = number of extra warps running an ALU+STS loop until the MMA warp is done
__global__ void __launch_bounds__(704) bench(int reps, int noise, int unroll8, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  __shared__ volatile int stop;
  __shared__ uint2 steps[18];
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); stop = 0; }
  if (threadIdx.x < 18) {
    const int tap = threadIdx.x / 2, j = threadIdx.x & 1;
    const uint32_t rowoff = (uint32_t)((tap / 3) * 66 + tap % 3);
    steps[threadIdx.x] = make_uint2((rowoff * 64u + (uint32_t)j * 32u) >> 4, (uint32_t)smem_desc(smem_u32(smem) + 64u * 1024u + threadIdx.x * 2048u, 64 * 16, 128));
  }
  if (warp == nwarps - 1) tmem_alloc(smem_u32(&tptr), 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  if (warp == nwarps - 1) {
    const uint64_t a_t = sw_desc(0, 64), b_t = smem_desc(0, 64 * 16, 128);
    const uint32_t a_hi32 = (uint32_t)(a_t >> 32), b_hi32 = (uint32_t)(b_t >> 32);
    const uint32_t abl = (uint32_t)a_t + (smem_u32(smem) >> 4), lo16 = (17408u) >> 4;
    const uint32_t idesc_2n = instr_desc(64), idesc_n = instr_desc(32);
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < reps; ++it) {
#pragma unroll 4
        for (int k = 0; k < 18; ++k) {
          const uint2 stp = steps[k];
          const uint32_t ah = abl + stp.x;
          umma_bf16_w(tmem, ah, a_hi32, stp.y, b_hi32, idesc_2n, (k | it) ? 1u : 0u);
          umma_bf16_acc(tmem, ah + lo16, a_hi32, stp.y, b_hi32, idesc_n);
        }
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { out[blockIdx.x] = t1 - t0; stop = 1; }
  } else if (warp < noise) {
    // noise: dependent FMAs + a swizzled 16-byte shared store per iteration (what a transform warp does)
    float a = threadIdx.x, b = 1.0001f;
    uint4* dst = reinterpret_cast<uint4*>(smem + 100 * 1024) + threadIdx.x;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 32; ++i) a = fmaf(a, b, 0.5f);
      *dst = make_uint4(__float_as_uint(a), 0, 0, 0);
    }
    if (a == 123.f) out[147] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == nwarps - 1) tmem_dealloc(tmem, 128);
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  const int reps = 200;
  for (int threads : {64, 704})
    for (int noise : {0, 4, 8, 16, 20}) {
      if (noise > threads / 32 - 1) continue;
      bench<<<148, threads, 128 * 1024>>>(reps, noise, 0, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < 146; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("threads %3d noise warps %2d : %6.1f cycles/MMA (N=64 + N=32 pairs, table-driven lean loop) (%s)\n", threads, noise,
             (double)mx / (reps * 36), cudaGetErrorString(e));
    }
  return 0;
}
