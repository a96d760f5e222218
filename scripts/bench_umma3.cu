// Micro-benchmark 3: cost of a ROW-SHIFTED A operand.  tc_conv addresses a filter tap as a row offset into the staged halo, so 8 of
// 9 taps start at a row that is not a multiple of 8 (the swizzle atom).  Same lean issue loop as bench_umma2.cu (M=128, K=16, SS).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hcmoco_b200/csrc scripts/bench_umma3.cu -o scripts/bench_umma3.bin
#include "tc_common.cuh"
#include <cstdio>

__device__ __forceinline__ uint64_t sw_desc(uint32_t saddr, uint32_t SW) {     // K-major, rows of SW bytes, SBO = 8*SW
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(((8 * SW) >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(SW == 128 ? 2 : 4) << 61;
  return d;
}
__global__ void __launch_bounds__(128) bench(int N, int reps, int SW, int shift_rows, int b_noswz, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tptr), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  if (warp == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 64u * 1024u;
    const uint32_t idesc = instr_desc(N);
    uint64_t ad[4], bd = b_noswz ? smem_desc(b0, (uint32_t)(2 * N) * 16, 128) : sw_desc(b0, 128);
    for (int k = 0; k < 4; ++k) ad[k] = sw_desc(a0 + (uint32_t)(k * shift_rows) * (uint32_t)SW + (uint32_t)(k & 1) * 32u, SW);
    __syncwarp();
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < reps; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tmem, ad[k & 3], bd, idesc, 1u);
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int reps = 4000;
  for (int SW : {128, 64})
    for (int shift : {0, 8, 1, 2, 3, 66})
      for (int bn : {0, 1})
        for (int N : {32, 64, 144, 256}) {
          bench<<<148, 128, 160 * 1024>>>(N, reps, SW, shift, bn, out);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[148];
          cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("A SW%-3d rows shifted by k*%-2d  B %s  N %3d : %6.1f cycles/MMA (%s)\n", SW, shift, bn ? "no-swizzle" : "SW128     ", N,
                 (double)mx / reps, cudaGetErrorString(e));
        }
  return 0;
}
