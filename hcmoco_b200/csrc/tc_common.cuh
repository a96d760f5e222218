// PTX building blocks shared by the tcgen05 kernels (tc_conv.cu, tc_wgrad.cu): mbarrier, cp.async.bulk (TMA),
// TMEM allocation, tcgen05.mma / commit / ld, shared-memory matrix descriptors, bf16 hi/lo split.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

__host__ __device__ inline int ceil_to(int a, int b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// one lane of a converged warp (the caller keeps the surrounding control flow warp-uniform so that descriptors and
// addresses stay in uniform registers: a divergent `if (lane == 0)` around tcgen05.mma costs a uniformisation loop
// of ~15 dependent instructions per MMA)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// The same MMA with the descriptors given as (low word, high word) pairs: every descriptor of a kernel shares its high word
// (layout / swizzle / SBO) and differs only in the 14-bit start address (and LBO) of the low word, so the issue loop needs ONE
// integer add per operand.  Measured on B200 (scripts/bench_umma2.cu): a lean issue loop retires an M=128 SS-mode MMA every
// 32 + N/4 cycles for N <= 128 (shared-memory operand reads: 4 KB of A + 32*N B of B at 128 B/clk) and N/2 above; the generic
// loop (64-bit descriptor arithmetic + schedule lookups per MMA, sharing its SMSP with busy warps) ran at ~100.
__device__ __forceinline__ void umma_bf16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum) : "memory");
}
// accumulate variant with a compile-time-true predicate (no setp in the loop)
__device__ __forceinline__ void umma_bf16_acc(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.eq.u32 p, 0, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// the same load without the wait (issue several, then tmem_ld_wait() once: one TMEM latency instead of one per load);
// the values land in `v` as raw bits — valid only after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major / MN-major no-swizzle shared-memory matrix descriptor (sm_100 format, version 1)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor (kind::f16): D fp32, A/B bf16, M=128, N; mn_major=1 -> both operands MN-major (else K-major)
__device__ __forceinline__ uint32_t instr_desc(int N, int mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)mn_major << 15) | ((uint32_t)mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// two fp32 -> packed bf16x2 (round to nearest even), `a` in the low half: ONE F2FP instruction on the ALU pipe
// (the scalar __float2bfloat16_rn path costs an F2F on the slow conversion unit per element plus the packing)
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// split two fp32 into bf16 hi / lo pairs: x = hi + lo up to 2^-17 relative
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(a, b);
  lo = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}
// split 8 fp32 into bf16 hi / lo, packed 2 per 32-bit word (element k in the low half)
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// byte offset of 16-byte chunk `kb` (multiple of 16, < SW) of row `row` inside a swizzled plane (rows of SW bytes, 1024-B aligned
// base): Swizzle<log2(SW/16),4,3> — the chunk index is XORed with address bits [7, 7+log2(SW/16))
__device__ __forceinline__ uint32_t swz16(uint32_t row, uint32_t kb, uint32_t SW) {
  const uint32_t off = row * SW;
  return off + ((((kb >> 4) ^ (off >> 7)) & (SW / 16 - 1)) << 4);
}

// Stage `nch` channels [c_first, c_first + nch) of `rows` positions (src_tab[row * tab_stride + tab_off] = source pixel, < 0:
// padding -> zeros) of a channels-last fp32 tensor into a swizzled bf16 hi / lo tile pair: rows of SW bytes per plane, the
// first channel at byte `byte0` (multiple of 16) of the staged row, planes `plane` bytes apart, lo tile `lo_off` after hi.
// The optional per-channel affine (+ReLU) is the producer's pending BatchNorm.
// Work item = (row, 8-channel chunk) = one 16-byte hi chunk + one 16-byte lo chunk.  A thread keeps ONE chunk for the whole
// call (channel offset, destination column and the 8+8 BN coefficients live in registers) and walks the rows, two rows
// (up to 8 loads) in flight: ~9 instructions per channel instead of the ~57 of a (row, 2-4 channel) unit with per-unit
// table lookups.  Consecutive lanes read consecutive chunks of a row, then the next row: fully coalesced.
// The kernels are instruction-issue-bound in these warps (ncu, 64x64 18->18: ~8.6k warp instructions per 128-position tile,
// 67 % of the issue slots), so the FULL chunks run a path without per-channel validity predicates, and the partial tail chunk
// of a row (18 channels = 2 full chunks + 2 channels, 36 = 4 + 4) is a separate cheap pass that loads, splits and stores only
// its valid channels: the rest of its 16 bytes keeps the zeros written at kernel start (or finite stale values that meet zero
// weights / never-read accumulator columns).  Before, a tail chunk cost as much as a full one: 33 % of the items of an 18-channel row.
// V = 4: 16-byte loads (C % 4 == 0), V = 2: 8-byte loads (C even).  Threads t in [0, TS) of one team call it together.
template <int V>
__device__ __forceinline__ void stage_rows8(const float* __restrict__ src, int C, int c_first, int nch, int rows,
                                            const int* __restrict__ src_tab, int tab_stride, int tab_off, uint8_t* hi_base,
                                            uint32_t lo_off, uint32_t plane, uint32_t SW, uint32_t byte0, const float* s_sc,
                                            const float* s_sh, bool affine, int relu, int t, int TS) {
  const int cpp = nch >> 3;                             // full chunks per row
  const int ntail = nch & 7;                            // channels of the partial chunk (even; a multiple of 4 when V == 4)
  const int* tab = src_tab + tab_off;
  if (cpp > 0) {
    const int dpos = TS / cpp;                          // rows advanced per step (TS >= 128 >= cpp)
    if (t < dpos * cpp) {
      const int pos0 = t / cpp, ch = t - pos0 * cpp;
      const int c = c_first + ch * 8;
      const uint32_t kb = byte0 + (uint32_t)ch * 16;
      uint8_t* hi = hi_base + (kb / SW) * plane;
      const uint32_t kbin = kb % SW;
      float sc[8], sh[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sc[i] = affine ? s_sc[c + i] : 0.f;
        sh[i] = affine ? s_sh[c + i] : 0.f;
      }
      const float* srcc = src + c;
      for (int pos = pos0; pos < rows; pos += 2 * dpos) {
        float v[2][8];
        int px[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int pp = pos + u * dpos;
          px[u] = pp < rows ? tab[pp * tab_stride] : -2;
#pragma unroll
          for (int i = 0; i < 8; ++i) v[u][i] = 0.f;
          if (px[u] >= 0) {
            const float* xp = srcc + (long)px[u] * C;
            if (V == 4) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(xp));
              const float4 b = __ldg(reinterpret_cast<const float4*>(xp + 4));
              v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w;
              v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 a = __ldg(reinterpret_cast<const float2*>(xp + 2 * j));
                v[u][2 * j] = a.x; v[u][2 * j + 1] = a.y;
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (px[u] == -2) continue;
          if (affine && px[u] >= 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float a = fmaf(v[u][i], sc[i], sh[i]);
              v[u][i] = relu ? fmaxf(a, 0.f) : a;
            }
          }
          uint4 h, l;
          split8(v[u], h, l);
          const uint32_t off = swz16((uint32_t)(pos + u * dpos), kbin, SW);
          *reinterpret_cast<uint4*>(hi + off) = h;
          *reinterpret_cast<uint4*>(hi + lo_off + off) = l;
        }
      }
    }
  }
  if (ntail) {
    // partial chunk: thread t takes rows t, t + TS, ... (two in flight), only the valid channels are loaded / written
    const int c = c_first + cpp * 8;
    const uint32_t kb = byte0 + (uint32_t)cpp * 16;
    uint8_t* hi = hi_base + (kb / SW) * plane;
    const uint32_t kbin = kb % SW;
    float sc[6], sh[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const bool ok = affine && i < ntail;
      sc[i] = ok ? s_sc[c + i] : 1.f;
      sh[i] = ok ? s_sh[c + i] : 0.f;
    }
    const float lo_clamp = relu ? 0.f : -INFINITY;
    const float* srcc = src + c;
    const int np = ntail >> 1;                          // channel pairs: 1..3 (V == 4: 2)
    for (int pos = t; pos < rows; pos += 2 * TS) {
      float v[2][6];
      int px[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int pp = pos + u * TS;
        px[u] = pp < rows ? tab[pp * tab_stride] : -2;
#pragma unroll
        for (int i = 0; i < 6; ++i) v[u][i] = 0.f;
        if (px[u] >= 0) {
          const float* xp = srcc + (long)px[u] * C;
          if (V == 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(xp));
            v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w;
          } else {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              if (j < np) {
                const float2 a = __ldg(reinterpret_cast<const float2*>(xp + 2 * j));
                v[u][2 * j] = a.x; v[u][2 * j + 1] = a.y;
              }
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (px[u] == -2) continue;
        const uint32_t off = swz16((uint32_t)(pos + u * TS), kbin, SW);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j < np) {
            uint32_t hw = 0, lw = 0;
            if (px[u] >= 0) {
              float a = v[u][2 * j], b = v[u][2 * j + 1];
              if (affine) { a = fmaxf(fmaf(a, sc[2 * j], sh[2 * j]), lo_clamp); b = fmaxf(fmaf(b, sc[2 * j + 1], sh[2 * j + 1]), lo_clamp); }
              split2(a, b, hw, lw);
            }
            *reinterpret_cast<uint32_t*>(hi + off + 4 * j) = hw;
            *reinterpret_cast<uint32_t*>(hi + lo_off + off + 4 * j) = lw;
          }
        }
      }
    }
  }
}

}  // namespace
