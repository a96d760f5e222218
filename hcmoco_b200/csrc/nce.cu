// Sample-level memory-bank NCE (CMCMem3, memory/mem_bank.py:157-205) as fused HBM kernels.
//
// Reference: three index_select gathers materialise w_m = bank_m[idx] ([B,K+1,128] each, 1.6 GB at
// B=64), six bmm's produce the logits, autograd keeps the gathered copies for backward.  Here one pass
// over the three gathered rows (16-byte vector loads, 8 lanes per 512-byte row, 4 rows per warp
// instruction) produces all six logit sets; nothing but the [6,B,K+1] logits is written.  Backward
// re-gathers the rows (recompute instead of a 1.6 GB save) and accumulates
//     dx_p[b] = sum_q coef_pq[b]/T * sum_k (softmax_pq[b,k] - [k==0]) * bank_q[idx[b,k]].
// Algorithmic bytes: 3*(K+1)*128*4 per sample per pass.
#include "common.cuh"

namespace {

constexpr int D = 128;           // feature dim (opt.feat_dim) — the kernels are specialised for 128
constexpr int NCE_THREADS = 256;
constexpr int ROWS_PER_CTA = 512;
// pair order of mem_bank.py:186-191: (query modality p, bank modality q), 0-based
// 0:(x1,w2) 1:(x2,w1) 2:(x2,w3) 3:(x3,w2) 4:(x1,w3) 5:(x3,w1)

__device__ __forceinline__ float dot16(const float4 (&w)[4], const float4 (&x)[4]) {
  float a = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    a = fmaf(w[j].x, x[j].x, a); a = fmaf(w[j].y, x[j].y, a);
    a = fmaf(w[j].z, x[j].z, a); a = fmaf(w[j].w, x[j].w, a);
  }
  return a;
}
__device__ __forceinline__ void axpy16(float4 (&acc)[4], float s, const float4 (&w)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    acc[j].x = fmaf(s, w[j].x, acc[j].x); acc[j].y = fmaf(s, w[j].y, acc[j].y);
    acc[j].z = fmaf(s, w[j].z, acc[j].z); acc[j].w = fmaf(s, w[j].w, acc[j].w);
  }
}
__device__ __forceinline__ float4 ldg_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

struct NceArgs {
  const float* bank[3];
  const float* x[3];   // query rows, row stride ldx
  long ldx;
  const long long* idx;  // [B,K1]
  int B, K1;
  float invT;
};

// grid (ceil(K1/ROWS_PER_CTA), B); logits [6][B][K1]
__global__ void __launch_bounds__(NCE_THREADS) nce_logits_kernel(const NceArgs a, float* __restrict__ logits) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, l8 = lane & 7;
  float4 x[3][4];
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) x[m][j] = *reinterpret_cast<const float4*>(a.x[m] + (long)b * a.ldx + 32 * j + 4 * l8);
  const int kbeg = blockIdx.x * ROWS_PER_CTA;
  const int kend = min(a.K1, kbeg + ROWS_PER_CTA);
  const long long* idx = a.idx + (long)b * a.K1;
  // warp-uniform trip count (the shuffles below need all 32 lanes); tail rows are predicated
  for (int kb = kbeg + warp * 4; kb < kend; kb += (NCE_THREADS / 32) * 4) {
    const int k = kb + sub;
    const bool valid = k < kend;
    const long long row = valid ? idx[k] : 0;
    float4 w[3][4];
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) w[m][j] = ldg_stream(a.bank[m] + row * D + 32 * j + 4 * l8);
    float v[6];
    v[0] = dot16(w[1], x[0]); v[1] = dot16(w[0], x[1]); v[2] = dot16(w[2], x[1]);
    v[3] = dot16(w[1], x[2]); v[4] = dot16(w[2], x[0]); v[5] = dot16(w[0], x[2]);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
#pragma unroll
      for (int q = 0; q < 6; ++q) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    if (valid && l8 < 6) {
      float out = v[0];
#pragma unroll
      for (int q = 1; q < 6; ++q) out = (l8 == q) ? v[q] : out;
      logits[((long)l8 * a.B + b) * a.K1 + k] = out * a.invT;
    }
  }
}

// one CTA per (pair, b): lse, logit of the positive (column 0), top-1 hit
__global__ void nce_rowstat_kernel(const float* __restrict__ logits, int B, int K1, float* lse, float* l0, float* hit) {
  __shared__ float red[32];
  const float* row = logits + (long)blockIdx.x * K1;
  float mx = -INFINITY;
  for (int k = threadIdx.x; k < K1; k += blockDim.x) mx = fmaxf(mx, row[k]);
  mx = block_max(mx, red);
  float s = 0.f;
  for (int k = threadIdx.x; k < K1; k += blockDim.x) s += __expf(row[k] - mx);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    lse[blockIdx.x] = mx + logf(s);
    l0[blockIdx.x] = row[0];
    hit[blockIdx.x] = (row[0] >= mx) ? 1.f : 0.f;
  }
}

// single CTA: masked means (contrast_trainer.py:212-253).  out: loss[6], acc[6]; coef [6][B] = sel/count
__global__ void nce_finish_kernel(const float* __restrict__ lse, const float* __restrict__ l0, const float* __restrict__ hit,
                                  const long long* use_depth, const long long* use_rgb, int B, float* loss, float* acc,
                                  float* coef) {
  __shared__ float red[32];
  __shared__ int s_any;
  // selection rule per pair
  int cnt_both = 0;
  if (threadIdx.x == 0) {
    int c = 0;
    for (int b = 0; b < B; ++b) {
      bool d = use_depth ? (use_depth[b] == 1) : true;
      bool r = use_rgb ? (use_rgb[b] == 1) : true;
      c += (d && r) ? 1 : 0;
    }
    s_any = c;
  }
  __syncthreads();
  cnt_both = s_any;
  for (int pair = 0; pair < 6; ++pair) {
    // which rows does this pair average over?
    //  use_rgb given : rows with both flags (all six pairs); if none -> pairs 0-3 zero, pairs 4-5 all rows
    //  use_depth only: pairs 0-3 rows with depth (none -> zero), pairs 4-5 all rows
    const bool masked_pair = (use_rgb != nullptr) ? (cnt_both > 0 || pair < 4) : (use_depth != nullptr && pair < 4);
    float ls = 0.f, hs = 0.f, cs = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
      bool sel = true;
      if (masked_pair) {
        bool d = use_depth ? (use_depth[b] == 1) : true;
        bool r = use_rgb ? (use_rgb[b] == 1) : true;
        sel = d && r;
      }
      if (sel) { ls += lse[pair * B + b] - l0[pair * B + b]; hs += hit[pair * B + b]; cs += 1.f; }
    }
    ls = block_sum(ls, red); hs = block_sum(hs, red); cs = block_sum(cs, red);
    const float inv = (cs > 0.f) ? 1.f / cs : 0.f;
    if (threadIdx.x == 0) { loss[pair] = ls * inv; acc[pair] = 100.f * hs * inv; }
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
      bool sel = true;
      if (masked_pair) {
        bool d = use_depth ? (use_depth[b] == 1) : true;
        bool r = use_rgb ? (use_rgb[b] == 1) : true;
        sel = d && r;
      }
      coef[pair * B + b] = sel ? inv : 0.f;
    }
    __syncthreads();
  }
}

// grid (ceil(K1/ROWS_PER_CTA), B); df [B][3*128] += (atomics over the K-slices)
__global__ void __launch_bounds__(NCE_THREADS, 2) nce_bwd_kernel(const NceArgs a, const float* __restrict__ logits,
                                                              const float* __restrict__ lse, const float* __restrict__ coef,
                                                              float gscale, float* df, long lddf) {
  __shared__ float4 sacc[NCE_THREADS / 32][3][4][8];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, l8 = lane & 7;
  // lse == nullptr: `logits` already holds d(loss)/d(logits) (generic autograd backward of the logits-returning API)
  const bool generic = (lse == nullptr);
  float cf[6], ls[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    cf[q] = (generic ? 1.f : coef[q * a.B + b]) * a.invT * gscale;
    ls[q] = generic ? 0.f : lse[q * a.B + b];
  }
  float4 acc[3][4];
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[m][j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int kbeg = blockIdx.x * ROWS_PER_CTA;
  const int kend = min(a.K1, kbeg + ROWS_PER_CTA);
  const long long* idx = a.idx + (long)b * a.K1;
  for (int kb = kbeg + warp * 4; kb < kend; kb += (NCE_THREADS / 32) * 4) {
    const int k = kb + sub;
    const bool valid = k < kend;
    const long long row = valid ? idx[k] : 0;
    float4 w[3][4];
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) w[m][j] = ldg_stream(a.bank[m] + row * D + 32 * j + 4 * l8);
    float s[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const float lg = valid ? logits[((long)q * a.B + b) * a.K1 + k] : (generic ? 0.f : -INFINITY);
      s[q] = !valid ? 0.f : (generic ? cf[q] * lg : cf[q] * (__expf(lg - ls[q]) - ((k == 0) ? 1.f : 0.f)));
    }
    // dx1 <- pairs 0 (w2), 4 (w3); dx2 <- pairs 1 (w1), 2 (w3); dx3 <- pairs 3 (w2), 5 (w1)
    axpy16(acc[0], s[0], w[1]); axpy16(acc[0], s[4], w[2]);
    axpy16(acc[1], s[1], w[0]); axpy16(acc[1], s[2], w[2]);
    axpy16(acc[2], s[3], w[1]); axpy16(acc[2], s[5], w[0]);
  }
  // fold the 4 row-subgroups of the warp (lanes l8, l8+8, l8+16, l8+24 hold the same columns)
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        acc[m][j].x += __shfl_xor_sync(0xffffffffu, acc[m][j].x, o);
        acc[m][j].y += __shfl_xor_sync(0xffffffffu, acc[m][j].y, o);
        acc[m][j].z += __shfl_xor_sync(0xffffffffu, acc[m][j].z, o);
        acc[m][j].w += __shfl_xor_sync(0xffffffffu, acc[m][j].w, o);
      }
      if (sub == 0) sacc[warp][m][j][l8] = acc[m][j];
    }
  __syncthreads();
  // 3*128 = 384 outputs; thread t < 96 owns one float4
  if (threadIdx.x < 96) {
    const int m = threadIdx.x / 32, r = threadIdx.x % 32, j = r / 8, l = r % 8;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int wv = 0; wv < NCE_THREADS / 32; ++wv) {
      const float4 u = sacc[wv][m][j][l];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    float* dst = df + (long)b * lddf + m * D + 32 * j + 4 * l;
    atomicAdd(dst + 0, t.x); atomicAdd(dst + 1, t.y); atomicAdd(dst + 2, t.z); atomicAdd(dst + 3, t.w);
  }
}

// one warp per update; duplicates of an index: all read the old row, the last one (largest i) writes
__global__ void bank_update_kernel(float* bank, const float* __restrict__ x, long ldx, const long long* __restrict__ y, int N,
                                   float m) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= N) return;
  const long long row = y[i];
  bool later = false;
  for (int j = i + 1 + lane; j < N; j += 32) later |= (y[j] == row);
  if (__any_sync(0xffffffffu, later)) return;
  float4 w = *reinterpret_cast<const float4*>(bank + row * D + 4 * lane);
  const float4 xv = *reinterpret_cast<const float4*>(x + (long)i * ldx + 4 * lane);
  w.x = w.x * m + xv.x * (1.f - m); w.y = w.y * m + xv.y * (1.f - m);
  w.z = w.z * m + xv.z * (1.f - m); w.w = w.w * m + xv.w * (1.f - m);
  float ss = warp_sum(w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  w.x *= inv; w.y *= inv; w.z *= inv; w.w *= inv;
  *reinterpret_cast<float4*>(bank + row * D + 4 * lane) = w;
}

}  // namespace

extern "C" {

// logits [6][B][K1]; x1/x2/x3 rows of stride ldx floats (16-byte aligned); idx [B][K1] int64 rows of the banks
int hcm_nce_logits(const float* bank1, const float* bank2, const float* bank3, const float* x1, const float* x2,
                   const float* x3, long ldx, const long long* idx, int B, int K1, int dim, float T, float* logits,
                   cudaStream_t stream) {
  HCM_CHECK_ARG(dim == D, "nce: feature dim %d unsupported (128 only)", dim);
  HCM_CHECK_ARG(bank1 && bank2 && bank3 && x1 && x2 && x3 && idx && logits && (ldx % 4) == 0, "nce_logits: bad args");
  NceArgs a;
  a.bank[0] = bank1; a.bank[1] = bank2; a.bank[2] = bank3;
  a.x[0] = x1; a.x[1] = x2; a.x[2] = x3; a.ldx = ldx; a.idx = idx; a.B = B; a.K1 = K1; a.invT = 1.f / T;
  dim3 grid(hcm_cdiv(K1, ROWS_PER_CTA), B);
  nce_logits_kernel<<<grid, NCE_THREADS, 0, stream>>>(a, logits);
  HCM_LAUNCH_CHECK("nce_logits");
  return HCM_OK;
}

// masked CE (target 0) + top-1 of the six logit sets.  lse/l0/hit/coef: [6][B] scratch owned by the caller.
int hcm_nce_loss(const float* logits, int B, int K1, const long long* use_depth, const long long* use_rgb, float* lse,
                 float* l0, float* hit, float* coef, float* loss6, float* acc6, cudaStream_t stream) {
  HCM_CHECK_ARG(logits && lse && l0 && hit && coef && loss6 && acc6, "nce_loss: null pointer");
  nce_rowstat_kernel<<<6 * B, 256, 0, stream>>>(logits, B, K1, lse, l0, hit);
  HCM_LAUNCH_CHECK("nce_rowstat");
  nce_finish_kernel<<<1, 256, 0, stream>>>(lse, l0, hit, use_depth, use_rgb, B, loss6, acc6, coef);
  HCM_LAUNCH_CHECK("nce_finish");
  return HCM_OK;
}

// df [B][lddf] (columns 0..383 = dx1|dx2|dx3) += gradient of gscale * sum of the six losses
int hcm_nce_bwd(const float* bank1, const float* bank2, const float* bank3, const float* x1, const float* x2,
                const float* x3, long ldx, const long long* idx, int B, int K1, int dim, float T, const float* logits,
                const float* lse, const float* coef, float gscale, float* df, long lddf, cudaStream_t stream) {
  HCM_CHECK_ARG(dim == D, "nce: feature dim %d unsupported (128 only)", dim);
  HCM_CHECK_ARG(bank1 && bank2 && bank3 && idx && logits && df && ((lse == nullptr) == (coef == nullptr)), "nce_bwd: null pointer");
  NceArgs a;
  a.bank[0] = bank1; a.bank[1] = bank2; a.bank[2] = bank3;
  a.x[0] = x1; a.x[1] = x2; a.x[2] = x3; a.ldx = ldx; a.idx = idx; a.B = B; a.K1 = K1; a.invT = 1.f / T;
  dim3 grid(hcm_cdiv(K1, ROWS_PER_CTA), B);
  nce_bwd_kernel<<<grid, NCE_THREADS, 0, stream>>>(a, logits, lse, coef, gscale, df, lddf);
  HCM_LAUNCH_CHECK("nce_bwd");
  return HCM_OK;
}

// bank[y[i]] <- normalize(m*bank[y[i]] + (1-m)*x[i])   (mem_bank.py:15-28)
int hcm_bank_update(float* bank, const float* x, long ldx, const long long* y, int N, int dim, float m,
                    cudaStream_t stream) {
  HCM_CHECK_ARG(dim == D, "bank_update: feature dim %d unsupported (128 only)", dim);
  HCM_CHECK_ARG(bank && x && y && (ldx % 4) == 0, "bank_update: bad args");
  bank_update_kernel<<<hcm_cdiv(N, 4), 128, 0, stream>>>(bank, x, ldx, y, N, m);
  HCM_LAUNCH_CHECK("bank_update");
  return HCM_OK;
}

}  // extern "C"
