// Segmentation fine-tuning head: the two pieces of pycontrast/learning/segment_trainer.py:722-745 and
// pycontrast/networks/fcn.py:35-111 that the pre-train kernels do not already cover.
//   l2norm_max   seg input  max(normalize(linear_merge1, dim=1), normalize(linear_merge2, dim=1))   (segment_trainer.py:724-729; one map
//                only for supervise_type 1 / 2, :732-741), channels-last [P,128], one warp per pixel; backward routes the gradient to
//                the larger branch (ties -> the first, as torch.max over the stacked dim) and through the L2 normalisation.
//   seg_ce       nn.CrossEntropyLoss(ignore_index, weight=class_weights) (main_segmentor.py:76-79) on the x4-upsampled logits
//                (fcn.py:108-110; the upsampling itself is hcm_fuse_sum, its adjoint hcm_upsample_adjoint) and eval_seg_aacc
//                (segment_trainer.py:375-379: argmax == label over ALL pixels, ignored ones included).
// The 1x1 convolutions, the BatchNorm and the classifier GEMM of FCNHead run on hcm_tc_conv / hcm_tc_wgrad / hcm_bn_* / hcm_gemm.
#include "common.cuh"

namespace {

constexpr float L2_EPS = 1e-12f;      // F.normalize: x / max(||x||, eps)

// one warp per pixel, C = 128: lane l holds channels 4l .. 4l+3
__global__ void l2norm_max_fwd_kernel(const float* __restrict__ m1, const float* __restrict__ m2, long P, float* __restrict__ out,
                                      float* __restrict__ inv1, float* __restrict__ inv2) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  for (long p = warp; p < P; p += nw) {
    const float4 a = *reinterpret_cast<const float4*>(m1 + p * 128 + 4 * lane);
    float s1 = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    float s2 = 0.f;
    if (m2) {
      b = *reinterpret_cast<const float4*>(m2 + p * 128 + 4 * lane);
      s2 = b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    const float i1 = 1.f / fmaxf(sqrtf(s1), L2_EPS), i2 = 1.f / fmaxf(sqrtf(s2), L2_EPS);
    float4 o4 = make_float4(a.x * i1, a.y * i1, a.z * i1, a.w * i1);
    if (m2) {
      o4.x = fmaxf(o4.x, b.x * i2); o4.y = fmaxf(o4.y, b.y * i2); o4.z = fmaxf(o4.z, b.z * i2); o4.w = fmaxf(o4.w, b.w * i2);
    }
    *reinterpret_cast<float4*>(out + p * 128 + 4 * lane) = o4;
    if (lane == 0) { inv1[p] = i1; if (m2) inv2[p] = i2; }
  }
}

// d m = inv * (g' - n * <g', n>) with n = m * inv and g' = the part of g routed to this branch (inactive when the norm hit eps: the
// reference's clamp passes no gradient through the norm there; with real features that never happens)
__global__ void l2norm_max_bwd_kernel(const float* __restrict__ g, const float* __restrict__ m1, const float* __restrict__ m2,
                                      const float* __restrict__ inv1, const float* __restrict__ inv2, long P, float gscale,
                                      float* d1, float* d2, int accumulate) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  for (long p = warp; p < P; p += nw) {
    const long o = p * 128 + 4 * lane;
    const float4 gg = *reinterpret_cast<const float4*>(g + o);
    const float4 a = *reinterpret_cast<const float4*>(m1 + o);
    const float i1 = inv1[p];
    float n1[4] = {a.x * i1, a.y * i1, a.z * i1, a.w * i1};
    float g1[4] = {gg.x * gscale, gg.y * gscale, gg.z * gscale, gg.w * gscale};
    float g2[4] = {0.f, 0.f, 0.f, 0.f}, n2[4] = {0.f, 0.f, 0.f, 0.f};
    float i2 = 0.f;
    if (m2) {
      const float4 b = *reinterpret_cast<const float4*>(m2 + o);
      i2 = inv2[p];
      n2[0] = b.x * i2; n2[1] = b.y * i2; n2[2] = b.z * i2; n2[3] = b.w * i2;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (n2[k] > n1[k]) { g2[k] = g1[k]; g1[k] = 0.f; }
      }
    }
    float t1 = g1[0] * n1[0] + g1[1] * n1[1] + g1[2] * n1[2] + g1[3] * n1[3];
    float t2 = g2[0] * n2[0] + g2[1] * n2[1] + g2[2] * n2[2] + g2[3] * n2[3];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) { t1 += __shfl_xor_sync(0xffffffffu, t1, s); t2 += __shfl_xor_sync(0xffffffffu, t2, s); }
    float4 r1 = make_float4(i1 * (g1[0] - n1[0] * t1), i1 * (g1[1] - n1[1] * t1), i1 * (g1[2] - n1[2] * t1), i1 * (g1[3] - n1[3] * t1));
    if (accumulate) { const float4 old = *reinterpret_cast<const float4*>(d1 + o); r1.x += old.x; r1.y += old.y; r1.z += old.z; r1.w += old.w; }
    *reinterpret_cast<float4*>(d1 + o) = r1;
    if (m2) {
      float4 r2 = make_float4(i2 * (g2[0] - n2[0] * t2), i2 * (g2[1] - n2[1] * t2), i2 * (g2[2] - n2[2] * t2), i2 * (g2[3] - n2[3] * t2));
      if (accumulate) { const float4 old = *reinterpret_cast<const float4*>(d2 + o); r2.x += old.x; r2.y += old.y; r2.z += old.z; r2.w += old.w; }
      *reinterpret_cast<float4*>(d2 + o) = r2;
    }
  }
}

// one thread per pixel: log-softmax over Cn <= 64 classes (two passes over the L1-resident row), weighted NLL, top-1
// acc[0] += w[y] * nll, acc[1] += w[y], acc[2] += [argmax == y]   (fp64 atomics: one per warp)
__global__ void seg_ce_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ label, const float* __restrict__ cw,
                                  long P, int Cn, int ignore_index, double* acc) {
  double s_l = 0.0, s_w = 0.0, s_c = 0.0;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
    const float* row = logits + p * Cn;
    float mx = row[0];
    int am = 0;
    for (int c = 1; c < Cn; ++c) { const float v = row[c]; if (v > mx) { mx = v; am = c; } }      // first maximum, as torch.argmax
    const long long y = label[p];
    if ((long long)am == y) s_c += 1.0;
    if (y == (long long)ignore_index || y < 0 || y >= Cn) continue;
    float se = 0.f;
    for (int c = 0; c < Cn; ++c) se += __expf(row[c] - mx);
    const float w = cw ? cw[y] : 1.f;
    s_l += (double)(w * (mx + __logf(se) - row[y]));
    s_w += (double)w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_l += __shfl_xor_sync(0xffffffffu, s_l, o); s_w += __shfl_xor_sync(0xffffffffu, s_w, o); s_c += __shfl_xor_sync(0xffffffffu, s_c, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (s_l != 0.0) atomicAdd(acc + 0, s_l);
    if (s_w != 0.0) atomicAdd(acc + 1, s_w);
    if (s_c != 0.0) atomicAdd(acc + 2, s_c);
  }
}

__global__ void seg_ce_finish_kernel(const double* acc, long P, float* out) {
  out[0] = acc[1] > 0.0 ? (float)(acc[0] / acc[1]) : 0.f;        // all pixels ignored: 0 (torch: NaN)
  out[1] = (float)(acc[2] / (double)P);
}

// dlogits[p][c] = gscale * w[y] / sum_w * (softmax_c - [c == y]);  ignored pixels: 0
__global__ void seg_ce_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ label, const float* __restrict__ cw,
                                  long P, int Cn, int ignore_index, const double* __restrict__ acc, float gscale, float* __restrict__ dl) {
  const double sw = acc[1];
  const float k = sw > 0.0 ? gscale / (float)sw : 0.f;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
    const float* row = logits + p * Cn;
    float* o = dl + p * Cn;
    const long long y = label[p];
    if (y == (long long)ignore_index || y < 0 || y >= Cn) {
      for (int c = 0; c < Cn; ++c) o[c] = 0.f;
      continue;
    }
    float mx = row[0];
    for (int c = 1; c < Cn; ++c) mx = fmaxf(mx, row[c]);
    float se = 0.f;
    for (int c = 0; c < Cn; ++c) se += __expf(row[c] - mx);
    const float f = k * (cw ? cw[y] : 1.f), inv = 1.f / se;
    for (int c = 0; c < Cn; ++c) o[c] = f * (__expf(row[c] - mx) * inv - (c == (int)y ? 1.f : 0.f));
  }
}

inline int grid_for(long items, int per_block) {
  long g = (items + per_block - 1) / per_block;
  if (g > 148L * 16) g = 148L * 16;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" {

// out [P,128] = max(m1 * inv1, m2 * inv2) per channel (m2 / inv2 null: one map); inv{1,2} [P] = 1 / max(||m||, 1e-12) are kept for the backward
int hcm_l2norm_max_fwd(const float* m1, const float* m2, long P, int C, float* out, float* inv1, float* inv2, cudaStream_t stream) {
  HCM_CHECK_ARG(m1 && out && inv1 && (!m2 || inv2) && C == 128 && P >= 1, "l2norm_max_fwd: bad args (C=%d, 128 expected)", C);
  l2norm_max_fwd_kernel<<<grid_for(P, 8), 256, 0, stream>>>(m1, m2, P, out, inv1, inv2);
  HCM_LAUNCH_CHECK("l2norm_max_fwd");
  return HCM_OK;
}

int hcm_l2norm_max_bwd(const float* dout, const float* m1, const float* m2, const float* inv1, const float* inv2, long P, int C,
                       float gscale, float* d1, float* d2, int accumulate, cudaStream_t stream) {
  HCM_CHECK_ARG(dout && m1 && inv1 && d1 && (!m2 || (inv2 && d2)) && C == 128 && P >= 1, "l2norm_max_bwd: bad args (C=%d)", C);
  l2norm_max_bwd_kernel<<<grid_for(P, 8), 256, 0, stream>>>(dout, m1, m2, inv1, inv2, P, gscale, d1, d2, accumulate);
  HCM_LAUNCH_CHECK("l2norm_max_bwd");
  return HCM_OK;
}

// logits [P,Cn] channels-last (already upsampled), label [P] int64; acc [4] fp64 scratch (zeroed here); out [2] = (loss, aAcc)
int hcm_seg_ce_fwd(const float* logits, const long long* label, const float* class_weight, long P, int Cn, int ignore_index,
                   double* acc, float* out, cudaStream_t stream) {
  HCM_CHECK_ARG(logits && label && acc && out && Cn >= 1 && Cn <= 64 && P >= 1, "seg_ce_fwd: bad args (Cn=%d)", Cn);
  cudaError_t e = cudaMemsetAsync(acc, 0, 4 * sizeof(double), stream);
  if (e != cudaSuccess) { hcm_set_error("seg_ce_fwd: memset: %s", cudaGetErrorString(e)); return HCM_ERR_CUDA; }
  seg_ce_fwd_kernel<<<grid_for(P, 256), 256, 0, stream>>>(logits, label, class_weight, P, Cn, ignore_index, acc);
  HCM_LAUNCH_CHECK("seg_ce_fwd");
  seg_ce_finish_kernel<<<1, 1, 0, stream>>>(acc, P, out);
  HCM_LAUNCH_CHECK("seg_ce_fwd (finish)");
  return HCM_OK;
}

int hcm_seg_ce_bwd(const float* logits, const long long* label, const float* class_weight, long P, int Cn, int ignore_index,
                   const double* acc, float gscale, float* dlogits, cudaStream_t stream) {
  HCM_CHECK_ARG(logits && label && acc && dlogits && Cn >= 1 && Cn <= 64 && P >= 1, "seg_ce_bwd: bad args (Cn=%d)", Cn);
  seg_ce_bwd_kernel<<<grid_for(P, 256), 256, 0, stream>>>(logits, label, class_weight, P, Cn, ignore_index, acc, gscale, dlogits);
  HCM_LAUNCH_CHECK("seg_ce_bwd");
  return HCM_OK;
}

}  // extern "C"
