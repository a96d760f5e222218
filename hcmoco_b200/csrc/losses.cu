// Dense intra-sample, sparse joint<->pixel and cross-subject joint (SCL) objectives
// (learning/contrast_trainer.py:642-723, 744-828, 830-892), forward and backward.
//
// Shared front end: gather C=128-channel pixels from the channels-last projection maps (512 B
// contiguous per pixel — the NHWC layout turns the reference's stride-h*w torch.gather into one
// coalesced row read per warp), L2-normalise, keep 1/norm for the backward.  The S x S (dense),
// J x J (sparse) and 2BJ x 2BJ (SCL) similarity matrices come from the batched GEMM in igemm.cu;
// the kernels here do the masked / soft-target log-softmax statistics, the in-place logit gradients
// and the scatter back into the map gradient.
#include "common.cuh"

namespace {

constexpr int D = 128;

// ---------------------------------------------------------------- gather + L2 normalise
// out[r][0:128] = src_row / max(||src_row||, 1e-12);  src_row = src[(b*HW + pix[r]) * 128] (pix != null)
//                                                      or src[r * lds]                      (pix == null)
__global__ void gather_l2norm_kernel(const float* __restrict__ src, long lds, const long long* __restrict__ pix, long HW,
                                     int rows_per_b, long nrows, float* out, long ldo, float* inv_norm) {
  const long r = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  const float* s = pix ? src + ((r / rows_per_b) * HW + pix[r]) * D : src + r * lds;
  float4 v = *reinterpret_cast<const float4*>(s + 4 * lane);
  const float ss = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
  *reinterpret_cast<float4*>(out + r * ldo + 4 * lane) = v;
  if (lane == 0 && inv_norm) inv_norm[r] = inv;
}

// dsrc_row += (dout - out*<dout,out>) * inv_norm   (atomic when gathered: pixels may repeat)
__global__ void gather_l2norm_bwd_kernel(const float* __restrict__ dout, long lddo, const float* __restrict__ out, long ldo,
                                         const float* __restrict__ inv_norm, const long long* __restrict__ pix, long HW,
                                         int rows_per_b, long nrows, float* dsrc, long lds, int accumulate) {
  const long r = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  const float4 g = *reinterpret_cast<const float4*>(dout + r * lddo + 4 * lane);
  const float4 o = *reinterpret_cast<const float4*>(out + r * ldo + 4 * lane);
  const float dot = warp_sum(g.x * o.x + g.y * o.y + g.z * o.z + g.w * o.w);
  const float inv = inv_norm[r];
  float4 d;
  d.x = (g.x - o.x * dot) * inv; d.y = (g.y - o.y * dot) * inv;
  d.z = (g.z - o.z * dot) * inv; d.w = (g.w - o.w * dot) * inv;
  if (pix) {
    float* p = dsrc + ((r / rows_per_b) * HW + pix[r]) * D + 4 * lane;
    atomicAdd(p + 0, d.x); atomicAdd(p + 1, d.y); atomicAdd(p + 2, d.z); atomicAdd(p + 3, d.w);
  } else {
    float4* p = reinterpret_cast<float4*>(dsrc + r * lds + 4 * lane);
    if (accumulate) { float4 t = *p; d.x += t.x; d.y += t.y; d.z += t.z; d.w += t.w; }
    *p = d;
  }
}

// ---------------------------------------------------------------- pixel index helpers
// clamp(floor(joint/4), 0, h-1) -> y*h + x   (contrast_trainer.py:754-760)
__global__ void joint_pixel_kernel(const float* __restrict__ joints_yx, long n, int h, long long* pix) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long y = (long long)floorf(joints_yx[2 * i] / 4.f), x = (long long)floorf(joints_yx[2 * i + 1] / 4.f);
  y = min(max(y, 0LL), (long long)h - 1);
  x = min(max(x, 0LL), (long long)h - 1);
  pix[i] = y * h + x;
}

// kept[b] = 1 if the nearest-resized depth mask (every `step`-th pixel) is non-empty (contrast_trainer.py:671-680)
__global__ void dense_kept_kernel(const float* __restrict__ mask, int R, int h, int step, float* kept) {
  __shared__ float red[32];
  const float* m = mask + (long)blockIdx.x * R * R;
  float a = 0.f;
  for (int e = threadIdx.x; e < h * h; e += blockDim.x) a += (m[(long)(e / h) * step * R + (e % h) * step] != 0.f) ? 1.f : 0.f;
  a = block_sum(a, red);
  if (threadIdx.x == 0) kept[blockIdx.x] = (a > 0.f) ? 1.f : 0.f;
}

// ---------------------------------------------------------------- dense soft-target InfoNCE statistics
// L [B][S][S] = <d_i, a_j>/T.  One thread per (b, j):
//   dir 0 (rgb2depth): statistics of column j over i      dir 1 (depth2rgb): statistics of row j over i
// stat [B][2][S][4] = (lse, Z, sum_i e^{-dist}*logit, first-argmax==j)
__global__ void dense_stats_kernel(const float* __restrict__ L, const long long* __restrict__ pix, int B, int S, int h,
                                   float* __restrict__ stat) {
  const int b = blockIdx.z, dir = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= S) return;
  const float* Lb = L + (long)b * S * S;
  const long long* pb = pix + (long)b * S;
  const float yj = (float)(pb[j] / h), xj = (float)(pb[j] % h);
  float mx = -INFINITY, se = 0.f, Z = 0.f, wl = 0.f;
  int am = 0;
  for (int i = 0; i < S; ++i) {
    const float l = dir == 0 ? Lb[(long)i * S + j] : Lb[(long)j * S + i];
    const float dy = (float)(pb[i] / h) - yj, dx = (float)(pb[i] % h) - xj;
    const float w = __expf(-sqrtf(dy * dy + dx * dx));
    Z += w; wl = fmaf(w, l, wl);
    if (l > mx) { se = se * __expf(mx - l) + 1.f; mx = l; am = i; }
    else se += __expf(l - mx);
  }
  float* o = stat + (((long)b * 2 + dir) * S + j) * 4;
  o[0] = mx + logf(se); o[1] = Z; o[2] = wl; o[3] = (am == j) ? 1.f : 0.f;
}

// single CTA: out[0..1] = losses (r2d, d2r), out[2..3] = accuracies, out[4] = B' (kept samples)
__global__ void dense_finish_kernel(const float* __restrict__ stat, const float* __restrict__ kept, const long long* use_depth,
                                    int B, int S, float* out) {
  __shared__ float red[32];
  float l0 = 0.f, l1 = 0.f, a0 = 0.f, a1 = 0.f, nk = 0.f, nd = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) { nk += kept[b]; nd += use_depth ? (use_depth[b] != 0 ? 1.f : 0.f) : 1.f; }
  for (long e = threadIdx.x; e < (long)B * S; e += blockDim.x) {
    const int b = (int)(e / S), j = (int)(e % S);
    if (kept[b] == 0.f) continue;
    const float* s0 = stat + (((long)b * 2 + 0) * S + j) * 4;
    const float* s1 = stat + (((long)b * 2 + 1) * S + j) * 4;
    l0 += s0[0] - s0[2] / s0[1]; a0 += s0[3];
    l1 += s1[0] - s1[2] / s1[1]; a1 += s1[3];
  }
  l0 = block_sum(l0, red); l1 = block_sum(l1, red); a0 = block_sum(a0, red); a1 = block_sum(a1, red);
  nk = block_sum(nk, red); nd = block_sum(nd, red);
  if (threadIdx.x == 0) {
    const bool on = (nd > 0.f) && (nk > 0.f);   // use_depth.sum()==0 -> exact zeros (contrast_trainer.py:663-665)
    const float inv = on ? 1.f / (nk * (float)S) : 0.f;
    out[0] = l0 * inv; out[1] = l1 * inv; out[2] = a0 * inv; out[3] = a1 * inv; out[4] = on ? nk : 0.f;
  }
}

// in place: L[b][r][c] <- gscale/(B'*S) * ( e^{L-collse_c} + e^{L-rowlse_r} - e^{-dist_rc} * (1/Z_c + 1/Z_r) )
__global__ void dense_grad_kernel(float* L, const long long* __restrict__ pix, const float* __restrict__ stat,
                                  const float* __restrict__ kept, const float* __restrict__ fin, int B, int S, int h,
                                  float gscale) {
  const long total = (long)B * S * S;
  const long stride = (long)gridDim.x * blockDim.x;
  const float nk = fin[4];
  const float coef = (nk > 0.f) ? gscale / (nk * (float)S) : 0.f;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c = (int)(e % S);
    const int r = (int)((e / S) % S);
    const int b = (int)(e / ((long)S * S));
    float g = 0.f;
    if (kept[b] != 0.f && coef != 0.f) {
      const float* sc = stat + (((long)b * 2 + 0) * S + c) * 4;
      const float* sr = stat + (((long)b * 2 + 1) * S + r) * 4;
      const long long pr = pix[(long)b * S + r], pc = pix[(long)b * S + c];
      const float dy = (float)(pr / h) - (float)(pc / h), dx = (float)(pr % h) - (float)(pc % h);
      const float w = __expf(-sqrtf(dy * dy + dx * dx));
      const float l = L[e];
      g = coef * (__expf(l - sc[0]) + __expf(l - sr[0]) - w * (1.f / sc[1] + 1.f / sr[1]));
    }
    L[e] = g;
  }
}

// ---------------------------------------------------------------- sparse joint <-> pixel CE
// Lr, Ld [B][J][J]: logits[b][k][j] = <skel_k, pixfeat_j>/T.  One warp per (b, which in {rgb, depth}).
// CE over k with target j; ignored where joints_vis==0 (and use_depth==0 for depth).
// rs [B][2][3] = (sum of valid CE, #valid, hits)
__global__ void joint_stats_kernel(const float* __restrict__ Lr, const float* __restrict__ Ld, const int* __restrict__ vis,
                                   const long long* use_depth, int B, int J, float* rs, float* lse_out) {
  const int b = blockIdx.x, which = threadIdx.x >> 5, j = threadIdx.x & 31;
  const float* Lb = (which == 0 ? Lr : Ld) + (long)b * J * J;
  float loss = 0.f, cnt = 0.f, hit = 0.f;
  if (j < J) {
    float mx = -INFINITY;
    int am = 0;
    for (int k = 0; k < J; ++k) { const float l = Lb[k * J + j]; if (l > mx) { mx = l; am = k; } }
    float se = 0.f;
    for (int k = 0; k < J; ++k) se += __expf(Lb[k * J + j] - mx);
    const float lse = mx + logf(se);
    lse_out[((long)b * 2 + which) * J + j] = lse;
    bool valid = vis[(long)b * J + j] != 0;
    if (which == 1 && use_depth) valid = valid && (use_depth[b] != 0);
    if (valid) { loss = lse - Lb[j * J + j]; cnt = 1.f; hit = (am == j) ? 1.f : 0.f; }
  }
  loss = warp_sum(loss); cnt = warp_sum(cnt); hit = warp_sum(hit);
  if (j == 0) { float* o = rs + ((long)b * 2 + which) * 3; o[0] = loss; o[1] = cnt; o[2] = hit; }
}

// single CTA: out[0..1] losses, out[2..3] accuracies, out[4..5] valid-target counts
__global__ void joint_finish_kernel(const float* __restrict__ rs, int B, float* out) {
  __shared__ float red[32];
  for (int which = 0; which < 2; ++which) {
    float ls = 0.f, cn = 0.f, ac = 0.f, ns = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
      const float* o = rs + ((long)b * 2 + which) * 3;
      ls += o[0]; cn += o[1];
      if (o[1] > 0.f) { ac += o[2] / o[1]; ns += 1.f; }
    }
    ls = block_sum(ls, red); cn = block_sum(cn, red); ac = block_sum(ac, red); ns = block_sum(ns, red);
    if (threadIdx.x == 0) {
      out[which] = (cn > 0.f) ? ls / cn : 0.f;          // reference gives NaN when every target is ignored
      out[2 + which] = (ns > 0.f) ? ac / ns : 0.f;
      out[4 + which] = cn;
    }
    __syncthreads();
  }
}

// in place: L[b][k][j] <- gscale/count * (softmax_k - [k==j]) for valid j, else 0
__global__ void joint_grad_kernel(float* Lr, float* Ld, const int* __restrict__ vis, const long long* use_depth,
                                  const float* __restrict__ lse, const float* __restrict__ fin, int B, int J, float gscale) {
  const long total = (long)B * 2 * J * J;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int j = (int)(e % J), k = (int)((e / J) % J);
    const int which = (int)((e / ((long)J * J)) % 2);
    const int b = (int)(e / ((long)2 * J * J));
    float* Lb = (which == 0 ? Lr : Ld) + ((long)b * J + k) * J + j;
    bool valid = vis[(long)b * J + j] != 0;
    if (which == 1 && use_depth) valid = valid && (use_depth[b] != 0);
    const float cn = fin[4 + which];
    float g = 0.f;
    if (valid && cn > 0.f) g = gscale / cn * (__expf(*Lb - lse[((long)b * 2 + which) * J + j]) - ((k == j) ? 1.f : 0.f));
    *Lb = g;
  }
}

// ---------------------------------------------------------------- cross-subject joint SCL
// Z [N][N], N = 2*B*J; rows [0,BJ) rgb, [BJ,2BJ) depth.  One warp per row.
__device__ __forceinline__ bool scl_off(int r, int B, int J, const long long* use_rgb, const long long* use_depth) {
  const int half = r / (B * J), b = (r % (B * J)) / J;
  if (half == 0) return use_rgb ? (use_rgb[b] == 0) : false;
  return use_depth ? (use_depth[b] == 0) : false;
}
// rowstat [N][3] = (lse, npos, sum_pos Z)
__global__ void scl_stats_kernel(const float* __restrict__ Z, int B, int J, const long long* use_rgb,
                                 const long long* use_depth, float* rowstat) {
  const int N = 2 * B * J;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= N) return;
  const float* z = Z + (long)r * N;
  float mx = -INFINITY;
  for (int c = lane; c < N; c += 32) mx = fmaxf(mx, z[c]);
  mx = warp_max(mx);
  float se = 0.f, np = 0.f, ps = 0.f;
  const bool roff = scl_off(r, B, J, use_rgb, use_depth);
  for (int c = lane; c < N; c += 32) {
    const float v = z[c];
    se += __expf(v - mx);
    if (!roff && c != r && (c % J) == (r % J) && !scl_off(c, B, J, use_rgb, use_depth)) { np += 1.f; ps += v; }
  }
  se = warp_sum(se); np = warp_sum(np); ps = warp_sum(ps);
  if (lane == 0) { rowstat[r * 3 + 0] = mx + logf(se); rowstat[r * 3 + 1] = np; rowstat[r * 3 + 2] = ps; }
}
__global__ void scl_finish_kernel(const float* __restrict__ rowstat, int N, const long long* use_depth, int B, float* out) {
  __shared__ float red[32];
  float ls = 0.f, nd = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) nd += use_depth ? (use_depth[b] != 0 ? 1.f : 0.f) : 1.f;
  for (int r = threadIdx.x; r < N; r += blockDim.x) {
    const float lse = rowstat[r * 3], np = rowstat[r * 3 + 1], ps = rowstat[r * 3 + 2];
    ls += -(ps - np * lse) / fmaxf(np, 1.f);
  }
  ls = block_sum(ls, red); nd = block_sum(nd, red);
  if (threadIdx.x == 0) { out[0] = (nd > 0.f) ? ls / (float)N : 0.f; out[1] = (nd > 0.f) ? 1.f : 0.f; }
}
// in place: Z[r][c] <- gscale/N * (npos_r*softmax[r,c] - pos[r,c]) / max(npos_r,1)
__global__ void scl_grad_kernel(float* Z, int B, int J, const long long* use_rgb, const long long* use_depth,
                                const float* __restrict__ rowstat, const float* __restrict__ fin, float gscale) {
  const int N = 2 * B * J;
  const long total = (long)N * N;
  const float on = fin[1];
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int r = (int)(e / N), c = (int)(e % N);
    const float lse = rowstat[r * 3], np = rowstat[r * 3 + 1];
    float g = 0.f;
    if (on != 0.f && np > 0.f) {
      const bool pos = c != r && (c % J) == (r % J) && !scl_off(r, B, J, use_rgb, use_depth) &&
                       !scl_off(c, B, J, use_rgb, use_depth);
      g = gscale / (float)N * (np * __expf(Z[e] - lse) - (pos ? 1.f : 0.f)) / np;
    }
    Z[e] = g;
  }
}

// ---------------------------------------------------------------- small helpers
__global__ void colsum_finalize_kernel(const float* __restrict__ part, int nparts, int C, float* out, int accumulate) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0;
  for (int i = lane; i < nparts; i += 32) s += (double)part[((long)i * 2) * C + c];
  s = warp_sum_d(s);
  if (lane == 0) out[c] = accumulate ? out[c] + (float)s : (float)s;
}
__global__ void colsum_small_kernel(const float* __restrict__ x, int R, int C, long ld, float* out, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int r = 0; r < R; ++r) s += x[(long)r * ld + c];
  out[c] = accumulate ? out[c] + s : s;
}
// out = sum_i w[i] * in[i]   (total loss; tiny)
__global__ void weighted_sum_kernel(const float* const* ptrs, const float* w, int n, float* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += w[i] * (*ptrs[i]);
    *out = s;
  }
}

inline int ew_grid(long total) {
  long g = (total + 1023) / 1024;
  if (g > 148L * 16) g = 148L * 16;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" {

int hcm_gather_l2norm(const float* src, long lds, const long long* pix, long HW, int rows_per_b, long nrows, int dim,
                      float* out, long ldo, float* inv_norm, cudaStream_t stream) {
  HCM_CHECK_ARG(dim == D && src && out, "gather_l2norm: bad args (dim=%d)", dim);
  gather_l2norm_kernel<<<hcm_cdiv(nrows, 8), 256, 0, stream>>>(src, lds, pix, HW, rows_per_b, nrows, out, ldo, inv_norm);
  HCM_LAUNCH_CHECK("gather_l2norm");
  return HCM_OK;
}

int hcm_gather_l2norm_bwd(const float* dout, long lddo, const float* out, long ldo, const float* inv_norm,
                          const long long* pix, long HW, int rows_per_b, long nrows, int dim, float* dsrc, long lds,
                          int accumulate, cudaStream_t stream) {
  HCM_CHECK_ARG(dim == D && dout && out && inv_norm && dsrc, "gather_l2norm_bwd: bad args (dim=%d)", dim);
  gather_l2norm_bwd_kernel<<<hcm_cdiv(nrows, 8), 256, 0, stream>>>(dout, lddo, out, ldo, inv_norm, pix, HW, rows_per_b,
                                                                   nrows, dsrc, lds, accumulate);
  HCM_LAUNCH_CHECK("gather_l2norm_bwd");
  return HCM_OK;
}

int hcm_joint_pixel_index(const float* joints_yx, long n, int h, long long* pix, cudaStream_t stream) {
  HCM_CHECK_ARG(joints_yx && pix, "joint_pixel_index: null pointer");
  joint_pixel_kernel<<<hcm_cdiv(n, 256), 256, 0, stream>>>(joints_yx, n, h, pix);
  HCM_LAUNCH_CHECK("joint_pixel_index");
  return HCM_OK;
}

int hcm_dense_kept(const float* depth_mask, int B, int R, int h, float* kept, cudaStream_t stream) {
  HCM_CHECK_ARG(depth_mask && kept && h >= 1 && R % h == 0, "dense_kept: bad args");
  dense_kept_kernel<<<B, 256, 0, stream>>>(depth_mask, R, h, R / h, kept);
  HCM_LAUNCH_CHECK("dense_kept");
  return HCM_OK;
}

// stat [B][2][S][4] scratch; fin[5] = loss_r2d, loss_d2r, acc_r2d, acc_d2r, B'
int hcm_dense_stats(const float* L, const long long* pix, const float* kept, const long long* use_depth, int B, int S,
                    int h, float* stat, float* fin, cudaStream_t stream) {
  HCM_CHECK_ARG(L && pix && kept && stat && fin, "dense_stats: null pointer");
  dim3 grid(hcm_cdiv(S, 64), 2, B);
  dense_stats_kernel<<<grid, 64, 0, stream>>>(L, pix, B, S, h, stat);
  HCM_LAUNCH_CHECK("dense_stats");
  dense_finish_kernel<<<1, 256, 0, stream>>>(stat, kept, use_depth, B, S, fin);
  HCM_LAUNCH_CHECK("dense_finish");
  return HCM_OK;
}

// fin[5] from stat [B][2][S][4] (shared by the unfused statistics above and the fused kernel in dense_affinity.cu)
int hcm_dense_finish(const float* stat, const float* kept, const long long* use_depth, int B, int S, float* fin,
                     cudaStream_t stream) {
  HCM_CHECK_ARG(stat && kept && fin, "dense_finish: null pointer");
  dense_finish_kernel<<<1, 1024, 0, stream>>>(stat, kept, use_depth, B, S, fin);
  HCM_LAUNCH_CHECK("dense_finish");
  return HCM_OK;
}

int hcm_dense_grad(float* L, const long long* pix, const float* stat, const float* kept, const float* fin, int B, int S,
                   int h, float gscale, cudaStream_t stream) {
  HCM_CHECK_ARG(L && pix && stat && kept && fin, "dense_grad: null pointer");
  dense_grad_kernel<<<ew_grid((long)B * S * S), 256, 0, stream>>>(L, pix, stat, kept, fin, B, S, h, gscale);
  HCM_LAUNCH_CHECK("dense_grad");
  return HCM_OK;
}

// rs [B][2][3], lse [B][2][J] scratch; fin[6] = loss_rgb, loss_d, acc_rgb, acc_d, count_rgb, count_d
int hcm_joint_stats(const float* Lr, const float* Ld, const int* joints_vis, const long long* use_depth, int B, int J,
                    float* rs, float* lse, float* fin, cudaStream_t stream) {
  HCM_CHECK_ARG(Lr && Ld && joints_vis && rs && lse && fin && J <= 32, "joint_stats: bad args (J=%d)", J);
  joint_stats_kernel<<<B, 64, 0, stream>>>(Lr, Ld, joints_vis, use_depth, B, J, rs, lse);
  HCM_LAUNCH_CHECK("joint_stats");
  joint_finish_kernel<<<1, 256, 0, stream>>>(rs, B, fin);
  HCM_LAUNCH_CHECK("joint_finish");
  return HCM_OK;
}

int hcm_joint_grad(float* Lr, float* Ld, const int* joints_vis, const long long* use_depth, const float* lse,
                   const float* fin, int B, int J, float gscale, cudaStream_t stream) {
  HCM_CHECK_ARG(Lr && Ld && joints_vis && lse && fin, "joint_grad: null pointer");
  joint_grad_kernel<<<ew_grid((long)B * 2 * J * J), 256, 0, stream>>>(Lr, Ld, joints_vis, use_depth, lse, fin, B, J, gscale);
  HCM_LAUNCH_CHECK("joint_grad");
  return HCM_OK;
}

// rowstat [N][3] scratch; fin[2] = loss, enabled flag
int hcm_scl_stats(const float* Z, int B, int J, const long long* use_rgb, const long long* use_depth, float* rowstat,
                  float* fin, cudaStream_t stream) {
  HCM_CHECK_ARG(Z && rowstat && fin, "scl_stats: null pointer");
  const int N = 2 * B * J;
  scl_stats_kernel<<<hcm_cdiv(N, 8), 256, 0, stream>>>(Z, B, J, use_rgb, use_depth, rowstat);
  HCM_LAUNCH_CHECK("scl_stats");
  scl_finish_kernel<<<1, 256, 0, stream>>>(rowstat, N, use_depth, B, fin);
  HCM_LAUNCH_CHECK("scl_finish");
  return HCM_OK;
}

int hcm_scl_grad(float* Z, int B, int J, const long long* use_rgb, const long long* use_depth, const float* rowstat,
                 const float* fin, float gscale, cudaStream_t stream) {
  HCM_CHECK_ARG(Z && rowstat && fin, "scl_grad: null pointer");
  const int N = 2 * B * J;
  scl_grad_kernel<<<ew_grid((long)N * N), 256, 0, stream>>>(Z, B, J, use_rgb, use_depth, rowstat, fin, gscale);
  HCM_LAUNCH_CHECK("scl_grad");
  return HCM_OK;
}

// out[c] (+)= sum of slot-0 partial rows written by hcm_bn_stats
int hcm_colsum_finalize(const float* part, int nparts, int C, float* out, int accumulate, cudaStream_t stream) {
  HCM_CHECK_ARG(part && out, "colsum_finalize: null pointer");
  colsum_finalize_kernel<<<hcm_cdiv(C, 4), 128, 0, stream>>>(part, nparts, C, out, accumulate);
  HCM_LAUNCH_CHECK("colsum_finalize");
  return HCM_OK;
}

// out[c] (+)= sum_r x[r*ld + c]   (R small)
int hcm_colsum_small(const float* x, int R, int C, long ld, float* out, int accumulate, cudaStream_t stream) {
  HCM_CHECK_ARG(x && out, "colsum_small: null pointer");
  colsum_small_kernel<<<hcm_cdiv(C, 128), 128, 0, stream>>>(x, R, C, ld, out, accumulate);
  HCM_LAUNCH_CHECK("colsum_small");
  return HCM_OK;
}

}  // extern "C"
