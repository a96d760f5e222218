// Layout / resampling kernels around the HRNet cross-resolution fuse and the model head:
//   * NCHW -> NHWC split of the [B,6,R,R] input (build_backbone.py:261 torch.split)
//   * fuse_sum: out = act( sum_t  affine_t( bilinear_up_{2^k}(term_t) ) + bias ) — the HR-module fuse
//     (official_hrnet.py:232-247, F.interpolate(mode='bilinear', align_corners=False)) and the 4-branch
//     merge feeding the 1x1 projection (build_backbone.py:247-254, 291-294).  BN's per-channel affine
//     commutes with bilinear interpolation (weights sum to 1), so the low-resolution raw conv output is
//     interpolated and the affine applied once at the high resolution.
//   * upsample_adjoint: exact transpose of the bilinear upsampling, as a deterministic gather
//   * global average pool (+ backward) (build_backbone.py:267-278)
#include "common.cuh"

namespace {

struct FuseTerm {
  const float* ptr;
  const float* scale;  // per channel, null -> 1
  const float* shift;  // per channel, null -> 0
  int log2f;           // 0: same resolution; k: source is (H>>k) x (W>>k), bilinear up by 2^k
};
struct FuseParams {
  FuseTerm t[4];
  int nterms;
  const float* bias;
  int relu;
  float* out;
  int B, H, W, C;
};

// PyTorch upsample_bilinear2d, align_corners=False, scale = in/out = 1/f
__device__ __forceinline__ void src_index(int dst, int f, int in_size, int& i0, int& i1, float& l1) {
  float s = ((float)dst + 0.5f) * (1.f / (float)f) - 0.5f;      // f is a power of two: the reciprocal and the product are exact
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
}

// One thread per (pixel, V channels): 32-bit index arithmetic with multiply-high divisions (the first version decoded every ELEMENT
// with six emulated 64-bit divisions: 52 us per launch in the step for a 19 MB output), the bilinear source rows / weights computed
// once per pixel and term, 8-byte loads and stores when C is even.  The per-element arithmetic (and its order) is unchanged.
template <int V>
__global__ void fuse_sum_kernel(const FuseParams p, FastDiv fd_cv, FastDiv fd_w, FastDiv fd_h, unsigned items) {
  const unsigned cv = (unsigned)p.C / V;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += stride) {
    const unsigned pix = fdiv(it, fd_cv), c = (it - pix * cv) * V;
    const unsigned t2 = fdiv(pix, fd_w), w = pix - t2 * (unsigned)p.W;
    const unsigned b = fdiv(t2, fd_h), h = t2 - b * (unsigned)p.H;
    const size_t e = (size_t)pix * p.C + c;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = p.bias ? p.bias[c + i] : 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (t >= p.nterms) break;
      const FuseTerm& T = p.t[t];
      float v[V];
      if (T.log2f == 0) {
        if (V == 2) { const float2 a = *reinterpret_cast<const float2*>(T.ptr + e); v[0] = a.x; v[V - 1] = a.y; }
        else v[0] = T.ptr[e];
      } else {
        const int f = 1 << T.log2f;
        const int Hs = p.H >> T.log2f, Ws = p.W >> T.log2f;
        int y0, y1, x0, x1;
        float ly, lx;
        src_index((int)h, f, Hs, y0, y1, ly);
        src_index((int)w, f, Ws, x0, x1, lx);
        const float* base = T.ptr + (size_t)b * Hs * Ws * p.C + c;
        const float* q00 = base + ((size_t)y0 * Ws + x0) * p.C;
        const float* q01 = base + ((size_t)y0 * Ws + x1) * p.C;
        const float* q10 = base + ((size_t)y1 * Ws + x0) * p.C;
        const float* q11 = base + ((size_t)y1 * Ws + x1) * p.C;
        float v00[V], v01[V], v10[V], v11[V];
        if (V == 2) {
          const float2 a = *reinterpret_cast<const float2*>(q00), bq = *reinterpret_cast<const float2*>(q01);
          const float2 cq = *reinterpret_cast<const float2*>(q10), d = *reinterpret_cast<const float2*>(q11);
          v00[0] = a.x; v00[V - 1] = a.y; v01[0] = bq.x; v01[V - 1] = bq.y;
          v10[0] = cq.x; v10[V - 1] = cq.y; v11[0] = d.x; v11[V - 1] = d.y;
        } else { v00[0] = *q00; v01[0] = *q01; v10[0] = *q10; v11[0] = *q11; }
        const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = hy * (hx * v00[i] + lx * v01[i]) + ly * (hx * v10[i] + lx * v11[i]);
      }
#pragma unroll
      for (int i = 0; i < V; ++i) {
        if (T.scale) v[i] = fmaf(v[i], T.scale[c + i], T.shift ? T.shift[c + i] : 0.f);
        acc[i] += v[i];
      }
    }
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = fmaxf(acc[i], 0.f);
    }
    if (V == 2) *reinterpret_cast<float2*>(p.out + e) = make_float2(acc[0], acc[V - 1]);
    else p.out[e] = acc[0];
  }
}

// out[b,Y,X,c] (+)= sum over high-res (h,w) of weight(h->Y)*weight(w->X) * g[b,h,w,c]
// f = 2^k, align_corners=False: s = (dst + 0.5)/f - 0.5 = num/(2f) with num = max(2*dst + 1 - f, 0), so i0 = num >> (k+1) and
// l1 = (num & (2f-1)) / 2f are exact integer expressions of the same fp32 values src_index() produces.  Only the 2f rows
// (columns) [f*Y - f/2, f*Y + 3f/2) can have i0 or i1 equal to Y.
__device__ __forceinline__ float adj_weight(int dst, int k, int in_size, int target) {
  const int f2 = 2 << k;
  int num = 2 * dst + 1 - (1 << k);
  num = max(num, 0);
  const int i0 = num >> (k + 1);
  const float l1 = (float)(num & (f2 - 1)) / (float)f2;
  const int i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  float w = 0.f;
  if (i0 == target) w += 1.f - l1;
  if (i1 == target) w += l1;
  return w;
}

__global__ void upsample_adjoint_kernel(const float* __restrict__ g, float* out, int accumulate, int B, int H, int W, int C,
                                        int log2f) {
  const int f = 1 << log2f;
  const int Hs = H >> log2f, Ws = W >> log2f;
  const long total = (long)B * Hs * Ws * C;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c = (int)(e % C);
    long pix = e / C;
    const int X = (int)(pix % Ws);
    pix /= Ws;
    const int Y = (int)(pix % Hs);
    const int b = (int)(pix / Hs);
    const int hlo = max(0, f * Y - f / 2), hhi = min(H - 1, f * Y + f + f / 2 - 1);
    const int wlo = max(0, f * X - f / 2), whi = min(W - 1, f * X + f + f / 2 - 1);
    float acc = 0.f;
    for (int h = hlo; h <= hhi; ++h) {
      const float wy = adj_weight(h, log2f, Hs, Y);
      const float* row = g + ((long)(b * H + h) * W) * C + c;
      float racc = 0.f;
      for (int w = wlo; w <= whi; ++w) racc = fmaf(adj_weight(w, log2f, Ws, X), __ldg(row + (long)w * C), racc);
      acc = fmaf(wy, racc, acc);
    }
    out[e] = accumulate ? out[e] + acc : acc;
  }
}

// Separable form of the same adjoint, one CTA per (b, low-res row Y):
//   phase 1  tmp[r][X][c] = sum_w wx(w -> X) * g[b, h_r, w, c]   for the <= 2f high-res rows h_r that reach Y (global loads
//            coalesced over c; 2f taps per element instead of (2f)^2 per output, and B*Hs*R*Ws*C independent sums instead
//            of B*Hs*Ws*C serial ones — the gather form is latency-bound at f = 4, 8: 78 us per launch measured);
//   phase 2  out[b, Y, X, c] (+)= sum_r wy(h_r -> Y) * tmp[r][X][c]   from shared memory.
__global__ void upsample_adjoint_rows_kernel(const float* __restrict__ g, float* out, int accumulate, int B, int H, int W, int C,
                                             int log2f) {
  extern __shared__ float tmp[];
  const int f = 1 << log2f, f2 = 2 * f;
  const int Hs = H >> log2f, Ws = W >> log2f;
  const int b = blockIdx.x / Hs, Y = blockIdx.x - b * Hs;
  const int hlo = max(0, f * Y - f / 2), hhi = min(H - 1, f * Y + f + f / 2 - 1);
  const int R = hhi - hlo + 1, WC = Ws * C;
  // tap weights once per CTA (they were recomputed, with a float division each, for every element and tap):
  // wxt[X][j] = weight of high-res column wlo(X) + j towards X (0 past the last contributing column), wyt[r] likewise for rows
  float* wxt = tmp + (size_t)f2 * WC;
  float* wyt = wxt + Ws * f2;
  for (int e = threadIdx.x; e < Ws * f2; e += blockDim.x) {
    const int X = e / f2, j = e - X * f2;
    const int wlo = max(0, f * X - f / 2), whi = min(W - 1, f * X + f + f / 2 - 1);
    wxt[e] = (wlo + j <= whi) ? adj_weight(wlo + j, log2f, Ws, X) : 0.f;
  }
  for (int r = threadIdx.x; r < f2; r += blockDim.x) wyt[r] = (r < R) ? adj_weight(hlo + r, log2f, Hs, Y) : 0.f;
  __syncthreads();
  for (int e = threadIdx.x; e < R * WC; e += blockDim.x) {
    const int r = e / WC, xc = e - r * WC;
    const int X = xc / C, c = xc - X * C;
    const int wlo = max(0, f * X - f / 2), whi = min(W - 1, f * X + f + f / 2 - 1);
    const float* row = g + ((long)(b * H + hlo + r) * W) * C + c;
    const float* wt = wxt + X * f2;
    float acc = 0.f;
    for (int w = wlo; w <= whi; ++w) acc = fmaf(wt[w - wlo], __ldg(row + (long)w * C), acc);
    tmp[e] = acc;
  }
  __syncthreads();
  float* orow = out + ((long)(b * Hs + Y) * Ws) * C;
  for (int e = threadIdx.x; e < WC; e += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < R; ++r) acc = fmaf(wyt[r], tmp[r * WC + e], acc);
    orow[e] = accumulate ? orow[e] + acc : acc;
  }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int Ctot, long HW, int coff,
                                    int Cn, int Cpad) {
  const long total = (long)B * HW;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const long b = e / HW, pix = e - b * HW;
    for (int c = 0; c < Cn; ++c) out[e * Cpad + c] = x[(b * Ctot + coff + c) * HW + pix];
    for (int c = Cn; c < Cpad; ++c) out[e * Cpad + c] = 0.f;
  }
}

// one CTA per sample; blockDim multiple of C so a thread's channel is fixed
__global__ void avgpool_kernel(const float* __restrict__ x, float* out, long HW, int C, int ldo, int coff) {
  __shared__ float sm[256];
  const int t = threadIdx.x, nt = blockDim.x;
  const float* xb = x + (long)blockIdx.x * HW * C;
  float a = 0.f;
  for (long e = t; e < HW * C; e += nt) a += xb[e];
  sm[t] = a;
  __syncthreads();
  if (t < C) {
    float r = 0.f;
    for (int j = t; j < nt; j += C) r += sm[j];
    out[(long)blockIdx.x * ldo + coff + t] = r / (float)HW;
  }
}

__global__ void avgpool_bwd_kernel(const float* __restrict__ dout, float* dx, int accumulate, int B, long HW, int C, int ldo,
                                   int coff) {
  const long total = (long)B * HW * C;
  const long stride = (long)gridDim.x * blockDim.x;
  const float inv = 1.f / (float)HW;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c = (int)(e % C);
    const long b = e / (HW * C);
    const float v = dout[b * ldo + coff + c] * inv;
    dx[e] = accumulate ? dx[e] + v : v;
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

// out[B,HW,Cn] (NHWC) = x[B, coff:coff+Cn, HW] (NCHW)
int hcm_nchw_to_nhwc(const float* x, float* out, int B, int Ctot, long HW, int coff, int Cn, cudaStream_t stream) {
  HCM_CHECK_ARG(x && out && coff + Cn <= Ctot, "nchw_to_nhwc: bad args");
  nchw_to_nhwc_kernel<<<ew_grid((long)B * HW * 4), 256, 0, stream>>>(x, out, B, Ctot, HW, coff, Cn, Cn);
  HCM_LAUNCH_CHECK("nchw_to_nhwc");
  return HCM_OK;
}

// the same with the channels of `out` zero-padded to Cpad (out [B,HW,Cpad]): the 3-channel RGB / depth planes become 4-channel
// rows (16-byte aligned, even channel count) so that the stem convolution runs on the tensor-core kernels
int hcm_nchw_to_nhwc_pad(const float* x, float* out, int B, int Ctot, long HW, int coff, int Cn, int Cpad, cudaStream_t stream) {
  HCM_CHECK_ARG(x && out && coff + Cn <= Ctot && Cpad >= Cn, "nchw_to_nhwc_pad: bad args");
  nchw_to_nhwc_kernel<<<ew_grid((long)B * HW * 4), 256, 0, stream>>>(x, out, B, Ctot, HW, coff, Cn, Cpad);
  HCM_LAUNCH_CHECK("nchw_to_nhwc_pad");
  return HCM_OK;
}

// terms: up to 4 (ptr, scale, shift, log2 factor); out [B,H,W,C]
int hcm_fuse_sum(int nterms, const float* const* ptrs, const float* const* scales, const float* const* shifts,
                 const int* log2f, const float* bias, int relu, float* out, int B, int H, int W, int C,
                 cudaStream_t stream) {
  HCM_CHECK_ARG(nterms >= 1 && nterms <= 4 && out, "fuse_sum: nterms=%d", nterms);
  FuseParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < nterms; ++i) {
    HCM_CHECK_ARG(ptrs[i] != nullptr, "fuse_sum: null term %d", i);
    HCM_CHECK_ARG(log2f[i] >= 0 && (H >> log2f[i]) >= 1 && ((H >> log2f[i]) << log2f[i]) == H &&
                  ((W >> log2f[i]) << log2f[i]) == W, "fuse_sum: size not divisible by 2^%d", log2f[i]);
    p.t[i].ptr = ptrs[i];
    p.t[i].scale = scales ? scales[i] : nullptr;
    p.t[i].shift = shifts ? shifts[i] : nullptr;
    p.t[i].log2f = log2f[i];
  }
  p.nterms = nterms; p.bias = bias; p.relu = relu; p.out = out; p.B = B; p.H = H; p.W = W; p.C = C;
  HCM_CHECK_ARG((long)B * H * W * C < (1L << 31), "fuse_sum: tensor too large for 32-bit indexing");
  const int V = (C % 2 == 0) ? 2 : 1;
  const unsigned items = (unsigned)((long)B * H * W * (C / V));
  const FastDiv fd_cv = make_fastdiv((unsigned)(C / V)), fd_w = make_fastdiv((unsigned)W), fd_h = make_fastdiv((unsigned)H);
  if (V == 2) fuse_sum_kernel<2><<<ew_grid((long)items * 4), 256, 0, stream>>>(p, fd_cv, fd_w, fd_h, items);
  else fuse_sum_kernel<1><<<ew_grid((long)items * 4), 256, 0, stream>>>(p, fd_cv, fd_w, fd_h, items);
  HCM_LAUNCH_CHECK("fuse_sum");
  return HCM_OK;
}

// g [B,H,W,C] high-res gradient -> out [B,H>>k,W>>k,C]
int hcm_upsample_adjoint(const float* g, float* out, int accumulate, int B, int H, int W, int C, int log2f,
                         cudaStream_t stream) {
  HCM_CHECK_ARG(g && out && log2f >= 1, "upsample_adjoint: bad args");
  const size_t smem = ((size_t)(2 << log2f) * (size_t)(W >> log2f) * C + (size_t)(2 << log2f) * ((W >> log2f) + 1)) * sizeof(float);
  if (smem <= 96 * 1024 && (H >> log2f) >= 1 && (W >> log2f) >= 1) {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(upsample_adjoint_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      attr = true;
    }
    upsample_adjoint_rows_kernel<<<B * (H >> log2f), 256, smem, stream>>>(g, out, accumulate, B, H, W, C, log2f);
  } else {
    upsample_adjoint_kernel<<<ew_grid(((long)B * H * W * C) >> (2 * log2f)), 256, 0, stream>>>(g, out, accumulate, B, H, W, C, log2f);
  }
  HCM_LAUNCH_CHECK("upsample_adjoint");
  return HCM_OK;
}

int hcm_avgpool(const float* x, float* out, int B, long HW, int C, int ldo, int coff, cudaStream_t stream) {
  HCM_CHECK_ARG(x && out && C >= 1 && C <= 256, "avgpool: bad args (C=%d)", C);
  avgpool_kernel<<<B, (256 / C) * C, 0, stream>>>(x, out, HW, C, ldo, coff);
  HCM_LAUNCH_CHECK("avgpool");
  return HCM_OK;
}

int hcm_avgpool_bwd(const float* dout, float* dx, int accumulate, int B, long HW, int C, int ldo, int coff,
                    cudaStream_t stream) {
  HCM_CHECK_ARG(dout && dx, "avgpool_bwd: null pointer");
  avgpool_bwd_kernel<<<ew_grid((long)B * HW * C), 256, 0, stream>>>(dout, dx, accumulate, B, HW, C, ldo, coff);
  HCM_LAUNCH_CHECK("avgpool_bwd");
  return HCM_OK;
}

}  // extern "C"
