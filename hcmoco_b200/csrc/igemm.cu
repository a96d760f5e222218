// fp32 implicit-GEMM family (CUDA cores): conv2d forward / data-gradient / weight-gradient over
// NHWC activations with reference-layout (OIHW) weights, plus a generic batched strided GEMM.
// This is the exact-fp32 path: it defines parity for every conv/GEMM shape (3->64 stem, 18..256
// channel branches, 1x1 fuse/projection convs) and is what the tensor-core kernels are checked
// against.  Replaces the cuDNN/cuBLAS calls behind nn.Conv2d / nn.Linear / torch.matmul at
// networks/official_hrnet/official_hrnet.py:26-29,68-75,187-216 and build_backbone.py:243-245.
//
// One CTA = 128 threads = 16 (tx, along M) x 8 (ty, along N); CTA tile (16*PM) x (8*CN);
// each thread owns PM x CN accumulators; K is consumed in chunks of <= 32 through shared memory.
//   FWD   : M = B*Ho*Wo output pixels, N = Cout, K = taps*Cin   (A gathered from x, on-load BN/ReLU)
//   DGRAD : M = B*H*W   input pixels,  N = Cin,  K = taps*Cout  (A gathered from dy)
//   WGRAD : M = taps*Cin,              N = Cout, K = output pixels (split over grid.z, fp32 RED)
//   GEMM  : C[b] = alpha*A[b]*B[b] (+bias) (+C), arbitrary element strides
#include "common.cuh"

namespace {

enum { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2, MODE_GEMM = 3 };
constexpr int KC = 32;
constexpr int NT = 128;

struct IgemmParams {
  const float* A;
  const float* Bm;
  float* C;
  int B, H, W, Cin, Ho, Wo, Cout, ks, stride, pad;
  const float* in_scale;
  const float* in_shift;
  int in_relu;
  float* stat_part;  // FWD: [gridDim.x][2][Cout] per-CTA column sums of y and y^2
  const float* bias;
  int accumulate;
  int M, N, K;
  long sAm, sAk, sBk, sBn, sCm, bsA, bsB, bsC;
  float alpha;
  int split_len;  // WGRAD: pixels per grid.z slice
};

template <int MODE, int PM, int CN>
__global__ void __launch_bounds__(NT) igemm_kernel(const IgemmParams p) {
  constexpr int TM = 16 * PM;
  constexpr int TN = 8 * CN;
  constexpr int LDA = TM + 4;
  __shared__ __align__(16) float As[KC][LDA];
  __shared__ __align__(16) float Bs[KC][TN];
  // row info: FWD/DGRAD -> per tile row (pixel); WGRAD -> per K row (pixel) of the current chunk
  __shared__ int ri_base[(MODE == MODE_WGRAD) ? KC : TM];
  __shared__ int ri_h[(MODE == MODE_WGRAD) ? KC : TM];
  __shared__ int ri_w[(MODE == MODE_WGRAD) ? KC : TM];
  // WGRAD column info (tap row/col, channel) per tile row
  __shared__ int ci_r[(MODE == MODE_WGRAD) ? TM : 1];
  __shared__ int ci_s[(MODE == MODE_WGRAD) ? TM : 1];
  __shared__ int ci_c[(MODE == MODE_WGRAD) ? TM : 1];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;

  float acc[PM][CN];
#pragma unroll
  for (int i = 0; i < PM; ++i)
#pragma unroll
    for (int j = 0; j < CN; ++j) acc[i][j] = 0.f;

  auto compute = [&](int kc) {
#pragma unroll 4
    for (int k = 0; k < kc; ++k) {
      float a[PM], b[CN];
      {
        const float4 v = *reinterpret_cast<const float4*>(&As[k][4 * tx]);
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
        if (PM == 8) {
          const float4 u = *reinterpret_cast<const float4*>(&As[k][64 + 4 * tx]);
          a[PM - 4] = u.x; a[PM - 3] = u.y; a[PM - 2] = u.z; a[PM - 1] = u.w;
        }
      }
#pragma unroll
      for (int j = 0; j < CN; ++j) b[j] = Bs[k][ty * CN + j];
#pragma unroll
      for (int i = 0; i < PM; ++i)
#pragma unroll
        for (int j = 0; j < CN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  };

  if (MODE == MODE_FWD || MODE == MODE_DGRAD) {
    // geometry of the tensor the A operand is gathered from, and of the M (pixel) index space
    const int Hm = (MODE == MODE_FWD) ? p.Ho : p.H, Wm = (MODE == MODE_FWD) ? p.Wo : p.W;
    const int Hs = (MODE == MODE_FWD) ? p.H : p.Ho, Ws = (MODE == MODE_FWD) ? p.W : p.Wo;
    const int Cs = (MODE == MODE_FWD) ? p.Cin : p.Cout;  // channels of the gathered tensor (K side)
    const int Cn = (MODE == MODE_FWD) ? p.Cout : p.Cin;  // N side
    const int M = p.B * Hm * Wm;
    for (int ml = tid; ml < TM; ml += NT) {
      int m = m0 + ml;
      if (m < M) {
        int n = m / (Hm * Wm);
        int rem = m - n * (Hm * Wm);
        int hh = rem / Wm, ww = rem - hh * Wm;
        ri_base[ml] = n * Hs * Ws;
        if (MODE == MODE_FWD) { ri_h[ml] = hh * p.stride - p.pad; ri_w[ml] = ww * p.stride - p.pad; }
        else { ri_h[ml] = hh + p.pad; ri_w[ml] = ww + p.pad; }
      } else {
        ri_base[ml] = -1; ri_h[ml] = 0; ri_w[ml] = 0;
      }
    }
    __syncthreads();
    const int taps = p.ks * p.ks;
    const int Ktot = taps * Cs;              // flattened K index kk = tap*Cs + c
    const int kl = tid & 31, mr = tid >> 5;  // A loader: lanes along K, 4 rows per pass
    const int bk = tid >> 2, bn = tid & 3;   // B loader: 4 threads per K row
    const bool tf = (MODE == MODE_FWD) && p.in_scale != nullptr;
    for (int k0 = 0; k0 < Ktot; k0 += KC) {
      const int kc = min(KC, Ktot - k0);
      // ---- A tile: As[k][ml] ----
      if (kl < kc) {
        const int kk = k0 + kl;
        const int tap = kk / Cs, c = kk - tap * Cs;
        const int r = tap / p.ks, s = tap - r * p.ks;
        float sc = 1.f, sh = 0.f;
        if (tf) { sc = p.in_scale[c]; sh = p.in_shift[c]; }
        for (int ml = mr; ml < TM; ml += NT / 32) {
          float v = 0.f;
          const int base = ri_base[ml];
          if (base >= 0) {
            int hi, wi;
            bool ok;
            if (MODE == MODE_FWD) {
              hi = ri_h[ml] + r; wi = ri_w[ml] + s;
              ok = (hi >= 0) && (hi < Hs) && (wi >= 0) && (wi < Ws);
            } else {
              int hn = ri_h[ml] - r, wn = ri_w[ml] - s;
              ok = (hn >= 0) && (wn >= 0);
              if (p.stride == 1) { hi = hn; wi = wn; }
              else { hi = hn / p.stride; wi = wn / p.stride; ok = ok && (hi * p.stride == hn) && (wi * p.stride == wn); }
              ok = ok && (hi < Hs) && (wi < Ws);
            }
            if (ok) {
              v = __ldg(p.A + ((long)(base + hi * Ws + wi)) * Cs + c);
              if (tf) { v = fmaf(v, sc, sh); if (p.in_relu) v = fmaxf(v, 0.f); }
            }
          }
          As[kl][ml] = v;
        }
      }
      // ---- B tile: Bs[k][n] from OIHW weights ----
      if (bk < kc) {
        const int kk = k0 + bk;
        const int tap = kk / Cs, c = kk - tap * Cs;
        for (int n = bn; n < TN; n += 4) {
          float v = 0.f;
          if (n0 + n < Cn) {
            long off = (MODE == MODE_FWD) ? ((long)(n0 + n) * p.Cin + c) * taps + tap
                                          : ((long)c * p.Cin + (n0 + n)) * taps + tap;
            v = __ldg(p.Bm + off);
          }
          Bs[bk][n] = v;
        }
      }
      __syncthreads();
      compute(kc);
      __syncthreads();
    }
    // ---- epilogue ----
    float cs[CN], cq[CN];
#pragma unroll
    for (int j = 0; j < CN; ++j) { cs[j] = 0.f; cq[j] = 0.f; }
#pragma unroll
    for (int i = 0; i < PM; ++i) {
      const int ml = (i < 4) ? (4 * tx + i) : (64 + 4 * tx + (i - 4));
      const int m = m0 + ml;
      if (m < M) {
        float* dst = p.C + (long)m * Cn + n0 + ty * CN;
#pragma unroll
        for (int j = 0; j < CN; ++j) {
          if (n0 + ty * CN + j < Cn) {
            float v = acc[i][j];
            if (MODE == MODE_FWD && p.bias) v += p.bias[n0 + ty * CN + j];
            if (MODE == MODE_DGRAD && p.accumulate) v += dst[j];
            dst[j] = v;
            cs[j] += v; cq[j] += v * v;
          }
        }
      }
    }
    if (MODE == MODE_FWD && p.stat_part) {
#pragma unroll
      for (int j = 0; j < CN; ++j) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], o);
          cq[j] += __shfl_xor_sync(0xffffffffu, cq[j], o);
        }
        const int n = n0 + ty * CN + j;
        if (tx == 0 && n < Cn) {
          p.stat_part[((long)blockIdx.x * 2 + 0) * Cn + n] = cs[j];
          p.stat_part[((long)blockIdx.x * 2 + 1) * Cn + n] = cq[j];
        }
      }
    }
  } else if (MODE == MODE_WGRAD) {
    const int taps = p.ks * p.ks;
    const int M = taps * p.Cin;
    const int P = p.B * p.Ho * p.Wo;
    for (int ml = tid; ml < TM; ml += NT) {
      int m = m0 + ml;
      if (m < M) {
        int tap = m / p.Cin;
        ci_c[ml] = m - tap * p.Cin;
        ci_r[ml] = tap / p.ks;
        ci_s[ml] = tap - (tap / p.ks) * p.ks;
      } else {
        ci_c[ml] = -1; ci_r[ml] = 0; ci_s[ml] = 0;
      }
    }
    const int pbeg = blockIdx.z * p.split_len;
    const int pend = min(P, pbeg + p.split_len);
    const bool tf = p.in_scale != nullptr;
    for (int p0 = pbeg; p0 < pend; p0 += KC) {
      const int kc = min(KC, pend - p0);
      if (tid < KC) {
        int pix = p0 + tid;
        if (tid < kc) {
          int n = pix / (p.Ho * p.Wo);
          int rem = pix - n * (p.Ho * p.Wo);
          int ho = rem / p.Wo, wo = rem - ho * p.Wo;
          ri_base[tid] = n * p.H * p.W;
          ri_h[tid] = ho * p.stride - p.pad;
          ri_w[tid] = wo * p.stride - p.pad;
        } else {
          ri_base[tid] = -1; ri_h[tid] = 0; ri_w[tid] = 0;
        }
      }
      __syncthreads();
      for (int e = tid; e < kc * TM; e += NT) {
        const int k = e / TM, ml = e - k * TM;
        float v = 0.f;
        const int c = ci_c[ml];
        if (c >= 0) {
          const int hi = ri_h[k] + ci_r[ml], wi = ri_w[k] + ci_s[ml];
          if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) {
            v = __ldg(p.A + ((long)(ri_base[k] + hi * p.W + wi)) * p.Cin + c);
            if (tf) { v = fmaf(v, p.in_scale[c], p.in_shift[c]); if (p.in_relu) v = fmaxf(v, 0.f); }
          }
        }
        As[k][ml] = v;
      }
      for (int e = tid; e < kc * TN; e += NT) {
        const int k = e / TN, n = e - k * TN;
        Bs[k][n] = (n0 + n < p.Cout) ? __ldg(p.Bm + (long)(p0 + k) * p.Cout + n0 + n) : 0.f;
      }
      __syncthreads();
      compute(kc);
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < PM; ++i) {
      const int ml = (i < 4) ? (4 * tx + i) : (64 + 4 * tx + (i - 4));
      const int c = ci_c[ml];
      if (c >= 0) {
        const int tap = ci_r[ml] * p.ks + ci_s[ml];
#pragma unroll
        for (int j = 0; j < CN; ++j) {
          const int n = n0 + ty * CN + j;
          if (n < p.Cout) atomicAdd(p.C + ((long)n * p.Cin + c) * taps + tap, acc[i][j]);
        }
      }
    }
  } else {  // MODE_GEMM
    const float* Ab = p.A + (long)blockIdx.z * p.bsA;
    const float* Bb = p.Bm + (long)blockIdx.z * p.bsB;
    float* Cb = p.C + (long)blockIdx.z * p.bsC;
    for (int k0 = 0; k0 < p.K; k0 += KC) {
      const int kc = min(KC, p.K - k0);
      if (p.sAk == 1) {  // K contiguous: lanes along K
        const int kl = tid & 31, mr = tid >> 5;
        if (kl < kc)
          for (int ml = mr; ml < TM; ml += NT / 32)
            As[kl][ml] = (m0 + ml < p.M) ? __ldg(Ab + (long)(m0 + ml) * p.sAm + (k0 + kl)) : 0.f;
      } else {           // lanes along M
        for (int e = tid; e < kc * TM; e += NT) {
          const int k = e / TM, ml = e - k * TM;
          As[k][ml] = (m0 + ml < p.M) ? __ldg(Ab + (long)(m0 + ml) * p.sAm + (long)(k0 + k) * p.sAk) : 0.f;
        }
      }
      for (int e = tid; e < kc * TN; e += NT) {
        const int k = e / TN, n = e - k * TN;
        Bs[k][n] = (n0 + n < p.N) ? __ldg(Bb + (long)(k0 + k) * p.sBk + (long)(n0 + n) * p.sBn) : 0.f;
      }
      __syncthreads();
      compute(kc);
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < PM; ++i) {
      const int m = m0 + ((i < 4) ? (4 * tx + i) : (64 + 4 * tx + (i - 4)));
      if (m < p.M) {
        float* dst = Cb + (long)m * p.sCm + n0 + ty * CN;
#pragma unroll
        for (int j = 0; j < CN; ++j) {
          const int n = n0 + ty * CN + j;
          if (n < p.N) {
            float v = p.alpha * acc[i][j];
            if (p.bias) v += p.bias[n];
            if (p.accumulate) v += dst[j];
            dst[j] = v;
          }
        }
      }
    }
  }
}

int pick_cn(int N) {
  if (N <= 24) return 3;
  if (N <= 32) return 4;
  if (N <= 40) return 5;
  if (N <= 48) return 6;
  if (N <= 64) return 8;
  int best = 8, bestw = 1 << 30;
  const int cand[3] = {8, 6, 5};
  for (int i = 0; i < 3; ++i) {
    int t = 8 * cand[i];
    int w = ((N + t - 1) / t) * t;
    if (w < bestw) { bestw = w; best = cand[i]; }
  }
  return best;
}

int pick_pm(long M) {
  if (M <= 64) return 4;
  long w8 = ((M + 127) / 128) * 128, w4 = ((M + 63) / 64) * 64;
  return (w4 < w8 && M < 1024) ? 4 : 8;
}

template <int MODE>
int launch(const IgemmParams& p, long M, int N, int gz, cudaStream_t st, const char* name, int* grid_m_out, bool small_tiles = false) {
  int pm = pick_pm(M), cn = pick_cn(N);
  // plain GEMMs only (the conv modes report their partial-row count from pick_pm): when the 128 x 64 tiling leaves most
  // SMs idle (SemGCN: 1024 x 128 outputs = 16 CTAs, 126 us per launch measured) use 64 x 32 tiles
  if (small_tiles && (long)hcm_cdiv(M, 16 * pm) * hcm_cdiv(N, 8 * cn) * gz < 148) { pm = 4; if (N > 24) cn = 4; }
  dim3 grid(hcm_cdiv(M, 16 * pm), hcm_cdiv(N, 8 * cn), gz);
  if (grid_m_out) *grid_m_out = grid.x;
#define HCM_CASE(PM_, CN_) \
  if (pm == PM_ && cn == CN_) { igemm_kernel<MODE, PM_, CN_><<<grid, NT, 0, st>>>(p); }
  HCM_CASE(4, 3) else HCM_CASE(4, 4) else HCM_CASE(4, 5) else HCM_CASE(4, 6) else HCM_CASE(4, 8)
  else HCM_CASE(8, 3) else HCM_CASE(8, 4) else HCM_CASE(8, 5) else HCM_CASE(8, 6) else HCM_CASE(8, 8)
#undef HCM_CASE
  HCM_LAUNCH_CHECK(name);
  return HCM_OK;
}

IgemmParams conv_params(int B, int H, int W, int Cin, int Cout, int ks, int stride) {
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ks = ks; p.stride = stride;
  p.pad = (ks - 1) / 2;
  p.Ho = (H + 2 * p.pad - ks) / stride + 1;
  p.Wo = (W + 2 * p.pad - ks) / stride + 1;
  p.alpha = 1.f;
  return p;
}

}  // namespace

extern "C" {

// number of per-CTA partial rows the forward conv writes into stat_part (each row = 2*Cout floats)
int hcm_conv2d_stat_rows(int B, int H, int W, int Cin, int Cout, int ks, int stride) {
  IgemmParams p = conv_params(B, H, W, Cin, Cout, ks, stride);
  long M = (long)B * p.Ho * p.Wo;
  return hcm_cdiv(M, 16 * pick_pm(M));
}

int hcm_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin,
                   int Cout, int ks, int stride, const float* in_scale, const float* in_shift, int in_relu,
                   float* stat_part, cudaStream_t stream) {
  HCM_CHECK_ARG(x && w && y, "conv2d_fwd: null pointer");
  HCM_CHECK_ARG((ks == 1 || ks == 3) && (stride == 1 || stride == 2), "conv2d_fwd: ks=%d stride=%d unsupported", ks, stride);
  HCM_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "conv2d_fwd: in_scale/in_shift must come together");
  IgemmParams p = conv_params(B, H, W, Cin, Cout, ks, stride);
  p.A = x; p.Bm = w; p.C = y; p.bias = bias;
  p.in_scale = in_scale; p.in_shift = in_shift; p.in_relu = in_relu; p.stat_part = stat_part;
  return launch<MODE_FWD>(p, (long)B * p.Ho * p.Wo, Cout, 1, stream, "conv2d_fwd", nullptr);
}

int hcm_conv2d_dgrad(const float* dy, const float* w, float* dx, int B, int H, int W, int Cin, int Cout, int ks,
                     int stride, int accumulate, cudaStream_t stream) {
  HCM_CHECK_ARG(dy && w && dx, "conv2d_dgrad: null pointer");
  HCM_CHECK_ARG((ks == 1 || ks == 3) && (stride == 1 || stride == 2), "conv2d_dgrad: ks=%d stride=%d unsupported", ks, stride);
  IgemmParams p = conv_params(B, H, W, Cin, Cout, ks, stride);
  p.A = dy; p.Bm = w; p.C = dx; p.accumulate = accumulate;
  return launch<MODE_DGRAD>(p, (long)B * H * W, Cin, 1, stream, "conv2d_dgrad", nullptr);
}

// dw (OIHW) += sum over pixels; the caller zeroes dw (or keeps earlier contributions) beforehand
int hcm_conv2d_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Cin, int Cout, int ks,
                     int stride, const float* in_scale, const float* in_shift, int in_relu, cudaStream_t stream) {
  HCM_CHECK_ARG(x && dy && dw, "conv2d_wgrad: null pointer");
  HCM_CHECK_ARG((ks == 1 || ks == 3) && (stride == 1 || stride == 2), "conv2d_wgrad: ks=%d stride=%d unsupported", ks, stride);
  IgemmParams p = conv_params(B, H, W, Cin, Cout, ks, stride);
  p.A = x; p.Bm = dy; p.C = dw;
  p.in_scale = in_scale; p.in_shift = in_shift; p.in_relu = in_relu;
  const long M = (long)ks * ks * Cin;
  const long P = (long)B * p.Ho * p.Wo;
  const int tiles = hcm_cdiv(M, 16 * pick_pm(M)) * hcm_cdiv(Cout, 8 * pick_cn(Cout));
  long splits = (148 * 4 + tiles - 1) / tiles;
  long maxs = (P + 127) / 128;
  if (splits > maxs) splits = maxs;
  if (splits < 1) splits = 1;
  long len = (P + splits - 1) / splits;
  len = ((len + KC - 1) / KC) * KC;
  splits = (P + len - 1) / len;
  p.split_len = (int)len;
  return launch<MODE_WGRAD>(p, M, Cout, (int)splits, stream, "conv2d_wgrad", nullptr);
}

// C[b][m][n] = alpha * sum_k A[b][m*sAm + k*sAk] * B[b][k*sBk + n*sBn] (+ bias[n]) (+ C if accumulate)
int hcm_gemm(const float* A, const float* Bm, const float* bias, float* C, int batch, int M, int N, int K, long sAm,
             long sAk, long sBk, long sBn, long sCm, long bsA, long bsB, long bsC, float alpha, int accumulate,
             cudaStream_t stream) {
  HCM_CHECK_ARG(A && Bm && C, "gemm: null pointer");
  HCM_CHECK_ARG(batch >= 1 && M >= 1 && N >= 1 && K >= 1, "gemm: bad shape %d %d %d %d", batch, M, N, K);
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.A = A; p.Bm = Bm; p.C = C; p.bias = bias; p.M = M; p.N = N; p.K = K;
  p.sAm = sAm; p.sAk = sAk; p.sBk = sBk; p.sBn = sBn; p.sCm = sCm; p.bsA = bsA; p.bsB = bsB; p.bsC = bsC;
  p.alpha = alpha; p.accumulate = accumulate;
  return launch<MODE_GEMM>(p, M, N, batch, stream, "gemm", nullptr, true);
}

}  // extern "C"
