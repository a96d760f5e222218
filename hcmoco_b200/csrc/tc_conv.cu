// Tensor-core (tcgen05 / TMEM) implicit-GEMM convolution for sm_100a: 3x3 and 1x1, stride 1, NHWC fp32
// activations, fp32-accurate through a two-term bf16 split (x = hi + lo; three MMAs per K step:
// hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM; SURVEY.md F8 — single-pass TF32/BF16 misses the
// 1e-3 parity bar, the split passes it).  Replaces the cuDNN calls behind nn.Conv2d in
// networks/official_hrnet/official_hrnet.py:26-29, 68-75, 187-216 for the stride-1 layers (forward and,
// with transposed+flipped packed weights, the data gradient).
//
// Formulation.  The zero-padded input is treated as one flat sequence of "virtual" positions
// [B][H+2][W+2]; an output tile is 128 consecutive virtual positions, so filter tap (r,s) is nothing
// but a constant offset r*(W+2)+s in that sequence.  Per tile the CTA
//   1. (transform warps) gathers the tile's halo once from global memory, applies the producer's
//      pending BatchNorm scale/shift(+ReLU) on the fly, zeroes padding positions, splits into bf16
//      hi/lo and stores them channel-chunk-major:  A[chunk(8 ch)][position][16 B]  (hi and lo planes).
//      In this layout a UMMA K-major/no-swizzle operand for ANY tap is just a different start address
//      (LBO = plane stride, SBO = 128 B): no im2col copy, every input element is staged once.
//   2. (TMA warp) streams the pre-packed bf16 hi/lo weight slabs (one per (tap, 16 channels)) through a
//      4-stage shared-memory ring with cp.async.bulk + mbarrier complete_tx.
//   3. (MMA warp, one thread) issues tcgen05.mma kind::f16 (bf16 x bf16 -> fp32) M=128, N=ceil16(Cout)
//      into TMEM, tcgen05.commit releases ring stages / signals the epilogue.
//   4. (epilogue = the transform warps) tcgen05.ld the accumulator rows, add bias / accumulate, store
//      the valid (interior) positions as fp32 NHWC.
// Outputs computed for padding positions are discarded (waste 2/(W+2) per row).
#include "tc_common.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int NTHREADS = 192;     // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: transform + epilogue
constexpr int WSTAGES = 4;
constexpr int MAXG = 8;
constexpr int HDR_BYTES = 4096;   // barriers, tmem pointer, scale/shift (2 x 256 floats)

struct Geo {
  int Hp, Wp, L, Lpad, Npad, Cin16, ngroups, cg[MAXG], cgmax, nsteps, tmem_cols;
  long Mv;
  size_t smem;
};

Geo make_geo(int B, int H, int W, int Cin, int Cout, int ks) {
  Geo g;
  g.Hp = (ks == 3) ? H + 2 : H;
  g.Wp = (ks == 3) ? W + 2 : W;
  g.Mv = (long)B * g.Hp * g.Wp;
  g.L = (ks == 3) ? TILE_M + 2 * (g.Wp + 1) : TILE_M;
  g.Lpad = ceil_to(g.L, 8);
  g.Npad = ceil_to(Cout, 16);
  g.Cin16 = ceil_to(Cin, 16);
  int max_cg = (int)((96 * 1024) / (4 * g.Lpad)) / 16 * 16;
  if (max_cg > 128) max_cg = 128;
  if (max_cg < 16) max_cg = 16;
  g.ngroups = (g.Cin16 + max_cg - 1) / max_cg;
  int per = ceil_to((g.Cin16 + g.ngroups - 1) / g.ngroups, 16);
  int left = g.Cin16;
  g.cgmax = 0;
  g.nsteps = 0;
  for (int i = 0; i < g.ngroups; ++i) {
    g.cg[i] = left < per ? left : per;
    left -= g.cg[i];
    if (g.cg[i] > g.cgmax) g.cgmax = g.cg[i];
    g.nsteps += ks * ks * (g.cg[i] / 16);
  }
  int c = 32;
  while (c < g.Npad) c <<= 1;
  g.tmem_cols = c;
  g.smem = HDR_BYTES + (size_t)4 * g.cgmax * g.Lpad + (size_t)WSTAGES * 64 * g.Npad;
  return g;
}

bool geo_ok(const Geo& g, int Cin, int Cout, int ks) {
  return (ks == 1 || ks == 3) && Cin <= 256 && Cout <= 256 && (Cout % 2) == 0 && (Cin % 2) == 0 && g.ngroups <= MAXG &&
         g.smem <= 220 * 1024 && g.Lpad * 16 < (1 << 18);
}

struct TcParams {
  const float* x;
  const float* in_scale;
  const float* in_shift;
  int in_relu;
  const __nv_bfloat16* wpack;
  const float* bias;
  float* y;
  int accumulate;
  int B, H, W, Cin, Cout, ks;
  Geo g;
};

// ------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(NTHREADS) tc_conv_kernel(const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Geo& g = p.g;
  // header
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);            // [0..3] wfull, [4..7] wempty, 8 a_ready, 9 a_free, 10 acc_ready
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 128);
  float* s_sc = reinterpret_cast<float*>(smem + 1024);
  float* s_sh = s_sc + 256;
  uint8_t* A_hi = smem + HDR_BYTES;
  const uint32_t plane = (uint32_t)g.Lpad * 16;                  // bytes per 8-channel plane
  uint8_t* A_lo = A_hi + (size_t)(g.cgmax / 8) * plane;
  uint8_t* Wring = A_hi + (size_t)4 * g.cgmax * g.Lpad;
  const uint32_t wslab = 64u * g.Npad;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < WSTAGES; ++s) { mbar_init(BAR(s), 1); mbar_init(BAR(4 + s), 1); }
    mbar_init(BAR(8), 128);
    mbar_init(BAR(9), 1);
    mbar_init(BAR(10), 1);
    fence_mbar_init();
  }
  for (int c = threadIdx.x; c < p.Cin; c += NTHREADS) {
    s_sc[c] = p.in_scale ? p.in_scale[c] : 1.f;
    s_sh[c] = p.in_scale ? p.in_shift[c] : 0.f;
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr), g.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  const long tile0 = (long)blockIdx.x * TILE_M;                  // first virtual position of this tile
  const int taps = p.ks * p.ks;
  const int center = (p.ks == 3) ? g.Wp + 1 : 0;                 // halo index of tile position 0

  if (warp == 0) {
    // ===== TMA producer: weight slabs, in the order the MMA warp consumes them =====
    if (lane == 0) {
      for (int it = 0; it < g.nsteps; ++it) {
        const int s = it % WSTAGES;
        mbar_wait(BAR(4 + s), ((it / WSTAGES) & 1) ^ 1);
        mbar_expect_tx(BAR(s), wslab);
        tma_bulk_g2s(smem_u32(Wring + (size_t)s * wslab), reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)it * wslab,
                     wslab, BAR(s));
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = instr_desc(g.Npad);
      const uint32_t a_hi0 = smem_u32(A_hi), a_lo0 = smem_u32(A_lo), w0 = smem_u32(Wring);
      const uint32_t b_lbo = (uint32_t)g.Npad * 16;
      int it = 0;
      for (int grp = 0; grp < g.ngroups; ++grp) {
        mbar_wait(BAR(8), grp & 1);
        tc_fence_after();
        for (int tap = 0; tap < taps; ++tap) {
          const int r = tap / p.ks, sft = tap - r * p.ks;
          const uint32_t pos_off = (p.ks == 3) ? (uint32_t)(r * g.Wp + sft) * 16u : 0u;
          for (int j = 0; j < g.cg[grp] / 16; ++j, ++it) {
            const int s = it % WSTAGES;
            mbar_wait(BAR(s), (it / WSTAGES) & 1);
            tc_fence_after();
            const uint32_t aoff = (uint32_t)(2 * j) * plane + pos_off;
            const uint64_t ah = smem_desc(a_hi0 + aoff, plane, 128), al = smem_desc(a_lo0 + aoff, plane, 128);
            const uint32_t wb = w0 + (uint32_t)s * wslab;
            const uint64_t bh = smem_desc(wb, b_lbo, 128), bl = smem_desc(wb + wslab / 2, b_lbo, 128);
            umma_bf16(tmem, ah, bh, idesc, it > 0 ? 1u : 0u);
            umma_bf16(tmem, ah, bl, idesc, 1u);
            umma_bf16(tmem, al, bh, idesc, 1u);
            umma_commit(BAR(4 + s));                 // ring stage free once these MMAs have read it
          }
        }
        if (grp + 1 < g.ngroups) umma_commit(BAR(9)); // A buffer may be overwritten with the next channel group
      }
      umma_commit(BAR(10));                           // accumulator complete
    }
  } else {
    // ===== transform warps (128 threads): gather + BN/ReLU-on-load + bf16 split -> A planes =====
    const int t = threadIdx.x - 64;
    const long HpWp = (long)g.Hp * g.Wp;
    int cbase = 0;
    for (int grp = 0; grp < g.ngroups; ++grp) {
      if (grp > 0) mbar_wait(BAR(9), (grp - 1) & 1);
      const int nchunk = g.cg[grp] / 8;
      for (int pos = t; pos < g.Lpad; pos += 128) {
        const long pv = tile0 - center + pos;
        bool valid = pos < g.L && pv >= 0 && pv < g.Mv;
        long src = 0;
        if (valid) {
          if (p.ks == 3) {
            const long b = pv / HpWp;
            const int rem = (int)(pv - b * HpWp);
            const int row = rem / g.Wp, col = rem - row * g.Wp;
            valid = row >= 1 && row <= p.H && col >= 1 && col <= p.W;
            src = ((b * p.H + row - 1) * p.W + (col - 1)) * (long)p.Cin;
          } else {
            src = pv * (long)p.Cin;
          }
        }
        const float* xp = p.x + src;
        for (int c8 = 0; c8 < nchunk; ++c8) {
          const int c0 = cbase + c8 * 8;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
          if (valid && c0 < p.Cin) {
            if (c0 + 8 <= p.Cin) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 u = __ldg(reinterpret_cast<const float2*>(xp + c0) + i);
                v[2 * i] = u.x; v[2 * i + 1] = u.y;
              }
            } else {
              for (int i = 0; i < p.Cin - c0; ++i) v[i] = __ldg(xp + c0 + i);
            }
            if (p.in_scale) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (c0 + i < p.Cin) {
                  float a = fmaf(v[i], s_sc[c0 + i], s_sh[c0 + i]);
                  v[i] = p.in_relu ? fmaxf(a, 0.f) : a;
                }
              }
            }
          }
          uint4 hi, lo;
          split8(v, hi, lo);
          *reinterpret_cast<uint4*>(A_hi + (size_t)c8 * plane + (size_t)pos * 16) = hi;
          *reinterpret_cast<uint4*>(A_lo + (size_t)c8 * plane + (size_t)pos * 16) = lo;
        }
      }
      cbase += g.cg[grp];
      fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      mbar_arrive(BAR(8));
    }
    // ===== epilogue: TMEM -> registers -> global (fp32 NHWC), interior positions only =====
    mbar_wait(BAR(10), 0);
    tc_fence_after();
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;
    const long pv = tile0 + m;
    bool valid = pv < g.Mv;
    long dst = 0;
    if (valid) {
      if (p.ks == 3) {
        const long b = pv / HpWp;
        const int rem = (int)(pv - b * HpWp);
        const int row = rem / g.Wp, col = rem - row * g.Wp;
        valid = row >= 1 && row <= p.H && col >= 1 && col <= p.W;
        dst = ((b * p.H + row - 1) * p.W + (col - 1)) * (long)p.Cout;
      } else {
        dst = pv * (long)p.Cout;
      }
    }
    float* yp = p.y + dst;
    for (int c0 = 0; c0 < g.Npad; c0 += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (valid) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const int c = c0 + i;
          if (c < p.Cout) {
            float2 o = make_float2(v[i], v[i + 1]);
            if (p.bias) { o.x += p.bias[c]; o.y += p.bias[c + 1]; }
            if (p.accumulate) { const float2 old = *reinterpret_cast<const float2*>(yp + c); o.x += old.x; o.y += old.y; }
            *reinterpret_cast<float2*>(yp + c) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, g.tmem_cols);
}

// ------------------------------------------------------------------------------------------ weight packing
// wpack: for every K step (channel group, tap, 16 channels) one slab of 64*Npad bytes:
//   hi[2 chunks][Npad][8 bf16], lo[2 chunks][Npad][8 bf16]   (the shared-memory image of the B operand)
// transpose = 0: B[n][c] = w[n][c][tap]                        (forward; w is OIHW [Cout][Cin][ks][ks])
// transpose = 1: B[n][c] = w[c][n][taps-1-tap]  with the conv seen from the gradient side:
//                n runs over the ORIGINAL Cin, c over the ORIGINAL Cout (data gradient)
__global__ void tc_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, Geo g, int Cin, int Cout, int ks,
                               int transpose, int wCin) {
  const int taps = ks * ks;
  const long total = (long)g.nsteps * 2 * g.Npad * 8;            // (step, chunk, n, k) tuples; hi and lo written together
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int k = (int)(e & 7);
    long r = e >> 3;
    const int n = (int)(r % g.Npad); r /= g.Npad;
    const int h = (int)(r & 1);
    int step = (int)(r >> 1);
    // decode step -> (group, tap, j)
    int grp = 0, cbase = 0, st = step;
    while (st >= taps * (g.cg[grp] / 16)) { st -= taps * (g.cg[grp] / 16); cbase += g.cg[grp]; ++grp; }
    const int tap = st / (g.cg[grp] / 16), j = st - tap * (g.cg[grp] / 16);
    const int c = cbase + j * 16 + h * 8 + k;
    float v = 0.f;
    if (n < Cout && c < Cin) {
      v = transpose ? w[((long)c * wCin + n) * taps + (taps - 1 - tap)] : w[((long)n * wCin + c) * taps + tap];
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* slab = out + (long)step * 32 * g.Npad;
    slab[((long)h * g.Npad + n) * 8 + k] = hi;
    slab[(long)16 * g.Npad + ((long)h * g.Npad + n) * 8 + k] = lo;
  }
}

}  // namespace

extern "C" {

// 1 if hcm_tc_conv can run this convolution (3x3 / 1x1, stride 1, even channel counts <= 256)
int hcm_tc_conv_supported(int B, int H, int W, int Cin, int Cout, int ks, int stride) {
  if (stride != 1 || (ks != 1 && ks != 3)) return 0;
  Geo g = make_geo(B, H, W, Cin, Cout, ks);
  return geo_ok(g, Cin, Cout, ks) ? 1 : 0;
}

// bytes of the packed-weight buffer for this geometry
long hcm_tc_conv_wpack_bytes(int B, int H, int W, int Cin, int Cout, int ks) {
  Geo g = make_geo(B, H, W, Cin, Cout, ks);
  return (long)g.nsteps * 64 * g.Npad;
}

// Pack OIHW fp32 weights into the bf16 hi/lo K-step slabs.  (Cin, Cout) describe the GEMM being run:
// transpose=0 -> the forward conv of w[Cout][Cin][ks][ks];  transpose=1 -> its data gradient, i.e. a conv with
// Cin' = Cout(w), Cout' = Cin(w): pass Cin = Cout(w), Cout = Cin(w).
int hcm_tc_conv_pack(const float* w, void* wpack, int B, int H, int W, int Cin, int Cout, int ks, int transpose,
                     cudaStream_t stream) {
  HCM_CHECK_ARG(w && wpack, "tc_conv_pack: null pointer");
  Geo g = make_geo(B, H, W, Cin, Cout, ks);
  HCM_CHECK_ARG(geo_ok(g, Cin, Cout, ks), "tc_conv_pack: unsupported geometry");
  const long total = (long)g.nsteps * 2 * g.Npad * 8;
  const int wCin = transpose ? Cout : Cin;                       // inner (second) dimension of the OIHW tensor
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  tc_pack_kernel<<<grid, 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(wpack), g, Cin, Cout, ks, transpose, wCin);
  HCM_LAUNCH_CHECK("tc_conv_pack");
  return HCM_OK;
}

// y[B,H,W,Cout] (+)= conv_{ks x ks, stride 1, pad (ks-1)/2}( T(x[B,H,W,Cin]) ) (+ bias), weights pre-packed by hcm_tc_conv_pack
int hcm_tc_conv(const float* x, const void* wpack, const float* bias, float* y, int B, int H, int W, int Cin, int Cout,
                int ks, const float* in_scale, const float* in_shift, int in_relu, int accumulate, cudaStream_t stream) {
  HCM_CHECK_ARG(x && wpack && y, "tc_conv: null pointer");
  HCM_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "tc_conv: in_scale/in_shift must come together");
  TcParams p;
  p.g = make_geo(B, H, W, Cin, Cout, ks);
  HCM_CHECK_ARG(geo_ok(p.g, Cin, Cout, ks), "tc_conv: unsupported geometry (Cin=%d Cout=%d ks=%d)", Cin, Cout, ks);
  p.x = x; p.in_scale = in_scale; p.in_shift = in_shift; p.in_relu = in_relu;
  p.wpack = reinterpret_cast<const __nv_bfloat16*>(wpack); p.bias = bias; p.y = y; p.accumulate = accumulate;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ks = ks;
  static size_t configured = 0;
  if (p.g.smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { hcm_set_error("tc_conv: smem attribute: %s", cudaGetErrorString(e)); return HCM_ERR_CUDA; }
    configured = 227 * 1024;
  }
  const long tiles = (p.g.Mv + TILE_M - 1) / TILE_M;
  tc_conv_kernel<<<(unsigned)tiles, NTHREADS, p.g.smem, stream>>>(p);
  HCM_LAUNCH_CHECK("tc_conv");
  return HCM_OK;
}

}  // extern "C"
