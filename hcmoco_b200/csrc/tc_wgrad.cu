// Tensor-core weight gradient of the stride-1 3x3 / 1x1 convolutions (tcgen05 / TMEM, bf16 hi/lo split):
//     dw[co][ci][r][s] += sum over output positions p of  dy[p][co] * T(x)[p + (r-1, s-1)][ci]
// the cuDNN wgrad behind nn.Conv2d backward (networks/official_hrnet/official_hrnet.py:26-29, 68-75, 187-216).
//
// Same virtual flat position space as tc_conv.cu.  Per 128-position tile the transform warps stage
//   dy  -> Dy[chunk(8 co)][128 positions][16 B]   (non-interior positions zeroed)
//   T(x)-> A [chunk(8 ci)][halo positions][16 B]  (BN scale/shift(+ReLU) applied on load, padding zeroed)
// as bf16 hi/lo planes.  With both operands read MN-major (contiguous 8-channel runs, positions = K), one
// tcgen05.mma M=128 (co rows) x N (ci) x K=16 positions per filter tap accumulates D_tap[co][ci] in TMEM;
// tap (r,s) is again only a start-address offset of the A operand.  A CTA keeps its 9 accumulators
// (9*N <= 512 TMEM columns) resident while it walks its range of tiles, double-buffering the staged operands
// against the MMAs, and finally adds its partial dw to global memory with fp32 atomics.
// Grid = (tile ranges, ci splits of <=48 (3x3) / <=256 (1x1) channels, co blocks of 128).
#include "tc_common.cuh"

namespace {

constexpr int TILE = 128;
constexpr int NTRANS = 256;                 // transform threads (warps 2..9); warps 2..5 also run the epilogue
constexpr int NTHREADS_W = 64 + NTRANS;     // warp 0 idle/reserved, warp 1 MMA issuer + TMEM owner
constexpr int HDR = 4096;

struct WGeo {
  int Hp, Wp, L, Lpad, taps, Nr, nsplit, nblk, ntr, tiles_per, tmem_cols, nstage;
  long Mv, T;
  size_t stage_bytes, smem;
};

WGeo make_wgeo(int B, int H, int W, int Cin, int Cout, int ks) {
  WGeo g;
  g.taps = ks * ks;
  g.Hp = (ks == 3) ? H + 2 : H;
  g.Wp = (ks == 3) ? W + 2 : W;
  g.Mv = (long)B * g.Hp * g.Wp;
  g.T = (g.Mv + TILE - 1) / TILE;
  g.L = (ks == 3) ? TILE + 2 * (g.Wp + 1) : TILE;
  g.Lpad = ceil_to(g.L, 8);
  const int Cin16 = ceil_to(Cin, 16);
  const int nr_max = (ks == 3) ? 48 : 128;
  g.nsplit = (Cin16 + nr_max - 1) / nr_max;
  g.Nr = ceil_to((Cin16 + g.nsplit - 1) / g.nsplit, 16);
  g.nblk = (Cout + 127) / 128;
  int c = 32;
  while (c < g.taps * g.Nr) c <<= 1;
  g.tmem_cols = c;
  const int cop8 = ceil_to(Cout < 128 ? Cout : 128, 8) / 8;       // dy planes per stage (largest co block)
  g.stage_bytes = (size_t)2 * cop8 * 2048 + (size_t)2 * (g.Nr / 8) * g.Lpad * 16;
  const size_t slack = 34 * 1024;                                  // M=128 reads 16 dy planes whatever Cout is
  g.nstage = (HDR + 2 * g.stage_bytes + slack <= 200 * 1024) ? 2 : 1;
  g.smem = HDR + g.nstage * g.stage_bytes + slack;
  const int per_sm = (g.tmem_cols <= 256 && g.smem <= 110 * 1024) ? 2 : 1;
  long want = (148L * per_sm) / ((long)g.nsplit * g.nblk);
  if (want < 1) want = 1;
  if (want > g.T) want = g.T;
  g.tiles_per = (int)((g.T + want - 1) / want);
  g.ntr = (int)((g.T + g.tiles_per - 1) / g.tiles_per);
  return g;
}

bool wgeo_ok(const WGeo& g, int Cin, int Cout, int ks) {
  return (ks == 1 || ks == 3) && Cin <= 256 && Cout <= 256 && (Cin % 2) == 0 && (Cout % 2) == 0 && g.tmem_cols <= 512 &&
         g.smem <= 225 * 1024;
}

struct WParams {
  const float* x;
  const float* in_scale;
  const float* in_shift;
  int in_relu;
  const float* dy;
  float* dw;
  int B, H, W, Cin, Cout, ks, lddw;
  WGeo g;
};

// gather 8 channels [c0, c0+8) of one position (zeros when !valid / beyond C), optional affine + ReLU, split, store
__device__ __forceinline__ void stage8(const float* __restrict__ src, bool valid, int c0, int C, const float* s_sc,
                                       const float* s_sh, bool affine, int relu, uint8_t* dst_hi, uint8_t* dst_lo) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (valid && c0 < C) {
    if (c0 + 8 <= C) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 u = __ldg(reinterpret_cast<const float2*>(src + c0) + i);
        v[2 * i] = u.x; v[2 * i + 1] = u.y;
      }
    } else {
      for (int i = 0; i < C - c0; ++i) v[i] = __ldg(src + c0 + i);
    }
    if (affine) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (c0 + i < C) {
          const float a = fmaf(v[i], s_sc[c0 + i], s_sh[c0 + i]);
          v[i] = relu ? fmaxf(a, 0.f) : a;
        }
      }
    }
  }
  uint4 hi, lo;
  split8(v, hi, lo);
  *reinterpret_cast<uint4*>(dst_hi) = hi;
  *reinterpret_cast<uint4*>(dst_lo) = lo;
}

__global__ void __launch_bounds__(NTHREADS_W) tc_wgrad_kernel(const WParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const WGeo& g = p.g;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);      // [0,1] full, [2,3] empty, [4] done
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 128);
  float* s_sc = reinterpret_cast<float*>(smem + 1024);
  float* s_sh = s_sc + 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  const int ci_lo = blockIdx.y * g.Nr;                     // this CTA's input-channel range
  const int co_lo = blockIdx.z * 128;
  const int co_n = min(128, p.Cout - co_lo);
  const int cop8 = (co_n + 7) / 8;                         // dy planes actually staged
  const int cop8_max = ceil_to(p.Cout < 128 ? p.Cout : 128, 8) / 8;
  const int nchunk = g.Nr / 8;
  const uint32_t a_plane = (uint32_t)g.Lpad * 16;
  // stage layout: [dy_hi cop8_max planes][dy_lo cop8_max planes][a_hi nchunk planes][a_lo nchunk planes]
  const uint32_t off_dylo = (uint32_t)cop8_max * 2048, off_ahi = 2 * off_dylo, off_alo = off_ahi + (uint32_t)nchunk * a_plane;

  if (threadIdx.x == 0) {
    mbar_init(BAR(0), NTRANS); mbar_init(BAR(1), NTRANS);
    mbar_init(BAR(2), 1); mbar_init(BAR(3), 1);
    mbar_init(BAR(4), 1);
    fence_mbar_init();
  }
  for (int c = threadIdx.x; c < p.Cin; c += NTHREADS_W) {
    s_sc[c] = p.in_scale ? p.in_scale[c] : 1.f;
    s_sh[c] = p.in_scale ? p.in_shift[c] : 0.f;
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr), g.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  const long t_beg = (long)blockIdx.x * g.tiles_per;
  const long t_end = min(g.T, t_beg + g.tiles_per);
  const int ntiles = (int)(t_end - t_beg);
  const int center = (p.ks == 3) ? g.Wp + 1 : 0;
  const long HpWp = (long)g.Hp * g.Wp;

  if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = instr_desc(g.Nr, 1);
      for (int it = 0; it < ntiles; ++it) {
        const int s = it % g.nstage;
        mbar_wait(BAR(s), (it / g.nstage) & 1);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + HDR + (size_t)s * g.stage_bytes);
        for (int tap = 0; tap < g.taps; ++tap) {
          const int r = tap / p.ks, sf = tap - r * p.ks;
          const uint32_t tap_off = (p.ks == 3) ? (uint32_t)(r * g.Wp + sf) * 16u : 0u;
          const uint32_t d = tmem + (uint32_t)(tap * g.Nr);
          for (int k = 0; k < TILE / 16; ++k) {
            const uint32_t kd = (uint32_t)k * 256u;          // 16 positions * 16 B
            const uint64_t dyh = smem_desc(st + kd, 128, 2048), dyl = smem_desc(st + off_dylo + kd, 128, 2048);
            const uint64_t ah = smem_desc(st + off_ahi + kd + tap_off, 128, a_plane);
            const uint64_t al = smem_desc(st + off_alo + kd + tap_off, 128, a_plane);
            umma_bf16(d, dyh, ah, idesc, (it > 0 || k > 0) ? 1u : 0u);
            umma_bf16(d, dyh, al, idesc, 1u);
            umma_bf16(d, dyl, ah, idesc, 1u);
          }
        }
        umma_commit(BAR(2 + s));
      }
      umma_commit(BAR(4));
    }
  } else if (warp >= 2) {
    // ===== transform warps =====
    const int t = threadIdx.x - 64;
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % g.nstage;
      mbar_wait(BAR(2 + s), ((it / g.nstage) & 1) ^ 1);
      uint8_t* st = smem + HDR + (size_t)s * g.stage_bytes;
      const long tile0 = (t_beg + it) * TILE;
      // T(x) halo, channels [ci_lo, ci_lo + Nr)
      for (int pos = t; pos < g.Lpad; pos += NTRANS) {
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
        for (int c8 = 0; c8 < nchunk; ++c8)
          stage8(p.x + src, valid, ci_lo + c8 * 8, p.Cin, s_sc, s_sh, p.in_scale != nullptr, p.in_relu,
                 st + off_ahi + (size_t)c8 * a_plane + (size_t)pos * 16, st + off_alo + (size_t)c8 * a_plane + (size_t)pos * 16);
      }
      // dy tile, channels [co_lo, co_lo + co_n)
      for (int e = t; e < TILE * cop8; e += NTRANS) {
        const int pos = e % TILE, c8 = e / TILE;
        const long pv = tile0 + pos;
        bool valid = pv < g.Mv;
        long src = 0;
        if (valid) {
          if (p.ks == 3) {
            const long b = pv / HpWp;
            const int rem = (int)(pv - b * HpWp);
            const int row = rem / g.Wp, col = rem - row * g.Wp;
            valid = row >= 1 && row <= p.H && col >= 1 && col <= p.W;
            src = ((b * p.H + row - 1) * p.W + (col - 1)) * (long)p.Cout;
          } else {
            src = pv * (long)p.Cout;
          }
        }
        stage8(p.dy + src, valid, co_lo + c8 * 8, p.Cout, nullptr, nullptr, false, 0,
               st + (size_t)c8 * 2048 + (size_t)pos * 16, st + off_dylo + (size_t)c8 * 2048 + (size_t)pos * 16);
      }
      fence_proxy_async();
      mbar_arrive(BAR(s));
    }
    // ===== epilogue (warps 2..5): D_tap[co][ci] -> dw (fp32 atomics) =====
    if (warp < 6) {
      mbar_wait(BAR(4), 0);
      tc_fence_after();
      const int q = warp & 3;
      const int co = co_lo + q * 32 + lane;
      const bool rowok = (q * 32 + lane) < co_n;
      for (int tap = 0; tap < g.taps; ++tap) {
        for (int c0 = 0; c0 < g.Nr; c0 += 16) {
          float v[16];
          tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(tap * g.Nr + c0), v);
          if (rowok && ntiles > 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int ci = ci_lo + c0 + i;
              if (ci < p.Cin) atomicAdd(p.dw + ((long)co * p.lddw + ci) * g.taps + tap, v[i]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, g.tmem_cols);
}

}  // namespace

extern "C" {

int hcm_tc_wgrad_supported(int B, int H, int W, int Cin, int Cout, int ks, int stride) {
  if (stride != 1 || (ks != 1 && ks != 3)) return 0;
  WGeo g = make_wgeo(B, H, W, Cin, Cout, ks);
  return wgeo_ok(g, Cin, Cout, ks) ? 1 : 0;
}

// dw[Cout,Cin,ks,ks] += sum_pixels dy * T(x)   (stride 1; fp32 atomics across CTAs; caller zeroes dw once per step).
// lddw > 0: dw is a column block of a wider [Cout][lddw][ks][ks] tensor
int hcm_tc_wgrad(const float* x, const float* dy, float* dw, int lddw, int B, int H, int W, int Cin, int Cout, int ks,
                 const float* in_scale, const float* in_shift, int in_relu, cudaStream_t stream) {
  HCM_CHECK_ARG(x && dy && dw, "tc_wgrad: null pointer");
  HCM_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "tc_wgrad: in_scale/in_shift must come together");
  WParams p;
  p.g = make_wgeo(B, H, W, Cin, Cout, ks);
  HCM_CHECK_ARG(wgeo_ok(p.g, Cin, Cout, ks), "tc_wgrad: unsupported geometry (Cin=%d Cout=%d ks=%d)", Cin, Cout, ks);
  p.x = x; p.in_scale = in_scale; p.in_shift = in_shift; p.in_relu = in_relu; p.dy = dy; p.dw = dw;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ks = ks; p.lddw = lddw > 0 ? lddw : Cin;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { hcm_set_error("tc_wgrad: smem attribute: %s", cudaGetErrorString(e)); return HCM_ERR_CUDA; }
    configured = true;
  }
  dim3 grid(p.g.ntr, p.g.nsplit, p.g.nblk);
  tc_wgrad_kernel<<<grid, NTHREADS_W, p.g.smem, stream>>>(p);
  HCM_LAUNCH_CHECK("tc_wgrad");
  return HCM_OK;
}

}  // extern "C"
