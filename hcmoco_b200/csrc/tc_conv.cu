// Tensor-core (tcgen05 / TMEM) implicit-GEMM convolution for sm_100a: 3x3 and 1x1, stride 1, NHWC fp32
// activations, fp32-accurate through a two-term bf16 split (x = hi + lo; three MMAs per K step:
// hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM; SURVEY.md F8 — single-pass TF32/BF16 misses the
// 1e-3 parity bar, the split passes it).  Replaces the cuDNN calls behind nn.Conv2d in
// networks/official_hrnet/official_hrnet.py:26-29, 68-75, 187-216 for the stride-1 layers (forward and,
// with transposed+flipped packed weights, the data gradient).
//
// Formulation.  The zero-padded input is treated as one flat sequence of "virtual" positions
// [B][H+2][W+2]; an output tile is 128 consecutive virtual positions, so filter tap (r,s) is nothing
// but a constant ROW offset r*(W+2)+s into the staged halo.  Persistent CTAs (one per SM) walk the tiles:
//   transform warps (256 thr)  gather the tile's halo once from global memory, apply the producer's pending
//                              BatchNorm scale/shift(+ReLU), zero the padding, split into bf16 hi/lo and store
//                              them as a swizzled channels-last tile  A[block of 32|64 ch][position][64|128 B]
//                              (the K-major SWIZZLE_64B/128B UMMA layout; a tap = start row + base_offset);
//   TMA warp                   weights, pre-packed + pre-swizzled (hi/lo slabs per (tap, channel block)):
//                              one cp.async.bulk set when they fit in shared memory, else a ring;
//   MMA warp (one thread)      tcgen05.mma kind::f16 M=128, N=ceil16(Cout), K=16 into a double-buffered TMEM
//                              accumulator; tcgen05.commit frees A stages / ring slots, signals the epilogue;
//   epilogue warps (128 thr)   tcgen05.ld, + bias / accumulate, store the interior positions (fp32 NHWC).
// Outputs computed for padding positions are discarded (waste 2/(W+2) per row).
#include "tc_common.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int NTRANS = 256;                  // transform threads: warps 0..7
constexpr int NTHREADS = NTRANS + 128 + 64;  // warps 8..11: epilogue, warp 12: TMA producer, warp 13: MMA issuer + TMEM owner
// (the SMSP arbiter favours the highest warp id: the single MMA-issuing thread must not starve behind busy transform warps)
constexpr int W_EPI = NTRANS / 32, W_TMA = W_EPI + 4, W_MMA = W_TMA + 1;
constexpr int MAXG = 8;
constexpr int HDR_BYTES = 8192;              // barriers, tmem pointer, scale/shift (2 x 256 floats), position tables (2 x 512 int)
constexpr int MAX_LPAD = 512;

struct Geo {
  int Hp, Wp, L, Lpad, Npad, Cin16, SW, KB, ngroups, cg[MAXG], cgmax, nblkmax, nslabs, tmem_cols;
  int nastage, acc_stages, nacc, w_resident, wst, grid;
  long Mv, tiles;
  size_t plane_bytes, a_stage_bytes, wslab, wbytes, smem;
};

Geo make_geo(int B, int H, int W, int Cin, int Cout, int ks) {
  Geo g;
  g.Hp = (ks == 3) ? H + 2 : H;
  g.Wp = (ks == 3) ? W + 2 : W;
  g.Mv = (long)B * g.Hp * g.Wp;
  g.tiles = (g.Mv + TILE_M - 1) / TILE_M;
  g.L = (ks == 3) ? TILE_M + 2 * (g.Wp + 1) : TILE_M;
  g.Lpad = ceil_to(g.L, 16);                 // Lpad * SW is a multiple of 1024 (swizzle pattern period)
  g.Npad = ceil_to(Cout, 16);
  g.Cin16 = ceil_to(Cin, 16);
  static int force_sw = -1;                   // debug knob: HCM_TC_SW=128 stages every layer with 128-byte rows
  if (force_sw < 0) { const char* e = getenv("HCM_TC_SW"); force_sw = e ? atoi(e) : 0; }
  g.SW = (g.Cin16 <= 32 && force_sw != 128) ? 64 : 128;   // bytes per staged row = swizzle span
  g.KB = g.SW / 2;                           // channels per row block
  g.plane_bytes = (size_t)g.Lpad * g.SW;
  g.wslab = (size_t)2 * g.Npad * g.SW;       // hi rows + lo rows of one (tap, channel block)
  // channel groups: one staged A buffer holds <= max_cg channels of the halo (hi + lo planes)
  int max_blk = (int)((80 * 1024) / (2 * g.plane_bytes));
  if (max_blk < 1) max_blk = 1;
  int max_cg = max_blk * g.KB;
  if (max_cg > 128) max_cg = 128;
  g.ngroups = (g.Cin16 + max_cg - 1) / max_cg;
  int per = ceil_to((g.Cin16 + g.ngroups - 1) / g.ngroups, 16);
  int left = g.Cin16;
  g.cgmax = 0;
  g.nslabs = 0;
  for (int i = 0; i < g.ngroups && i < MAXG; ++i) {
    g.cg[i] = left < per ? left : per;
    left -= g.cg[i];
    if (g.cg[i] > g.cgmax) g.cgmax = g.cg[i];
    g.nslabs += ks * ks * ((g.cg[i] + g.KB - 1) / g.KB);
  }
  g.nblkmax = (g.cgmax + g.KB - 1) / g.KB;
  g.a_stage_bytes = (size_t)2 * g.nblkmax * g.plane_bytes;
  // independent TMEM accumulators for the split-precision passes (hi*hi | hi*lo | lo*hi): back-to-back MMAs into the
  // SAME accumulator serialise on its read-modify-write latency (~250 cycles measured at N<=64); the epilogue adds them
  g.nacc = (3 * g.Npad <= 512) ? 3 : 2;
  g.acc_stages = (2 * g.nacc * g.Npad <= 512) ? 2 : 1;
  int c = 32;
  while (c < g.acc_stages * g.nacc * g.Npad) c <<= 1;
  g.tmem_cols = c;
  g.wbytes = (size_t)g.nslabs * g.wslab;
  const size_t budget = 222 * 1024 - HDR_BYTES;
  const size_t wmin = g.wbytes < 4 * g.wslab ? g.wbytes : 4 * g.wslab;
  g.nastage = (2 * g.a_stage_bytes + wmin <= budget) ? 2 : 1;
  const size_t left_b = budget > g.nastage * g.a_stage_bytes ? budget - g.nastage * g.a_stage_bytes : 0;
  g.w_resident = g.wbytes <= left_b ? 1 : 0;
  g.wst = g.w_resident ? 0 : (int)(left_b / g.wslab > 8 ? 8 : left_b / g.wslab);
  g.smem = HDR_BYTES + g.nastage * g.a_stage_bytes + (g.w_resident ? g.wbytes : (size_t)g.wst * g.wslab);
  g.grid = (int)(g.tiles < 148 ? g.tiles : 148);
  return g;
}

bool geo_ok(const Geo& g, int Cin, int Cout, int ks) {
  return (ks == 1 || ks == 3) && Cin <= 256 && Cout <= 256 && (Cout % 2) == 0 && (Cin % 2) == 0 && g.ngroups <= MAXG &&
         g.Lpad <= MAX_LPAD && g.Mv < (1L << 31) && (g.w_resident || g.wst >= 2) && g.smem <= 226 * 1024;
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
  long long* dbg;              // optional per-CTA cycle counters (HCM_TC_DEBUG), else null
  int base_offset_mode;        // 1: descriptor base_offset = (start >> 7) & 7 (row-shifted swizzled operands)
  Geo g;
};

// virtual position -> pixel index of the unpadded [B,H,W] tensor, or -1 for padding / out of range
__device__ __forceinline__ int virt_to_pixel(long pv, const TcParams& p) {
  const Geo& g = p.g;
  if (pv < 0 || pv >= g.Mv) return -1;
  if (p.ks != 3) return (int)pv;
  const unsigned v = (unsigned)pv, hw = (unsigned)(g.Hp * g.Wp);
  const unsigned b = v / hw, rem = v - b * hw;
  const unsigned row = rem / (unsigned)g.Wp, col = rem - row * (unsigned)g.Wp;
  if (row < 1 || row > (unsigned)p.H || col < 1 || col > (unsigned)p.W) return -1;
  return (int)((b * p.H + row - 1) * p.W + (col - 1));
}

// byte offset of 16-byte chunk `c16` of row `row` inside a swizzled plane (rows of SW bytes, 1024-B aligned base):
// Swizzle<log2(SW/16),4,3>: the chunk index is XORed with address bits [7, 7+log2(SW/16))
__host__ __device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t c16, uint32_t SW) {
  const uint32_t off = row * SW;
  return off + (((c16 ^ (off >> 7)) & (SW / 16 - 1)) << 4);
}

// K-major swizzled operand descriptor: rows of SW bytes, 8-row groups SBO = 8*SW apart
__device__ __forceinline__ uint64_t sw_desc(uint32_t saddr, uint32_t SW, int base_offset_mode) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                                   // LBO (unused for swizzled K-major; CUTLASS writes 1)
  d |= (uint64_t)(((8 * SW) >> 4) & 0x3FFFu) << 32;         // SBO
  d |= (uint64_t)1 << 46;                                   // descriptor version (sm_100)
  if (base_offset_mode) d |= (uint64_t)((saddr >> 7) & 7u) << 49;
  d |= (uint64_t)(SW == 128 ? 2 : (SW == 64 ? 4 : 6)) << 61;
  return d;
}

// ------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(NTHREADS, 1) tc_conv_kernel(const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Geo& g = p.g;
  // barriers: 0,1 a_full  2,3 a_empty  4,5 acc_full  6,7 acc_empty  8 w_full(resident)  16.. ring full, 24.. ring empty
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 512);
  float* s_sc = reinterpret_cast<float*>(smem + 1024);
  float* s_sh = s_sc + 256;
  int* s_src = reinterpret_cast<int*>(smem + 3072);               // [2][MAX_LPAD]
  uint8_t* Abase = smem + HDR_BYTES;
  const uint32_t SW = (uint32_t)g.SW;
  const uint32_t plane = (uint32_t)g.plane_bytes;
  const uint32_t lo_off = (uint32_t)g.nblkmax * plane;            // A_lo planes follow the A_hi planes of a stage
  uint8_t* Wbase = Abase + (size_t)g.nastage * g.a_stage_bytes;
  const uint32_t wslab = (uint32_t)g.wslab;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  if (threadIdx.x == 0) {
    mbar_init(BAR(0), NTRANS); mbar_init(BAR(1), NTRANS);
    mbar_init(BAR(2), 1); mbar_init(BAR(3), 1);
    mbar_init(BAR(4), 1); mbar_init(BAR(5), 1);
    mbar_init(BAR(6), 128); mbar_init(BAR(7), 128);
    mbar_init(BAR(8), 1);
    for (int s = 0; s < 8; ++s) { mbar_init(BAR(16 + s), 1); mbar_init(BAR(24 + s), 1); }
    fence_mbar_init();
  }
  for (int c = threadIdx.x; c < 256; c += NTHREADS) {
    s_sc[c] = (p.in_scale && c < p.Cin) ? p.in_scale[c] : 1.f;
    s_sh[c] = (p.in_scale && c < p.Cin) ? p.in_shift[c] : 0.f;
  }
  if (warp == W_MMA) tmem_alloc(smem_u32(tmem_ptr), g.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  const int taps = p.ks * p.ks;
  const int center = (p.ks == 3) ? g.Wp + 1 : 0;                  // halo row of tile position 0
  const int my_tiles = (int)((g.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  if (warp == W_TMA) {
    // ===== TMA producer: weights =====
    if (lane == 0) {
      if (g.w_resident) {
        mbar_expect_tx(BAR(8), (uint32_t)g.wbytes);
        for (size_t off = 0; off < g.wbytes; off += 32768) {
          const uint32_t n = (uint32_t)(g.wbytes - off < 32768 ? g.wbytes - off : 32768);
          tma_bulk_g2s(smem_u32(Wbase + off), reinterpret_cast<const uint8_t*>(p.wpack) + off, n, BAR(8));
        }
      } else {
        long it = 0;
        for (int ti = 0; ti < my_tiles; ++ti)
          for (int sl = 0; sl < g.nslabs; ++sl, ++it) {
            const int s = (int)(it % g.wst);
            mbar_wait(BAR(24 + s), (uint32_t)(((it / g.wst) & 1) ^ 1));
            mbar_expect_tx(BAR(16 + s), wslab);
            tma_bulk_g2s(smem_u32(Wbase + (size_t)s * wslab), reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)sl * wslab,
                         wslab, BAR(16 + s));
          }
      }
    }
  } else if (warp == W_MMA) {
    // ===== MMA issuer: the whole warp runs the (warp-uniform) loop, one elected lane issues tcgen05 =====
    {
      const uint32_t idesc = instr_desc(g.Npad);
      const uint32_t a0 = smem_u32(Abase), w0 = smem_u32(Wbase);
      const uint64_t desc_t = sw_desc(0, SW, 0);        // everything but the start-address field
      long long c_w = 0, c_acc = 0, c_a = 0, c_all = clock64(), tq;
      tq = clock64();
      if (g.w_resident) { mbar_wait(BAR(8), 0); }
      c_w += clock64() - tq;
      long it = 0, f = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int as = ti % g.acc_stages;
        tq = clock64();
        mbar_wait(BAR(6 + as), (uint32_t)(((ti / g.acc_stages) & 1) ^ 1));     // epilogue has drained this accumulator
        c_acc += clock64() - tq;
        tc_fence_after();
        const uint32_t d0 = tmem + (uint32_t)(as * g.nacc * g.Npad);
        const uint32_t d1 = d0 + (uint32_t)g.Npad, d2 = d0 + (uint32_t)((g.nacc - 1) * g.Npad);
        int sl = 0;
        uint32_t first = 1;
        for (int grp = 0; grp < g.ngroups; ++grp, ++f) {
          const int s = (int)(f % g.nastage);
          tq = clock64();
          mbar_wait(BAR(s), (uint32_t)((f / g.nastage) & 1));
          c_a += clock64() - tq;
          tc_fence_after();
          const uint32_t ab = a0 + (uint32_t)s * (uint32_t)g.a_stage_bytes;
          const int nblk = (g.cg[grp] + g.KB - 1) / g.KB;
          for (int tap = 0; tap < taps; ++tap) {
            const int r = tap / p.ks, sft = tap - r * p.ks;
            const uint32_t row_off = (p.ks == 3) ? (uint32_t)(r * g.Wp + sft) * SW : 0u;
            for (int blk = 0; blk < nblk; ++blk, ++sl, ++it) {
              uint32_t wb;
              int rs = 0;
              if (g.w_resident) {
                wb = w0 + (uint32_t)sl * wslab;
              } else {
                rs = (int)(it % g.wst);
                mbar_wait(BAR(16 + rs), (uint32_t)((it / g.wst) & 1));
                tc_fence_after();
                wb = w0 + (uint32_t)rs * wslab;
              }
              const int ksteps = min(g.KB, g.cg[grp] - blk * g.KB) / 16;
              const uint32_t arow = ab + (uint32_t)blk * plane + row_off;
              for (int j = 0; j < ksteps; ++j) {
                const uint64_t ah = desc_t | (uint64_t)(((arow + 32u * j) & 0x3FFFFu) >> 4);
                const uint64_t al = desc_t | (uint64_t)(((arow + lo_off + 32u * j) & 0x3FFFFu) >> 4);
                const uint64_t bh = desc_t | (uint64_t)(((wb + 32u * j) & 0x3FFFFu) >> 4);
                const uint64_t bl = desc_t | (uint64_t)(((wb + wslab / 2 + 32u * j) & 0x3FFFFu) >> 4);
                if (elect_one()) {
                  umma_bf16(d0, ah, bh, idesc, first ? 0u : 1u);
                  umma_bf16(d1, ah, bl, idesc, first ? 0u : 1u);
                  umma_bf16(d2, al, bh, idesc, (first && g.nacc == 3) ? 0u : 1u);
                }
                first = 0;
              }
              if (!g.w_resident && elect_one()) umma_commit(BAR(24 + rs));
            }
          }
          if (elect_one()) umma_commit(BAR(2 + s));     // staged A buffer free
        }
        if (elect_one()) umma_commit(BAR(4 + as));      // accumulator complete
      }
      if (p.dbg && lane == 0) {
        long long* o = p.dbg + (long)blockIdx.x * 16;
        o[0] = clock64() - c_all; o[1] = c_w; o[2] = c_acc; o[3] = c_a;
      }
    }
  } else if (warp < W_EPI) {
    // ===== transform warps: gather + BN/ReLU-on-load + bf16 split -> swizzled channels-last tile =====
    const int t = threadIdx.x;
    const uint32_t cpb = SW / 16;                       // 16-byte chunks per row block
    long f = 0;
    long long c_wait = 0, c_all = clock64(), tq;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const long tile0 = ((long)blockIdx.x + (long)ti * gridDim.x) * TILE_M;
      int cbase = 0;
      for (int grp = 0; grp < g.ngroups; ++grp, ++f) {
        const int s = (int)(f % g.nastage);
        tq = clock64();
        mbar_wait(BAR(2 + s), (uint32_t)(((f / g.nastage) & 1) ^ 1));
        c_wait += clock64() - tq;
        int* src_tab = s_src + s * MAX_LPAD;
        for (int pos = t; pos < g.Lpad; pos += NTRANS)
          src_tab[pos] = (pos < g.L) ? virt_to_pixel(tile0 - center + pos, p) : -1;
        asm volatile("bar.sync 1, %0;" ::"n"(NTRANS) : "memory");
        uint8_t* A_hi = Abase + (size_t)s * g.a_stage_bytes;
        const int nchunk = g.cg[grp] / 8;
        const int total = nchunk * g.Lpad;
        for (int e0 = t; e0 < total; e0 += 2 * NTRANS) {
          float v[2][8];
          uint32_t dstoff[2];
          bool act[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int e = e0 + u * NTRANS;
            act[u] = e < total;
            const int c8 = act[u] ? e / g.Lpad : 0;
            const int pos = act[u] ? e - c8 * g.Lpad : 0;
            dstoff[u] = (uint32_t)(c8 / cpb) * plane + swz((uint32_t)pos, (uint32_t)c8 % cpb, SW);
            const int c0 = cbase + c8 * 8;
            const int px = act[u] ? src_tab[pos] : -1;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[u][i] = 0.f;
            if (px >= 0 && c0 < p.Cin) {
              const float* xp = p.x + (long)px * p.Cin + c0;
              if (c0 + 8 <= p.Cin) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 w2 = __ldg(reinterpret_cast<const float2*>(xp) + i);
                  v[u][2 * i] = w2.x; v[u][2 * i + 1] = w2.y;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (c0 + i < p.Cin) v[u][i] = __ldg(xp + i);
              }
              if (p.in_scale) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int cc = (c0 + i) & 255;
                  const float a = fmaf(v[u][i], s_sc[cc], s_sh[cc]);
                  v[u][i] = (c0 + i < p.Cin) ? (p.in_relu ? fmaxf(a, 0.f) : a) : 0.f;
                }
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (act[u]) {
              uint4 hi, lo;
              split8(v[u], hi, lo);
              *reinterpret_cast<uint4*>(A_hi + dstoff[u]) = hi;
              *reinterpret_cast<uint4*>(A_hi + lo_off + dstoff[u]) = lo;
            }
          }
        }
        cbase += g.cg[grp];
        fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        mbar_arrive(BAR(s));
      }
    }
    if (p.dbg && t == 0) { long long* o = p.dbg + (long)blockIdx.x * 16; o[4] = clock64() - c_all; o[5] = c_wait; }
  } else {
    // ===== epilogue warps: TMEM -> registers -> global (fp32 NHWC), interior positions only =====
    const int q = warp & 3;                             // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;
    long long c_wait = 0, c_all = clock64(), tq;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const long tile0 = ((long)blockIdx.x + (long)ti * gridDim.x) * TILE_M;
      const int as = ti % g.acc_stages;
      const int px = virt_to_pixel(tile0 + m, p);
      float* yp = p.y + (long)(px < 0 ? 0 : px) * p.Cout;
      tq = clock64();
      mbar_wait(BAR(4 + as), (uint32_t)((ti / g.acc_stages) & 1));
      c_wait += clock64() - tq;
      tc_fence_after();
      for (int c0 = 0; c0 < g.Npad; c0 += 16) {
        float v[16];
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * g.nacc * g.Npad + c0);
        tmem_ld16(tbase, v);
        for (int a = 1; a < g.nacc; ++a) {
          float u[16];
          tmem_ld16(tbase + (uint32_t)(a * g.Npad), u);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += u[i];
        }
        if (px >= 0) {
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
      tc_fence_before();
      mbar_arrive(BAR(6 + as));                         // accumulator stage may be overwritten
    }
    if (p.dbg && m == 0) { long long* o = p.dbg + (long)blockIdx.x * 16; o[6] = clock64() - c_all; o[7] = c_wait; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, g.tmem_cols);
}

// ------------------------------------------------------------------------------------------ weight packing
// wpack: one slab per (channel group, tap, channel block of KB = SW/2 channels):
//   hi[Npad rows][SW bytes], lo[Npad rows][SW bytes]   — the swizzled shared-memory image of the K-major B operand
// transpose = 0: B[n][c] = w[n][c][tap]                        (forward; w is OIHW [Cout][Cin][ks][ks])
// transpose = 1: B[n][c] = w[c][n][taps-1-tap]  with the conv seen from the gradient side:
//                n runs over the ORIGINAL Cin, c over the ORIGINAL Cout (data gradient)
__global__ void tc_pack_kernel(const float* __restrict__ w, uint8_t* __restrict__ out, Geo g, int Cin, int Cout, int ks,
                               int transpose, int wCin) {
  const int taps = ks * ks;
  const long total = (long)g.nslabs * g.Npad * g.KB;             // (slab, n, k) tuples; hi and lo written together
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int k = (int)(e % g.KB);
    long r = e / g.KB;
    const int n = (int)(r % g.Npad);
    int sl = (int)(r / g.Npad);
    // decode slab -> (group, tap, block)
    int grp = 0, cbase = 0, st = sl;
    for (;;) {
      const int nb = (g.cg[grp] + g.KB - 1) / g.KB;
      if (st < taps * nb) break;
      st -= taps * nb; cbase += g.cg[grp]; ++grp;
    }
    const int nb = (g.cg[grp] + g.KB - 1) / g.KB;
    const int tap = st / nb, blk = st - tap * nb;
    const int cl = blk * g.KB + k;                                // channel within the group
    const int c = cbase + cl;
    float v = 0.f;
    if (n < Cout && c < Cin && cl < g.cg[grp]) {
      v = transpose ? w[((long)c * wCin + n) * taps + (taps - 1 - tap)] : w[((long)n * wCin + c) * taps + tap];
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    uint8_t* slab = out + (size_t)sl * g.wslab;
    const uint32_t off = swz((uint32_t)n, (uint32_t)(k >> 3), (uint32_t)g.SW) + (uint32_t)(k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(slab + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(slab + g.wslab / 2 + off) = lo;
  }
}

int g_base_offset_mode = -1;

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
  return (long)g.wbytes;
}

// Pack OIHW fp32 weights into the bf16 hi/lo slabs.  `w` may point at a column block of a wider [O][ldw][ks][ks]
// tensor (the per-branch blocks of the 1x1 projection).  (Cin, Cout) describe the GEMM being run:
// transpose=0 -> the forward conv of w[Cout][Cin][ks][ks];  transpose=1 -> its data gradient, i.e. a conv with
// Cin' = Cout(w), Cout' = Cin(w): pass Cin = Cout(w), Cout = Cin(w).
int hcm_tc_conv_pack(const float* w, int ldw, void* wpack, int B, int H, int W, int Cin, int Cout, int ks, int transpose,
                     cudaStream_t stream) {
  HCM_CHECK_ARG(w && wpack, "tc_conv_pack: null pointer");
  Geo g = make_geo(B, H, W, Cin, Cout, ks);
  HCM_CHECK_ARG(geo_ok(g, Cin, Cout, ks), "tc_conv_pack: unsupported geometry");
  const long total = (long)g.nslabs * g.Npad * g.KB;
  // ldw: second dimension of the OIHW tensor the slice lives in (its row stride / ks^2); 0 = contiguous
  const int wCin = ldw > 0 ? ldw : (transpose ? Cout : Cin);
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  tc_pack_kernel<<<grid, 256, 0, stream>>>(w, reinterpret_cast<uint8_t*>(wpack), g, Cin, Cout, ks, transpose, wCin);
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
  if (g_base_offset_mode < 0) {
    const char* e = getenv("HCM_TC_BASE_OFFSET");
    g_base_offset_mode = e ? atoi(e) : 0;   // measured on B200: the swizzle XOR uses absolute smem address bits
  }
  p.base_offset_mode = g_base_offset_mode;
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    dbg_on = getenv("HCM_TC_DEBUG") ? 1 : 0;
    if (dbg_on) cudaMalloc(&dbg, 148 * 16 * sizeof(long long));
  }
  p.dbg = dbg;
  if (dbg_on) cudaMemsetAsync(dbg, 0, 148 * 16 * sizeof(long long), stream);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { hcm_set_error("tc_conv: smem attribute: %s", cudaGetErrorString(e)); return HCM_ERR_CUDA; }
    configured = true;
  }
  tc_conv_kernel<<<(unsigned)p.g.grid, NTHREADS, p.g.smem, stream>>>(p);
  HCM_LAUNCH_CHECK("tc_conv");
  if (dbg_on) {
    long long h[148 * 16];
    cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
    const long long* o = h;      // CTA 0
    fprintf(stderr, "[tc_conv dbg] %dx%d %d->%d k%d tiles/cta %ld | mma: total %lld wait_w %lld wait_acc %lld wait_a %lld | "
            "transform: total %lld wait %lld | epilogue: total %lld wait %lld\n", H, W, Cin, Cout, ks,
            (p.g.tiles + p.g.grid - 1) / p.g.grid, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
  }
  return HCM_OK;
}

}  // extern "C"
