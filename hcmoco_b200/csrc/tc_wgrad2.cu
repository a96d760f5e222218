// Tensor-core weight gradient (tcgen05 / TMEM, bf16 hi/lo split) of the 3x3 (stride 1 or 2) and 1x1 convolutions:
//     dw[co][ci][r][s] += sum over output positions p of  dy[p][co] * T(x)[src(p, r, s)][ci]
// the cuDNN wgrad behind nn.Conv2d backward (networks/official_hrnet/official_hrnet.py:26-29, 68-75, 187-216, 336-357).
//
// Same virtual flat position space and the same swizzled channels-last staged tiles as tc_conv.cu.  Per 128-position
// tile the transform teams stage  dy -> Dy[pos][co]  (non-interior positions zeroed)  and  T(x) -> A[halo pos][ci]
// (BN scale/shift(+ReLU) applied on load, padding zeroed; stride 2: four parity planes side by side).  Both are read
// MN-major (channels contiguous, positions = K): one tcgen05.mma M=128 (co) x N x K=16 positions.
// For stride 1 the three taps of a filter row are ONE MMA: the N blocks of the B operand are LBO = one staged row apart
// (tap s+1 = next halo row), N = 3 x 32 channels; the accumulators D[co][tap*32 + ci] (288 TMEM columns) stay
// resident while the CTA walks its range of tiles; the partial dw is added to global memory with fp32 atomics.
// Grid = (tile ranges, splits of 32 (3x3) / 64 (1x1) input channels, output-channel blocks of 128).
#include "tc_common.cuh"

namespace {

constexpr int TILE = 128;
constexpr int NTRANS = 512;
constexpr int NTHREADS_W = NTRANS + 128 + 32;   // 16 transform warps, 4 epilogue warps, last warp: MMA issuer + TMEM owner
constexpr int W_EPI = NTRANS / 32, W_MMA = W_EPI + 4;
constexpr int HDR = 4096;
constexpr int MAX_ASTAGE = 4;

struct WGeo {
  int stride, ks, taps, nq, Hp, Wp, Ho, Wo, center, L, Lpad;
  int CI, nsplit, nblk, SWa, SWd, a_blocks, d_blocks, ntr, tiles_per, tmem_cols, nstage, Va, Vd;
  long Mv, T;
  FastDiv fd_hw, fd_wp;   // position decode: / (Hp*Wp), / Wp
  size_t a_plane, d_plane, a_bytes, d_bytes, stage_bytes, tab_bytes, smem;
};

WGeo make_wgeo(int B, int H, int W, int Cin, int Cout, int ks, int stride) {
  WGeo g;
  g.stride = stride; g.ks = ks; g.taps = ks * ks;
  g.nq = (stride == 2) ? 4 : 1;
  g.Ho = (stride == 2) ? H / 2 : H;
  g.Wo = (stride == 2) ? W / 2 : W;
  if (ks == 1) { g.Hp = H; g.Wp = W; g.center = 0; g.L = TILE; }
  else if (stride == 1) { g.Hp = H + 2; g.Wp = W + 2; g.center = g.Wp + 1; g.L = TILE + 2 * (g.Wp + 1); }
  else { g.Hp = g.Ho + 1; g.Wp = g.Wo + 1; g.center = g.Wp + 1; g.L = TILE + g.Wp + 1; }
  g.Mv = (long)B * g.Hp * g.Wp;
  g.fd_hw = make_fastdiv((uint32_t)(g.Hp * g.Wp));
  g.fd_wp = make_fastdiv((uint32_t)g.Wp);
  g.T = (g.Mv + TILE - 1) / TILE;
  g.Lpad = ceil_to(g.L, 16);
  g.CI = (ks == 3) ? 32 : 64;                          // input channels per CTA split
  g.nsplit = (Cin + g.CI - 1) / g.CI;
  g.nblk = (Cout + 127) / 128;
  const int sc = g.nq * g.CI;                          // staged channels per halo row
  g.SWa = (sc <= 32) ? 64 : 128;
  g.a_blocks = (sc * 2 + g.SWa - 1) / g.SWa;
  g.a_plane = (size_t)g.Lpad * g.SWa;
  g.a_bytes = 2 * g.a_blocks * g.a_plane;              // hi + lo
  const int cd16 = ceil_to(Cout < 128 ? Cout : 128, 16);
  g.SWd = (cd16 <= 32) ? 64 : 128;
  g.d_blocks = (cd16 * 2 + g.SWd - 1) / g.SWd;         // blocks actually staged (M=128 reads 256/SWd blocks)
  g.d_plane = (size_t)TILE * g.SWd;
  g.d_bytes = 2 * g.d_blocks * g.d_plane;
  g.stage_bytes = g.d_bytes + g.a_bytes;
  g.Va = (Cin % 4 == 0) ? 4 : 2;
  g.Vd = (Cout % 4 == 0) ? 4 : 2;
  int c = 32;
  while (c < g.taps * g.CI) c <<= 1;
  g.tmem_cols = c;
  const size_t tab1 = (size_t)ceil_to((g.Lpad * g.nq + TILE) * 4, 1024);
  // M=128 reads 256/SWd dy planes whatever Cout is (rows >= Cout are garbage and never read back): keep those reads in bounds
  const size_t reach = (size_t)g.d_blocks * g.d_plane + (size_t)(256 / g.SWd) * g.d_plane;   // from the dy_lo start
  const size_t slack = reach > g.stage_bytes ? reach - g.stage_bytes : 0;
  const size_t total = 226 * 1024 - HDR - slack;
  g.nstage = 1;
  for (int n = MAX_ASTAGE; n >= 2; n >>= 1)
    if (n * (g.stage_bytes + tab1) <= total) { g.nstage = n; break; }
  g.tab_bytes = g.nstage * tab1;
  g.smem = HDR + g.tab_bytes + g.nstage * g.stage_bytes + slack;
  long want = 148L / ((long)g.nsplit * g.nblk);
  if (want < 1) want = 1;
  if (want > g.T) want = g.T;
  g.tiles_per = (int)((g.T + want - 1) / want);
  g.ntr = (int)((g.T + g.tiles_per - 1) / g.tiles_per);
  return g;
}

bool wgeo_ok(const WGeo& g, int H, int W, int Cin, int Cout, int ks, int stride) {
  if (!((ks == 3 && (stride == 1 || stride == 2)) || (ks == 1 && stride == 1))) return false;
  if (stride == 2 && ((H | W) & 1)) return false;
  return Cin <= 256 && Cout <= 256 && (Cin % 2) == 0 && (Cout % 2) == 0 && g.tmem_cols <= 512 && g.smem <= 227 * 1024 &&
         g.Mv < (1L << 31) && g.Lpad <= 512;
}

struct WParams {
  const float* x;
  const float* in_scale;
  const float* in_shift;
  int in_relu;
  const float* dy;
  float* dw;
  int B, H, W, Cin, Cout, lddw;
  long long* dbg;              // HCM_TC_DEBUG=1: per-CTA cycle counters of the three roles
  WGeo g;
};

__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t kb, uint32_t SW) {
  const uint32_t off = row * SW;
  return off + ((((kb >> 4) ^ (off >> 7)) & (SW / 16 - 1)) << 4) + (kb & 15);
}
// MN-major swizzled operand descriptor: rows (K = positions) of SW bytes, 8-row groups SBO = 8*SW apart, MN blocks LBO apart
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t SW, uint32_t lbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(((8 * SW) >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(SW == 128 ? 2 : (SW == 64 ? 4 : 6)) << 61;
  return d;
}

// virtual position -> input pixel of plane q (or -1), and -> output pixel (or -1); same grids as tc_conv.cu
__device__ __forceinline__ void virt_decode(long pv, const WParams& p, int& src0, int& src1, int& src2, int& src3, int& dst) {
  const WGeo& g = p.g;
  src0 = src1 = src2 = src3 = dst = -1;
  if (pv < 0 || pv >= g.Mv) return;
  if (g.ks == 1) { src0 = (int)pv; dst = (int)pv; return; }
  const unsigned v = (unsigned)pv, hw = (unsigned)(g.Hp * g.Wp);
  const unsigned b = fdiv(v, g.fd_hw), rem = v - b * hw;
  const unsigned row = fdiv(rem, g.fd_wp), col = rem - row * (unsigned)g.Wp;
  if (g.stride == 1) {
    if (row < 1 || row > (unsigned)p.H || col < 1 || col > (unsigned)p.W) return;
    src0 = (int)((b * p.H + row - 1) * p.W + (col - 1));
    dst = src0;
    return;
  }
  if (row < 1 || col < 1) return;
  const unsigned yi = 2 * (row - 1), xi = 2 * (col - 1);
  src0 = (int)((b * p.H + yi) * p.W + xi);
  src1 = src0 + 1; src2 = src0 + p.W; src3 = src0 + p.W + 1;
  dst = (int)((b * g.Ho + row - 1) * g.Wo + (col - 1));
}

__global__ void __launch_bounds__(NTHREADS_W, 1) tc_wgrad2_kernel(const WParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const WGeo& g = p.g;
  // barriers: 0..3 full, 4..7 empty, 8 done
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 512);
  float* s_sc = reinterpret_cast<float*>(smem + 1024);
  float* s_sh = s_sc + 256;
  int* s_tab = reinterpret_cast<int*>(smem + HDR);           // per stage: [Lpad][nq] source pixels, then [128] output pixels
  const int tab_stride = (int)(g.tab_bytes / g.nstage / 4);
  uint8_t* Sbase = smem + HDR + g.tab_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  const int ci_lo = blockIdx.y * g.CI;
  const int ci_n = min(g.CI, p.Cin - ci_lo);                 // real input channels of this split
  const int co_lo = blockIdx.z * 128;
  const int co_n = min(128, p.Cout - co_lo);
  const uint32_t SWa = (uint32_t)g.SWa, SWd = (uint32_t)g.SWd;
  const uint32_t d_lo = (uint32_t)g.d_blocks * (uint32_t)g.d_plane;          // dy_lo planes follow dy_hi planes
  const uint32_t a_off = (uint32_t)g.d_bytes;                                 // a-tile follows the dy tile in a stage
  const uint32_t a_lo = (uint32_t)g.a_blocks * (uint32_t)g.a_plane;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_ASTAGE; ++s) { mbar_init(BAR(s), NTRANS / g.nstage); mbar_init(BAR(4 + s), 1); }
    mbar_init(BAR(8), 1);
    fence_mbar_init();
  }
  // zero all staged tiles once (channel padding is never written afterwards and must read as 0)
  for (size_t i = threadIdx.x; i < ((size_t)g.nstage * g.stage_bytes) / 16; i += NTHREADS_W)
    reinterpret_cast<uint4*>(Sbase)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == W_MMA) tmem_alloc(smem_u32(tmem_ptr), g.tmem_cols);
  pdl_wait();                    // (PDL protocol, common.cuh: only shared memory / TMEM / parameters were touched so far)
  pdl_trigger();
  for (int c = threadIdx.x; c < 256; c += NTHREADS_W) {
    s_sc[c] = (p.in_scale && c < p.Cin) ? p.in_scale[c] : 1.f;
    s_sh[c] = (p.in_scale && c < p.Cin) ? p.in_shift[c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  const long t_beg = (long)blockIdx.x * g.tiles_per;
  const long t_end = min(g.T, t_beg + g.tiles_per);
  const int ntiles = (int)(t_end - t_beg);

  if (warp == W_MMA) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
    const uint32_t s0 = smem_u32(Sbase);
    const uint64_t dy_t = mn_desc(0, SWd, (uint32_t)g.d_plane);
    const bool concat = (g.stride == 1 && g.ks == 3);          // 3 taps of a filter row = N blocks one row apart
    const uint64_t a_t = mn_desc(0, SWa, concat ? SWa : (uint32_t)g.a_plane);
    const uint32_t idesc = instr_desc(concat ? 3 * g.CI : g.CI, 1);
    const int nmma = concat ? 3 : g.taps;                       // MMAs per K step and pass
    // One staged dy plane (Cout block <= SWd/2 channels): the M=128 operand read through the dy_hi descriptor covers the dy_lo
    // plane as its next MN block, so accumulator rows [SWd/2, SWd/2 + co_n) already hold dy_lo*(a_hi + a_lo).  The third
    // MMA (dy_lo*a_hi) is then redundant: two MMAs per step, and the epilogue folds the two row blocks.
    const bool fold = (g.d_blocks == 1);
    uint32_t* tap_boff = reinterpret_cast<uint32_t*>(smem + 3072);  // byte offset of the B operand / TMEM column per MMA
    uint32_t* tap_col = tap_boff + 16;
    for (int m = lane; m < 9; m += 32) {
      tap_boff[m] = 0; tap_col[m] = 0;
      if (m >= nmma) continue;
      if (concat) { tap_boff[m] = (uint32_t)(m * g.Wp) * SWa; tap_col[m] = (uint32_t)(m * 3 * g.CI); }
      else if (g.ks == 3) {
        const int r = m / 3, sx = m - 3 * r;
        const int q = ((r == 1) ? 0 : 2) + ((sx == 1) ? 0 : 1);
        const uint32_t rows = (uint32_t)(((r == 0) ? 0 : 1) * g.Wp + ((sx == 0) ? 0 : 1));
        const uint32_t kb = (uint32_t)(q * g.CI * 2);           // plane q = bytes [q*CI*2, (q+1)*CI*2) of the staged row
        tap_boff[m] = rows * SWa + (kb / SWa) * (uint32_t)g.a_plane + (kb % SWa);
        tap_col[m] = (uint32_t)(m * g.CI);
      }
    }
    __syncwarp();
    long long c_all = clock64(), c_wait = 0, tq = 0;
    // lean issue loop (see umma_bf16_w in tc_common.cuh): descriptor low words = stage base + K-step offset + tap offset
    const uint32_t dy_hi32 = (uint32_t)(dy_t >> 32), a_hi32 = (uint32_t)(a_t >> 32);
    const uint32_t dlo16 = d_lo >> 4, alo16 = a_lo >> 4, kd16 = (16u * SWd) >> 4, ka16 = (16u * SWa) >> 4;
    uint32_t boff16[9], dcolv[9];
#pragma unroll
    for (int m = 0; m < 9; ++m) { boff16[m] = tap_boff[m] >> 4; dcolv[m] = tmem + tap_col[m]; }
    const bool dbg = p.dbg != nullptr;
    int s = 0;                                                   // stage and its phase, advanced incrementally (no divisions)
    uint32_t sph = 0;
    for (int it = 0; it < ntiles; ++it) {
      if (dbg) tq = clock64();
      mbar_wait(BAR(s), sph);
      if (dbg) c_wait += clock64() - tq;
      tc_fence_after();
      const uint32_t st = s0 + (uint32_t)s * (uint32_t)g.stage_bytes;
      if (elect_one()) {
        uint32_t dyh = (uint32_t)dy_t + (st >> 4), ab = (uint32_t)a_t + ((st + a_off) >> 4);
        for (int k = 0; k < TILE / 16; ++k, dyh += kd16, ab += ka16) {
          const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
          if (nmma == 3) {
#pragma unroll
            for (int m = 0; m < 3; ++m) {
              const uint32_t ah = ab + boff16[m];
              umma_bf16_w(dcolv[m], dyh, dy_hi32, ah, a_hi32, idesc, acc);
              umma_bf16_acc(dcolv[m], dyh, dy_hi32, ah + alo16, a_hi32, idesc);
              if (!fold) umma_bf16_acc(dcolv[m], dyh + dlo16, dy_hi32, ah, a_hi32, idesc);
            }
          } else {
#pragma unroll
            for (int m = 0; m < 9; ++m) {
              if (m < nmma) {
                const uint32_t ah = ab + boff16[m];
                umma_bf16_w(dcolv[m], dyh, dy_hi32, ah, a_hi32, idesc, acc);
                umma_bf16_acc(dcolv[m], dyh, dy_hi32, ah + alo16, a_hi32, idesc);
                if (!fold) umma_bf16_acc(dcolv[m], dyh + dlo16, dy_hi32, ah, a_hi32, idesc);
              }
            }
          }
        }
        umma_commit(BAR(4 + s));
      }
      __syncwarp();
      if (++s == g.nstage) { s = 0; sph ^= 1u; }
    }
    if (elect_one()) umma_commit(BAR(8));
    if (p.dbg && lane == 0) { long long* o = p.dbg + (long)blockIdx.x * 8; o[0] = clock64() - c_all; o[1] = c_wait; }
  } else if (warp < W_EPI) {
    // ===== transform teams: team k stages tiles k, k+nstage, ... into stage k =====
    const int TS = NTRANS / g.nstage;
    const int team = threadIdx.x / TS, t = threadIdx.x - team * TS;
    long long c_all = clock64(), c_wait = 0, c_tab = 0, tq = 0;
    const bool dbg = p.dbg != nullptr;
    uint32_t eph = 1;
    for (int it = team; it < ntiles; it += g.nstage, eph ^= 1u) {
      const int s = team;
      if (dbg) tq = clock64();
      mbar_wait(BAR(4 + s), eph);
      if (dbg) { c_wait += clock64() - tq; tq = clock64(); }
      const long tile0 = (t_beg + it) * TILE;
      int* tab = s_tab + s * tab_stride;
      int* dtab = tab + g.Lpad * g.nq;
      for (int pos = t; pos < g.Lpad; pos += TS) {
        int s0, s1, s2, s3, d;
        virt_decode((pos < g.L) ? tile0 - g.center + pos : -1, p, s0, s1, s2, s3, d);
        if (g.nq == 1) tab[pos] = s0;
        else { tab[pos * 4 + 0] = s0; tab[pos * 4 + 1] = s1; tab[pos * 4 + 2] = s2; tab[pos * 4 + 3] = s3; }
        const int m = pos - g.center;
        if (m >= 0 && m < TILE) dtab[m] = d;
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(TS) : "memory");
      if (dbg) c_tab += clock64() - tq;
      uint8_t* st = Sbase + (size_t)s * g.stage_bytes;
      // dy tile: channels [co_lo, co_lo + co_n)
      if (g.Vd == 4) stage_rows8<4>(p.dy, p.Cout, co_lo, co_n, TILE, dtab, 1, 0, st, d_lo, (uint32_t)g.d_plane, SWd, 0, nullptr, nullptr, false, 0, t, TS);
      else stage_rows8<2>(p.dy, p.Cout, co_lo, co_n, TILE, dtab, 1, 0, st, d_lo, (uint32_t)g.d_plane, SWd, 0, nullptr, nullptr, false, 0, t, TS);
      // T(x) halo: channels [ci_lo, ci_lo + ci_n) of each parity plane
      for (int q = 0; q < g.nq; ++q) {
        if (g.Va == 4) stage_rows8<4>(p.x, p.Cin, ci_lo, ci_n, g.Lpad, tab, g.nq, q, st + a_off, a_lo, (uint32_t)g.a_plane, SWa,
                                     (uint32_t)(q * g.CI * 2), s_sc, s_sh, p.in_scale != nullptr, p.in_relu, t, TS);
        else stage_rows8<2>(p.x, p.Cin, ci_lo, ci_n, g.Lpad, tab, g.nq, q, st + a_off, a_lo, (uint32_t)g.a_plane, SWa,
                           (uint32_t)(q * g.CI * 2), s_sc, s_sh, p.in_scale != nullptr, p.in_relu, t, TS);
      }
      fence_proxy_async();
      mbar_arrive(BAR(s));
    }
    if (p.dbg && threadIdx.x == 0) { long long* o = p.dbg + (long)blockIdx.x * 8; o[2] = clock64() - c_all; o[3] = c_wait; o[4] = c_tab; }
  }
  // ===== drain: D[co][tap*CI + ci] -> dw.  EVERY warp helps (a warp may read the TMEM lane quarter warp%4): the 4 dedicated
  // warps alone need ~6k dependent instructions each for this, 13-26 us of a 45-75 us kernel (measured); 20 warps share it.
  // The CTA's block of dw is assembled in shared memory in the GLOBAL layout [co][ci][tap] (the staging tiles are free once
  // every MMA has completed), then added to global memory with coalesced atomics.
  {
    const long long c_all = clock64();
    mbar_wait(BAR(8), 0);
    const long long c_w = clock64() - c_all;
    if (p.dbg && warp == W_EPI && lane == 0) { long long* o = p.dbg + (long)blockIdx.x * 8; o[6] = c_w; o[7] = clock64(); }
    tc_fence_after();
    const int q = warp & 3, wsub = warp >> 2;                     // lane quarter; helper index among the warps of that quarter
    constexpr int NSUB = (NTHREADS_W / 32) / 4;                   // 5 full helper sets (warps 0..19); the MMA warp sits out
    const bool helper = wsub < NSUB;
    const int m = q * 32 + lane, bw = (int)SWd / 2;
    // accumulator row -> output channel: rows [0, co_n) are dy_hi products; with a single dy plane rows [bw, bw + co_n) are
    // the dy_lo products of the same channels (see `fold` in the MMA warp)
    const int cr = (m < co_n) ? m : ((g.d_blocks == 1 && m >= bw && m - bw < co_n) ? m - bw : -1);
    const bool rowok = cr >= 0 && ntiles > 0;
    // channels of this split that exist in dw (lddw < Cin: the input was stored with zero-padded channels, e.g. the 3-channel stem)
    const int ci_w = max(0, min(ci_n, p.lddw - ci_lo));
    const int RL = ci_w * g.taps, RLp = RL | 1;                  // row length / odd row pitch (floats)
    float* red = reinterpret_cast<float*>(Sbase);
    const int nch = (ci_n + 15) / 16;                            // 16-column chunks per tap
    if ((size_t)co_n * RLp * 4 <= (size_t)g.nstage * g.stage_bytes) {
      // pass 0: the dy_hi rows store their values; pass 1 (single dy plane only): the dy_lo rows add theirs.  Every slot has
      // exactly one writer per pass, so plain shared-memory stores suffice (a shared fp32 atomicAdd is a CAS loop)
      const int npass = (g.d_blocks == 1) ? 2 : 1;
      for (int pass = 0; pass < npass; ++pass) {
        const bool mine = rowok && ((pass == 0) == (m < co_n));
        if (helper) {
          for (int j = wsub; j < g.taps * nch; j += NSUB) {
            const int tap = j / nch, c0 = (j - tap * nch) * 16;
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(tap * g.CI + c0), v);
            if (mine) {
              float* o = red + cr * RLp + c0 * g.taps + tap;
              if (c0 + 16 <= ci_w) {
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i * g.taps] = (pass == 0) ? v[i] : o[i * g.taps] + v[i];
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (c0 + i < ci_w) o[i * g.taps] = (pass == 0) ? v[i] : o[i * g.taps] + v[i];
              }
            }
          }
        }
        __syncthreads();
      }
      if (ntiles > 0) {
        for (int e = threadIdx.x; e < co_n * RL; e += NTHREADS_W) {
          const int row = e / RL, col = e - row * RL;
          atomicAdd(p.dw + ((long)(co_lo + row) * p.lddw + ci_lo) * g.taps + col, red[row * RLp + col]);
        }
      }
    } else if (helper) {
      // (block too large for the staging area: direct scatter)
      const int co = co_lo + cr;
      for (int j = wsub; j < g.taps * nch; j += NSUB) {
        const int tap = j / nch, c0 = (j - tap * nch) * 16;
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(tap * g.CI + c0), v);
        if (rowok) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int ci = ci_lo + c0 + i;
            if (c0 + i < ci_w) atomicAdd(p.dw + ((long)co * p.lddw + ci) * g.taps + tap, v[i]);
          }
        }
      }
    }
  }
  if (p.dbg && warp == W_EPI && lane == 0) { long long* o = p.dbg + (long)blockIdx.x * 8; o[5] = clock64(); }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, g.tmem_cols);
}

}  // namespace

extern "C" {

int hcm_tc_wgrad_supported(int B, int H, int W, int Cin, int Cout, int ks, int stride) {
  if (!((ks == 3 && (stride == 1 || stride == 2)) || (ks == 1 && stride == 1))) return 0;
  WGeo g = make_wgeo(B, H, W, Cin, Cout, ks, stride);
  return wgeo_ok(g, H, W, Cin, Cout, ks, stride) ? 1 : 0;
}

// dw[Cout,Cin,ks,ks] += sum_pixels dy * T(x)   (fp32 atomics across CTAs; the caller zeroes dw once per step).
// lddw > 0: dw is a column block of a wider [Cout][lddw][ks][ks] tensor; lddw < Cin: dw has only lddw input channels (x is stored
// with zero-padded channels): the gradient of the padding channels is dropped
int hcm_tc_wgrad(const float* x, const float* dy, float* dw, int lddw, int B, int H, int W, int Cin, int Cout, int ks, int stride,
                 const float* in_scale, const float* in_shift, int in_relu, cudaStream_t stream) {
  HCM_CHECK_ARG(x && dy && dw, "tc_wgrad: null pointer");
  HCM_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "tc_wgrad: in_scale/in_shift must come together");
  WParams p;
  p.g = make_wgeo(B, H, W, Cin, Cout, ks, stride);
  HCM_CHECK_ARG(wgeo_ok(p.g, H, W, Cin, Cout, ks, stride), "tc_wgrad: unsupported geometry (Cin=%d Cout=%d ks=%d stride=%d)", Cin,
                Cout, ks, stride);
  p.x = x; p.in_scale = in_scale; p.in_shift = in_shift; p.in_relu = in_relu; p.dy = dy; p.dw = dw;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.lddw = lddw > 0 ? lddw : Cin;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { hcm_set_error("tc_wgrad: smem attribute: %s", cudaGetErrorString(e)); return HCM_ERR_CUDA; }
    configured = true;
  }
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    dbg_on = getenv("HCM_TC_DEBUG") ? 1 : 0;
    if (dbg_on) cudaMalloc(&dbg, 1024 * 8 * sizeof(long long));
  }
  p.dbg = dbg;
  dim3 grid(p.g.ntr, p.g.nsplit, p.g.nblk);
  hcm_launch_pdl(tc_wgrad2_kernel, grid, dim3(NTHREADS_W), p.g.smem, stream, p);
  HCM_LAUNCH_CHECK("tc_wgrad");
  if (dbg_on) {
    long long h[8];
    cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc_wgrad dbg] %dx%d %d->%d k%d s%d tiles/cta %d stages %d | mma: total %lld wait %lld | transform(team 0): total %lld "
            "wait %lld table %lld | epilogue: wait %lld drain %lld\n", H, W, Cin, Cout, ks, stride, p.g.tiles_per, p.g.nstage, h[0], h[1],
            h[2], h[3], h[4], h[6], h[5] - h[7]);
  }
  return HCM_OK;
}

}  // extern "C"
