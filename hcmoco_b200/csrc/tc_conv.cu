// Tensor-core (tcgen05 / TMEM) implicit-GEMM convolution for sm_100a: 3x3 (stride 1 or 2) and 1x1, NHWC fp32
// activations, fp32-accurate through a two-term bf16 split (x = hi + lo: hi*hi + hi*lo + lo*hi, fp32 accumulation in
// TMEM; SURVEY.md F8 — single-pass TF32/BF16 misses the 1e-3 parity bar, the split passes it).  Replaces the cuDNN
// calls behind nn.Conv2d in networks/official_hrnet/official_hrnet.py:26-29, 68-75, 187-216, 336-357 (forward and,
// for stride 1 with transposed+flipped packed weights, the data gradient).
//
// Formulation.  Output positions live on a flat "virtual" grid: [B][H+2][W+2] for 3x3/s1 (zero padding included),
// [B][Ho+1][Wo+1] for 3x3/s2 (one padding row/column, top/left), [B][H][W] for 1x1.  A tile is 128 consecutive virtual
// positions, so a filter tap is a constant ROW offset into the staged halo.  For stride 2 the halo is staged
// space-to-depth: 4 parity planes of the input side by side on the channel axis, tap (r,s) = (plane, row offset).
// Persistent CTAs (one per SM) walk the tiles:
//   transform warps (256 thr)  gather the halo once from global memory with fully coalesced 8/16-byte accesses,
//                              apply the producer's pending BatchNorm scale/shift(+ReLU), zero the padding, split to bf16
//                              hi/lo and store a swizzled channels-last tile (K-major SWIZZLE_64B/128B UMMA layout);
//   TMA warp                   weights, pre-packed per K=16 step as [w_hi rows ; w_lo rows]: one cp.async.bulk set when
//                              they fit in shared memory, else a ring;
//   MMA warp (one lane)        tcgen05.mma kind::f16 M=128.  An M=128 MMA costs ~100 cycles for any N <= 128 (measured),
//                              so when 2*ceil16(Cout) <= 256 the hi and lo weights are ONE operand (N = 2*Np):
//                              D[:, :Np] += A_hi*w_hi + A_lo*w_hi, D[:, Np:] += A_hi*w_lo  — two MMAs per K step, not three;
//   epilogue warps (128 thr)   tcgen05.ld (double-buffered accumulators), add the column halves, + bias / accumulate,
//                              store the interior positions (fp32 NHWC).
#include "tc_common.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int NTRANS = 512;                  // transform threads: warps 0..15 (the gather is latency-bound: 4 loads in flight each)
constexpr int NTHREADS = NTRANS + 128 + 64;  // then 4 epilogue warps, the TMA producer warp, and last the MMA issuer + TMEM owner
// (the SMSP arbiter favours the highest warp id: the single MMA-issuing warp must not starve behind busy transform warps)
constexpr int W_EPI = NTRANS / 32, W_TMA = W_EPI + 4, W_MMA = W_TMA + 1;
constexpr int MAXG = 16;
constexpr int MAX_LPAD = 1024;
constexpr int HDR_BYTES = 4096 + 6144 + 2048;   // barriers / tmem ptr / scale / shift | row-concat boundary exchange (2 x 3 warps x 3 rows x 80 floats) | issue schedule (descriptor low words); the per-stage source tables follow
constexpr int MAX_ASTAGE = 4;
constexpr int MAX_UNITS = 256;
constexpr int MAX_STEPS = 144;                // 9 taps x 256/16 channels

struct Geo {
  int mode;          // 0: convolution; 1: data gradient of a 3x3 stride-2 convolution (see hcm_tc_dgrad_s2)
  int Cp;            // mode 1: padded original Cin = width of one output parity block
  int nqs;           // mode 1: output parities handled per launch (4, 2 or 1 so that N = nqs*Cp <= 256)
  int stride, ks, taps, nq, Hp, Wp, Ho, Wo, center, L, Lpad, Npad, Cin16, SC, SW, KB, cg, ngroups, nblk, nsteps;
  int concat, acc_cols, acc_stages, tmem_cols, nastage, w_resident, wst, spb, grid, V;
  int rc;            // row-concatenated taps (3x3 stride 1, 3*Npad <= 256): see make_geo
  int NB;            // N of one weight operand: 3*Npad (rc) or Npad
  int TM;            // output positions per tile: 126 (rc) or 128
  int st;            // tiles per staged fill ("supertile", see make_geo): 1, or 2..4 consecutive tiles sharing one halo
  int ksplit;        // 1, or 2: the filter taps of a tile are split over two CTAs (work unit = (tile, part)); see make_geo
  long Mv, tiles, units;
  FastDiv fd_hw, fd_wp;   // position decode: / (Hp*Wp), / Wp
  size_t plane_bytes, a_stage_bytes, wslab, wbytes, smem, tab_bytes;
};

// Row concatenation of the filter-row taps is implemented and parity-green (tests/test_kernels_gpu.py::test_tc_conv_rowcat) but
// OFF by default.  Measured on B200 (64x64 18->18, B=64, HCM_TC_DEBUG counters): the MMA warp's busy time drops from 53k to 32k
// cycles per CTA (12 instead of 36 MMAs per tile), but the 4 epilogue warps, which now read 3 accumulator blocks per chunk and
// exchange boundary rows through a named barrier, become latency-bound at 9.7k cycles per tile (4.4k before): 87 us vs 42 us.
// Enabling it needs two epilogue warp groups (one per accumulator stage) and TMEM loads issued a chunk ahead.
// HCM_TC_ROWCAT=1 switches it on for experiments.
bool rowcat_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("HCM_TC_ROWCAT"); on = (e && e[0] == '1') ? 1 : 0; }
  return on == 1;
}
bool rowcat_ok(int Cout, int ks, int stride, int mode) {
  return rowcat_enabled() && mode == 0 && ks == 3 && stride == 1 && 3 * ceil_to(Cout, 16) <= 256;
}

Geo make_geo1(int B, int H, int W, int Cin, int Cout, int ks, int stride, int mode, int st);

// Supertiles.  A 3x3 tile of 128 flat positions needs a halo of 2*(Wp+1) more staged rows: 262 rows for 128 outputs on the 64x64
// maps (2.05x), and the transform warps — the instruction-issue hogs of the narrow layers — pay for every one of them.  When one
// channel group holds the layer (Cin16 * nq <= 64) and shared memory allows, a fill stages `st` CONSECUTIVE tiles at once
// (st*128 + halo rows: 1.26x at st = 4) and the MMA warp issues the st tiles from row offsets 0, 128, .. of the same staged block;
// the CTAs then own contiguous tile ranges.  HCM_TC_ST=1 switches it off, =2/3 caps it (experiments).
Geo make_geo(int B, int H, int W, int Cin, int Cout, int ks, int stride, int mode = 0) {
  static int st_max = -1;
  if (st_max < 0) { const char* e = getenv("HCM_TC_ST"); st_max = e ? atoi(e) : 4; if (st_max < 1) st_max = 1; if (st_max > 4) st_max = 4; }
  Geo g1 = make_geo1(B, H, W, Cin, Cout, ks, stride, mode, 1);
  const bool halo = (mode == 1) || ks == 3;
  if (!halo || g1.rc || g1.ngroups != 1 || g1.ksplit != 1) return g1;
  const long tpc = (g1.tiles + 147) / 148;                    // tiles per CTA
  for (int st = st_max; st >= 2; --st) {
    if (2 * st > tpc) continue;                               // at least two fills per CTA, or nothing overlaps
    Geo g = make_geo1(B, H, W, Cin, Cout, ks, stride, mode, st);
    if (g.ngroups == 1 && g.nastage >= 2 && g.w_resident == g1.w_resident && g.Lpad <= MAX_LPAD && g.smem <= 227 * 1024) return g;
  }
  return g1;
}

Geo make_geo1(int B, int H, int W, int Cin, int Cout, int ks, int stride, int mode, int st) {
  Geo g;
  g.st = st;
  // Row concatenation: the three taps (r, 0..2) of a filter row share ONE A operand start (row offset r*Wp) and are the
  // N blocks of one weight operand: E_s[m'] = sum_r A[m' + r*Wp] * W_(r,s).  The column shift moves to the epilogue,
  // out[m] = E_0[m] + E_1[m+1] + E_2[m+2] (warp shuffles + a 3-row exchange between the epilogue warps), so a tile yields
  // 126 outputs from 128 accumulator rows.  An M=128 tcgen05.mma costs ~100 cycles for any N <= 128 and N/2 above
  // (measured), so narrow layers need 3x fewer MMA slots: 18->18 goes from 36 to 12 MMAs per tile.
  g.rc = rowcat_ok(Cout, ks, stride, mode) ? 1 : 0;
  g.TM = g.rc ? TILE_M - 2 : TILE_M;
  g.mode = mode; g.Cp = 0; g.nqs = 0;
  g.stride = stride; g.ks = ks; g.taps = ks * ks;
  g.nq = (stride == 2) ? 4 : 1;
  g.Ho = (stride == 2) ? H / 2 : H;
  g.Wo = (stride == 2) ? W / 2 : W;
  if (ks == 1) { g.Hp = H; g.Wp = W; g.center = 0; g.L = TILE_M; }
  else if (stride == 1) { g.Hp = H + 2; g.Wp = W + 2; g.center = g.Wp + 1; g.L = TILE_M + 2 * (g.Wp + 1) - (g.rc ? 2 : 0); }
  else { g.Hp = g.Ho + 1; g.Wp = g.Wo + 1; g.center = g.Wp + 1; g.L = TILE_M + g.Wp + 1; }
  if (mode == 1) {
    // H, W are the OUTPUT (= conv input) size; the staged tensor is dy [B,H/2,W/2,Cin]; virtual grid padded bottom/right
    g.stride = 1; g.ks = 2; g.taps = 4; g.nq = 1;
    g.Ho = H / 2; g.Wo = W / 2;
    g.Hp = g.Ho + 1; g.Wp = g.Wo + 1; g.center = 0; g.L = TILE_M + g.Wp + 1;
    g.Cp = ceil_to(Cout, 16);
    g.nqs = (4 * g.Cp <= 256) ? 4 : ((2 * g.Cp <= 256) ? 2 : 1);
  }
  g.Mv = (long)B * g.Hp * g.Wp;
  g.fd_hw = make_fastdiv((uint32_t)(g.Hp * g.Wp));
  g.fd_wp = make_fastdiv((uint32_t)g.Wp);
  g.tiles = (g.Mv + g.TM - 1) / g.TM;
  g.Lpad = ceil_to((st - 1) * g.TM + g.L, 16);
  g.Npad = (mode == 1) ? g.nqs * g.Cp : ceil_to(Cout, 16);
  g.Cin16 = ceil_to(Cin, 16);
  g.SC = g.nq * g.Cin16;                      // staged channels per position
  g.V = (Cin % 4 == 0) ? 4 : 2;
  g.SW = (g.SC <= 32) ? 64 : 128;             // bytes per staged row block = swizzle span
  g.KB = g.SW / 2;
  g.plane_bytes = (size_t)g.Lpad * g.SW;
  g.NB = g.rc ? 3 * g.Npad : g.Npad;
  g.nsteps = (g.rc ? 3 : g.taps) * (g.Cin16 / 16);
  g.wslab = (size_t)64 * g.NB;                // [2 K-chunks][2*NB rows: hi then lo][16 B]
  g.wbytes = (size_t)g.nsteps * g.wslab;
  const size_t tab1 = (size_t)ceil_to(g.Lpad * g.nq * 4, 1024);           // source-pixel table of one stage
  const size_t wmin = g.wbytes < 8 * g.wslab ? g.wbytes : 8 * g.wslab;
  const size_t total = 225 * 1024 - HDR_BYTES;
  // channel groups: one A stage holds `cg` staged channels (hi + lo planes), a multiple of KB when there are several.
  // As many blocks per group as 84 KB hold, but never so many that only ONE stage fits beside a minimal weight ring: with a
  // single stage the fill of group g+1 waits for the MMAs of group g (measured, 8x8 144->144: fill 5k -> MMA 15k -> fill 5k ->
  // MMA 2k cycles, strictly serial); two smaller stages let the fills run side by side / under the MMAs.
  int max_blk = (int)((84 * 1024) / (2 * g.plane_bytes));
  if (max_blk < 1) max_blk = 1;
  const int nblk_all = (g.SC + g.KB - 1) / g.KB;
  for (;; --max_blk) {
    g.ngroups = (nblk_all + max_blk - 1) / max_blk;
    g.nblk = (nblk_all + g.ngroups - 1) / g.ngroups;
    g.a_stage_bytes = (size_t)2 * g.nblk * g.plane_bytes;
    if (max_blk == 1 || 2 * (g.a_stage_bytes + tab1) + wmin <= total) break;
  }
  g.cg = g.nblk * g.KB;
  // (row-concat keeps hi and lo weights as separate operands: the epilogue then reads 3, not 6, accumulator blocks per chunk)
  g.concat = (!g.rc && 2 * g.NB <= 256) ? 1 : 0;
  g.acc_cols = g.concat ? 2 * g.NB : g.NB;
  g.acc_stages = (2 * g.acc_cols <= 512) ? 2 : 1;
  int c = 32;
  while (c < g.acc_stages * g.acc_cols) c <<= 1;
  g.tmem_cols = c;
  // A stages: each is filled by its own team of transform warps, so several halo gathers are in flight at once (the
  // gather of a narrow layer is pure latency: ~5 loads per thread); 2..4 stages as shared memory allows
  g.nastage = 1;
  for (int n = MAX_ASTAGE; n >= 2; n >>= 1) {      // 4 or 2: teams of NTRANS/n threads (whole warps)
    const size_t w = (g.wbytes + n * (g.a_stage_bytes + tab1) <= total) ? g.wbytes : wmin;   // prefer resident weights
    if (n * (g.a_stage_bytes + tab1) + w <= total && (n == 2 || g.wbytes + n * (g.a_stage_bytes + tab1) <= total)) { g.nastage = n; break; }
  }
  g.tab_bytes = g.nastage * tab1;
  const size_t budget = total - g.tab_bytes;
  const size_t left_b = budget > g.nastage * g.a_stage_bytes ? budget - g.nastage * g.a_stage_bytes : 0;
  g.w_resident = g.wbytes <= left_b ? 1 : 0;
  // ring: one barrier per chunk of `spb` consecutive schedule steps (an mbarrier wait costs ~90 cycles: per-step waits
  // would dominate the ~100 cycles of MMA work a step carries)
  g.spb = 1;
  g.wst = 0;
  if (!g.w_resident) {
    int spb = (int)((24 * 1024) / g.wslab);
    if (spb < 1) spb = 1;
    if (spb > 8) spb = 8;
    while (spb > 1 && left_b / (spb * g.wslab) < 3) --spb;
    g.spb = spb;
    const size_t slots = left_b / (spb * g.wslab);
    g.wst = (int)(slots > 16 ? 16 : slots);
  }
  g.smem = HDR_BYTES + g.tab_bytes + g.nastage * g.a_stage_bytes + (g.w_resident ? g.wbytes : (size_t)g.wst * g.spb * g.wslab);
  // Split-K over the filter taps for launches that under-fill the 148 SMs (stage-4 144->144 at 8x8: 50 tiles, 81 K steps each;
  // 72->72 at 16x16: 162 tiles = two rounds for 14 CTAs): a work unit is (tile, half of the taps), the two halves are added with
  // fp32 atomics onto a zeroed output.  With exactly TWO addends the result does not depend on their order (0 + a = a exactly,
  // a + b = b + a), so the forward stays bit-reproducible.  Taken when it shortens the critical path in rounds of work units.
  // MEASURED (B200, round 2): not a win as built — 8x8 144->144 B=64 34.7 us (36.6 without), 16x16 72->72 41.3 us (32.7
  // without), step 90.7 ms vs 87.0: both halves re-stage the whole halo, the scalar fp32 atomics of the epilogue and the extra
  // memset cost more than the shorter K loop saves, and the weight ring (not the MMA count) bounds these layers: every CTA
  // streams its slabs from L2 at ~21 B/clk (5 chunks in flight).  Off unless HCM_TC_KSPLIT=1 (kept for experiments).
  g.ksplit = 1;
  static int ks_on = -1;
  if (ks_on < 0) { const char* e = getenv("HCM_TC_KSPLIT"); ks_on = (e && e[0] == '1') ? 1 : 0; }
  if (ks_on == 1 && mode == 0 && !g.rc && g.taps == 9 && g.nsteps >= 18) {
    const long r1 = (g.tiles + 147) / 148, r2 = (2 * g.tiles + 147) / 148;
    if (r2 < 2 * r1) g.ksplit = 2;
  }
  g.units = g.tiles * g.ksplit;
  g.grid = (int)(g.units < 148 ? g.units : 148);
  return g;
}

bool geo_ok(const Geo& g, int H, int W, int Cin, int Cout, int ks, int stride) {
  if (!((ks == 3 && (stride == 1 || stride == 2)) || (ks == 1 && stride == 1))) return false;
  if (stride == 2 && ((H | W) & 1)) return false;
  return Cin <= 256 && Cout <= 256 && (Cout % 2) == 0 && (Cin % 2) == 0 && g.ngroups <= MAXG && g.Lpad <= MAX_LPAD &&
         g.Mv < (1L << 31) && (g.w_resident || g.wst >= 2) && g.smem <= 227 * 1024 && (g.cg / g.V) <= MAX_UNITS && g.nsteps <= MAX_STEPS;
}

struct TcParams {
  const float* x;
  const float* in_scale;
  const float* in_shift;
  int in_relu;
  const uint8_t* wpack;
  const float* bias;
  float* y;
  int accumulate;
  int B, H, W, Cin, Cout;
  int q0;                      // mode 1: first output parity of this launch
  long long* dbg;
  int dbg_mode;                // HCM_TC_DBGMODE experiments: 1 transform skips staging, 2 epilogue skips stores, 4 epilogue skips TMEM loads too
  Geo g;
  // host-built issue schedule: steps ordered (group, tap, K16 step); x = byte offset of the A operand inside a stage,
  // y = weight slab index.  The MMA warp reads it from the constant bank (uniform loads): the issue loop must stay lean,
  // a tcgen05.mma of N<=64 retires in ~40-48 cycles (measured) and integer address math per step would dominate.
  int gcount[2][MAXG];         // steps of (part, group); part = half of the filter taps when g.ksplit == 2
  int pstart[2], plen[2];      // first schedule step / number of steps of a part
  uint2 steps[MAX_STEPS];
};

// tap -> parity plane (stride 2) and row offset into the staged halo
__host__ __device__ __forceinline__ void tap_info(const Geo& g, int tap, int& q, int& rowoff) {
  const int r = tap / g.ks, s = tap - r * g.ks;
  if (g.mode == 1) { q = 0; rowoff = r * g.Wp + s; }
  else if (g.ks == 1) { q = 0; rowoff = 0; }
  else if (g.stride == 1) { q = 0; rowoff = r * g.Wp + s; }
  else {
    const int py = (r == 1) ? 0 : 1, px = (s == 1) ? 0 : 1;
    q = py * 2 + px;
    rowoff = ((r == 0) ? 0 : 1) * g.Wp + ((s == 0) ? 0 : 1);
  }
}

void build_steps(TcParams& p) {
  const Geo& g = p.g;
  const int nj = g.Cin16 / 16;
  const int ntap = g.rc ? 3 : g.taps;
  int n = 0;
  for (int part = 0; part < 2; ++part) {
    // taps of this part: all of them, or the first 5 / last 4 of the 9
    const int t_lo = (g.ksplit == 2) ? (part == 0 ? 0 : (ntap + 1) / 2) : (part == 0 ? 0 : ntap);
    const int t_hi = (g.ksplit == 2) ? (part == 0 ? (ntap + 1) / 2 : ntap) : ntap;
    p.pstart[part] = n;
    for (int grp = 0; grp < g.ngroups; ++grp) {
      const int c_lo = grp * g.cg, c_hi = c_lo + g.cg;
      p.gcount[part][grp] = 0;
      for (int tap = t_lo; tap < t_hi; ++tap) {
        int q, rowoff;
        if (g.rc) { q = 0; rowoff = tap * g.Wp; }          // one step per filter row: its three taps are N blocks of the operand
        else tap_info(g, tap, q, rowoff);
        for (int j = 0; j < nj; ++j) {
          const int cs = q * g.Cin16 + 16 * j;
          if (cs < c_lo || cs >= c_hi) continue;
          const int cl = cs - c_lo;
          p.steps[n].x = (uint32_t)(cl / g.KB) * (uint32_t)g.plane_bytes + (uint32_t)rowoff * g.SW + (uint32_t)(cl % g.KB) * 2;
          p.steps[n].y = (uint32_t)(tap * nj + j);
          ++n; ++p.gcount[part][grp];
        }
      }
    }
    p.plen[part] = n - p.pstart[part];
  }
}

// virtual position (+ parity plane for stride 2) -> pixel index of the [B,H,W] input, or -1 for padding / out of range
__device__ __forceinline__ int virt_to_src(long pv, int q, const TcParams& p) {
  const Geo& g = p.g;
  if (pv < 0 || pv >= g.Mv) return -1;
  if (g.mode == 0 && g.ks == 1) return (int)pv;
  const unsigned v = (unsigned)pv, hw = (unsigned)(g.Hp * g.Wp);
  const unsigned b = fdiv(v, g.fd_hw), rem = v - b * hw;
  const unsigned row = fdiv(rem, g.fd_wp), col = rem - row * (unsigned)g.Wp;
  if (g.mode == 1) {
    if (row >= (unsigned)g.Ho || col >= (unsigned)g.Wo) return -1;
    return (int)((b * g.Ho + row) * g.Wo + col);
  }
  if (g.stride == 1) {
    if (row < 1 || row > (unsigned)p.H || col < 1 || col > (unsigned)p.W) return -1;
    return (int)((b * p.H + row - 1) * p.W + (col - 1));
  }
  if (row < 1 || col < 1) return -1;
  const unsigned yi = 2 * (row - 1) + (unsigned)(q >> 1), xi = 2 * (col - 1) + (unsigned)(q & 1);
  return (int)((b * p.H + yi) * p.W + xi);
}
// virtual position -> output pixel index of [B,Ho,Wo], or -1
__device__ __forceinline__ int virt_to_dst(long pv, const TcParams& p) {
  const Geo& g = p.g;
  if (pv < 0 || pv >= g.Mv) return -1;
  if (g.mode == 0 && g.ks == 1) return (int)pv;
  const unsigned v = (unsigned)pv, hw = (unsigned)(g.Hp * g.Wp);
  const unsigned b = fdiv(v, g.fd_hw), rem = v - b * hw;
  const unsigned row = fdiv(rem, g.fd_wp), col = rem - row * (unsigned)g.Wp;
  if (g.mode == 1) {
    if (row >= (unsigned)g.Ho || col >= (unsigned)g.Wo) return -1;
    return (int)((b * p.H + 2 * row) * p.W + 2 * col);
  }
  if (row < 1 || col < 1 || row > (unsigned)g.Ho || col > (unsigned)g.Wo) return -1;
  return (int)((b * g.Ho + row - 1) * g.Wo + (col - 1));
}

// byte offset of byte `kb` (multiple of 4) of row `row` inside a swizzled plane (rows of SW bytes, 1024-B aligned base):
// Swizzle<log2(SW/16),4,3>: the 16-byte chunk index is XORed with address bits [7, 7+log2(SW/16))
__host__ __device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t kb, uint32_t SW) {
  const uint32_t off = row * SW;
  return off + ((((kb >> 4) ^ (off >> 7)) & (SW / 16 - 1)) << 4) + (kb & 15);
}

// K-major swizzled operand descriptor template (everything but the start address): rows of SW bytes, SBO = 8*SW.
// base_offset stays 0: measured on B200, the swizzle XOR uses absolute shared-memory address bits.
__device__ __forceinline__ uint64_t sw_desc_template(uint32_t SW) {
  uint64_t d = (uint64_t)1 << 16;
  d |= (uint64_t)(((8 * SW) >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(SW == 128 ? 2 : (SW == 64 ? 4 : 6)) << 61;
  return d;
}
__device__ __forceinline__ uint64_t with_addr(uint64_t templ, uint32_t saddr) { return templ | (uint64_t)((saddr & 0x3FFFFu) >> 4); }

// Accumulating epilogue (the data gradient of a layer whose input already holds another gradient contribution — every block's
// first convolution — and the partial sums of the projection): `y += v` as a vector reduction at the L2 (`red.global.add.v4.f32`,
// fire-and-forget) instead of load + add + store.  Each output element has exactly one writer per launch, so the result is the same
// single fp32 addition; but the load version exposed a full L2 / HBM round trip per 16-column chunk in the four epilogue warps —
// measured (B200, B=64): 64x64 18->18 41 -> 88 us, 64x64 64->256 (1x1) 117 -> 406 us with accumulate = 1 (scripts/time_acc.py).
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float2 v) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// ------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(NTHREADS, 1) tc_conv_kernel(const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Geo& g = p.g;
  if (p.dbg && threadIdx.x == 0) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); p.dbg[(long)blockIdx.x * 16 + 8] = t; }
  // barriers: 0..3 a_full  4..7 a_empty  8,9 acc_full  10,11 acc_empty  12 w_full(resident)  16..31 ring full, 32..47 ring empty
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 512);
  float* s_sc = reinterpret_cast<float*>(smem + 1024);
  float* s_sh = s_sc + 256;
  uint2* s_steps = reinterpret_cast<uint2*>(smem + 4096 + 6144);         // resident weights: per step (A offset >> 4, B descriptor low word)
  int* s_src = reinterpret_cast<int*>(smem + HDR_BYTES);                 // [nastage][Lpad][nq] source pixel or -1
  const int tab_stride = (int)(g.tab_bytes / g.nastage / 4);
  uint8_t* Abase = smem + HDR_BYTES + g.tab_bytes;
  const uint32_t SW = (uint32_t)g.SW;
  const uint32_t plane = (uint32_t)g.plane_bytes;
  const uint32_t lo_off = (uint32_t)g.nblk * plane;                      // A_lo planes follow the A_hi planes of a stage
  uint8_t* Wbase = Abase + (size_t)g.nastage * g.a_stage_bytes;
  const uint32_t wslab = (uint32_t)g.wslab;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  if (threadIdx.x < 48) {
    // one barrier per thread (a single thread initialising all 45 serialised ~0.5 us into the prologue of every launch)
    const int i = threadIdx.x;
    const uint32_t count = i < 4 ? (uint32_t)(NTRANS / g.nastage) : ((i == 10 || i == 11) ? 128u : 1u);
    if (i < 13 || i >= 16) mbar_init(BAR(i), count);
    fence_mbar_init();
  }
  {
    // issue schedule as descriptor LOW WORDS (see umma_bf16_w): the MMA warp then needs one add per operand and step
    // (.y = B descriptor of the step's slab when the weights are resident; the ring computes it from the slot)
    const uint32_t b_lo32 = (uint32_t)smem_desc(0, (uint32_t)(2 * g.NB) * 16, 128);
    const uint32_t w0s = smem_u32(smem + HDR_BYTES + g.tab_bytes + (size_t)g.nastage * g.a_stage_bytes);
    for (int i = threadIdx.x; i < g.nsteps; i += NTHREADS)
      s_steps[i] = make_uint2(p.steps[i].x >> 4, b_lo32 + ((w0s + p.steps[i].y * (uint32_t)g.wslab) >> 4));
  }
  // zero the staged A buffers once: K-padding channels are never written afterwards and must read as 0
  for (size_t i = threadIdx.x; i < (size_t)g.nastage * g.a_stage_bytes / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(Abase)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == W_MMA) tmem_alloc(smem_u32(tmem_ptr), g.tmem_cols);
  // everything above touched only shared memory / TMEM / kernel parameters: with a programmatic dependent launch it ran while the
  // stream predecessor was still finishing.  From here on the predecessor's results are read (common.cuh: PDL protocol).
  pdl_wait();
  pdl_trigger();
  for (int c = threadIdx.x; c < 256; c += NTHREADS) {
    s_sc[c] = (p.in_scale && c < p.Cin) ? p.in_scale[c] : 1.f;
    s_sh[c] = (p.in_scale && c < p.Cin) ? p.in_shift[c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (p.dbg && threadIdx.x == 0) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); p.dbg[(long)blockIdx.x * 16 + 9] = t; }

  const int ksplit = g.ksplit;
  const int st = g.st;                                                   // tiles per fill; > 1: this CTA owns a contiguous tile range
  const long t_first = (st > 1) ? ((long)blockIdx.x * g.tiles) / gridDim.x : 0;
  const int my_tiles = (st > 1) ? (int)(((long)(blockIdx.x + 1) * g.tiles) / gridDim.x - t_first)
                                : (int)((g.units - blockIdx.x + gridDim.x - 1) / gridDim.x);   // work units (tile, part) of this CTA
  // supertile fills ramp up (1, 2, st, st, .. tiles) so that the first MMA does not wait for a whole st-tile halo
  const int st_r1 = min(2, st);
  auto fill_first = [&](int f) -> int { return f == 0 ? 0 : (f == 1 ? 1 : 1 + st_r1 + (f - 2) * st); };
  auto fill_size = [&](int f) -> int { return min(f == 0 ? 1 : (f == 1 ? st_r1 : st), my_tiles - fill_first(f)); };
  // first virtual position of the CTA's ti-th tile
  auto tile_pos = [&](int ti) -> long {
    if (st > 1) return (t_first + ti) * g.TM;
    const unsigned unit = blockIdx.x + (unsigned)ti * gridDim.x;
    return (long)((ksplit == 2) ? (unit >> 1) : unit) * g.TM;
  };
  const int nj = g.Cin16 / 16;                                           // K=16 steps per tap
  (void)nj;

  if (warp == W_TMA) {
    // ===== TMA producer: weight slabs in the order the MMA warp consumes them =====
    if (lane == 0) {
      if (g.w_resident) {
        mbar_expect_tx(BAR(12), (uint32_t)g.wbytes);
        for (size_t off = 0; off < g.wbytes; off += 32768) {
          const uint32_t n = (uint32_t)(g.wbytes - off < 32768 ? g.wbytes - off : 32768);
          tma_bulk_g2s(smem_u32(Wbase + off), p.wpack + off, n, BAR(12));
        }
      } else {
        int s = 0;                                                       // ring slot and its phase (no per-chunk divisions)
        uint32_t ph = 1;
        const uint32_t chunk_bytes = (uint32_t)g.spb * wslab;
        for (int ti = 0; ti < my_tiles; ++ti) {
          const int part = (ksplit == 2) ? (int)((blockIdx.x + (unsigned)ti * gridDim.x) & 1u) : 0;
          const int st_end = p.pstart[part] + p.plen[part];
          for (int st0 = p.pstart[part]; st0 < st_end; st0 += g.spb) {
            const int n = min(g.spb, st_end - st0);
            mbar_wait(BAR(32 + s), ph);
            mbar_expect_tx(BAR(16 + s), (uint32_t)n * wslab);
            for (int k = 0; k < n; ++k)
              tma_bulk_g2s(smem_u32(Wbase + (size_t)s * chunk_bytes + (size_t)k * wslab),
                           p.wpack + (size_t)p.steps[st0 + k].y * wslab, wslab, BAR(16 + s));
            if (++s == g.wst) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ===== MMA issuer: the whole warp runs the (warp-uniform) loop, one elected lane issues tcgen05 =====
    const uint32_t idesc_n = instr_desc(g.NB), idesc_2n = instr_desc(2 * g.NB);
    const uint32_t a0 = smem_u32(Abase), w0 = smem_u32(Wbase);
    const uint64_t a_t = sw_desc_template(SW);
    const uint32_t b_lbo = (uint32_t)(2 * g.NB) * 16;
    const uint64_t b_t = smem_desc(0, b_lbo, 128);                         // no-swizzle K-major weight slab
    long long c_acc = 0, c_a = 0, c_all = clock64(), tq = 0;
    if (g.w_resident) mbar_wait(BAR(12), 0);
    // stage / phase counters advance incrementally: this warp is the serial resource of the kernel, and the emulated 64-bit
    // divisions that used to derive them per tile cost ~1k of its ~3k cycles per tile (HCM_TC_DBGMODE=7 experiment, round 2)
    const bool dbg = p.dbg != nullptr;
    int s = 0, as = 0;               // A stage / accumulator stage and their phases
    uint32_t sph = 0, aph = 1;
    int within = 0, rs = 0;          // weight ring: step inside the current chunk, slot, slot phase
    uint32_t rph = 0;
    // st == 1: per tile { wait accumulator; per channel group { wait its staged fill; issue } ; commit accumulator }
    // st  > 1: one channel group; per fill { wait it; per tile of the fill { wait accumulator; issue from row offset j*128; commit } }
    for (int ti = 0, fi = 0; ti < my_tiles; ++fi) {
      const int nsub = (st > 1) ? fill_size(fi) : 1;
      uint32_t d = 0, first = 0;          // accumulator; accumulate flag of the next MMA (0 = overwrite: first MMA of the tile)
      const int part = (ksplit == 2) ? (int)((blockIdx.x + (unsigned)ti * gridDim.x) & 1u) : 0;
      int sidx = p.pstart[part];
      const int st_end = sidx + p.plen[part];
      if (st == 1) {
        if (dbg) tq = clock64();
        mbar_wait(BAR(10 + as), aph);                                     // epilogue has drained this accumulator
        if (dbg) c_acc += clock64() - tq;
        tc_fence_after();
        d = tmem + (uint32_t)(as * g.acc_cols);
      }
      for (int grp = 0; grp < g.ngroups; ++grp) {
        if (dbg) tq = clock64();
        mbar_wait(BAR(s), sph);
        if (dbg) c_a += clock64() - tq;
        tc_fence_after();
       for (int j = 0; j < nsub; ++j) {
        if (st > 1) {
          if (dbg) tq = clock64();
          mbar_wait(BAR(10 + as), aph);
          if (dbg) c_acc += clock64() - tq;
          tc_fence_after();
          d = tmem + (uint32_t)(as * g.acc_cols);
          first = 0; sidx = p.pstart[0];
        }
        const uint32_t ab = a0 + (uint32_t)s * (uint32_t)g.a_stage_bytes + (uint32_t)(j * TILE_M) * SW;
        const int n = p.gcount[part][grp];
        if (g.w_resident) {
          if (elect_one()) {
            // lean issue loop: descriptor low words = per-stage base + per-step offset (one add per operand), shared high words
            const uint32_t a_hi32 = (uint32_t)(a_t >> 32), b_hi32 = (uint32_t)(b_t >> 32);
            const uint32_t abl = (uint32_t)a_t + (ab >> 4), lo16 = lo_off >> 4, nb = (uint32_t)g.NB;
            const uint2* st = s_steps + sidx;
            if (g.concat) {
#pragma unroll 4
              for (int k = 0; k < n; ++k) {
                const uint2 stp = st[k];
                const uint32_t ah = abl + stp.x;
                umma_bf16_w(d, ah, a_hi32, stp.y, b_hi32, idesc_2n, (k | first) ? 1u : 0u);   // [A_hi*w_hi | A_hi*w_lo]
                umma_bf16_acc(d, ah + lo16, a_hi32, stp.y, b_hi32, idesc_n);                   // += A_lo*w_hi (first Np rows of the slab)
              }
            } else {
#pragma unroll 2
              for (int k = 0; k < n; ++k) {
                const uint2 stp = st[k];
                const uint32_t ah = abl + stp.x;
                umma_bf16_w(d, ah, a_hi32, stp.y, b_hi32, idesc_n, (k | first) ? 1u : 0u);
                umma_bf16_acc(d, ah, a_hi32, stp.y + nb, b_hi32, idesc_n);                     // lo rows start Npad*16 bytes further
                umma_bf16_acc(d, ah + lo16, a_hi32, stp.y, b_hi32, idesc_n);
              }
            }
          }
          sidx += n;
          if (n) first = 1u;
        } else {
          // weight ring, lean: one barrier wait per chunk of `spb` steps, then the chunk's steps are issued by the elected lane
          // with descriptor low words (A: stage base + per-step offset from s_steps; B: ring slot, advancing one slab per step)
          const uint32_t a_hi32 = (uint32_t)(a_t >> 32), b_hi32 = (uint32_t)(b_t >> 32);
          const uint32_t abl = (uint32_t)a_t + (ab >> 4), lo16 = lo_off >> 4, nb = (uint32_t)g.NB, ws16 = wslab >> 4;
          const uint32_t bbase = (uint32_t)b_t + (w0 >> 4);
          int k = 0;
          while (k < n) {
            mbar_wait(BAR(16 + rs), rph);                                  // (re-waiting a completed phase is harmless)
            tc_fence_after();
            const int m = min(n - k, g.spb - within);                      // steps of this group inside the current chunk
            const bool done = (within + m == g.spb) || (sidx + m == st_end);
            if (elect_one()) {
              uint32_t wbl = bbase + (uint32_t)(rs * g.spb + within) * ws16;
              const uint2* st = s_steps + sidx;
              if (g.concat) {
                for (int i = 0; i < m; ++i, wbl += ws16) {
                  const uint32_t ah = abl + st[i].x;
                  umma_bf16_w(d, ah, a_hi32, wbl, b_hi32, idesc_2n, (i | k | first) ? 1u : 0u);
                  umma_bf16_acc(d, ah + lo16, a_hi32, wbl, b_hi32, idesc_n);
                }
              } else {
                for (int i = 0; i < m; ++i, wbl += ws16) {
                  const uint32_t ah = abl + st[i].x;
                  umma_bf16_w(d, ah, a_hi32, wbl, b_hi32, idesc_n, (i | k | first) ? 1u : 0u);
                  umma_bf16_acc(d, ah, a_hi32, wbl + nb, b_hi32, idesc_n);
                  umma_bf16_acc(d, ah + lo16, a_hi32, wbl, b_hi32, idesc_n);
                }
              }
              if (done) umma_commit(BAR(32 + rs));
            }
            __syncwarp();
            k += m; sidx += m; within += m;
            if (done) { within = 0; if (++rs == g.wst) { rs = 0; rph ^= 1u; } }
          }
          if (n) first = 1u;
        }
        if (st > 1) {
          if (elect_one()) umma_commit(BAR(8 + as));                      // accumulator of this tile complete
          if (++as == g.acc_stages) { as = 0; aph ^= 1u; }
        }
       }
        if (elect_one()) umma_commit(BAR(4 + s));                         // staged A buffer free
        if (++s == g.nastage) { s = 0; sph ^= 1u; }
      }
      if (st == 1) {
        if (elect_one()) umma_commit(BAR(8 + as));                        // accumulator complete
        if (++as == g.acc_stages) { as = 0; aph ^= 1u; }
      }
      ti += nsub;
    }
    if (p.dbg && lane == 0) {
      long long* o = p.dbg + (long)blockIdx.x * 16;
      o[0] = clock64() - c_all; o[1] = 0; o[2] = c_acc; o[3] = c_a;
    }
  } else if (warp < W_EPI) {
    // ===== transform warps: coalesced gather + BN/ReLU-on-load + bf16 split -> swizzled channels-last tile =====
    const int TS = NTRANS / g.nastage;                                    // team size: team k owns A stage k
    const int team = threadIdx.x / TS, t = threadIdx.x - team * TS;
    const int V = g.V;
    long long c_wait = 0, c_all = clock64(), tq = 0;
    const bool dbg = p.dbg != nullptr;
    int nfills = my_tiles * g.ngroups;
    if (st > 1) nfills = my_tiles <= 1 ? my_tiles : (my_tiles <= 1 + st_r1 ? 2 : 2 + (my_tiles - 1 - st_r1 + st - 1) / st);
    const int qsh = (g.nq == 4) ? 2 : 0;                                  // nq is 1 or 4
    // fill f = team + k * nastage -> (tile ti, channel group grp), advanced incrementally (no per-fill divisions);
    // supertiles (st > 1, one channel group): fill f = tiles [f*st, f*st + nsub)
    int ti = (st > 1) ? fill_first(team) : team / g.ngroups, grp = (st > 1) ? 0 : team - ti * g.ngroups;
    uint32_t eph = 1;
    for (int f = team; f < nfills; f += g.nastage, eph ^= 1u) {
      {
        const long tile0 = tile_pos(ti);
        const int nsub = (st > 1) ? fill_size(f) : 1;
        const int Lf = (nsub - 1) * g.TM + g.L;                           // staged rows of this fill, and padded to the K = 16 steps
        const int rows_f = min(g.Lpad, (Lf + 15) & ~15);
        const int s = team;
        if (dbg) tq = clock64();
        mbar_wait(BAR(4 + s), eph);
        if (dbg) c_wait += clock64() - tq;
        int* src_tab = s_src + s * tab_stride;
        for (int e = t; e < rows_f * g.nq; e += TS) {
          const int pos = e >> qsh, q = e - (pos << qsh);
          src_tab[e] = (pos < Lf) ? virt_to_src(tile0 - g.center + pos, q, p) : -1;
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(TS) : "memory");
        // the real (non-padding) channels of this group, plane by plane (stride 2: four parity planes side by side)
        const int c_lo = grp * g.cg, c_hi = c_lo + g.cg;
        uint8_t* A_hi = Abase + (size_t)s * g.a_stage_bytes;
        for (int q = 0; q < g.nq; ++q) {
          const int a = max(c_lo, q * g.Cin16), b = min(c_hi, q * g.Cin16 + p.Cin);
          if (b <= a) continue;
          if (p.dbg_mode & 1) continue;
          if (V == 4) stage_rows8<4>(p.x, p.Cin, a - q * g.Cin16, b - a, rows_f, src_tab, g.nq, q, A_hi, lo_off, plane, SW,
                                     (uint32_t)(a - c_lo) * 2, s_sc, s_sh, p.in_scale != nullptr, p.in_relu, t, TS);
          else stage_rows8<2>(p.x, p.Cin, a - q * g.Cin16, b - a, rows_f, src_tab, g.nq, q, A_hi, lo_off, plane, SW,
                              (uint32_t)(a - c_lo) * 2, s_sc, s_sh, p.in_scale != nullptr, p.in_relu, t, TS);
        }
        fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        mbar_arrive(BAR(s));
      }
      if (st > 1) ti = fill_first(f + g.nastage);
      else { grp += g.nastage; while (grp >= g.ngroups) { grp -= g.ngroups; ++ti; } }
    }
    if (p.dbg && threadIdx.x == 0) { long long* o = p.dbg + (long)blockIdx.x * 16; o[4] = clock64() - c_all; o[5] = c_wait; }
  } else {
    // ===== epilogue warps: TMEM -> registers -> global (fp32 NHWC), interior positions only =====
    const int q = warp & 3;                             // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;
    long long c_wait = 0, c_all = clock64(), tq = 0;
    const bool dbg = p.dbg != nullptr;
    const bool vec4 = (p.Cout % 4) == 0;
    int as = 0;
    uint32_t aph = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const unsigned unit = blockIdx.x + (unsigned)ti * gridDim.x;
      const long tile0 = tile_pos(ti);
      const bool atomic_out = ksplit > 1;                 // split-K: both halves are added onto the zeroed / accumulated output
      const bool add_bias = p.bias && (ksplit == 1 || (unit & 1u) == 0);
      const int px = (m < g.TM) ? virt_to_dst(tile0 + m, p) : -1;   // row-concat tiles: rows 126, 127 belong to the next tile
      float* yp = p.y + (long)(px < 0 ? 0 : px) * p.Cout;   // indexed [c0 + i] below
      if (dbg) tq = clock64();
      mbar_wait(BAR(8 + as), aph);
      if (dbg) c_wait += clock64() - tq;
      tc_fence_after();
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * g.acc_cols);
      // row-concat boundary exchange: [tile parity][chunk][warp 1..3][E_1 row 0, E_2 row 0, E_2 row 1][16]
      float* xch = reinterpret_cast<float*>(smem + 4096) + (ti & 1) * (5 * 3 * 3 * 16);
      for (int c0 = 0; c0 < g.Npad; c0 += 16) {
        if (p.dbg_mode & 4) break;
        if (g.mode == 1) {
          // parity block qq = c0 / Cp of the 2x2 output pixels of this position; channel offset inside the block
          const int qq = p.q0 + c0 / g.Cp, cc = c0 - (c0 / g.Cp) * g.Cp;
          yp = p.y + ((long)(px < 0 ? 0 : px) + (long)(qq >> 1) * p.W + (qq & 1)) * p.Cout - c0 + cc;
        }
        float v[16];
        if (g.rc) {
          // out[m] = E_0[m] + E_1[m+1] + E_2[m+2]: the column taps are row shifts of the accumulator.  Within a warp the
          // shift is a shuffle; rows 0 and 1 of warps 1..3 publish E_1[row 0], E_2[row 0], E_2[row 1] for lanes 30 / 31 of
          // the previous warp (one 128-thread named barrier per chunk).
          float e1[16], e2[16];
          tmem_ld16_issue(trow + (uint32_t)c0, v);
          tmem_ld16_issue(trow + (uint32_t)(g.Npad + c0), e1);
          tmem_ld16_issue(trow + (uint32_t)(2 * g.Npad + c0), e2);
          tmem_ld_wait();
          float* xc = xch + (c0 >> 4) * (3 * 3 * 16);
          if (q > 0 && lane < 2) {
            float* o = xc + (q - 1) * 3 * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (lane == 0) o[i] = e1[i];
              o[(1 + lane) * 16 + i] = e2[i];
            }
          }
          asm volatile("bar.sync 6, 128;" ::: "memory");
          const float* nx = xc + q * 3 * 16;                  // published by warp q+1
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float t1 = __shfl_down_sync(0xffffffffu, e1[i], 1), t2 = __shfl_down_sync(0xffffffffu, e2[i], 2);
            if (q < 3) {
              if (lane == 31) t1 = nx[i];
              if (lane >= 30) t2 = nx[(1 + lane - 30) * 16 + i];
            }
            v[i] += t1 + t2;
          }
        } else {
          const uint32_t tbase = trow + (uint32_t)c0;
          tmem_ld16(tbase, v);
          if (g.concat) {
            float u[16];
            tmem_ld16(tbase + (uint32_t)g.Npad, u);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += u[i];
          }
        }
        const int cend = (g.mode == 1) ? (c0 / g.Cp) * g.Cp + p.Cout : p.Cout;      // first invalid column
        if (px >= 0 && !(p.dbg_mode & 2)) {
          if (add_bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (c0 + i < cend) v[i] += p.bias[c0 + i];
          }
          if (atomic_out) {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (c0 + i < cend) atomicAdd(yp + c0 + i, v[i]);
          } else if (vec4) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              if (c0 + i < cend) {
                float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                if (p.accumulate) red_add_v4(yp + c0 + i, o);
                else *reinterpret_cast<float4*>(yp + c0 + i) = o;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              if (c0 + i < cend) {
                float2 o = make_float2(v[i], v[i + 1]);
                if (p.accumulate) red_add_v2(yp + c0 + i, o);
                else *reinterpret_cast<float2*>(yp + c0 + i) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(BAR(10 + as));                        // accumulator stage may be overwritten
      if (++as == g.acc_stages) { as = 0; aph ^= 1u; }
    }
    if (p.dbg && m == 0) { long long* o = p.dbg + (long)blockIdx.x * 16; o[6] = clock64() - c_all; o[7] = c_wait; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, g.tmem_cols);
  if (p.dbg && threadIdx.x == 0) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); p.dbg[(long)blockIdx.x * 16 + 10] = t; }
}

// ------------------------------------------------------------------------------------------ weight packing
// wpack: one slab of 64*Npad bytes per K=16 step, step = tap * (Cin16/16) + j (channels 16j..16j+15 of the source):
//   [2 K-chunks of 8][2*Npad rows][8 bf16]  with rows [0,Npad) = hi(w), rows [Npad, 2*Npad) = lo(w)
// (K-major, no swizzle: LBO = 2*Npad*16 B between the K chunks, SBO = 128 B between 8-row groups).
// transpose = 0: B[n][c] = w[n][c][tap]                        (forward; w is OIHW [Cout][ldw][ks][ks])
// transpose = 1: B[n][c] = w[c][n][taps-1-tap]  with the conv seen from the gradient side (stride 1 only):
//                n runs over the ORIGINAL Cin, c over the ORIGINAL Cout
__global__ void tc_pack_kernel(const float* __restrict__ w, uint8_t* __restrict__ out, Geo g, int Cin, int Cout, int transpose,
                               int wCin) {
  const int nj = g.Cin16 / 16;
  const int step = blockIdx.x;                   // one block per (tap, K step), whatever the layout
  const int tap = step / nj, j = step - tap * nj;
  uint8_t* slab = out + (size_t)step * 64 * g.Npad;
  for (int e = threadIdx.x; e < g.Npad * 16; e += blockDim.x) {
    const int n = e >> 4, k = e & 15;
    const int c = 16 * j + k;
    float v = 0.f;
    // (forward packs only: wCin < Cin means the weight tensor has wCin input channels and the GEMM's K is zero-padded to Cin —
    //  the 3-channel stem run on a 4-channel padded input)
    if (n < Cout && c < Cin && (transpose || c < wCin))
      v = transpose ? w[((long)c * wCin + n) * g.taps + (g.taps - 1 - tap)] : w[((long)n * wCin + c) * g.taps + tap];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    if (g.rc) {
      // row-concat layout: one slab per (filter row r, K step j); rows [s*Npad + n] = hi of tap (r,s), then the lo rows
      const int r = tap / 3, s = tap - 3 * r;
      uint8_t* rslab = out + (size_t)(r * nj + j) * g.wslab;
      const size_t chunk = (size_t)(k >> 3) * (2 * g.NB) * 16;
      *reinterpret_cast<__nv_bfloat16*>(rslab + chunk + (size_t)(s * g.Npad + n) * 16 + (k & 7) * 2) = hi;
      *reinterpret_cast<__nv_bfloat16*>(rslab + chunk + (size_t)(g.NB + s * g.Npad + n) * 16 + (k & 7) * 2) = lo;
      continue;
    }
    const size_t chunk = (size_t)(k >> 3) * (2 * g.Npad) * 16;
    *reinterpret_cast<__nv_bfloat16*>(slab + chunk + (size_t)n * 16 + (k & 7) * 2) = hi;
    *reinterpret_cast<__nv_bfloat16*>(slab + chunk + (size_t)(g.Npad + n) * 16 + (k & 7) * 2) = lo;
  }
}

// Data gradient of a 3x3 / stride-2 / pad-1 convolution as ONE 2x2-tap stride-1 GEMM over dy whose N axis is the four
// output-pixel parities: slab(tap=(du,dv), K step j): B[n = q*Cp + ci][c = co] = w[co][ci][r][s] with
// (py,du) -> r : (0,0)->1, (1,0)->2, (1,1)->0, (0,1)->none ; same for (px,dv) -> s.
__global__ void tc_pack_dgrad2_kernel(const float* __restrict__ w, uint8_t* __restrict__ out, Geo g, int CoutW, int CinW, int q0) {
  const int nj = g.Cin16 / 16;
  const int step = blockIdx.x;
  const int tap = step / nj, j = step - tap * nj;
  const int du = tap >> 1, dv = tap & 1;
  uint8_t* slab = out + (size_t)step * g.wslab;
  for (int e = threadIdx.x; e < g.Npad * 16; e += blockDim.x) {
    const int n = e >> 4, k = e & 15;
    const int q = q0 + n / g.Cp, ci = n - (n / g.Cp) * g.Cp;
    const int py = q >> 1, px = q & 1;
    const int r = (py == 0) ? (du == 0 ? 1 : -1) : (du == 0 ? 2 : 0);
    const int sx = (px == 0) ? (dv == 0 ? 1 : -1) : (dv == 0 ? 2 : 0);
    const int co = 16 * j + k;
    float v = 0.f;
    if (r >= 0 && sx >= 0 && ci < CinW && co < CoutW) v = w[(((long)co * CinW + ci) * 3 + r) * 3 + sx];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const size_t chunk = (size_t)(k >> 3) * (2 * g.Npad) * 16;
    *reinterpret_cast<__nv_bfloat16*>(slab + chunk + (size_t)n * 16 + (k & 7) * 2) = hi;
    *reinterpret_cast<__nv_bfloat16*>(slab + chunk + (size_t)(g.Npad + n) * 16 + (k & 7) * 2) = lo;
  }
}

// All weight packs of a step in ONE launch (the weights only change at the optimiser step): blockIdx.x = global K-step index,
// jobs[j] = {w, out, Cin, Cout, ks, mode (0 forward, 1 transposed stride-1 dgrad, 2 stride-2 dgrad), ldw, first_step}
struct PackJob { long long w, out, Cin, Cout, ks, mode, ldw, first; };
__global__ void tc_pack_batch_kernel(const PackJob* __restrict__ jobs, int njobs) {
  __shared__ PackJob jb;
  if (threadIdx.x == 0) {
    int lo = 0, hi = njobs - 1;                                   // last job with first <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].first <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    jb = jobs[lo];
  }
  __syncthreads();
  const float* w = reinterpret_cast<const float*>(jb.w);
  const int Cin = (int)jb.Cin, Cout = (int)jb.Cout, ks = (int)jb.ks, mode = (int)jb.mode & 3;
  const bool rc = ((int)jb.mode & 4) != 0;                        // row-concat layout (see tc_pack_kernel)
  const int step = (int)((long long)blockIdx.x - jb.first);
  int Npad, Cin16, taps, Cp = 0;
  int q0 = 0;
  if (mode == 2) {                                               // (Cin, Cout) of the conv weight; ldw = q0 + 16*nqs
    Cp = ceil_to(Cin, 16); Npad = (int)(jb.ldw >> 4) * Cp; q0 = (int)(jb.ldw & 15); Cin16 = ceil_to(Cout, 16); taps = 4;
  }
  else { Npad = ceil_to(Cout, 16); Cin16 = ceil_to(Cin, 16); taps = ks * ks; }
  const int nj = Cin16 / 16;
  const int tap = step / nj, j = step - tap * nj;
  uint8_t* slab = reinterpret_cast<uint8_t*>(jb.out) + (size_t)step * 64 * Npad;
  const int wCin = (mode != 2 && jb.ldw > 0) ? (int)jb.ldw : (mode == 1 ? Cout : Cin);
  for (int e = threadIdx.x; e < Npad * 16; e += blockDim.x) {
    const int n = e >> 4, k = e & 15;
    float v = 0.f;
    if (mode == 2) {
      const int du = tap >> 1, dv = tap & 1;
      const int q = q0 + n / Cp, ci = n - (n / Cp) * Cp;
      const int py = q >> 1, px = q & 1;
      const int r = (py == 0) ? (du == 0 ? 1 : -1) : (du == 0 ? 2 : 0);
      const int sx = (px == 0) ? (dv == 0 ? 1 : -1) : (dv == 0 ? 2 : 0);
      const int co = 16 * j + k;
      if (r >= 0 && sx >= 0 && ci < Cin && co < Cout) v = w[(((long)co * Cin + ci) * 3 + r) * 3 + sx];
    } else {
      const int c = 16 * j + k;
      if (n < Cout && c < Cin && (mode == 1 || c < wCin))
        v = mode == 1 ? w[((long)c * wCin + n) * taps + (taps - 1 - tap)] : w[((long)n * wCin + c) * taps + tap];
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    if (rc) {
      const int r = tap / 3, s = tap - 3 * r, NB = 3 * Npad;
      uint8_t* rslab = reinterpret_cast<uint8_t*>(jb.out) + (size_t)(r * nj + j) * 64 * NB;
      const size_t chunk = (size_t)(k >> 3) * (2 * NB) * 16;
      *reinterpret_cast<__nv_bfloat16*>(rslab + chunk + (size_t)(s * Npad + n) * 16 + (k & 7) * 2) = hi;
      *reinterpret_cast<__nv_bfloat16*>(rslab + chunk + (size_t)(NB + s * Npad + n) * 16 + (k & 7) * 2) = lo;
      continue;
    }
    const size_t chunk = (size_t)(k >> 3) * (2 * Npad) * 16;
    *reinterpret_cast<__nv_bfloat16*>(slab + chunk + (size_t)n * 16 + (k & 7) * 2) = hi;
    *reinterpret_cast<__nv_bfloat16*>(slab + chunk + (size_t)(Npad + n) * 16 + (k & 7) * 2) = lo;
  }
}

bool dgrad2_ok(const Geo& g, int H, int W, int Cin, int Cout) {
  return ((H | W) & 1) == 0 && (Cin % 2) == 0 && (Cout % 2) == 0 && Cout <= 256 && g.Npad <= 256 && g.ngroups <= MAXG &&
         g.Lpad <= MAX_LPAD && g.Mv < (1L << 31) && (g.w_resident || g.wst >= 2) && g.smem <= 227 * 1024 &&
         (g.cg / g.V) <= MAX_UNITS && g.nsteps <= MAX_STEPS;
}

int launch_tc(TcParams& p, cudaStream_t stream, const char* what) {
  build_steps(p);
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    dbg_on = getenv("HCM_TC_DEBUG") ? 1 : 0;
    if (dbg_on) cudaMalloc(&dbg, 148 * 16 * sizeof(long long));
  }
  p.dbg = dbg;
  { static int dm = -1; if (dm < 0) { const char* e = getenv("HCM_TC_DBGMODE"); dm = e ? atoi(e) : 0; } p.dbg_mode = dm; }
  if (dbg_on) cudaMemsetAsync(dbg, 0, 148 * 16 * sizeof(long long), stream);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { hcm_set_error("%s: smem attribute: %s", what, cudaGetErrorString(e)); return HCM_ERR_CUDA; }
    configured = true;
  }
  hcm_launch_pdl(tc_conv_kernel, dim3((unsigned)p.g.grid), dim3(NTHREADS), p.g.smem, stream, p);
  HCM_LAUNCH_CHECK(what);
  if (dbg_on) {
    long long h[148 * 16];
    cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
    const long long* o = h;      // CTA 0
    {
      long long s_min = 0, s_max = 0, e_max = 0, pro = 0;
      for (int c = 0; c < p.g.grid; ++c) {
        const long long* q = h + c * 16;
        if (c == 0 || q[8] < s_min) s_min = q[8];
        if (c == 0 || q[8] > s_max) s_max = q[8];
        if (c == 0 || q[10] > e_max) e_max = q[10];
        if (q[9] - q[8] > pro) pro = q[9] - q[8];
      }
      fprintf(stderr, "[%s time] span %.2f us (first CTA start -> last CTA end), CTA starts spread %.2f us, prologue (max) %.2f us, CTA0 %.2f us, st %d\n",
              what, (e_max - s_min) * 1e-3, (s_max - s_min) * 1e-3, pro * 1e-3, (h[10] - h[8]) * 1e-3, p.g.st);
    }
    fprintf(stderr, "[%s dbg] %dx%d %d->%d mode %d tiles/cta %ld steps %d astages %d resident %d | mma: total %lld wait_acc %lld "
            "wait_a %lld | transform: total %lld wait %lld | epilogue: total %lld wait %lld\n", what, p.H, p.W, p.Cin, p.Cout,
            p.g.mode, (p.g.units + p.g.grid - 1) / p.g.grid, p.g.nsteps, p.g.nastage, p.g.w_resident, o[0], o[2], o[3], o[4],
            o[5], o[6], o[7]);
  }
  return HCM_OK;
}

}  // namespace

extern "C" {

// ---- data gradient of a 3x3 stride-2 pad-1 convolution w[Cout][Cin][3][3]: dx[B,H,W,Cin] (+)= conv_transpose(dy[B,H/2,W/2,Cout])
int hcm_tc_dgrad_s2_supported(int B, int H, int W, int Cin, int Cout) {
  if ((H | W) & 1) return 0;
  Geo g = make_geo(B, H, W, Cout, Cin, 3, 2, 1);          // GEMM: K = Cout (channels of dy), N = parities x Cin
  return dgrad2_ok(g, H, W, Cout, Cin) ? 1 : 0;
}
// all parity groups' packs, consecutively (4/nqs groups of nsteps*wslab bytes)
long hcm_tc_dgrad_s2_wpack_bytes(int B, int H, int W, int Cin, int Cout) {
  Geo g = make_geo(B, H, W, Cout, Cin, 3, 2, 1);
  return (long)g.wbytes * (4 / g.nqs);
}
// parities per launch (4, 2 or 1): the pack buffer holds 4/nqs groups
int hcm_tc_dgrad_s2_nqs(int B, int H, int W, int Cin, int Cout) {
  Geo g = make_geo(B, H, W, Cout, Cin, 3, 2, 1);
  return g.nqs;
}
int hcm_tc_dgrad_s2_pack(const float* w, void* wpack, int B, int H, int W, int Cin, int Cout, cudaStream_t stream) {
  HCM_CHECK_ARG(w && wpack, "tc_dgrad_s2_pack: null pointer");
  Geo g = make_geo(B, H, W, Cout, Cin, 3, 2, 1);
  HCM_CHECK_ARG(dgrad2_ok(g, H, W, Cout, Cin), "tc_dgrad_s2_pack: unsupported geometry");
  for (int q0 = 0; q0 < 4; q0 += g.nqs) {
    tc_pack_dgrad2_kernel<<<g.nsteps, 256, 0, stream>>>(w, reinterpret_cast<uint8_t*>(wpack) + (size_t)(q0 / g.nqs) * g.wbytes, g, Cout,
                                                          Cin, q0);
    HCM_LAUNCH_CHECK("tc_dgrad_s2_pack");
  }
  return HCM_OK;
}
int hcm_tc_dgrad_s2(const float* dy, const void* wpack, float* dx, int B, int H, int W, int Cin, int Cout, int accumulate,
                    cudaStream_t stream) {
  HCM_CHECK_ARG(dy && wpack && dx, "tc_dgrad_s2: null pointer");
  TcParams p;
  p.g = make_geo(B, H, W, Cout, Cin, 3, 2, 1);
  HCM_CHECK_ARG(dgrad2_ok(p.g, H, W, Cout, Cin), "tc_dgrad_s2: unsupported geometry (Cin=%d Cout=%d)", Cin, Cout);
  p.x = dy; p.in_scale = nullptr; p.in_shift = nullptr; p.in_relu = 0;
  p.bias = nullptr; p.y = dx; p.accumulate = accumulate;
  p.B = B; p.H = H; p.W = W; p.Cin = Cout; p.Cout = Cin;   // as seen by the GEMM: staged channels = Cout(w), output channels = Cin(w)
  for (int q0 = 0; q0 < 4; q0 += p.g.nqs) {                 // disjoint output pixels per parity group
    p.q0 = q0;
    p.wpack = reinterpret_cast<const uint8_t*>(wpack) + (size_t)(q0 / p.g.nqs) * p.g.wbytes;
    int rc = launch_tc(p, stream, "tc_dgrad_s2");
    if (rc != HCM_OK) return rc;
  }
  return HCM_OK;
}

// One launch for all weight packs of a step.  jobs (device): njobs x 8 int64 {w ptr, out ptr, Cin, Cout, ks, mode, ldw, first_step}
// with mode 0 = hcm_tc_conv_pack(flags 0), 1 = (flags 1), +4 = row-concatenated layout, 2 = hcm_tc_dgrad_s2_pack; (Cin, Cout) as passed to those calls;
// first_step = running sum of the jobs' K-step counts (ks*ks*ceil16(Cin)/16, mode 2: 4*ceil16(Cout)/16); total_steps = their sum.
int hcm_tc_pack_batch(const long long* jobs, int njobs, int total_steps, cudaStream_t stream) {
  HCM_CHECK_ARG(jobs && njobs >= 1 && total_steps >= 1, "tc_pack_batch: bad args");
  tc_pack_batch_kernel<<<total_steps, 128, 0, stream>>>(reinterpret_cast<const PackJob*>(jobs), njobs);
  HCM_LAUNCH_CHECK("tc_pack_batch");
  return HCM_OK;
}

// 1 if hcm_tc_conv can run this convolution: 3x3 stride 1/2 (even H,W for stride 2) or 1x1 stride 1, even channels <= 256
int hcm_tc_conv_supported(int B, int H, int W, int Cin, int Cout, int ks, int stride) {
  if (!((ks == 3 && (stride == 1 || stride == 2)) || (ks == 1 && stride == 1))) return 0;
  Geo g = make_geo(B, H, W, Cin, Cout, ks, stride);
  return geo_ok(g, H, W, Cin, Cout, ks, stride) ? 1 : 0;
}

// bytes of the packed-weight buffer (depends on channels and kernel size only)
long hcm_tc_conv_wpack_bytes(int B, int H, int W, int Cin, int Cout, int ks) {
  Geo g = make_geo(B, H, W, Cin, Cout, ks, 1);
  return (long)g.wbytes;
}

// 1 if hcm_tc_conv runs this convolution with row-concatenated taps (3x3, stride 1, 3*ceil16(Cout) <= 256): its weights
// must then be packed with flag 4 (the layout differs).  For the data gradient pass the GEMM's Cout (= Cin of the weight).
int hcm_tc_conv_rowcat_supported(int Cout, int ks, int stride) { return rowcat_ok(Cout, ks, stride, 0) ? 1 : 0; }

// Pack OIHW fp32 weights into the bf16 hi/lo K-step slabs.  `w` may point at a column block of a wider
// [O][ldw][ks][ks] tensor (ldw = 0: contiguous; forward packs with 0 < ldw < Cin: the tensor has only ldw input channels and the
// GEMM's K is zero-padded to Cin, for an input stored with padded channels).  (Cin, Cout) describe the GEMM being run.  flags:
//   bit 0 (1): transpose -> the data gradient of a STRIDE-1 conv, i.e. a conv with Cin' = Cout(w), Cout' = Cin(w): pass
//              Cin = Cout(w), Cout = Cin(w);  clear -> the forward conv of w[Cout][Cin][ks][ks] (any stride)
//   bit 2 (4): row-concatenated layout, required iff hcm_tc_conv_rowcat_supported(Cout, ks, stride of the consumer)
int hcm_tc_conv_pack(const float* w, int ldw, void* wpack, int B, int H, int W, int Cin, int Cout, int ks, int flags,
                     cudaStream_t stream) {
  HCM_CHECK_ARG(w && wpack, "tc_conv_pack: null pointer");
  HCM_CHECK_ARG(Cin <= 256 && Cout <= 256 && (ks == 1 || ks == 3), "tc_conv_pack: unsupported geometry");
  const int transpose = flags & 1;
  Geo g = make_geo(B, H, W, Cin, Cout, ks, 1);
  HCM_CHECK_ARG(!(flags & 4) || g.rc, "tc_conv_pack: row-concatenated layout requested for an ineligible convolution");
  if (!(flags & 4)) { g.rc = 0; g.NB = g.Npad; g.wslab = (size_t)64 * g.Npad; }
  const int wCin = ldw > 0 ? ldw : (transpose ? Cout : Cin);
  tc_pack_kernel<<<g.taps * (g.Cin16 / 16), 256, 0, stream>>>(w, reinterpret_cast<uint8_t*>(wpack), g, Cin, Cout, transpose, wCin);
  HCM_LAUNCH_CHECK("tc_conv_pack");
  return HCM_OK;
}

// y[B,Ho,Wo,Cout] (+)= conv_{ks x ks, stride, pad (ks-1)/2}( T(x[B,H,W,Cin]) ) (+ bias), weights pre-packed by hcm_tc_conv_pack
int hcm_tc_conv(const float* x, const void* wpack, const float* bias, float* y, int B, int H, int W, int Cin, int Cout,
                int ks, int stride, const float* in_scale, const float* in_shift, int in_relu, int accumulate,
                cudaStream_t stream) {
  HCM_CHECK_ARG(x && wpack && y, "tc_conv: null pointer");
  HCM_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "tc_conv: in_scale/in_shift must come together");
  TcParams p;
  p.g = make_geo(B, H, W, Cin, Cout, ks, stride);
  HCM_CHECK_ARG(geo_ok(p.g, H, W, Cin, Cout, ks, stride), "tc_conv: unsupported geometry (Cin=%d Cout=%d ks=%d stride=%d)", Cin,
                Cout, ks, stride);
  p.x = x; p.in_scale = in_scale; p.in_shift = in_shift; p.in_relu = in_relu;
  p.wpack = reinterpret_cast<const uint8_t*>(wpack); p.bias = bias; p.y = y; p.accumulate = accumulate;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.q0 = 0;
  if (p.g.ksplit > 1 && !accumulate) {
    cudaError_t e = cudaMemsetAsync(y, 0, (size_t)B * p.g.Ho * p.g.Wo * Cout * sizeof(float), stream);
    if (e != cudaSuccess) { hcm_set_error("tc_conv: memset: %s", cudaGetErrorString(e)); return HCM_ERR_CUDA; }
  }
  return launch_tc(p, stream, "tc_conv");
}

}  // extern "C"
