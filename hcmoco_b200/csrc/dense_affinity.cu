// Fused dense intra-sample affinity objective (learning/contrast_trainer.py:642-723), forward and backward.
//
//   a_j = norm(G1[b, pix_j, :])   d_i = norm(G2[b, pix_i, :])   L[i][j] = <d_i, a_j> / T      (S sampled pixels, C = 128)
//   loss_r2d = -mean_j sum_i W[i][j] log_softmax_i L[i][j]      loss_d2r = the same on L^T     W = softmax_i(-|c_i - c_j|)
//
// Two kernels per direction of the chain rule:
//   dense_prep_kernel   gathers the S sampled pixels of both maps ONCE per sample (512-B coalesced rows of the channels-last
//                       maps), L2-normalises, splits into bf16 hi / lo and writes them as ready-made UMMA operand slabs
//                       ([chunk of NC rows][channel/8][row][8 channels], K-major no-swizzle) into a caller-owned workspace;
//   dense_affinity_kernel  one CTA per (sample, side, strip of NC <= 128 rows).  side 0: strip rows are the a_j (statistics of
//                       the COLUMNS of L), the other operand is all d_i; side 1: strip rows are the d_i (statistics of the ROWS
//                       of L).  Slabs arrive by TMA (cp.async.bulk, one per slab, double-buffered in the forward); S x S x 128
//                       on tcgen05 (hi*hi + lo*hi + hi*lo, fp32 in TMEM, two accumulators in the forward so the MMAs of chunk
//                       c+1 run under the epilogue of chunk c); epilogue on 16 warps: soft-target log-softmax statistics
//                       (forward) or the logit gradient G, which goes back to shared memory as the A operand of a second MMA
//                       dXn = G * Y (the Y slab re-read in place as an MN-major B operand), then the L2-norm backward and a
//                       coalesced atomic scatter into the map gradient (sampled pixels repeat).
// Computing both L-strips and L^T-strips makes every softmax statistic a per-row (= per TMEM lane) reduction: no cross-lane
// shuffles, no cross-CTA combine, and nothing S x S is ever written to HBM.
// Round 2: before the split into prep + TMA-fed main kernel every strip re-gathered and re-normalised the whole other operand
// (40 % of the 18 M warp instructions of a forward launch, ncu), and the epilogue used the accurate sqrtf / a branchy online
// softmax (115 instructions per 32 logits); forward 103 -> see DESIGN.md for the measured times.
//
// Algorithmic HBM bytes per depth-bearing triplet: 2*S*128*4 gathered features (+2*S*8 indices); the slabs (same size, bf16 hi+lo)
// are written once and re-read from L2.
#include "tc_common.cuh"

namespace {

constexpr int DA_C = 128;          // feature channels
constexpr int DA_THREADS = 512;         // 16 warps: 4 per TMEM lane quarter
constexpr int DA_PARTS = DA_THREADS / 128;  // column parts of the epilogue (one per warp of a lane quarter)
constexpr int DA_HDR = 1024;

struct DaGeo {
  int S, h, HW, nch, NC, Spad;
  uint32_t lbo, lbo_g;              // K-direction core-matrix strides (bytes): operand slabs (NC rows) / gradient slab (128 rows)
  uint32_t half, slab;              // bytes of the hi (= lo) part / of one slab
  uint32_t off_tab, off_red, off_x, off_y, off_g, smem_bytes;
};

DaGeo da_geo(int S, int h, bool bwd) {
  DaGeo g;
  g.S = S; g.h = h; g.HW = h * h;
  g.nch = (S + 127) / 128;                                        // strips of rows = chunks of columns
  g.NC = ceil_to((S + g.nch - 1) / g.nch, 16);
  g.Spad = g.nch * g.NC;
  g.lbo = (uint32_t)g.NC * 16 + 16;
  g.lbo_g = 128 * 16 + 16;
  g.half = 16 * g.lbo;
  g.slab = 2 * g.half;
  uint32_t o = DA_HDR;
  g.off_tab = o; o += (uint32_t)g.Spad * 5 * 4;                  // cy, cx, lse_other, zinv_other, pixel offset
  g.off_red = o; o += DA_PARTS * 128 * 8 * 4;
  o = (o + 127) / 128 * 128;
  g.off_x = o; o += g.slab;                                      // X hi | X lo   (+ the following Y slab: fp32 staging of the scatter)
  g.off_y = o; o += (bwd ? 1 : 2) * g.slab;                      // Y hi | Y lo, double-buffered in the forward
  g.off_g = o; if (bwd) o += 2 * (uint32_t)(g.NC / 8) * g.lbo_g; // G hi | G lo
  if (bwd && o < g.off_x + 128 * 129 * 4) o = g.off_x + 128 * 129 * 4;
  g.smem_bytes = o;
  return g;
}

struct DaParams {
  const float* G1; const float* G2;
  const long long* pix; const float* kept; const float* fin;
  float* stat;                     // [B][2][S][4] = (lse, Z, sum w*logit, first-argmax hit)
  float* dG1; float* dG2;
  uint8_t* work;                   // [B][2][nch][slab]: operand slabs written by dense_prep_kernel
  float inv_T, gscale, gscale_o;     // gscale: d(total)/d(loss_r2d), gscale_o: d(total)/d(loss_d2r)
  DaGeo g;
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, float (&v)[8]) {      // values valid after tmem_ld_wait()
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "r"(taddr));
}

// The epilogue is INSTRUCTION-bound (ncu, round 2: 30 M warp instructions per launch, 115 per 32 logits; the accurate sqrtf and
// the branchy online softmax were most of it), so it works in log2 units with the approximate MUFU forms (2^-22 relative):
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrta(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float DA_LOG2E = 1.4426950408889634f, DA_LN2 = 0.6931471805599453f;

// kind::f16 instruction descriptor: D fp32, A/B bf16, M = 128; b_mn = 1 -> B operand MN-major
__device__ __forceinline__ uint32_t da_idesc(int N, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// One warp per sampled pixel q of (sample b, map m): 512-B coalesced row -> F.normalize (eps 1e-12) -> bf16 hi / lo -> slab.
__global__ void __launch_bounds__(256) dense_prep_kernel(const DaParams p) {
  const DaGeo& g = p.g;
  const int b = blockIdx.z, m = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 8 + warp;
  if (q >= g.Spad || p.kept[b] == 0.f) return;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q < g.S) {
    const float* map = (m == 0 ? p.G1 : p.G2) + ((long)b * g.HW + p.pix[(long)b * g.S + q]) * DA_C;
    v = __ldg(reinterpret_cast<const float4*>(map) + lane);
  }
  const float ss = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
  const float inv = ss > 1e-24f ? rsqrtf(ss) : 1e12f;            // 1 / max(|x|, 1e-12)
  uint32_t h01, l01, h23, l23;
  split2(v.x * inv, v.y * inv, h01, l01);
  split2(v.z * inv, v.w * inv, h23, l23);
  uint8_t* slab = p.work + (((size_t)b * 2 + m) * g.nch + q / g.NC) * g.slab;
  const uint32_t off = (uint32_t)(lane >> 1) * g.lbo + (uint32_t)(q % g.NC) * 16 + (uint32_t)(lane & 1) * 8;
  *reinterpret_cast<uint2*>(slab + off) = make_uint2(h01, h23);
  *reinterpret_cast<uint2*>(slab + g.half + off) = make_uint2(l01, l23);
}

// three-product MMA over the 8 K=16 steps of the 128 channels: D (+)= A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (descriptor low words
// advance by two 8-channel core-matrix columns per step; see umma_bf16_w)
__device__ __forceinline__ void da_mma_xy(uint32_t d, uint32_t xs, uint32_t ys, const DaGeo& g, uint32_t idesc) {
  const uint64_t xt = smem_desc(xs, g.lbo, 128), yt = smem_desc(ys, g.lbo, 128);
  const uint32_t x_hi32 = (uint32_t)(xt >> 32), y_hi32 = (uint32_t)(yt >> 32);
  const uint32_t lo16 = g.half >> 4, kk = (2u * g.lbo) >> 4;
  uint32_t xa = (uint32_t)xt, ya = (uint32_t)yt;
  umma_bf16_w(d, xa, x_hi32, ya, y_hi32, idesc, 0u);
  umma_bf16_acc(d, xa + lo16, x_hi32, ya, y_hi32, idesc);
  umma_bf16_acc(d, xa, x_hi32, ya + lo16, y_hi32, idesc);
#pragma unroll
  for (int k = 1; k < DA_C / 16; ++k) {
    xa += kk; ya += kk;
    umma_bf16_acc(d, xa, x_hi32, ya, y_hi32, idesc);
    umma_bf16_acc(d, xa + lo16, x_hi32, ya, y_hi32, idesc);
    umma_bf16_acc(d, xa, x_hi32, ya + lo16, y_hi32, idesc);
  }
}

template <bool BWD>
__global__ void __launch_bounds__(DA_THREADS, 1) dense_affinity_kernel(const DaParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const DaGeo& g = p.g;
  const int b = blockIdx.z, side = blockIdx.y, strip = blockIdx.x;
  if (p.kept[b] == 0.f) return;
  // side 0 strips hold the statistics of loss_r2d (columns of L), side 1 those of loss_d2r: `own` / `other` swap with the side
  float coef = 0.f, coef_o = 0.f;
  if (BWD) {
    const float nk = p.fin[4];
    if (!(nk > 0.f)) return;
    coef = (side == 0 ? p.gscale : p.gscale_o) / (nk * (float)g.S);
    coef_o = (side == 0 ? p.gscale_o : p.gscale) / (nk * (float)g.S);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = g.S;
  const int row0 = strip * g.NC, row_end = min(S, row0 + g.NC);
  // side 0: strip = rgb rows (map 0), other = depth (map 1); side 1: strip = depth rows, other = rgb
  const float* Xmap = (side == 0 ? p.G1 : p.G2) + (long)b * g.HW * DA_C;
  const uint8_t* Xslabs = p.work + ((size_t)b * 2 + side) * g.nch * g.slab;
  const uint8_t* Yslabs = p.work + ((size_t)b * 2 + (1 - side)) * g.nch * g.slab;
  const long long* pix_b = p.pix + (long)b * S;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 64);
  float* s_cy = reinterpret_cast<float*>(smem + g.off_tab);
  float* s_cx = s_cy + g.Spad;
  float* s_lse_o = s_cx + g.Spad;
  float* s_zinv_o = s_lse_o + g.Spad;
  int* s_off = reinterpret_cast<int*>(s_zinv_o + g.Spad);
  float* s_red = reinterpret_cast<float*>(smem + g.off_red);
  uint8_t* Xs = smem + g.off_x;
  uint8_t* Ys = smem + g.off_y;                                  // forward: two slabs
  uint8_t* Ghi = smem + g.off_g;
  uint8_t* Glo = Ghi + (uint32_t)(g.NC / 8) * g.lbo_g;
  // barriers: 0 X slab, 1-2 Y slab buffers (TMA complete_tx), 3-4 affinity MMAs of accumulator 0 / 1, 5 gradient MMAs
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const uint32_t tmem_cols = 256u;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(BAR(i), 1);
    fence_mbar_init();
    // operand slabs: the strip and the first chunk(s) of the other operand, one bulk copy each
    mbar_expect_tx(BAR(0), g.slab);
    tma_bulk_g2s(smem_u32(Xs), Xslabs + (size_t)strip * g.slab, g.slab, BAR(0));
    mbar_expect_tx(BAR(1), g.slab);
    tma_bulk_g2s(smem_u32(Ys), Yslabs, g.slab, BAR(1));
    if (!BWD && g.nch > 1) {
      mbar_expect_tx(BAR(2), g.slab);
      tma_bulk_g2s(smem_u32(Ys + g.slab), Yslabs + g.slab, g.slab, BAR(2));
    }
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_ptr), tmem_cols);
  for (int q = threadIdx.x; q < g.Spad; q += DA_THREADS) {
    const int px = q < S ? (int)pix_b[q] : 0;
    const int py = px / g.h;
    s_cy[q] = (float)py;
    s_cx[q] = (float)(px - py * g.h);
    s_off[q] = px * DA_C;
    if (BWD) {
      const float* so = p.stat + (((long)b * 2 + (1 - side)) * S + (q < S ? q : 0)) * 4;
      s_lse_o[q] = so[0] * DA_LOG2E;                       // log2 units (see the epilogue)
      s_zinv_o[q] = 1.f / so[1];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t tdX = tmem + 128;                             // backward: gradient accumulator; forward: second affinity accumulator

  // this thread's strip row (TMEM lane) and column part (8-column groups part, part + DA_PARTS, ...)
  const int qtr = warp & 3, part = warp >> 2;
  const int row = qtr * 32 + lane;
  const int grow = row0 + row;
  const bool row_ok = grow < row_end;
  const float my_y = s_cy[row_ok ? grow : 0], my_x = s_cx[row_ok ? grow : 0];
  float lse_own = 0.f, zinv_own = 0.f;
  if (BWD && row_ok) {
    const float* so = p.stat + (((long)b * 2 + side) * S + grow) * 4;
    lse_own = so[0] * DA_LOG2E;
    zinv_own = 1.f / so[1];
  }
  // Logits are <unit vector, unit vector> / T, so |l| <= 1/T: the soft-max sum needs no running maximum — every term is
  // exp(l - 1/T) <= 1 and the sum of S of them stays far inside fp32 (>= S * e^(-2/T)).  No branch, no rescaling.
  const float k2 = p.inv_T * DA_LOG2E;                         // logit in log2 units: l2 = <.,.> * k2, bounded by k2
  float mx = -INFINITY, se = 0.f, Z = 0.f, wl = 0.f;
  int am = 0;
  const uint32_t idesc1 = da_idesc(g.NC, 0), idesc2 = da_idesc(128, 1);
  const uint32_t xs = smem_u32(Xs), ys = smem_u32(Ys);

  if (warp == 0) {                                             // affinity MMAs of chunk 0
    mbar_wait(BAR(0), 0);
    mbar_wait(BAR(1), 0);
    tc_fence_after();
    if (elect_one()) {
      da_mma_xy(tmem, xs, ys, g, idesc1);
      umma_commit(BAR(3));
    }
    __syncwarp();
  }

  for (int c = 0; c < g.nch; ++c) {
    const int buf = BWD ? 0 : (c & 1);                         // Y buffer / accumulator of chunk c
    const uint32_t tP = tmem + (uint32_t)(buf * 128);
    if (!BWD && c + 1 < g.nch && warp == 0) {
      // the MMAs of chunk c+1 run under the epilogue of chunk c (its accumulator was drained before the last __syncthreads)
      const int nb = (c + 1) & 1;
      mbar_wait(BAR(1 + nb), (uint32_t)(((c + 1) >> 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        da_mma_xy(tmem + (uint32_t)(nb * 128), xs, ys + (uint32_t)nb * g.slab, g, idesc1);
        umma_commit(BAR(3 + nb));
      }
      __syncwarp();
    }
    mbar_wait(BAR(3 + buf), (uint32_t)(BWD ? (c & 1) : ((c >> 1) & 1)));
    tc_fence_after();
    if (!BWD && c + 2 < g.nch && threadIdx.x == 0) {           // Y buffer of chunk c is free: fetch chunk c+2 into it
      mbar_expect_tx(BAR(1 + buf), g.slab);
      tma_bulk_g2s(smem_u32(Ys + (size_t)buf * g.slab), Yslabs + (size_t)(c + 2) * g.slab, g.slab, BAR(1 + buf));
    }

    // ---- epilogue over this thread's 8-column groups of the chunk, TWO groups per iteration: the work of a group is one
    // dependent chain per logit (TMEM load -> MUFU sqrt -> MUFU ex2 -> ...; ncu: short-scoreboard + fixed-latency stalls dominate
    // with 4 warps per scheduler), so two groups in flight double the independent instructions the scheduler can pick from
    auto group = [&](const float (&v)[8], int c0) {
      const int q0 = c * g.NC + c0;
      const float4 cya = *reinterpret_cast<const float4*>(s_cy + q0), cyb = *reinterpret_cast<const float4*>(s_cy + q0 + 4);
      const float4 cxa = *reinterpret_cast<const float4*>(s_cx + q0), cxb = *reinterpret_cast<const float4*>(s_cx + q0 + 4);
      const float cy[8] = {cya.x, cya.y, cya.z, cya.w, cyb.x, cyb.y, cyb.z, cyb.w};
      const float cx[8] = {cxa.x, cxa.y, cxa.z, cxa.w, cxb.x, cxb.y, cxb.z, cxb.w};
      if (!BWD) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int q = q0 + i;
          const bool ok = q < S;
          const float l2 = v[i] * k2;
          const float dy = cy[i] - my_y, dx = cx[i] - my_x;
          const float w = ok ? ex2a(-DA_LOG2E * sqrta(fmaf(dy, dy, dx * dx))) : 0.f;
          Z += w;
          wl = fmaf(w, l2, wl);
          se += ok ? ex2a(l2 - k2) : 0.f;
          if (ok && l2 > mx) { mx = l2; am = q; }              // first maximal index (q ascends)
        }
      } else {
        float gv[8];
        const float4 la = *reinterpret_cast<const float4*>(s_lse_o + q0), lb = *reinterpret_cast<const float4*>(s_lse_o + q0 + 4);
        const float4 za = *reinterpret_cast<const float4*>(s_zinv_o + q0), zb = *reinterpret_cast<const float4*>(s_zinv_o + q0 + 4);
        const float lo_[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
        const float zo_[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int q = q0 + i;
          const float l2 = v[i] * k2;
          const float dy = cy[i] - my_y, dx = cx[i] - my_x;
          const float w = ex2a(-DA_LOG2E * sqrta(fmaf(dy, dy, dx * dx)));
          const float gg = coef * (ex2a(l2 - lse_own) - w * zinv_own) + coef_o * (ex2a(l2 - lo_[i]) - w * zo_[i]);
          gv[i] = (q < S && row_ok) ? gg : 0.f;
        }
        uint4 gh, gl;
        split8(gv, gh, gl);
        const uint32_t off = (uint32_t)(c0 >> 3) * g.lbo_g + (uint32_t)row * 16;
        *reinterpret_cast<uint4*>(Ghi + off) = gh;
        *reinterpret_cast<uint4*>(Glo + off) = gl;
      }
    };
    const uint32_t trow = tP + ((uint32_t)(qtr * 32) << 16);
    for (int c0 = part * 8; c0 < g.NC; c0 += 2 * DA_PARTS * 8) {
      const int c1 = c0 + DA_PARTS * 8;
      float va[8], vb[8];
      if (c1 < g.NC) {
        tmem_ld8_issue(trow + (uint32_t)c0, va);
        tmem_ld8_issue(trow + (uint32_t)c1, vb);
        tmem_ld_wait();
        group(va, c0);
        group(vb, c1);
      } else {
        tmem_ld8(trow + (uint32_t)c0, va);
        group(va, c0);
      }
    }
    tc_fence_before();
    if (BWD) {
      // ---- dXn[row][ch] += sum_q G[row][q] * Y[q][ch]   (A = G K-major, B = the staged Y slab read MN-major)
      fence_proxy_async();
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          const uint32_t gh = smem_u32(Ghi), gl = smem_u32(Glo), yh = ys, yl = ys + g.half;
          // MN-major no-swizzle canonical layout ((1,n),(8,k)):((X,SBO),(1,LBO)) in 16-byte units: the 16-byte channel
          // blocks are SBO = lbo apart, the 8-row K groups LBO = 128 bytes apart (roles swapped w.r.t. K-major; verified
          // on B200 against the fp32 statement)
          const uint32_t b_lbo = 128u, b_sbo = g.lbo;
#pragma unroll 1
          for (int ks = 0; ks < g.NC / 16; ++ks) {
            const uint64_t ah = smem_desc(gh + 2 * ks * g.lbo_g, g.lbo_g, 128), al = smem_desc(gl + 2 * ks * g.lbo_g, g.lbo_g, 128);
            const uint64_t bh = smem_desc(yh + ks * 256, b_lbo, b_sbo), bl = smem_desc(yl + ks * 256, b_lbo, b_sbo);
            umma_bf16(tdX, ah, bh, idesc2, (c > 0 || ks > 0) ? 1u : 0u);
            umma_bf16(tdX, al, bh, idesc2, 1u);
            umma_bf16(tdX, ah, bl, idesc2, 1u);
          }
          umma_commit(BAR(5));
        }
        __syncwarp();
      }
      mbar_wait(BAR(5), (uint32_t)(c & 1));
      tc_fence_after();
      if (c + 1 < g.nch && warp == 0) {
        // the Y slab and the G slab are free: fetch the next chunk and issue its affinity MMAs
        if (lane == 0) {
          mbar_expect_tx(BAR(1), g.slab);
          tma_bulk_g2s(ys, Yslabs + (size_t)(c + 1) * g.slab, g.slab, BAR(1));
        }
        __syncwarp();
        mbar_wait(BAR(1), (uint32_t)((c + 1) & 1));
        tc_fence_after();
        if (elect_one()) {
          da_mma_xy(tmem, xs, ys, g, idesc1);
          umma_commit(BAR(3));
        }
        __syncwarp();
      }
    } else {
      __syncthreads();         // every warp has drained this accumulator before the MMAs of chunk c+2 overwrite it
    }
  }

  if (!BWD) {
    // ---- combine the column parts of every row; the first maximal index wins (torch.argmax)
    if (part > 0) {
      float* r = s_red + ((part - 1) * 128 + row) * 8;
      r[0] = mx; r[1] = se; r[2] = Z; r[3] = wl; r[4] = __int_as_float(am);
    }
    __syncthreads();
    if (part == 0 && row_ok) {
#pragma unroll
      for (int pp = 0; pp < DA_PARTS - 1; ++pp) {
        const float* r = s_red + (pp * 128 + row) * 8;
        const float mx2 = r[0];
        const int am2 = __float_as_int(r[4]);
        if (mx2 > mx || (mx2 == mx && am2 < am)) { am = am2; mx = mx2; }
        se += r[1];                                      // same fixed shift in every part
        Z += r[2];
        wl += r[3];
      }
      float* o = p.stat + (((long)b * 2 + side) * S + grow) * 4;
      o[0] = p.inv_T + logf(se);                         // lse = 1/T + ln sum exp(l - 1/T)
      o[1] = Z;
      o[2] = wl * DA_LN2;                                // sum w * l, back from log2 units
      o[3] = (am == grow) ? 1.f : 0.f;
    }
  } else {
    // ---- L2-norm backward of the strip rows: dx = inv_T * (dXn - xn <dXn, xn>) / |x|, then atomic scatter
    constexpr int CW = DA_C / DA_PARTS;                 // channels per thread
    const float* xrow = Xmap + s_off[row_ok ? grow : 0] + part * CW;
    float dot = 0.f, ss = 0.f;
    for (int c0 = 0; c0 < CW; c0 += 8) {
      float v[8];
      tmem_ld8(tdX + ((uint32_t)(qtr * 32) << 16) + (uint32_t)(part * CW + c0), v);
      const float4 xa = __ldg(reinterpret_cast<const float4*>(xrow + c0)), xb = __ldg(reinterpret_cast<const float4*>(xrow + c0 + 4));
      const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) { dot = fmaf(v[i], x[i], dot); ss = fmaf(x[i], x[i], ss); }
    }
    s_red[(part * 128 + row) * 2 + 0] = dot;
    s_red[(part * 128 + row) * 2 + 1] = ss;
    __syncthreads();                                  // also: all MMAs are complete, the X / Y slabs are free
    dot = 0.f; ss = 0.f;
#pragma unroll
    for (int pp = 0; pp < DA_PARTS; ++pp) { dot += s_red[(pp * 128 + row) * 2]; ss += s_red[(pp * 128 + row) * 2 + 1]; }
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    const float dotn = dot * inv;                     // <dXn, xn>
    float* stage = reinterpret_cast<float*>(Xs);      // [128][129] fp32 (X slab + the head of the Y slab)
    for (int c0 = 0; c0 < CW; c0 += 8) {
      float v[8];
      tmem_ld8(tdX + ((uint32_t)(qtr * 32) << 16) + (uint32_t)(part * CW + c0), v);
      const float4 xa = __ldg(reinterpret_cast<const float4*>(xrow + c0)), xb = __ldg(reinterpret_cast<const float4*>(xrow + c0 + 4));
      const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) stage[row * 129 + part * CW + c0 + i] = p.inv_T * (v[i] - x[i] * inv * dotn) * inv;
    }
    __syncthreads();
    float* dmap = (side == 0 ? p.dG1 : p.dG2) + (long)b * g.HW * DA_C;
    for (int r = warp; r < row_end - row0; r += DA_THREADS / 32) {
      float* dst = dmap + s_off[row0 + r];
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(dst + lane + 32 * i, stage[r * 129 + lane + 32 * i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace

extern "C" {

int hcm_dense_finish(const float* stat, const float* kept, const long long* use_depth, int B, int S, float* fin,
                     cudaStream_t stream);

// bytes of the operand-slab workspace of hcm_dense_affinity_fwd / _bwd
long hcm_dense_affinity_work_bytes(int B, int S) {
  DaGeo g = da_geo(S, 1, false);
  return (long)B * 2 * g.nch * g.slab;
}

static int da_prep(const DaParams& p, int B, cudaStream_t stream) {
  dim3 grid((p.g.Spad + 7) / 8, 2, B);
  dense_prep_kernel<<<grid, 256, 0, stream>>>(p);
  HCM_LAUNCH_CHECK("dense_affinity (operand slabs)");
  return HCM_OK;
}

// stat [B][2][S][4] scratch (kept for the backward); fin[5] = loss_r2d, loss_d2r, acc_r2d, acc_d2r, B';
// work: hcm_dense_affinity_work_bytes(B, S) bytes, 128-byte aligned (the normalised bf16 hi/lo operand slabs; the backward
// re-uses them when called with prepared = 1 on the same G1, G2, pix)
int hcm_dense_affinity_fwd(const float* G1, const float* G2, const long long* pix, const float* kept,
                           const long long* use_depth, int B, int S, int h, int dim, float inv_T, float* stat, float* fin,
                           void* work, cudaStream_t stream) {
  HCM_CHECK_ARG(G1 && G2 && pix && kept && stat && fin && work, "dense_affinity_fwd: null pointer");
  HCM_CHECK_ARG(dim == DA_C && B >= 1 && S >= 1 && h >= 1, "dense_affinity_fwd: bad args (dim=%d B=%d S=%d h=%d)", dim, B, S, h);
  HCM_CHECK_ARG(((size_t)work & 127) == 0, "dense_affinity_fwd: workspace must be 128-byte aligned");
  DaParams p = {};
  p.G1 = G1; p.G2 = G2; p.pix = pix; p.kept = kept; p.stat = stat; p.inv_T = inv_T; p.work = reinterpret_cast<uint8_t*>(work);
  p.g = da_geo(S, h, false);
  HCM_CHECK_ARG(p.g.smem_bytes <= 227 * 1024, "dense_affinity_fwd: shared memory (%u bytes)", p.g.smem_bytes);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(dense_affinity_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr = true;
  }
  int rc = da_prep(p, B, stream);
  if (rc != HCM_OK) return rc;
  dim3 grid(p.g.nch, 2, B);
  dense_affinity_kernel<false><<<grid, DA_THREADS, p.g.smem_bytes, stream>>>(p);
  HCM_LAUNCH_CHECK("dense_affinity_fwd");
  return hcm_dense_finish(stat, kept, use_depth, B, S, fin, stream);
}

// dG1, dG2 [B][h*h][128] are ACCUMULATED into (atomics: sampled pixels repeat); the caller zeroes them.
// prepared = 1: `work` still holds the slabs written by hcm_dense_affinity_fwd for the same G1, G2, pix; 0: they are rebuilt.
int hcm_dense_affinity_bwd(const float* G1, const float* G2, const long long* pix, const float* stat, const float* kept,
                           const float* fin, int B, int S, int h, int dim, float inv_T, float gscale_r2d, float gscale_d2r, float* dG1,
                           float* dG2, void* work, int prepared, cudaStream_t stream) {
  HCM_CHECK_ARG(G1 && G2 && pix && stat && kept && fin && dG1 && dG2 && work, "dense_affinity_bwd: null pointer");
  HCM_CHECK_ARG(dim == DA_C && B >= 1 && S >= 1 && h >= 1, "dense_affinity_bwd: bad args (dim=%d B=%d S=%d h=%d)", dim, B, S, h);
  HCM_CHECK_ARG(((size_t)work & 127) == 0, "dense_affinity_bwd: workspace must be 128-byte aligned");
  DaParams p = {};
  p.G1 = G1; p.G2 = G2; p.pix = pix; p.kept = kept; p.fin = fin; p.stat = const_cast<float*>(stat);
  p.dG1 = dG1; p.dG2 = dG2; p.inv_T = inv_T; p.gscale = gscale_r2d; p.gscale_o = gscale_d2r; p.work = reinterpret_cast<uint8_t*>(work);
  p.g = da_geo(S, h, true);
  HCM_CHECK_ARG(p.g.smem_bytes <= 227 * 1024, "dense_affinity_bwd: shared memory (%u bytes)", p.g.smem_bytes);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(dense_affinity_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr = true;
  }
  if (!prepared) {
    int rc = da_prep(p, B, stream);
    if (rc != HCM_OK) return rc;
  }
  dim3 grid(p.g.nch, 2, B);
  dense_affinity_kernel<true><<<grid, DA_THREADS, p.g.smem_bytes, stream>>>(p);
  HCM_LAUNCH_CHECK("dense_affinity_bwd");
  return HCM_OK;
}

}  // extern "C"
