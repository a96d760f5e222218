// Train-mode batch-norm over channels-last activations ([P, C] row-major, P = B*H*W pixels or B*J
// joints), forward and backward, as HBM-bound column reductions + elementwise passes.
// Replaces nn.BatchNorm2d(momentum=0.01) (networks/official_hrnet/official_hrnet.py:22-23) and
// nn.BatchNorm1d (networks/SGCN/sem_gcn.py:13) together with the ReLU / residual-add that follow
// them (official_hrnet.py:44-60, 86-101).
//
// Layout trick: every kernel runs with blockDim.x = the largest multiple of C/VEC <= 256 and walks
// the flat array with a stride that is a multiple of C, so a thread's channel never changes and no
// per-element modulo is needed; VEC-wide vector loads keep the accesses coalesced.
// Reductions are two-level and deterministic: per-CTA fp32 partials -> fp64 finalize.
#include "common.cuh"

namespace {

constexpr int MAXT = 256;

inline int vec_for(long total, int C) { return (C % 4 == 0 && total % 4 == 0) ? 4 : ((C % 2 == 0 && total % 2 == 0) ? 2 : 1); }
inline int threads_for(int C, int vec) { int cv = C / vec; return (MAXT / cv) * cv; }

template <int VEC> struct VecT;
template <> struct VecT<1> { typedef float T; };
template <> struct VecT<2> { typedef float2 T; };
template <> struct VecT<4> { typedef float4 T; };

template <int VEC>
__device__ __forceinline__ void vload(const float* p, float (&v)[VEC]) {
  typename VecT<VEC>::T t = *reinterpret_cast<const typename VecT<VEC>::T*>(p);
  const float* f = reinterpret_cast<const float*>(&t);
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = f[i];
}
template <int VEC>
__device__ __forceinline__ void vstore(float* p, const float (&v)[VEC]) {
  typename VecT<VEC>::T t;
  float* f = reinterpret_cast<float*>(&t);
#pragma unroll
  for (int i = 0; i < VEC; ++i) f[i] = v[i];
  *reinterpret_cast<typename VecT<VEC>::T*>(p) = t;
}

struct FwdFin {      // forward finalize outputs / state (bn_finalize)
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; long long* nbt;
  float momentum, eps;
  float* scale; float* shift; float* mean_out; float* invstd_out;
};
struct BwdFin {      // backward finalize (bn_bwd_finalize)
  const float* gamma; const float* mean; const float* invstd;
  float* dgamma; float* dbeta; float* k1; float* k2; float* k3;
};

// channel c from its fp64 sums s = sum y, q = sum y^2
__device__ __forceinline__ void fwd_finalize_channel(const FwdFin& f, int c, double s, double q, double count) {
  const double m = s / count;
  double var = q / count - m * m;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)f.eps));
  const float g = f.gamma ? f.gamma[c] : 1.f, b = f.beta ? f.beta[c] : 0.f;
  f.scale[c] = g * is;
  f.shift[c] = b - (float)m * g * is;
  f.mean_out[c] = (float)m;
  f.invstd_out[c] = is;
  if (f.running_mean) {
    const double unb = (count > 1.0) ? var * count / (count - 1.0) : var;
    f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * (float)m;
    f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)unb;
  }
}
// channel c from s = sum g, q = sum g*yhat: dgamma, dbeta and dy = k1*g + k2*y + k3
__device__ __forceinline__ void bwd_finalize_channel(const BwdFin& f, int c, double s, double q, double count) {
  const float g = f.gamma ? f.gamma[c] : 1.f;
  const double is = f.invstd[c], mu = f.mean[c];
  if (f.dgamma) f.dgamma[c] = (float)q;
  if (f.dbeta) f.dbeta[c] = (float)s;
  const double m1 = s / count, m2 = q / count;
  const double a = (double)g * is;
  f.k1[c] = (float)a;
  f.k2[c] = (float)(-a * is * m2);
  f.k3[c] = (float)(-a * m1 + a * is * mu * m2);
}

// Finalize inside the statistics kernel: the CTA that takes the last ticket reduces all partial rows (fp64, fixed order ->
// deterministic) and writes the per-channel coefficients, saving a ~5 us kernel + its launch gap per BatchNorm.
// Lanes run over channels (coalesced rows of `part`), warps over partial rows; 32 channels per round.
template <class Fin, int MODE>
__device__ __forceinline__ void finalize_in_cta(const float* part, int nparts, int C, double count, const Fin& f, double* s_acc) {
  // thread t -> (channel c = t % C, row slice r = t / C of R = blockDim.x / C slices): consecutive threads read consecutive
  // channels of a partial row (coalesced); 8 rows (16 loads) in flight per thread — the tail is pure L2 latency
  const int t = threadIdx.x;
  const int R = blockDim.x / C;                          // >= 1 (blockDim.x >= C for every launch plan)
  const int c = t % C, r = t / C;
  double s = 0.0, q = 0.0;
  if (r < R) {
    int i = r;
    for (; i + 7 * R < nparts; i += 8 * R) {
      float a[8], b[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a[k] = __ldcg(part + ((long)(i + k * R) * 2 + 0) * C + c);
        b[k] = __ldcg(part + ((long)(i + k * R) * 2 + 1) * C + c);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { s += (double)a[k]; q += (double)b[k]; }
    }
    for (; i < nparts; i += R) {
      s += (double)__ldcg(part + ((long)i * 2 + 0) * C + c);
      q += (double)__ldcg(part + ((long)i * 2 + 1) * C + c);
    }
  }
  s_acc[t] = s;
  s_acc[256 + t] = q;
  __syncthreads();
  if (t < C) {
    double ss = 0.0, qq = 0.0;
    for (int k = 0; k < R; ++k) { ss += s_acc[t + k * C]; qq += s_acc[256 + t + k * C]; }
    if constexpr (MODE == 0) fwd_finalize_channel(f, t, ss, qq, count);
    else bwd_finalize_channel(f, t, ss, qq, count);
  }
}

// ---- column statistics: part[blk][0][c] = sum y, part[blk][1][c] = sum y^2 over the CTA's slice ----
// MODE 0: plain (y).  MODE 1: backward sums (g, g*yhat) with g = dz * [mask > 0].
template <int VEC, int MODE, class Fin>
__global__ void colstat_kernel(const float* __restrict__ a, const float* __restrict__ y, const float* __restrict__ mask,
                               const float* __restrict__ mean, const float* __restrict__ invstd,
                               const float* __restrict__ msc, const float* __restrict__ msh, long total, int C,
                               long per_cta, float* part, unsigned int* counter, double count, const Fin fin) {
  __shared__ float s0[MAXT * 4], s1[MAXT * 4];
  pdl_wait();
  pdl_trigger();                 // single-wave grid (<= 4 CTAs per SM): dependents may become resident right away
  const int t = threadIdx.x, nt = blockDim.x;
  const int cv = C / VEC;
  const int c0 = (t % cv) * VEC;
  float mu[VEC], is[VEC], ms[VEC], mh[VEC];
  if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      mu[i] = mean[c0 + i]; is[i] = invstd[c0 + i];
      ms[i] = msc ? msc[c0 + i] : 0.f; mh[i] = msc ? msh[c0 + i] : 0.f;
    }
  }
  float a0[VEC], a1[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) { a0[i] = 0.f; a1[i] = 0.f; }
  const long beg = (long)blockIdx.x * per_cta;
  const long end = min(total, beg + per_cta);
  for (long e = beg + (long)t * VEC; e < end; e += (long)nt * VEC) {
    float v[VEC];
    vload<VEC>(a + e, v);
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) { a0[i] += v[i]; a1[i] = fmaf(v[i], v[i], a1[i]); }
    } else {
      float yy[VEC];
      vload<VEC>(y + e, yy);
      if (mask) {
        float m[VEC];
        vload<VEC>(mask + e, m);
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = (m[i] > 0.f) ? v[i] : 0.f;
      } else if (msc) {  // ReLU mask recomputed from the raw conv output: relu(y*scale+shift) > 0
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = (fmaf(yy[i], ms[i], mh[i]) > 0.f) ? v[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) { a0[i] += v[i]; a1[i] = fmaf(v[i], (yy[i] - mu[i]) * is[i], a1[i]); }
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) { s0[t * VEC + i] = a0[i]; s1[t * VEC + i] = a1[i]; }
  __syncthreads();
  // threads 0..C-1 fold the copies of their channel: element index j*C + c  <->  thread (j*cv + c/VEC), slot c%VEC
  if (t < C) {
    float r0 = 0.f, r1 = 0.f;
    const int reps = nt / cv;
    for (int j = 0; j < reps; ++j) {
      const int idx = (j * cv + t / VEC) * VEC + (t % VEC);
      r0 += s0[idx]; r1 += s1[idx];
    }
    part[((long)blockIdx.x * 2 + 0) * C + t] = r0;
    part[((long)blockIdx.x * 2 + 1) * C + t] = r1;
  }
  if (counter) {
    // last-ticket CTA finalizes (threadfence reduction pattern); it also re-arms the counter for the next launch
    __shared__ int s_last;
    __shared__ double s_acc[8 * 2 * 32];
    __threadfence();
    __syncthreads();
    if (t == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {
      __threadfence();
      finalize_in_cta<Fin, MODE>(part, (int)gridDim.x, C, count, fin, s_acc);
      if (t == 0) {
        *counter = 0u;
        if constexpr (MODE == 0) { if (fin.nbt) *fin.nbt += 1; }
      }
    }
  }
}

// one CTA (128 threads) per channel: fp64 reduction of the partial rows, then the per-channel coefficients
__device__ __forceinline__ void block_sum2_d(double& s, double& q, double* red) {
  s = warp_sum_d(s); q = warp_sum_d(q);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[w] = s; red[4 + w] = q; }
  __syncthreads();
  s = red[0] + red[1] + red[2] + red[3];
  q = red[4] + red[5] + red[6] + red[7];
}

__global__ void __launch_bounds__(128) bn_finalize_kernel(const float* __restrict__ part, int nparts, int C, double count,
                                                          const FwdFin f) {
  __shared__ double red[8];
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0 && f.nbt) *f.nbt += 1;
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 128) {
    s += (double)part[((long)i * 2 + 0) * C + c];
    q += (double)part[((long)i * 2 + 1) * C + c];
  }
  block_sum2_d(s, q, red);
  if (threadIdx.x == 0) fwd_finalize_channel(f, c, s, q, count);
}

// backward finalize: dgamma, dbeta and dy = k1*g + k2*y + k3
__global__ void __launch_bounds__(128) bn_bwd_finalize_kernel(const float* __restrict__ part, int nparts, int C, double count,
                                                              const BwdFin f) {
  __shared__ double red[8];
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x;
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 128) {
    s += (double)part[((long)i * 2 + 0) * C + c];
    q += (double)part[((long)i * 2 + 1) * C + c];
  }
  block_sum2_d(s, q, red);
  if (threadIdx.x == 0) bwd_finalize_channel(f, c, s, q, count);
}

// z = act(y*scale + shift + (res*res_scale + res_shift))
template <int VEC>
__global__ void bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                                const float* __restrict__ res, const float* __restrict__ res_scale,
                                const float* __restrict__ res_shift, int relu, float* __restrict__ out, long total, int C) {
  pdl_wait();
  const int t = threadIdx.x, nt = blockDim.x;
  const int cv = C / VEC;
  const int c0 = (t % cv) * VEC;
  float sc[VEC], sh[VEC], rs[VEC], rh[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    sc[i] = scale ? scale[c0 + i] : 1.f; sh[i] = shift ? shift[c0 + i] : 0.f;
    rs[i] = res_scale ? res_scale[c0 + i] : 1.f; rh[i] = res_shift ? res_shift[c0 + i] : 0.f;
  }
  const long stride = (long)gridDim.x * nt * VEC;
  for (long e = ((long)blockIdx.x * nt + t) * VEC; e < total; e += stride) {
    float v[VEC];
    vload<VEC>(y + e, v);
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = fmaf(v[i], sc[i], sh[i]);
    if (res) {
      float r[VEC];
      vload<VEC>(res + e, r);
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[i] += fmaf(r[i], rs[i], rh[i]);
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    vstore<VEC>(out + e, v);
  }
  pdl_trigger();
}

// g = dz*[mask>0];  dy = k1*g + k2*y + k3;  optionally g_out (+)= g
template <int VEC>
__global__ void bn_bwd_apply_kernel(const float* dz, const float* __restrict__ mask, const float* __restrict__ y,
                                    const float* __restrict__ msc, const float* __restrict__ msh,
                                    const float* __restrict__ k1, const float* __restrict__ k2, const float* __restrict__ k3,
                                    float* dy, float* g_out, int g_accumulate, long total, int C) {
  pdl_wait();
  const int t = threadIdx.x, nt = blockDim.x;
  const int cv = C / VEC;
  const int c0 = (t % cv) * VEC;
  float a[VEC], b[VEC], c[VEC], ms[VEC], mh[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    a[i] = k1[c0 + i]; b[i] = k2[c0 + i]; c[i] = k3[c0 + i];
    ms[i] = msc ? msc[c0 + i] : 0.f; mh[i] = msc ? msh[c0 + i] : 0.f;
  }
  const long stride = (long)gridDim.x * nt * VEC;
  for (long e = ((long)blockIdx.x * nt + t) * VEC; e < total; e += stride) {
    float g[VEC], yy[VEC], o[VEC];
    vload<VEC>(dz + e, g);
    vload<VEC>(y + e, yy);
    if (mask) {
      float m[VEC];
      vload<VEC>(mask + e, m);
#pragma unroll
      for (int i = 0; i < VEC; ++i) g[i] = (m[i] > 0.f) ? g[i] : 0.f;
    } else if (msc) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) g[i] = (fmaf(yy[i], ms[i], mh[i]) > 0.f) ? g[i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) o[i] = fmaf(a[i], g[i], fmaf(b[i], yy[i], c[i]));
    vstore<VEC>(dy + e, o);
    if (g_out) {
      if (g_accumulate) {
        float old[VEC];
        vload<VEC>(g_out + e, old);
#pragma unroll
        for (int i = 0; i < VEC; ++i) g[i] += old[i];
      }
      vstore<VEC>(g_out + e, g);
    }
  }
  pdl_trigger();
}

// generic elementwise helpers (float4 main body + scalar tail)
__global__ void relu_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, float* g, int accumulate,
                                long total) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    float v = (out[e] > 0.f) ? dout[e] : 0.f;
    g[e] = accumulate ? g[e] + v : v;
  }
}
__global__ void axpy_kernel(float* __restrict__ dst, const float* __restrict__ src, float alpha, long total) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) dst[e] = fmaf(alpha, src[e], dst[e]);
}

inline int ew_grid(long total, int per_thread_elems, int threads) {
  long g = (total + (long)threads * per_thread_elems - 1) / ((long)threads * per_thread_elems);
  long cap = 148L * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

inline void colstat_plan(long total, int C, int* vec, int* threads, int* nparts, long* per_cta) {
  *vec = vec_for(total, C);
  *threads = threads_for(C, *vec);
  long unit = (long)(*threads) * (*vec);           // multiple of C
  long want = 148L * 4;
  long units = (total + unit - 1) / unit;
  long per = (units + want - 1) / want;
  if (per < 4) per = (units < 4) ? units : 4;
  if (per < 1) per = 1;
  *per_cta = per * unit;
  *nparts = (int)((total + *per_cta - 1) / *per_cta);
}

template <int MODE, class Fin>
int launch_colstat(const float* a, const float* y, const float* mask, const float* mean, const float* invstd, const float* msc,
                   const float* msh, long P, int C, float* part, unsigned int* counter, const Fin& fin, cudaStream_t stream) {
  int vec, threads, nparts;
  long per;
  const long total = P * C;
  colstat_plan(total, C, &vec, &threads, &nparts, &per);
  if (vec == 4) hcm_launch_pdl(colstat_kernel<4, MODE, Fin>, nparts, threads, 0, stream, a, y, mask, mean, invstd, msc, msh, total, C, per, part, counter, (double)P, fin);
  else if (vec == 2) hcm_launch_pdl(colstat_kernel<2, MODE, Fin>, nparts, threads, 0, stream, a, y, mask, mean, invstd, msc, msh, total, C, per, part, counter, (double)P, fin);
  else hcm_launch_pdl(colstat_kernel<1, MODE, Fin>, nparts, threads, 0, stream, a, y, mask, mean, invstd, msc, msh, total, C, per, part, counter, (double)P, fin);
  return HCM_OK;
}

}  // namespace

extern "C" {

int hcm_colstat_rows(long P, int C) {
  int vec, threads, nparts;
  long per;
  colstat_plan(P * C, C, &vec, &threads, &nparts, &per);
  return nparts;
}

// part [hcm_colstat_rows][2][C]: per-CTA sums of y and y^2
int hcm_bn_stats(const float* y, long P, int C, float* part, cudaStream_t stream) {
  HCM_CHECK_ARG(y && part && C >= 1 && C <= 256, "bn_stats: bad args (C=%d)", C);
  launch_colstat<0>(y, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, P, C, part, nullptr, FwdFin{}, stream);
  HCM_LAUNCH_CHECK("bn_stats");
  return HCM_OK;
}

int hcm_bn_finalize(const float* part, int nparts, int C, long count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                    float* scale, float* shift, float* mean, float* invstd, cudaStream_t stream) {
  HCM_CHECK_ARG(part && scale && shift && mean && invstd && nparts >= 1, "bn_finalize: bad args");
  const FwdFin f = {gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, scale, shift, mean, invstd};
  hcm_launch_pdl(bn_finalize_kernel, C, 128, 0, stream, part, nparts, C, (double)count, f);
  HCM_LAUNCH_CHECK("bn_finalize");
  return HCM_OK;
}

// hcm_bn_stats + hcm_bn_finalize in ONE launch (count = P): the last CTA to finish reduces the partial rows.
// `counter` is one zero-initialised uint32 owned by the caller (per stream); the kernel leaves it at zero.
int hcm_bn_stats_finalize(const float* y, long P, int C, float* part, unsigned int* counter, const float* gamma,
                          const float* beta, float* running_mean, float* running_var, long long* num_batches_tracked,
                          float momentum, float eps, float* scale, float* shift, float* mean, float* invstd,
                          cudaStream_t stream) {
  HCM_CHECK_ARG(y && part && counter && scale && shift && mean && invstd && C >= 1 && C <= 256,
                "bn_stats_finalize: bad args (C=%d)", C);
  const FwdFin f = {gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, scale, shift, mean, invstd};
  launch_colstat<0>(y, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, P, C, part, counter, f, stream);
  HCM_LAUNCH_CHECK("bn_stats_finalize");
  return HCM_OK;
}

int hcm_bn_apply(const float* y, const float* scale, const float* shift, const float* res, const float* res_scale,
                 const float* res_shift, int relu, float* out, long P, int C, cudaStream_t stream) {
  HCM_CHECK_ARG(y && out && C >= 1 && C <= 256, "bn_apply: bad args (C=%d)", C);
  const long total = P * C;
  const int vec = vec_for(total, C), threads = threads_for(C, vec);
  const int grid = ew_grid(total, vec * 4, threads);
  if (vec == 4) hcm_launch_pdl(bn_apply_kernel<4>, grid, threads, 0, stream, y, scale, shift, res, res_scale, res_shift, relu, out, total, C);
  else if (vec == 2) hcm_launch_pdl(bn_apply_kernel<2>, grid, threads, 0, stream, y, scale, shift, res, res_scale, res_shift, relu, out, total, C);
  else hcm_launch_pdl(bn_apply_kernel<1>, grid, threads, 0, stream, y, scale, shift, res, res_scale, res_shift, relu, out, total, C);
  HCM_LAUNCH_CHECK("bn_apply");
  return HCM_OK;
}

// part [hcm_colstat_rows][2][C]: per-CTA sums of g and g*yhat, g = dz*[mask>0] (mask may be null)
int hcm_bn_bwd_reduce(const float* dz, const float* mask, const float* mask_scale, const float* mask_shift,
                      const float* y, const float* mean, const float* invstd, long P, int C, float* part,
                      cudaStream_t stream) {
  HCM_CHECK_ARG(dz && y && mean && invstd && part && C >= 1 && C <= 256, "bn_bwd_reduce: bad args (C=%d)", C);
  launch_colstat<1>(dz, y, mask, mean, invstd, mask_scale, mask_shift, P, C, part, nullptr, BwdFin{}, stream);
  HCM_LAUNCH_CHECK("bn_bwd_reduce");
  return HCM_OK;
}

int hcm_bn_bwd_finalize(const float* part, int nparts, int C, long count, const float* gamma, const float* mean,
                        const float* invstd, float* dgamma, float* dbeta, float* k1, float* k2, float* k3,
                        cudaStream_t stream) {
  HCM_CHECK_ARG(part && mean && invstd && k1 && k2 && k3, "bn_bwd_finalize: bad args");
  const BwdFin f = {gamma, mean, invstd, dgamma, dbeta, k1, k2, k3};
  hcm_launch_pdl(bn_bwd_finalize_kernel, C, 128, 0, stream, part, nparts, C, (double)count, f);
  HCM_LAUNCH_CHECK("bn_bwd_finalize");
  return HCM_OK;
}

// hcm_bn_bwd_reduce + hcm_bn_bwd_finalize in ONE launch (count = P); `counter` as in hcm_bn_stats_finalize
int hcm_bn_bwd_reduce_finalize(const float* dz, const float* mask, const float* mask_scale, const float* mask_shift,
                               const float* y, const float* mean, const float* invstd, long P, int C, float* part,
                               unsigned int* counter, const float* gamma, float* dgamma, float* dbeta, float* k1, float* k2,
                               float* k3, cudaStream_t stream) {
  HCM_CHECK_ARG(dz && y && mean && invstd && part && counter && k1 && k2 && k3 && C >= 1 && C <= 256,
                "bn_bwd_reduce_finalize: bad args (C=%d)", C);
  const BwdFin f = {gamma, mean, invstd, dgamma, dbeta, k1, k2, k3};
  launch_colstat<1>(dz, y, mask, mean, invstd, mask_scale, mask_shift, P, C, part, counter, f, stream);
  HCM_LAUNCH_CHECK("bn_bwd_reduce_finalize");
  return HCM_OK;
}

int hcm_bn_bwd_apply(const float* dz, const float* mask, const float* mask_scale, const float* mask_shift,
                     const float* y, const float* k1, const float* k2, const float* k3, float* dy, float* g_out,
                     int g_accumulate, long P, int C, cudaStream_t stream) {
  HCM_CHECK_ARG(dz && y && k1 && k2 && k3 && dy && C >= 1 && C <= 256, "bn_bwd_apply: bad args (C=%d)", C);
  const long total = P * C;
  const int vec = vec_for(total, C), threads = threads_for(C, vec);
  const int grid = ew_grid(total, vec * 4, threads);
  if (vec == 4) hcm_launch_pdl(bn_bwd_apply_kernel<4>, grid, threads, 0, stream, dz, mask, y, mask_scale, mask_shift, k1, k2, k3, dy, g_out, g_accumulate, total, C);
  else if (vec == 2) hcm_launch_pdl(bn_bwd_apply_kernel<2>, grid, threads, 0, stream, dz, mask, y, mask_scale, mask_shift, k1, k2, k3, dy, g_out, g_accumulate, total, C);
  else hcm_launch_pdl(bn_bwd_apply_kernel<1>, grid, threads, 0, stream, dz, mask, y, mask_scale, mask_shift, k1, k2, k3, dy, g_out, g_accumulate, total, C);
  HCM_LAUNCH_CHECK("bn_bwd_apply");
  return HCM_OK;
}

// g (+)= dout * [out > 0]
int hcm_relu_bwd(const float* dout, const float* out, float* g, int accumulate, long total, cudaStream_t stream) {
  HCM_CHECK_ARG(dout && out && g, "relu_bwd: null pointer");
  relu_bwd_kernel<<<ew_grid(total, 4, 256), 256, 0, stream>>>(dout, out, g, accumulate, total);
  HCM_LAUNCH_CHECK("relu_bwd");
  return HCM_OK;
}

// dst += alpha * src
int hcm_axpy(float* dst, const float* src, float alpha, long total, cudaStream_t stream) {
  HCM_CHECK_ARG(dst && src, "axpy: null pointer");
  axpy_kernel<<<ew_grid(total, 4, 256), 256, 0, stream>>>(dst, src, alpha, total);
  HCM_LAUNCH_CHECK("axpy");
  return HCM_OK;
}

}  // extern "C"
