// PointNet++ set-abstraction / feature-propagation primitives: the nine entry points of the
// reference's ONLY native extension, `pointnet2_cuda` (pycontrast/networks/pointnet2/src/pointnet2_api.cpp:10-23), consumed by
// networks/pointnet2/pointnet2_utils.py:10-228 (FurthestPointSampling, GatherOperation, ThreeNN, ThreeInterpolate,
// GroupingOperation, BallQuery) inside Pointnet2MSG — the depth branch of the HRNetPN variant (networks/build_backbone.py:305-514).
// Written from the semantics of those calls, sm_100a-first:
//   fps             one CTA per cloud, the cloud's coordinates AND running minimum distances live in registers (the reference
//                   streams both through global memory every iteration), argmax by redux.sync + one shared-memory exchange:
//                   ONE __syncthreads per selected point instead of eleven;
//   ball_query      one warp per query: 32 candidates per step, in-radius lanes ranked with a ballot (index order preserved), early exit;
//   three_nn        known points streamed through shared-memory tiles and broadcast to 256 queries per CTA;
//   group / gather / three_interpolate (+ their gradients): flat index kernels, coalesced along the point axis.
// Results are IDENTICAL to the reference kernels, including which of several equidistant points wins: the reference's choice
// depends on its thread layout (thread t scans points t, t+T, ..; its shared-memory tree prefers the lower thread on ties), which
// is restated here as an explicit total order (see fps_better) so that any reduction shape reproduces it.
// tests/test_pointnet2_gpu.py checks bit-equality against the reference kernels compiled from /root/reference (oracle/_ref).
#include "common.cuh"

namespace {

__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
  return (ax - bx) * (ax - bx) + (ay - by) * (ay - by) + (az - bz) * (az - bz);
}

// ------------------------------------------------------------------------------------------ farthest point sampling
// Candidate ordering of the reference reduction: larger value wins.  Ties: the reference's thread t = k mod T scans k = t, t+T, ..
// keeping the first maximum, and its shared-memory tree (pairs (p, p + h), h = T/2 .. 1, the lower position kept on ties) prefers,
// level by level from the LAST one, the candidate whose thread index has a 0 in bit 0, then bit 1, ..: the smaller BIT-REVERSED
// thread index.  So: value desc, bitrev(k mod T) asc, k asc — packed below into one 32-bit tie key (larger = preferred):
// 0xFFFFFF - ((bitrev_T(k mod T) << 14) | k)   (T <= 1024, k < 16384).
__device__ __forceinline__ unsigned fps_tie_key(int k, int tmask, int tshift) {
  return 0xFFFFFFu - (((__brev((unsigned)(k & tmask)) >> tshift) << 14) | (unsigned)k);
}

// One CTA per cloud; thread t owns points t, t + NT, ..: coordinates and running minimum distances stay in registers for all M
// iterations.  Per iteration: PTS distance updates, then the argmax as TWO warp-wide redux.sync (max of the distance bits — the
// distances are >= 0, so their bit patterns order like the values — then max of the tie key among the lanes that hold it), one
// shared-memory exchange between the warps (double-buffered: ONE __syncthreads per selected point) and the same two redux again.
template <int NT, int PTS>
__global__ void __launch_bounds__(NT) fps_kernel(const float* __restrict__ xyz, int N, int M, int tmask, int tshift,
                                                 int* __restrict__ idx) {
  extern __shared__ float s_xyz[];                       // [N][3] copy of the cloud (the winner's coordinates are read from here)
  __shared__ unsigned s_v[2][32], s_t[2][32];
  constexpr int NW = NT / 32;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const float* cloud = xyz + (size_t)blockIdx.x * N * 3;
  int* out = idx + (size_t)blockIdx.x * M;
  for (int i = t; i < N * 3; i += NT) s_xyz[i] = cloud[i];
  __syncthreads();
  float px[PTS], py[PTS], pz[PTS], md[PTS];
#pragma unroll
  for (int j = 0; j < PTS; ++j) {
    const int k = t + j * NT;
    px[j] = k < N ? s_xyz[3 * k] : 0.f; py[j] = k < N ? s_xyz[3 * k + 1] : 0.f; pz[j] = k < N ? s_xyz[3 * k + 2] : 0.f;
    md[j] = 1e10f;                                        // pointnet2_utils.py:27 `temp.fill_(1e10)`
  }
  int old = 0;
  if (t == 0) out[0] = 0;
  for (int it = 1; it < M; ++it) {
    const float ox = s_xyz[3 * old], oy = s_xyz[3 * old + 1], oz = s_xyz[3 * old + 2];
    float bv = -1.f;
    int bk = -1;
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
      const int k = t + j * NT;
      if (k < N) {
        const float d = fminf(sqdist(px[j], py[j], pz[j], ox, oy, oz), md[j]);
        md[j] = d;
        // (equal distances inside one thread are rare: the tie key is only computed then)
        if (d > bv || (d == bv && fps_tie_key(k, tmask, tshift) > fps_tie_key(bk, tmask, tshift))) { bv = d; bk = k; }
      }
    }
    unsigned vb = bk >= 0 ? __float_as_uint(bv) : 0u;      // d >= 0: the bit pattern is monotone in the value
    unsigned tk = bk >= 0 ? fps_tie_key(bk, tmask, tshift) : 0u;
    unsigned vmax = __reduce_max_sync(0xffffffffu, vb);
    unsigned tmax = __reduce_max_sync(0xffffffffu, (vb == vmax) ? tk : 0u);
    const int buf = it & 1;
    if (lane == 0) { s_v[buf][warp] = vmax; s_t[buf][warp] = tmax; }
    __syncthreads();
    vb = lane < NW ? s_v[buf][lane] : 0u;
    tk = lane < NW ? s_t[buf][lane] : 0u;
    vmax = __reduce_max_sync(0xffffffffu, vb);
    tmax = __reduce_max_sync(0xffffffffu, (vb == vmax) ? tk : 0u);
    old = (int)((0xFFFFFFu - tmax) & 0x3FFFu);
    if (t == 0) out[it] = old;
  }
}

// ------------------------------------------------------------------------------------------ ball query
__global__ void ball_query_kernel(const float* __restrict__ new_xyz, const float* __restrict__ xyz, int N, int M, float radius2,
                                  int nsample, int* __restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (q >= M) return;
  const float* c = new_xyz + ((size_t)b * M + q) * 3;
  const float cx = c[0], cy = c[1], cz = c[2];
  const float* cloud = xyz + (size_t)b * N * 3;
  int* out = idx + ((size_t)b * M + q) * nsample;
  int cnt = 0, first = 0;
  for (int base = 0; base < N && cnt < nsample; base += 32) {
    const int k = base + lane;
    bool in = false;
    if (k < N) in = sqdist(cx, cy, cz, cloud[3 * k], cloud[3 * k + 1], cloud[3 * k + 2]) < radius2;
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (m) {
      if (cnt == 0) first = base + __ffs(m) - 1;
      const int pos = cnt + __popc(m & ((1u << lane) - 1u));
      if (in && pos < nsample) out[pos] = k;
      cnt += __popc(m);
    }
  }
  cnt = min(cnt, nsample);
  // fewer than nsample neighbours: the rest repeats the first one; none at all: the caller's zero fill (pointnet2_utils.py:218)
  for (int l = cnt + lane; l < nsample; l += 32) out[l] = first;
}

// ------------------------------------------------------------------------------------------ three nearest neighbours
constexpr int NN_TILE = 1024;
__global__ void __launch_bounds__(256) three_nn_kernel(const float* __restrict__ unknown, const float* __restrict__ known, int n, int m,
                                                       float* __restrict__ dist2, int* __restrict__ idx) {
  __shared__ float s_k[NN_TILE * 3];
  const int b = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const bool live = p < n;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (live) { const float* u = unknown + ((size_t)b * n + p) * 3; ux = u[0]; uy = u[1]; uz = u[2]; }
  const float* kb = known + (size_t)b * m * 3;
  float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;      // (the reference starts from 1e40 held in a double: above every float)
  int i1 = 0, i2 = 0, i3 = 0;
  for (int base = 0; base < m; base += NN_TILE) {
    const int cntk = min(NN_TILE, m - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cntk * 3; i += 256) s_k[i] = kb[(size_t)base * 3 + i];
    __syncthreads();
    if (live) {
      for (int j = 0; j < cntk; ++j) {
        const float d = sqdist(ux, uy, uz, s_k[3 * j], s_k[3 * j + 1], s_k[3 * j + 2]);
        const int k = base + j;
        if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
        else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
        else if (d < b3) { b3 = d; i3 = k; }
      }
    }
  }
  if (live) {
    float* d = dist2 + ((size_t)b * n + p) * 3;
    int* o = idx + ((size_t)b * n + p) * 3;
    // fewer than three known points: the reference leaves 1e40 (as float: +inf) and index 0
    d[0] = b1; d[1] = b2; d[2] = b3;
    o[0] = i1; o[1] = i2; o[2] = i3;
  }
}

// ------------------------------------------------------------------------------------------ gathers
// out[b][c][e] = points[b][c][idx[b][e]]   (e over npoint, or npoint*nsample for grouping)
__global__ void gather_kernel(const float* __restrict__ points, const int* __restrict__ idx, int C, int N, long E, float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const float* src = points + ((size_t)b * C + c) * N;
  const int* id = idx + (size_t)b * E;
  float* dst = out + ((size_t)b * C + c) * E;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long)gridDim.x * blockDim.x) dst[e] = src[id[e]];
}
// grad_points[b][c][idx[b][e]] += grad_out[b][c][e]
__global__ void gather_grad_kernel(const float* __restrict__ grad_out, const int* __restrict__ idx, int C, int N, long E,
                                   float* __restrict__ grad_points) {
  const int b = blockIdx.z, c = blockIdx.y;
  float* dst = grad_points + ((size_t)b * C + c) * N;
  const int* id = idx + (size_t)b * E;
  const float* src = grad_out + ((size_t)b * C + c) * E;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long)gridDim.x * blockDim.x) atomicAdd(dst + id[e], src[e]);
}

// out[b][c][p] = sum_j weight[b][p][j] * points[b][c][idx[b][p][j]]
__global__ void three_interpolate_kernel(const float* __restrict__ points, const int* __restrict__ idx, const float* __restrict__ weight,
                                         int C, int m, int n, float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const float* src = points + ((size_t)b * C + c) * m;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const int* id = idx + ((size_t)b * n + p) * 3;
    const float* w = weight + ((size_t)b * n + p) * 3;
    out[((size_t)b * C + c) * n + p] = w[0] * src[id[0]] + w[1] * src[id[1]] + w[2] * src[id[2]];
  }
}
__global__ void three_interpolate_grad_kernel(const float* __restrict__ grad_out, const int* __restrict__ idx,
                                              const float* __restrict__ weight, int C, int n, int m, float* __restrict__ grad_points) {
  const int b = blockIdx.z, c = blockIdx.y;
  float* dst = grad_points + ((size_t)b * C + c) * m;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const int* id = idx + ((size_t)b * n + p) * 3;
    const float* w = weight + ((size_t)b * n + p) * 3;
    const float g = grad_out[((size_t)b * C + c) * n + p];
    atomicAdd(dst + id[0], g * w[0]);
    atomicAdd(dst + id[1], g * w[1]);
    atomicAdd(dst + id[2], g * w[2]);
  }
}

inline unsigned blocks_for(long items, int threads) {
  long g = (items + threads - 1) / threads;
  if (g > 65535) g = 65535;
  if (g < 1) g = 1;
  return (unsigned)g;
}

template <int NT, int PTS>
int launch_fps(const float* xyz, int B, int N, int M, int tmask, int tshift, int* idx, cudaStream_t stream) {
  const size_t smem = (size_t)N * 3 * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(fps_kernel<NT, PTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) { hcm_set_error("pn2_fps: smem attribute: %s", cudaGetErrorString(e)); return HCM_ERR_CUDA; }
  fps_kernel<NT, PTS><<<B, NT, smem, stream>>>(xyz, N, M, tmask, tshift, idx);
  return HCM_OK;
}

}  // namespace

extern "C" {

// furthest_point_sampling_wrapper (pointnet2_api.cpp:19; pointnet2_utils.py:12-29): xyz [B,N,3] -> idx [B,M] int32, starting from
// point 0, distances initialised to 1e10 (the reference's `temp` tensor lives in registers here).  N <= 16384.
int hcm_pn2_furthest_point_sampling(const float* xyz, int B, int N, int M, int* idx, cudaStream_t stream) {
  HCM_CHECK_ARG(xyz && idx && B >= 1 && N >= 1 && M >= 1 && N <= 16384, "pn2_fps: bad args (N=%d; N <= 16384)", N);
  // thread count of the reference launch (largest power of two <= min(N, 1024), cuda_utils.h:10-14): fixes which of several
  // equidistant points wins (see fps_better)
  int T = 1, bits = 0;
  while (2 * T <= N && 2 * T <= 1024) { T *= 2; ++bits; }
  const int tmask = T - 1, tshift = 32 - bits;            // (bits == 0: T = 1, k & 0 = 0, the shifted value is never used)
  int rc;
  if (N <= 512) rc = launch_fps<512, 1>(xyz, B, N, M, tmask, tshift & 31, idx, stream);
  else if (N <= 1024) rc = launch_fps<512, 2>(xyz, B, N, M, tmask, tshift, idx, stream);
  else if (N <= 2048) rc = launch_fps<512, 4>(xyz, B, N, M, tmask, tshift, idx, stream);
  else if (N <= 4096) rc = launch_fps<512, 8>(xyz, B, N, M, tmask, tshift, idx, stream);
  else if (N <= 8192) rc = launch_fps<1024, 8>(xyz, B, N, M, tmask, tshift, idx, stream);
  else rc = launch_fps<1024, 16>(xyz, B, N, M, tmask, tshift, idx, stream);
  if (rc != HCM_OK) return rc;
  HCM_LAUNCH_CHECK("pn2_fps");
  return HCM_OK;
}

// ball_query_wrapper (pointnet2_api.cpp:11; pointnet2_utils.py:203-221): idx [B,M,nsample] = the first nsample points of xyz
// [B,N,3] (in index order) closer than radius to new_xyz [B,M,3], padded with the first; zeros when there is none
int hcm_pn2_ball_query(const float* new_xyz, const float* xyz, int B, int N, int M, float radius, int nsample, int* idx,
                       cudaStream_t stream) {
  HCM_CHECK_ARG(new_xyz && xyz && idx && B >= 1 && N >= 1 && M >= 1 && nsample >= 1, "pn2_ball_query: bad args");
  cudaError_t e = cudaMemsetAsync(idx, 0, (size_t)B * M * nsample * sizeof(int), stream);
  if (e != cudaSuccess) { hcm_set_error("pn2_ball_query: memset: %s", cudaGetErrorString(e)); return HCM_ERR_CUDA; }
  ball_query_kernel<<<dim3((M + 7) / 8, B), 256, 0, stream>>>(new_xyz, xyz, N, M, radius * radius, nsample, idx);
  HCM_LAUNCH_CHECK("pn2_ball_query");
  return HCM_OK;
}

// three_nn_wrapper (pointnet2_api.cpp:21; pointnet2_utils.py:79-98): squared distances and indices of the three points of known
// [B,m,3] nearest to each point of unknown [B,n,3] (the caller takes the square root, pointnet2_utils.py:99)
int hcm_pn2_three_nn(const float* unknown, const float* known, int B, int n, int m, float* dist2, int* idx, cudaStream_t stream) {
  HCM_CHECK_ARG(unknown && known && dist2 && idx && B >= 1 && n >= 1 && m >= 1, "pn2_three_nn: bad args");
  three_nn_kernel<<<dim3((n + 255) / 256, B), 256, 0, stream>>>(unknown, known, n, m, dist2, idx);
  HCM_LAUNCH_CHECK("pn2_three_nn");
  return HCM_OK;
}

// three_interpolate_wrapper / three_interpolate_grad_wrapper (pointnet2_api.cpp:22-23; pointnet2_utils.py:111-151):
// points [B,C,m], idx / weight [B,n,3] -> out [B,C,n];  grad_points [B,C,m] += (caller zeroes it)
int hcm_pn2_three_interpolate(const float* points, const int* idx, const float* weight, int B, int C, int m, int n, float* out,
                              cudaStream_t stream) {
  HCM_CHECK_ARG(points && idx && weight && out && B >= 1 && C >= 1 && C <= 65535 && m >= 1 && n >= 1, "pn2_three_interpolate: bad args");
  three_interpolate_kernel<<<dim3(blocks_for(n, 256), C, B), 256, 0, stream>>>(points, idx, weight, C, m, n, out);
  HCM_LAUNCH_CHECK("pn2_three_interpolate");
  return HCM_OK;
}
int hcm_pn2_three_interpolate_grad(const float* grad_out, const int* idx, const float* weight, int B, int C, int n, int m,
                                   float* grad_points, cudaStream_t stream) {
  HCM_CHECK_ARG(grad_out && idx && weight && grad_points && B >= 1 && C >= 1 && C <= 65535 && m >= 1 && n >= 1,
                "pn2_three_interpolate_grad: bad args");
  three_interpolate_grad_kernel<<<dim3(blocks_for(n, 256), C, B), 256, 0, stream>>>(grad_out, idx, weight, C, n, m, grad_points);
  HCM_LAUNCH_CHECK("pn2_three_interpolate_grad");
  return HCM_OK;
}

// group_points_wrapper / group_points_grad_wrapper (pointnet2_api.cpp:13-14; pointnet2_utils.py:159-194):
// points [B,C,N], idx [B,npoint,nsample] -> out [B,C,npoint,nsample];  grad_points [B,C,N] += (caller zeroes it)
int hcm_pn2_group_points(const float* points, const int* idx, int B, int C, int N, int npoint, int nsample, float* out,
                         cudaStream_t stream) {
  HCM_CHECK_ARG(points && idx && out && B >= 1 && C >= 1 && C <= 65535 && N >= 1 && npoint >= 1 && nsample >= 1, "pn2_group_points: bad args");
  const long E = (long)npoint * nsample;
  gather_kernel<<<dim3(blocks_for(E, 256), C, B), 256, 0, stream>>>(points, idx, C, N, E, out);
  HCM_LAUNCH_CHECK("pn2_group_points");
  return HCM_OK;
}
int hcm_pn2_group_points_grad(const float* grad_out, const int* idx, int B, int C, int N, int npoint, int nsample, float* grad_points,
                              cudaStream_t stream) {
  HCM_CHECK_ARG(grad_out && idx && grad_points && B >= 1 && C >= 1 && C <= 65535 && N >= 1 && npoint >= 1 && nsample >= 1,
                "pn2_group_points_grad: bad args");
  const long E = (long)npoint * nsample;
  gather_grad_kernel<<<dim3(blocks_for(E, 256), C, B), 256, 0, stream>>>(grad_out, idx, C, N, E, grad_points);
  HCM_LAUNCH_CHECK("pn2_group_points_grad");
  return HCM_OK;
}

// gather_points_wrapper / gather_points_grad_wrapper (pointnet2_api.cpp:16-17; pointnet2_utils.py:42-70):
// points [B,C,N], idx [B,npoint] -> out [B,C,npoint];  grad_points [B,C,N] += (caller zeroes it)
int hcm_pn2_gather_points(const float* points, const int* idx, int B, int C, int N, int npoint, float* out, cudaStream_t stream) {
  HCM_CHECK_ARG(points && idx && out && B >= 1 && C >= 1 && C <= 65535 && N >= 1 && npoint >= 1, "pn2_gather_points: bad args");
  gather_kernel<<<dim3(blocks_for(npoint, 256), C, B), 256, 0, stream>>>(points, idx, C, N, (long)npoint, out);
  HCM_LAUNCH_CHECK("pn2_gather_points");
  return HCM_OK;
}
int hcm_pn2_gather_points_grad(const float* grad_out, const int* idx, int B, int C, int N, int npoint, float* grad_points,
                               cudaStream_t stream) {
  HCM_CHECK_ARG(grad_out && idx && grad_points && B >= 1 && C >= 1 && C <= 65535 && N >= 1 && npoint >= 1, "pn2_gather_points_grad: bad args");
  gather_grad_kernel<<<dim3(blocks_for(npoint, 256), C, B), 256, 0, stream>>>(grad_out, idx, C, N, (long)npoint, grad_points);
  HCM_LAUNCH_CHECK("pn2_gather_points_grad");
  return HCM_OK;
}

}  // extern "C"
