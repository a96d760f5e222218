// SemGCN 2D-keypoint encoder pieces (networks/SGCN/sem_graph_conv.py:34-48) and the fused SGD step
// (torch.optim.SGD as configured at main_contrast.py:78-81).
//
// A SemGraphConv is out = diag(A) * (x W0) + offdiag(A) * (x W1) + b with A = row-softmax of the
// masked learnable edge logits e.  Aggregating *before* the channel contraction turns it into one
// GEMM with K = 2*Cin:  out = [A_jj x_j , sum_{k!=j} A_jk x_k] * [W0; W1] + b, and W's [2,Cin,Cout]
// storage already is that stacked matrix.  The kernels here do the (tiny) graph side; the GEMM is
// hcm_gemm.
#include "common.cuh"

namespace {

// A [J][J] = softmax over the row's edges; rows/cols [nnz] list the mask's non-zeros in row-major order
__global__ void sgcn_adj_kernel(const float* __restrict__ e, const int* __restrict__ rows, const int* __restrict__ cols,
                                int nnz, int J, float* A) {
  for (int i = threadIdx.x; i < J * J; i += blockDim.x) A[i] = 0.f;
  __syncthreads();
  const int r = threadIdx.x;
  if (r < J) {
    float mx = -INFINITY;
    for (int i = 0; i < nnz; ++i) if (rows[i] == r) mx = fmaxf(mx, e[i]);
    float s = 0.f;
    for (int i = 0; i < nnz; ++i) if (rows[i] == r) s += __expf(e[i] - mx);
    for (int i = 0; i < nnz; ++i) if (rows[i] == r) A[r * J + cols[i]] = __expf(e[i] - mx) / s;
  }
}

// de[i] (+)= A[r,c] * (dA[r,c] - sum_c' A[r,c'] dA[r,c'])
__global__ void sgcn_adj_bwd_kernel(const float* __restrict__ A, const float* __restrict__ dA, const int* __restrict__ rows,
                                    const int* __restrict__ cols, int nnz, int J, float* de, int accumulate) {
  const int i = threadIdx.x;
  if (i >= nnz) return;
  const int r = rows[i];
  float dot = 0.f;
  for (int c = 0; c < J; ++c) dot = fmaf(A[r * J + c], dA[r * J + c], dot);
  const float v = A[r * J + cols[i]] * (dA[r * J + cols[i]] - dot);
  de[i] = accumulate ? de[i] + v : v;
}

// xa [B][J][2*Cin]: first half A_jj * x_j, second half sum_{k != j} A_jk * x_k
__global__ void sgcn_aggregate_kernel(const float* __restrict__ x, const float* __restrict__ A, int B, int J, int Cin,
                                      float* __restrict__ xa) {
  const long total = (long)B * J * Cin;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cin), j = (int)((e / Cin) % J);
    const long b = e / ((long)Cin * J);
    const float* xb = x + b * J * Cin + c;
    float off = 0.f;
    for (int k = 0; k < J; ++k) if (k != j) off = fmaf(A[j * J + k], xb[(long)k * Cin], off);
    float* o = xa + (b * J + j) * 2 * Cin;
    o[c] = A[j * J + j] * xb[(long)j * Cin];
    o[Cin + c] = off;
  }
}

// dx[b][k][c] (+)= A_kk * dxa[b][k][c] + sum_{j != k} A_jk * dxa[b][j][Cin + c]
__global__ void sgcn_aggregate_bwd_kernel(const float* __restrict__ dxa, const float* __restrict__ A, int B, int J, int Cin,
                                          float* dx, int accumulate) {
  const long total = (long)B * J * Cin;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cin), k = (int)((e / Cin) % J);
    const long b = e / ((long)Cin * J);
    const float* gb = dxa + b * J * 2 * Cin;
    float v = A[k * J + k] * gb[(long)k * 2 * Cin + c];
    for (int j = 0; j < J; ++j) if (j != k) v = fmaf(A[j * J + k], gb[(long)j * 2 * Cin + Cin + c], v);
    dx[e] = accumulate ? dx[e] + v : v;
  }
}

// dA[j][k] = sum_{b,c} dxa[b][j][(k==j ? c : Cin + c)] * x[b][k][c]; one CTA per (j,k)
__global__ void sgcn_dadj_kernel(const float* __restrict__ dxa, const float* __restrict__ x, int B, int J, int Cin, float* dA) {
  __shared__ float red[32];
  const int j = blockIdx.x / J, k = blockIdx.x % J;
  const int off = (j == k) ? 0 : Cin;
  float a = 0.f;
  for (long e = threadIdx.x; e < (long)B * Cin; e += blockDim.x) {
    const long b = e / Cin;
    const int c = (int)(e % Cin);
    a = fmaf(dxa[(b * J + j) * 2 * Cin + off + c], x[(b * J + k) * Cin + c], a);
  }
  a = block_sum(a, red);
  if (threadIdx.x == 0) dA[j * J + k] = a;
}

// mean over joints: out[b][c] = mean_j x[b][j][c]   (build_backbone.py:279) and its backward
__global__ void joint_mean_kernel(const float* __restrict__ x, int B, int J, int C, float* out) {
  const long total = (long)B * C;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long b = e / C;
    const int c = (int)(e % C);
    float s = 0.f;
    for (int j = 0; j < J; ++j) s += x[(b * J + j) * C + c];
    out[e] = s / (float)J;
  }
}
__global__ void joint_mean_bwd_kernel(const float* __restrict__ dout, int B, int J, int C, float* dx, int accumulate) {
  const long total = (long)B * J * C;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const long b = e / ((long)J * C);
    const float v = dout[b * C + c] / (float)J;
    dx[e] = accumulate ? dx[e] + v : v;
  }
}

// p -= lr * buf, buf = momentum*buf + (g + wd*p)   (first step: buf = g + wd*p)
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long n, float lr,
                           float momentum, float wd, int first, float gscale) {
  const long stride = (long)gridDim.x * blockDim.x;
  const long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* b4 = reinterpret_cast<float4*>(buf);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = p4[i], gv = g4[i], bv = b4[i];
    gv.x = fmaf(wd, pv.x, gv.x * gscale); gv.y = fmaf(wd, pv.y, gv.y * gscale);
    gv.z = fmaf(wd, pv.z, gv.z * gscale); gv.w = fmaf(wd, pv.w, gv.w * gscale);
    if (first) bv = gv;
    else { bv.x = fmaf(momentum, bv.x, gv.x); bv.y = fmaf(momentum, bv.y, gv.y); bv.z = fmaf(momentum, bv.z, gv.z); bv.w = fmaf(momentum, bv.w, gv.w); }
    pv.x -= lr * bv.x; pv.y -= lr * bv.y; pv.z -= lr * bv.z; pv.w -= lr * bv.w;
    p4[i] = pv; b4[i] = bv;
  }
  for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gv = fmaf(wd, p[i], g[i] * gscale);
    const float bv = first ? gv : fmaf(momentum, buf[i], gv);
    p[i] -= lr * bv;
    buf[i] = bv;
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

int hcm_sgcn_adj(const float* e, const int* rows, const int* cols, int nnz, int J, float* A, cudaStream_t stream) {
  HCM_CHECK_ARG(e && rows && cols && A && J <= 32, "sgcn_adj: bad args (J=%d)", J);
  sgcn_adj_kernel<<<1, 64, 0, stream>>>(e, rows, cols, nnz, J, A);
  HCM_LAUNCH_CHECK("sgcn_adj");
  return HCM_OK;
}

int hcm_sgcn_adj_bwd(const float* A, const float* dA, const int* rows, const int* cols, int nnz, int J, float* de,
                     int accumulate, cudaStream_t stream) {
  HCM_CHECK_ARG(A && dA && rows && cols && de && nnz <= 256, "sgcn_adj_bwd: bad args (nnz=%d)", nnz);
  sgcn_adj_bwd_kernel<<<1, 256, 0, stream>>>(A, dA, rows, cols, nnz, J, de, accumulate);
  HCM_LAUNCH_CHECK("sgcn_adj_bwd");
  return HCM_OK;
}

int hcm_sgcn_aggregate(const float* x, const float* A, int B, int J, int Cin, float* xa, cudaStream_t stream) {
  HCM_CHECK_ARG(x && A && xa, "sgcn_aggregate: null pointer");
  sgcn_aggregate_kernel<<<ew_grid((long)B * J * Cin * 4), 256, 0, stream>>>(x, A, B, J, Cin, xa);
  HCM_LAUNCH_CHECK("sgcn_aggregate");
  return HCM_OK;
}

// dx (+)= adjoint of the aggregation; dA [J][J] = gradient wrt the adjacency
int hcm_sgcn_aggregate_bwd(const float* dxa, const float* x, const float* A, int B, int J, int Cin, float* dx,
                           int accumulate, float* dA, cudaStream_t stream) {
  HCM_CHECK_ARG(dxa && A, "sgcn_aggregate_bwd: null pointer");
  if (dx) {
    sgcn_aggregate_bwd_kernel<<<ew_grid((long)B * J * Cin * 4), 256, 0, stream>>>(dxa, A, B, J, Cin, dx, accumulate);
    HCM_LAUNCH_CHECK("sgcn_aggregate_bwd");
  }
  if (dA) {
    HCM_CHECK_ARG(x != nullptr, "sgcn_aggregate_bwd: x needed for dA");
    sgcn_dadj_kernel<<<J * J, 256, 0, stream>>>(dxa, x, B, J, Cin, dA);
    HCM_LAUNCH_CHECK("sgcn_dadj");
  }
  return HCM_OK;
}

int hcm_joint_mean(const float* x, int B, int J, int C, float* out, cudaStream_t stream) {
  HCM_CHECK_ARG(x && out, "joint_mean: null pointer");
  joint_mean_kernel<<<ew_grid((long)B * C * 4), 256, 0, stream>>>(x, B, J, C, out);
  HCM_LAUNCH_CHECK("joint_mean");
  return HCM_OK;
}

int hcm_joint_mean_bwd(const float* dout, int B, int J, int C, float* dx, int accumulate, cudaStream_t stream) {
  HCM_CHECK_ARG(dout && dx, "joint_mean_bwd: null pointer");
  joint_mean_bwd_kernel<<<ew_grid((long)B * J * C * 4), 256, 0, stream>>>(dout, B, J, C, dx, accumulate);
  HCM_LAUNCH_CHECK("joint_mean_bwd");
  return HCM_OK;
}

// fused SGD over a flat parameter / gradient / momentum buffer (16-byte aligned); gscale scales g first
// (1/world_size after a sum all-reduce)
int hcm_sgd_step(float* p, const float* g, float* buf, long n, float lr, float momentum, float wd, int first,
                 float gscale, cudaStream_t stream) {
  HCM_CHECK_ARG(p && g && buf, "sgd_step: null pointer");
  HCM_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)buf) & 15) == 0, "sgd_step: buffers must be 16-byte aligned");
  sgd_kernel<<<ew_grid(n / 4 + 1), 256, 0, stream>>>(p, g, buf, n, lr, momentum, wd, first, gscale);
  HCM_LAUNCH_CHECK("sgd_step");
  return HCM_OK;
}

int hcm_zero(void* p, long bytes, cudaStream_t stream) {
  HCM_CHECK_ARG(p != nullptr || bytes == 0, "zero: null pointer");
  if (bytes == 0) return HCM_OK;
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, stream);
  if (e != cudaSuccess) { hcm_set_error("zero: %s", cudaGetErrorString(e)); return HCM_ERR_CUDA; }
  return HCM_OK;
}

}  // extern "C"
