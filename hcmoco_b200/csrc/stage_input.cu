// Input staging for real data: the GPU side of what the reference's CPU data pipeline does per
// NTU RGB-D sample after decoding — pycontrast/datasets/dataset.py:104-160 (NTURGBD.__getitem__: resized crop of the RGB frame
// (bilinear) and of the 16-bit depth frame (nearest), horizontal flip, /255 + ImageNet mean/std, depth mm -> m replicated x3) and
// :594-602 (NTUMPIIRGBD3D2DSkeletonGCN.__getitem__: depth_mask = depth > 0, depth <- depth - mean(depth over the mask), 0 outside).
// Fed with decoded frames (uint8 RGB, uint16 depth in millimetres, e.g. from pinned host buffers) and the crop parameters the
// sampler drew, it writes the step's input x [B,6,R,R] and depth_mask [B,R,R] directly: at B200 step rates (~90 ms for 64
// triplets) the 40-worker CPU loader of the reference is the next bottleneck, and moving 512x424 uint8/uint16 frames (0.65 + 0.43
// MB per sample) instead of 6 fp32 planes (1.57 MB at 256^2) also cuts the H2D bytes.
// Integer work (nearest-neighbour source indices, the mask, the millimetre sum and the pixel count behind the mean) is exact;
// the bilinear RGB resample is plain fp32 (align_corners = False, no antialiasing — PIL's antialiased fixed-point filter is
// NOT reproduced; samples without depth (MPII / COCO: affine warp with rotation, dataset.py:500-560) are out of scope and are
// passed with has_depth = 0: zero depth planes and mask, RGB through the same crop path).
#include "common.cuh"

namespace {

struct StageParams {
  const uint8_t* rgb;            // [B][Hs][Ws][3]
  const uint16_t* depth;         // [B][Hs][Ws] millimetres
  const int* crop;               // [B][4] = top i, left j, height h, width w of the crop window in source pixels (may leave the frame: zeros)
  const int* flip;               // [B] horizontal flip
  const long long* has_depth;    // [B] (true_depth); null = all ones
  unsigned long long* sums;      // [B][2] = sum of the cropped depth in mm over the mask, pixel count of the mask
  float* x;                      // [B][6][R][R]
  float* mask;                   // [B][R][R]
  int B, Hs, Ws, R;
};

// destination pixel -> nearest source pixel of the crop window (PIL NEAREST: centre of the destination pixel), -1 if outside the frame
__device__ __forceinline__ int nearest_src(int y, int x, int i, int j, int h, int w, int R, int Hs, int Ws) {
  const int sy = i + min(h - 1, (int)(((long long)(2 * y + 1) * h) / (2 * R)));
  const int sx = j + min(w - 1, (int)(((long long)(2 * x + 1) * w) / (2 * R)));
  if (sy < 0 || sy >= Hs || sx < 0 || sx >= Ws) return -1;
  return sy * Ws + sx;
}

__global__ void stage_depth_stats_kernel(const StageParams p) {
  const int b = blockIdx.y;
  const int* c = p.crop + 4 * b;
  const int i = c[0], j = c[1], h = c[2], w = c[3];
  const uint16_t* d = p.depth + (size_t)b * p.Hs * p.Ws;
  unsigned long long s = 0, n = 0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < p.R * p.R; e += gridDim.x * blockDim.x) {
    const int src = nearest_src(e / p.R, e % p.R, i, j, h, w, p.R, p.Hs, p.Ws);      // (the flip permutes pixels: sums unchanged)
    const unsigned v = src >= 0 ? d[src] : 0u;
    s += v;
    n += v > 0 ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); n += __shfl_xor_sync(0xffffffffu, n, o); }
  if ((threadIdx.x & 31) == 0 && (s | n)) { atomicAdd(p.sums + 2 * b, s); atomicAdd(p.sums + 2 * b + 1, n); }
}

__global__ void stage_input_kernel(const StageParams p) {
  const int b = blockIdx.y;
  const int* c = p.crop + 4 * b;
  const int i = c[0], j = c[1], h = c[2], w = c[3];
  const bool flip = p.flip && p.flip[b] != 0;
  const bool hd = !p.has_depth || p.has_depth[b] != 0;
  const uint8_t* rgb = p.rgb + (size_t)b * p.Hs * p.Ws * 3;
  const uint16_t* dep = p.depth + (size_t)b * p.Hs * p.Ws;
  const unsigned long long cnt = p.sums[2 * b + 1];
  const float mean = cnt ? (float)((double)p.sums[2 * b] / (double)cnt / 1000.0) : 0.f;
  const long RR = (long)p.R * p.R;
  float* xb = p.x + (size_t)b * 6 * RR;
  const float sh = (float)h / (float)p.R, sw = (float)w / (float)p.R;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < RR; e += gridDim.x * blockDim.x) {
    const int y = e / p.R, xo = e - y * p.R;
    const int x = flip ? p.R - 1 - xo : xo;                     // source column of this output pixel before the flip
    // ---- RGB: bilinear (align_corners = False), taps clamped to the crop window, zero outside the frame
    const float fy = fminf(fmaxf(((float)y + 0.5f) * sh - 0.5f, 0.f), (float)(h - 1));
    const float fx = fminf(fmaxf(((float)x + 0.5f) * sw - 0.5f, 0.f), (float)(w - 1));
    const int y0 = (int)fy, x0 = (int)fx, y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float wy = fy - (float)y0, wx = fx - (float)x0;
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int sy = i + ((t & 2) ? y1 : y0), sx = j + ((t & 1) ? x1 : x0);
      const float wt = ((t & 2) ? wy : 1.f - wy) * ((t & 1) ? wx : 1.f - wx);
      if (sy >= 0 && sy < p.Hs && sx >= 0 && sx < p.Ws) {
        const uint8_t* px = rgb + ((size_t)sy * p.Ws + sx) * 3;
        acc[0] = fmaf(wt, (float)px[0], acc[0]);
        acc[1] = fmaf(wt, (float)px[1], acc[1]);
        acc[2] = fmaf(wt, (float)px[2], acc[2]);
      }
    }
    xb[e] = (acc[0] / 255.f - 0.485f) / 0.229f;
    xb[RR + e] = (acc[1] / 255.f - 0.456f) / 0.224f;
    xb[2 * RR + e] = (acc[2] / 255.f - 0.406f) / 0.225f;
    // ---- depth: nearest, mm -> m, mask, mean-centred inside the mask
    const int src = nearest_src(y, x, i, j, h, w, p.R, p.Hs, p.Ws);
    const unsigned mm = (hd && src >= 0) ? dep[src] : 0u;
    const float d = mm ? (float)mm / 1000.f - mean : 0.f;
    xb[3 * RR + e] = d;
    xb[4 * RR + e] = d;
    xb[5 * RR + e] = d;
    p.mask[(size_t)b * RR + e] = mm ? 1.f : 0.f;
  }
}

}  // namespace

extern "C" {

// sums [B][2] (uint64: millimetre sum over the mask, pixel count) is zeroed here, filled by the first kernel and read by the second
int hcm_stage_input(const unsigned char* rgb, const unsigned short* depth, const int* crop, const int* flip, const long long* has_depth,
                    int B, int Hs, int Ws, int R, unsigned long long* sums, float* x, float* depth_mask, cudaStream_t stream) {
  HCM_CHECK_ARG(rgb && depth && crop && sums && x && depth_mask, "stage_input: null pointer");
  HCM_CHECK_ARG(B >= 1 && Hs >= 1 && Ws >= 1 && R >= 1 && (long)Hs * Ws < (1L << 30), "stage_input: bad sizes");
  StageParams p;
  p.rgb = rgb; p.depth = depth; p.crop = crop; p.flip = flip; p.has_depth = has_depth; p.sums = sums; p.x = x; p.mask = depth_mask;
  p.B = B; p.Hs = Hs; p.Ws = Ws; p.R = R;
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)B * 2 * sizeof(unsigned long long), stream);
  if (e != cudaSuccess) { hcm_set_error("stage_input: memset: %s", cudaGetErrorString(e)); return HCM_ERR_CUDA; }
  const int nb = hcm_cdiv((long)R * R, 256 * 4);
  stage_depth_stats_kernel<<<dim3(nb, B), 256, 0, stream>>>(p);
  HCM_LAUNCH_CHECK("stage_input (depth statistics)");
  stage_input_kernel<<<dim3(nb, B), 256, 0, stream>>>(p);
  HCM_LAUNCH_CHECK("stage_input");
  return HCM_OK;
}

}  // extern "C"
