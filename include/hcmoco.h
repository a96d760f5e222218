/* hcmoco.h — C-ABI of libhcmoco_sm100.so: the sm_100a kernels behind the HCMoCo pre-train step.
 *
 * The reference has no FFI on this path (it is PyTorch eager end to end); this ABI is the seam one
 * level below the Python surface main_contrast.py uses.  Each entry point names the
 * reference call site it replaces (paths relative to pycontrast/ in hongfz16/HCMoCo).
 *
 * Conventions
 *   - all tensors are device pointers owned by the caller; the library never allocates, frees or
 *     keeps a pointer after the call returns; scratch buffers are passed in (sizes documented);
 *   - activations are channels-last fp32: [B, H, W, C] (or [rows, C]); conv weights keep the
 *     reference / checkpoint layout OIHW; indices are int64 as torch produces them;
 *   - every call only enqueues work on `stream` (no sync, no default-stream use) => composes with
 *     CUDA graphs and side streams;
 *   - return 0 on success, negative on error (HCM_ERR_*); hcm_last_error() gives a thread-local text.
 */
#ifndef HCMOCO_H_
#define HCMOCO_H_

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

int hcm_abi_version(void);
const char* hcm_last_error(void);

/* ---- convolutions / GEMM (igemm.cu) : nn.Conv2d in networks/official_hrnet/official_hrnet.py:26-29,
 *      68-75, 187-216, 336-357 ; nn.Linear / 1x1 projection in networks/build_backbone.py:226-245 ---- */
/* y[B,Ho,Wo,Cout] = conv(T(x[B,H,W,Cin]), w[Cout,Cin,ks,ks]) (+bias); pad=(ks-1)/2; ks in {1,3}; stride in {1,2}.
 * T = optional per-channel affine (+ReLU) applied on load (= the previous layer's BatchNorm).
 * stat_part (optional) [hcm_conv2d_stat_rows()][2][Cout]: per-CTA column sums of y and y^2 for train-mode BN. */
int hcm_conv2d_stat_rows(int B, int H, int W, int Cin, int Cout, int ks, int stride);
int hcm_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin,
                   int Cout, int ks, int stride, const float* in_scale, const float* in_shift, int in_relu,
                   float* stat_part, cudaStream_t stream);
/* dx[B,H,W,Cin] (+)= conv_transpose(dy[B,Ho,Wo,Cout], w) */
int hcm_conv2d_dgrad(const float* dy, const float* w, float* dx, int B, int H, int W, int Cin, int Cout, int ks,
                     int stride, int accumulate, cudaStream_t stream);
/* dw[Cout,Cin,ks,ks] += sum_pixels dy * T(x)   (fp32 reductions; caller zeroes dw once per step) */
int hcm_conv2d_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Cin, int Cout, int ks,
                     int stride, const float* in_scale, const float* in_shift, int in_relu, cudaStream_t stream);
/* C[b][m][n] = alpha * sum_k A[b*bsA + m*sAm + k*sAk] * B[b*bsB + k*sBk + n*sBn] (+bias[n]) (+C) ; C row stride sCm */
int hcm_gemm(const float* A, const float* Bm, const float* bias, float* C, int batch, int M, int N, int K, long sAm,
             long sAk, long sBk, long sBn, long sCm, long bsA, long bsB, long bsC, float alpha, int accumulate,
             cudaStream_t stream);

/* ---- tensor-core convolution (tc_conv.cu): tcgen05.mma (bf16 hi/lo split, fp32 accumulate in TMEM), weights
 *      streamed by cp.async.bulk (TMA) — the stride-1 3x3 / 1x1 nn.Conv2d layers of official_hrnet.py:26-29,
 *      68-75, 187-216 and the 1x1 projection build_backbone.py:243-245; forward and data gradient ---- */
int hcm_tc_conv_supported(int B, int H, int W, int Cin, int Cout, int ks, int stride);
long hcm_tc_conv_wpack_bytes(int B, int H, int W, int Cin, int Cout, int ks);
/* 1 if hcm_tc_conv runs the convolution with row-concatenated taps (3x3, stride 1, 3*ceil16(Cout) <= 256: the three taps of a
 * filter row are the N blocks of ONE tcgen05.mma, the column shift is resolved in the epilogue); its weights must then be packed
 * with flag 4.  For the data gradient pass the GEMM's Cout (= Cin of the weight) */
int hcm_tc_conv_rowcat_supported(int Cout, int ks, int stride);
/* flags bit 0 clear: pack w[Cout][Cin][ks][ks] for the forward conv; set: pack the same tensor for its data gradient seen as
 * a conv with Cin' = Cout(w), Cout' = Cin(w) (pass those as Cin, Cout).  flags bit 2 (4): row-concatenated layout, required iff
 * hcm_tc_conv_rowcat_supported(Cout, ks, stride of the consuming conv).  ldw > 0: `w` is a column block
 * of a wider [O][ldw][ks][ks] tensor (per-branch blocks of the 1x1 projection); lddw likewise for hcm_tc_wgrad.
 * 0 < ldw < Cin (forward packs) / 0 < lddw < Cin (hcm_tc_wgrad): the weight tensor has only ldw input channels and the input is
 * stored with its channels zero-padded to Cin (the 3-channel stem on 4-channel rows, hcm_nchw_to_nhwc_pad): the GEMM's K is
 * zero-padded, the gradient of the padding channels is dropped */
int hcm_tc_conv_pack(const float* w, int ldw, void* wpack, int B, int H, int W, int Cin, int Cout, int ks, int flags,
                     cudaStream_t stream);
/* all weight packs of a step in one launch: jobs (device) = njobs x 8 int64 {w ptr, out ptr, Cin, Cout, ks, mode, ldw, first_step};
 * mode 0/1 = hcm_tc_conv_pack(transpose 0/1), 2 = hcm_tc_dgrad_s2_pack, (Cin, Cout) as passed to those calls */
int hcm_tc_pack_batch(const long long* jobs, int njobs, int total_steps, cudaStream_t stream);
/* y[B,Ho,Wo,Cout] (+)= conv(T(x[B,H,W,Cin])) (+bias), 3x3 stride 1|2 or 1x1, pad (ks-1)/2.  accumulate = 1 adds onto y with one
 * vector reduction per element at the L2 (red.global.add.v4.f32: each element has exactly one writer per launch, the result is the
 * same single fp32 addition; subnormal sums flush to zero) */
int hcm_tc_conv(const float* x, const void* wpack, const float* bias, float* y, int B, int H, int W, int Cin, int Cout,
                int ks, int stride, const float* in_scale, const float* in_shift, int in_relu, int accumulate,
                cudaStream_t stream);

/* data gradient of a 3x3 / stride-2 / pad-1 convolution w[Cout][Cin][3][3] on tensor cores (one 2x2-tap GEMM over dy whose
 * N axis is the four output-pixel parities): dx[B,H,W,Cin] (+)= conv_transpose(dy[B,H/2,W/2,Cout], w) */
int hcm_tc_dgrad_s2_supported(int B, int H, int W, int Cin, int Cout);
long hcm_tc_dgrad_s2_wpack_bytes(int B, int H, int W, int Cin, int Cout);
int hcm_tc_dgrad_s2_nqs(int B, int H, int W, int Cin, int Cout);   /* output parities per launch: 4, 2 or 1 */
int hcm_tc_dgrad_s2_pack(const float* w, void* wpack, int B, int H, int W, int Cin, int Cout, cudaStream_t stream);
int hcm_tc_dgrad_s2(const float* dy, const void* wpack, float* dx, int B, int H, int W, int Cin, int Cout, int accumulate,
                    cudaStream_t stream);

/* tensor-core weight gradient (tc_wgrad2.cu): dw[Cout,Cin,ks,ks] += sum_pixels dy * T(x); 3x3 stride 1|2, 1x1 */
int hcm_tc_wgrad_supported(int B, int H, int W, int Cin, int Cout, int ks, int stride);
int hcm_tc_wgrad(const float* x, const float* dy, float* dw, int lddw, int B, int H, int W, int Cin, int Cout, int ks,
                 int stride, const float* in_scale, const float* in_shift, int in_relu, cudaStream_t stream);

/* ---- train-mode batch norm (bn.cu) : nn.BatchNorm2d(momentum=0.01) official_hrnet.py:22-23 (+ReLU /
 *      residual add :44-60, :86-101) and nn.BatchNorm1d networks/SGCN/sem_gcn.py:13 ---- */
int hcm_colstat_rows(long P, int C);
int hcm_bn_stats(const float* y, long P, int C, float* part, cudaStream_t stream);
int hcm_bn_finalize(const float* part, int nparts, int C, long count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                    float* scale, float* shift, float* mean, float* invstd, cudaStream_t stream);
/* out = act(y*scale + shift + (res*res_scale + res_shift)) */
/* hcm_bn_stats + hcm_bn_finalize (count = P) in ONE launch: the CTA taking the last ticket reduces the partial rows (fp64,
   fixed order).  `counter`: one zero-initialised uint32 owned by the caller (one per stream); the kernel leaves it at zero. */
int hcm_bn_stats_finalize(const float* y, long P, int C, float* part, unsigned int* counter, const float* gamma,
                          const float* beta, float* running_mean, float* running_var, long long* num_batches_tracked,
                          float momentum, float eps, float* scale, float* shift, float* mean, float* invstd,
                          cudaStream_t stream);
/* hcm_bn_bwd_reduce + hcm_bn_bwd_finalize (count = P) in ONE launch; `counter` as above */
int hcm_bn_bwd_reduce_finalize(const float* dz, const float* mask, const float* mask_scale, const float* mask_shift,
                               const float* y, const float* mean, const float* invstd, long P, int C, float* part,
                               unsigned int* counter, const float* gamma, float* dgamma, float* dbeta, float* k1, float* k2,
                               float* k3, cudaStream_t stream);
int hcm_bn_apply(const float* y, const float* scale, const float* shift, const float* res, const float* res_scale,
                 const float* res_shift, int relu, float* out, long P, int C, cudaStream_t stream);
/* g = dz * [m > 0] with m = mask (if given) else y*mask_scale+mask_shift (if given) else 1 */
int hcm_bn_bwd_reduce(const float* dz, const float* mask, const float* mask_scale, const float* mask_shift,
                      const float* y, const float* mean, const float* invstd, long P, int C, float* part,
                      cudaStream_t stream);
int hcm_bn_bwd_finalize(const float* part, int nparts, int C, long count, const float* gamma, const float* mean,
                        const float* invstd, float* dgamma, float* dbeta, float* k1, float* k2, float* k3,
                        cudaStream_t stream);
/* dy = k1*g + k2*y + k3 (dy may alias dz); g_out (+)= g (gradient of the residual branch) */
int hcm_bn_bwd_apply(const float* dz, const float* mask, const float* mask_scale, const float* mask_shift,
                     const float* y, const float* k1, const float* k2, const float* k3, float* dy, float* g_out,
                     int g_accumulate, long P, int C, cudaStream_t stream);
int hcm_relu_bwd(const float* dout, const float* out, float* g, int accumulate, long total, cudaStream_t stream);
int hcm_axpy(float* dst, const float* src, float alpha, long total, cudaStream_t stream);

/* ---- layout / resampling (resample.cu) : torch.split build_backbone.py:261; HR-module fuse
 *      official_hrnet.py:232-247; merge_all_res build_backbone.py:247-254; avg-pool :267-278 ---- */
int hcm_nchw_to_nhwc(const float* x, float* out, int B, int Ctot, long HW, int coff, int Cn, cudaStream_t stream);
int hcm_nchw_to_nhwc_pad(const float* x, float* out, int B, int Ctot, long HW, int coff, int Cn, int Cpad,
                         cudaStream_t stream);   /* out [B,HW,Cpad], channels Cn..Cpad-1 zero */
int hcm_fuse_sum(int nterms, const float* const* ptrs, const float* const* scales, const float* const* shifts,
                 const int* log2f, const float* bias, int relu, float* out, int B, int H, int W, int C,
                 cudaStream_t stream);
int hcm_upsample_adjoint(const float* g, float* out, int accumulate, int B, int H, int W, int C, int log2f,
                         cudaStream_t stream);
int hcm_avgpool(const float* x, float* out, int B, long HW, int C, int ldo, int coff, cudaStream_t stream);
int hcm_avgpool_bwd(const float* dout, float* dx, int accumulate, int B, long HW, int C, int ldo, int coff,
                    cudaStream_t stream);

/* ---- memory-bank NCE (nce.cu) : CMCMem3.forward memory/mem_bank.py:172-205, _update_memory :15-28,
 *      _compute_loss_accuracy learning/contrast_trainer.py:212-253 ---- */
int hcm_nce_logits(const float* bank1, const float* bank2, const float* bank3, const float* x1, const float* x2,
                   const float* x3, long ldx, const long long* idx, int B, int K1, int dim, float T, float* logits,
                   cudaStream_t stream);
int hcm_nce_loss(const float* logits, int B, int K1, const long long* use_depth, const long long* use_rgb, float* lse,
                 float* l0, float* hit, float* coef, float* loss6, float* acc6, cudaStream_t stream);
/* lse == coef == NULL: `logits` holds d(loss)/d(logits) (backward of the logits-returning CMCMem3.forward) */
int hcm_nce_bwd(const float* bank1, const float* bank2, const float* bank3, const float* x1, const float* x2,
                const float* x3, long ldx, const long long* idx, int B, int K1, int dim, float T, const float* logits,
                const float* lse, const float* coef, float gscale, float* df, long lddf, cudaStream_t stream);
int hcm_bank_update(float* bank, const float* x, long ldx, const long long* y, int N, int dim, float m,
                    cudaStream_t stream);

/* ---- dense / sparse / SCL objectives (losses.cu) : contrast_trainer.py:642-723, 744-828, 830-892 ---- */
int hcm_gather_l2norm(const float* src, long lds, const long long* pix, long HW, int rows_per_b, long nrows, int dim,
                      float* out, long ldo, float* inv_norm, cudaStream_t stream);
int hcm_gather_l2norm_bwd(const float* dout, long lddo, const float* out, long ldo, const float* inv_norm,
                          const long long* pix, long HW, int rows_per_b, long nrows, int dim, float* dsrc, long lds,
                          int accumulate, cudaStream_t stream);
int hcm_joint_pixel_index(const float* joints_yx, long n, int h, long long* pix, cudaStream_t stream);
int hcm_dense_kept(const float* depth_mask, int B, int R, int h, float* kept, cudaStream_t stream);
int hcm_dense_stats(const float* L, const long long* pix, const float* kept, const long long* use_depth, int B, int S,
                    int h, float* stat, float* fin, cudaStream_t stream);
int hcm_dense_grad(float* L, const long long* pix, const float* stat, const float* kept, const float* fin, int B, int S,
                   int h, float gscale, cudaStream_t stream);
int hcm_dense_finish(const float* stat, const float* kept, const long long* use_depth, int B, int S, float* fin,
                     cudaStream_t stream);
/* Fused form of the dense objective (dense_affinity.cu; replaces contrast_trainer.py:684-723 in one kernel per
   direction of the chain rule): gather S sampled pixels of both projection maps G1, G2 [B][h*h][dim=128] (channels-last),
   L2-normalise (:692-693), S x S x 128 affinity on tcgen05 tensor cores (bf16 hi/lo split, fp32 accumulation in TMEM;
   :695-699), soft-target log-softmax statistics (:702-721) in the epilogue.  Nothing S x S is written to memory.
   stat [B][2][S][4] is scratch kept for the backward; fin[5] = loss_r2d, loss_d2r, acc_r2d, acc_d2r, B'.
   The backward recomputes the affinity and ACCUMULATES into dG1, dG2 (atomics; the caller zeroes them);
   gscale_r2d / gscale_d2r = d(total)/d(loss_r2d), d(total)/d(loss_d2r) (1, 1 for the reference's plain sum, :980).
   A first kernel gathers, L2-normalises and splits the S sampled pixels of both maps ONCE per sample into bf16 hi/lo operand
   slabs in `work`; the main kernel streams them with cp.async.bulk (TMA).  prepared = 1: `work` still holds the forward's slabs. */
long hcm_dense_affinity_work_bytes(int B, int S);   /* operand-slab workspace (128-byte aligned, caller-owned) */
int hcm_dense_affinity_fwd(const float* G1, const float* G2, const long long* pix, const float* kept,
                           const long long* use_depth, int B, int S, int h, int dim, float inv_T, float* stat, float* fin,
                           void* work, cudaStream_t stream);
int hcm_dense_affinity_bwd(const float* G1, const float* G2, const long long* pix, const float* stat, const float* kept,
                           const float* fin, int B, int S, int h, int dim, float inv_T, float gscale_r2d, float gscale_d2r,
                           float* dG1, float* dG2, void* work, int prepared, cudaStream_t stream);
int hcm_joint_stats(const float* Lr, const float* Ld, const int* joints_vis, const long long* use_depth, int B, int J,
                    float* rs, float* lse, float* fin, cudaStream_t stream);
int hcm_joint_grad(float* Lr, float* Ld, const int* joints_vis, const long long* use_depth, const float* lse,
                   const float* fin, int B, int J, float gscale, cudaStream_t stream);
int hcm_scl_stats(const float* Z, int B, int J, const long long* use_rgb, const long long* use_depth, float* rowstat,
                  float* fin, cudaStream_t stream);
int hcm_scl_grad(float* Z, int B, int J, const long long* use_rgb, const long long* use_depth, const float* rowstat,
                 const float* fin, float gscale, cudaStream_t stream);
int hcm_colsum_finalize(const float* part, int nparts, int C, float* out, int accumulate, cudaStream_t stream);
int hcm_colsum_small(const float* x, int R, int C, long ld, float* out, int accumulate, cudaStream_t stream);

/* ---- SemGCN graph side + optimiser (sgcn.cu) : networks/SGCN/sem_graph_conv.py:34-48;
 *      torch.optim.SGD main_contrast.py:78-81 ---- */
int hcm_sgcn_adj(const float* e, const int* rows, const int* cols, int nnz, int J, float* A, cudaStream_t stream);
int hcm_sgcn_adj_bwd(const float* A, const float* dA, const int* rows, const int* cols, int nnz, int J, float* de,
                     int accumulate, cudaStream_t stream);
int hcm_sgcn_aggregate(const float* x, const float* A, int B, int J, int Cin, float* xa, cudaStream_t stream);
int hcm_sgcn_aggregate_bwd(const float* dxa, const float* x, const float* A, int B, int J, int Cin, float* dx,
                           int accumulate, float* dA, cudaStream_t stream);
int hcm_joint_mean(const float* x, int B, int J, int C, float* out, cudaStream_t stream);
int hcm_joint_mean_bwd(const float* dout, int B, int J, int C, float* dx, int accumulate, cudaStream_t stream);
/* ---- input staging for real data (stage_input.cu): datasets/dataset.py:104-160 (resized crop of the decoded RGB / depth frame,
 *      flip, /255 + ImageNet mean/std, mm -> m) and :594-602 (depth_mask = depth > 0, depth -= its mean over the mask).
 *      rgb [B,Hs,Ws,3] uint8, depth [B,Hs,Ws] uint16 (mm), crop [B,4] = (top, left, height, width), flip [B] (may be null),
 *      has_depth [B] (null = all) -> x [B,6,R,R] fp32, depth_mask [B,R,R]; sums [B,2] uint64 scratch (exact mm sum / count). ---- */
int hcm_stage_input(const unsigned char* rgb, const unsigned short* depth, const int* crop, const int* flip,
                    const long long* has_depth, int B, int Hs, int Ws, int R, unsigned long long* sums, float* x,
                    float* depth_mask, cudaStream_t stream);
/* ---- segmentation fine-tuning head (seg_head.cu): learning/segment_trainer.py:722-745 (max of the L2-normalised projection maps),
 *      networks/fcn.py:108-110 + main_segmentor.py:76-79 (weighted CrossEntropyLoss(ignore_index) on the x4-upsampled logits),
 *      segment_trainer.py:375-379 (aAcc).  m1, m2, out, d1, d2 [P,128] channels-last (m2 null: one map); inv1, inv2 [P];
 *      logits / dlogits [P,Cn] (upsampled, Cn <= 64), label [P] int64, acc [4] fp64 scratch, out2 [2] = (loss, aAcc). ---- */
int hcm_l2norm_max_fwd(const float* m1, const float* m2, long P, int C, float* out, float* inv1, float* inv2,
                       cudaStream_t stream);
int hcm_l2norm_max_bwd(const float* dout, const float* m1, const float* m2, const float* inv1, const float* inv2, long P,
                       int C, float gscale, float* d1, float* d2, int accumulate, cudaStream_t stream);
int hcm_seg_ce_fwd(const float* logits, const long long* label, const float* class_weight, long P, int Cn, int ignore_index,
                   double* acc, float* out2, cudaStream_t stream);
int hcm_seg_ce_bwd(const float* logits, const long long* label, const float* class_weight, long P, int Cn, int ignore_index,
                   const double* acc, float gscale, float* dlogits, cudaStream_t stream);
/* ---- PointNet++ primitives (pointnet2.cu): the nine entry points of the reference's native extension `pointnet2_cuda`
 *      (networks/pointnet2/src/pointnet2_api.cpp:10-23), same argument meaning, raw device pointers instead of at::Tensor, int32
 *      indices, results identical to the reference kernels (ties included).  xyz / new_xyz / unknown / known [B,*,3];
 *      points [B,C,N] channel-major; the *_grad entry points ADD into a buffer the caller zeroed (pointnet2_utils.py:66,146,190). ---- */
int hcm_pn2_furthest_point_sampling(const float* xyz, int B, int N, int M, int* idx, cudaStream_t stream);
int hcm_pn2_ball_query(const float* new_xyz, const float* xyz, int B, int N, int M, float radius, int nsample, int* idx,
                       cudaStream_t stream);
int hcm_pn2_three_nn(const float* unknown, const float* known, int B, int n, int m, float* dist2, int* idx, cudaStream_t stream);
int hcm_pn2_three_interpolate(const float* points, const int* idx, const float* weight, int B, int C, int m, int n, float* out,
                              cudaStream_t stream);
int hcm_pn2_three_interpolate_grad(const float* grad_out, const int* idx, const float* weight, int B, int C, int n, int m,
                                   float* grad_points, cudaStream_t stream);
int hcm_pn2_group_points(const float* points, const int* idx, int B, int C, int N, int npoint, int nsample, float* out,
                         cudaStream_t stream);
int hcm_pn2_group_points_grad(const float* grad_out, const int* idx, int B, int C, int N, int npoint, int nsample,
                              float* grad_points, cudaStream_t stream);
int hcm_pn2_gather_points(const float* points, const int* idx, int B, int C, int N, int npoint, float* out, cudaStream_t stream);
int hcm_pn2_gather_points_grad(const float* grad_out, const int* idx, int B, int C, int N, int npoint, float* grad_points,
                               cudaStream_t stream);
int hcm_sgd_step(float* p, const float* g, float* buf, long n, float lr, float momentum, float wd, int first,
                 float gscale, cudaStream_t stream);
int hcm_zero(void* p, long bytes, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HCMOCO_H_ */
