// TEST INFRASTRUCTURE (oracle/): C entry points around the kernel launchers of the REFERENCE's own pointnet2 extension
// (/root/reference/pycontrast/networks/pointnet2/src/*_gpu.cu, compiled where they lie by oracle/build_ref.py into
// oracle/_ref/libpn2_ref.so — never copied into this repository).  tests/test_pointnet2_gpu.py runs hcm_pn2_* against these on the
// same device buffers and requires identical results.  Nothing under hcmoco_b200/ links or loads this.
#include "ball_query_gpu.h"
#include "group_points_gpu.h"
#include "interpolate_gpu.h"
#include "sampling_gpu.h"

extern "C" {
void ref_furthest_point_sampling(int b, int n, int m, const float* dataset, float* temp, int* idxs) {
  furthest_point_sampling_kernel_launcher(b, n, m, dataset, temp, idxs, 0);
}
void ref_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz, int* idx) {
  ball_query_kernel_launcher_fast(b, n, m, radius, nsample, new_xyz, xyz, idx, 0);
}
void ref_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2, int* idx) {
  three_nn_kernel_launcher_fast(b, n, m, unknown, known, dist2, idx, 0);
}
void ref_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx, const float* weight, float* out) {
  three_interpolate_kernel_launcher_fast(b, c, m, n, points, idx, weight, out, 0);
}
void ref_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out, const int* idx, const float* weight, float* grad_points) {
  three_interpolate_grad_kernel_launcher_fast(b, c, n, m, grad_out, idx, weight, grad_points, 0);
}
void ref_group_points(int b, int c, int n, int npoints, int nsample, const float* points, const int* idx, float* out) {
  group_points_kernel_launcher_fast(b, c, n, npoints, nsample, points, idx, out, 0);
}
void ref_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out, const int* idx, float* grad_points) {
  group_points_grad_kernel_launcher_fast(b, c, n, npoints, nsample, grad_out, idx, grad_points, 0);
}
void ref_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx, float* out) {
  gather_points_kernel_launcher_fast(b, c, n, npoints, points, idx, out, 0);
}
void ref_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out, const int* idx, float* grad_points) {
  gather_points_grad_kernel_launcher_fast(b, c, n, npoints, grad_out, idx, grad_points, 0);
}
}
