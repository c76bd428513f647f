// Scoring stage: ADD, 2-D reprojection error and box IoU of estimated against ground-truth
// poses, batched on the GPU.  One CTA per image, threads over the model's vertices, fp64 like the reference's numpy.
// Reference: 3_6Dpose_estimator/utils/metrics.py:10-22 (add_err), :77-93 (iou), :96-127 (projection_error_2d) and
// the loop at betapose_evaluate.py:203-266.
#include <cuda_runtime.h>

#include "betapose_b200.h"
#include "engine.h"

namespace {

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += s_red[i];  // fixed order: identical on every thread, run to run
  return t;
}

__device__ __forceinline__ double dot3(const double* r, double x, double y, double z, double t) {
  return __fma_rn(r[0], x, __fma_rn(r[1], y, __fma_rn(r[2], z, t)));
}

__global__ void __launch_bounds__(256)
score_poses_kernel(const double* __restrict__ R_est, const double* __restrict__ t_est, const int32_t* __restrict__ status,
                   const float* __restrict__ box_est, const double* __restrict__ R_gt, const double* __restrict__ t_gt,
                   const float* __restrict__ box_gt, const double* __restrict__ model, const int32_t* __restrict__ model_idx,
                   int V, double fx, double fy, double cx, double cy, double* __restrict__ add_err,
                   double* __restrict__ proj_err, float* __restrict__ iou_out, uint8_t* __restrict__ scored) {
  __shared__ double s_red[8];
  const int i = blockIdx.x;
  const double* Re = R_est + 9 * (long)i;
  const double* te = t_est + 3 * (long)i;
  const double* Rg = R_gt + 9 * (long)i;
  const double* tg = t_gt + 3 * (long)i;
  const double* mv = model + (long)(model_idx ? model_idx[i] : 0) * V * 3;
  // iou(gt_box, est_box): metrics.py:77-93, boxes as corners (x1, y1, x2, y2)
  float iou = 0.f;
  {
    const float* g = box_gt + 4 * (long)i;
    const float* e = box_est + 4 * (long)i;
    const double xA = fmax((double)g[0], (double)e[0]), yA = fmax((double)g[1], (double)e[1]);
    const double xB = fmin((double)g[2], (double)e[2]), yB = fmin((double)g[3], (double)e[3]);
    if (!(xB <= xA || yB <= yA)) {
      const double inter = (xB - xA) * (yB - yA);
      const double aA = ((double)g[2] - g[0]) * ((double)g[3] - g[1]);
      const double aB = ((double)e[2] - e[0]) * ((double)e[3] - e[1]);
      iou = (float)(inter / (aA + aB - inter));
    }
  }
  const bool have_pose = !status || status[i] == 1;
  // cam @ pose (3 x 4), as projection_error_2d builds it
  double Mg[12], Me[12];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double g0 = c < 3 ? Rg[c] : tg[0], g1 = c < 3 ? Rg[3 + c] : tg[1], g2 = c < 3 ? Rg[6 + c] : tg[2];
    const double e0 = c < 3 ? Re[c] : te[0], e1 = c < 3 ? Re[3 + c] : te[1], e2 = c < 3 ? Re[6 + c] : te[2];
    Mg[c] = __fma_rn(fx, g0, __dmul_rn(cx, g2)); Mg[4 + c] = __fma_rn(fy, g1, __dmul_rn(cy, g2)); Mg[8 + c] = g2;
    Me[c] = __fma_rn(fx, e0, __dmul_rn(cx, e2)); Me[4 + c] = __fma_rn(fy, e1, __dmul_rn(cy, e2)); Me[8 + c] = e2;
  }
  double sum_add = 0.0, sum_proj = 0.0;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const double x = mv[3 * v], y = mv[3 * v + 1], z = mv[3 * v + 2];
    // both poses go through the same explicitly fused expression, so identical poses score exactly zero (left to the
    // compiler, the subtraction may be contracted into one side's last multiply-add)
    const double ax = dot3(Rg, x, y, z, tg[0]) - dot3(Re, x, y, z, te[0]);
    const double ay = dot3(Rg + 3, x, y, z, tg[1]) - dot3(Re + 3, x, y, z, te[1]);
    const double az = dot3(Rg + 6, x, y, z, tg[2]) - dot3(Re + 6, x, y, z, te[2]);
    sum_add += sqrt(ax * ax + ay * ay + az * az);
    const double gw = dot3(Mg + 8, x, y, z, Mg[11]);
    const double gu = __ddiv_rn(dot3(Mg, x, y, z, Mg[3]), gw), gv = __ddiv_rn(dot3(Mg + 4, x, y, z, Mg[7]), gw);
    const double ew = dot3(Me + 8, x, y, z, Me[11]);
    const double eu = __ddiv_rn(dot3(Me, x, y, z, Me[3]), ew), ev = __ddiv_rn(dot3(Me + 4, x, y, z, Me[7]), ew);
    sum_proj += sqrt((gu - eu) * (gu - eu) + (gv - ev) * (gv - ev));
  }
  sum_add = block_sum(sum_add, s_red);
  sum_proj = block_sum(sum_proj, s_red);
  if (threadIdx.x == 0) {
    // the reference scores ADD / reprojection only for frames with a pose whose box overlaps the ground truth by
    // IoU >= 0.5 (betapose_evaluate.py:244-257); `scored` marks them, the errors are written for every posed frame
    add_err[i] = have_pose ? sum_add / (double)V : 0.0;
    proj_err[i] = have_pose ? sum_proj / (double)V : 0.0;
    iou_out[i] = have_pose ? iou : 0.f;
    scored[i] = (have_pose && iou >= 0.5f) ? 1 : 0;
  }
}

}  // namespace

extern "C" int bp_score_poses(bp_engine* e, int n, const double* R_est, const double* t_est, const int32_t* status,
                              const float* box_est, const double* R_gt, const double* t_gt, const float* box_gt,
                              const double* model, const int32_t* model_idx, int n_vertices, const double* cam, double* add_err,
                              double* proj_err, float* iou, uint8_t* scored, void* stream) {
  if (!e || n <= 0 || !R_est || !t_est || !box_est || !R_gt || !t_gt || !box_gt || !model || n_vertices <= 0 || !cam || !add_err ||
      !proj_err || !iou || !scored)
    return bp_fail(BP_ERR_INVALID, "bp_score_poses: bad arguments");
  score_poses_kernel<<<n, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(R_est, t_est, status, box_est, R_gt, t_gt, box_gt, model,
                                                                           model_idx, n_vertices, cam[0], cam[1], cam[2], cam[3],
                                                                           add_err, proj_err, iou, scored);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? BP_OK : bp_fail(BP_ERR_CUDA, cudaGetErrorString(err));
}
