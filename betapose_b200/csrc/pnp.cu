// a9 (single-proposal pose-NMS) + a10 (key-point selection) + a11 (PnP) in one launch: one CTA per detection.
//   threads k < K      : score floor / threshold / -0.3 shift / rank-based selection
//   threads h < n_hyp  : one 5-point EPnP hypothesis each (fp64), scored against all selected points
//   warp 0             : picks the consensus set (shuffle arg-max) and runs the LM refit lane-parallel over points
// Reference: pPose_nms.py:24-122 (n = 1 branch), dataloader.py:715-726, utils/utils.py:17-41.
#include <cuda_runtime.h>

#include "betapose_b200.h"
#include "engine.h"
#include "pnp_math.cuh"

namespace {

constexpr int kMaxK = 64;
constexpr int kMaxHyp = 128;

struct WarpLanes {
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int count() const { return 32; }
  __device__ __forceinline__ void allreduce(double* v, int n) const {
    for (int i = 0; i < n; ++i) {
      double x = v[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      v[i] = x;
    }
  }
};

__global__ void __launch_bounds__(kMaxHyp)
pose_pnp_kernel(const float* __restrict__ preds_img, const float* __restrict__ maxval, const float* __restrict__ det_score,
                const uint8_t* __restrict__ valid, int K, const double* __restrict__ kp3d, const int32_t* __restrict__ model_idx,
                double fx, double fy, double cx, double cy, int left_number, int mode, int flags, double thr2, int n_hyp, uint32_t seed,
                float* __restrict__ keypoints, float* __restrict__ kp_score, float* __restrict__ proposal,
                uint8_t* __restrict__ selected, double* __restrict__ R_out, double* __restrict__ t_out,
                uint8_t* __restrict__ inlier, int32_t* __restrict__ status) {
  __shared__ double s_pw[kMaxK * 3];
  __shared__ double s_uv[kMaxK * 2];
  __shared__ float s_sc[kMaxK];
  __shared__ uint8_t s_sel[kMaxK];
  __shared__ uint8_t s_inl[kMaxK];
  __shared__ int s_cnt[kMaxHyp];
  __shared__ double s_tot[kMaxHyp];
  __shared__ double s_R[kMaxHyp * 9];
  __shared__ double s_t[kMaxHyp * 3];
  __shared__ int s_state;  // 1 = run PnP, 0 = rejected
  __shared__ int s_nsel;

  const int i = blockIdx.x;
  const int tid = threadIdx.x;
  const bool ok_in = !valid || valid[i];

  // ---- stage A: pose-NMS (n = 1) and selection
  if (tid < K) {
    const bool raw = flags & BP_PNP_RAW_POINTS;  // inputs are final key-points: no pose-NMS arithmetic
    float sc = ok_in ? (raw ? 1.f : maxval[(long)i * K + tid]) : 0.f;
    if (sc == 0.f) sc = 1e-5f;
    s_sc[tid] = sc;
    const double* mp = kp3d + ((long)(model_idx ? model_idx[i] : 0) * K + tid) * 3;
    s_pw[3 * tid] = mp[0];
    s_pw[3 * tid + 1] = mp[1];
    s_pw[3 * tid + 2] = mp[2];
    const float shift = raw ? 0.f : 0.3f;
    const float kx = ok_in ? __fsub_rn(preds_img[((long)i * K + tid) * 2], shift) : 0.f;
    const float ky = ok_in ? __fsub_rn(preds_img[((long)i * K + tid) * 2 + 1], shift) : 0.f;
    s_uv[2 * tid] = (double)kx;
    s_uv[2 * tid + 1] = (double)ky;
    keypoints[((long)i * K + tid) * 2] = kx;
    keypoints[((long)i * K + tid) * 2 + 1] = ky;
    kp_score[(long)i * K + tid] = sc;
  }
  __syncthreads();
  if (tid < K) {
    // delete arg-min (first on ties) until left_number remain  <=>  drop the (K - left) lowest by (score, index)
    int rank = 0;
    const float me = s_sc[tid];
    for (int j = 0; j < K; ++j) rank += (s_sc[j] < me || (s_sc[j] == me && j < tid)) ? 1 : 0;
    const int drop = K > left_number ? K - left_number : 0;
    s_sel[tid] = rank >= drop ? 1 : 0;
  }
  if (tid == 0) {
    float mx = s_sc[0], sum = 0.f;
    for (int j = 0; j < K; ++j) {
      mx = fmaxf(mx, s_sc[j]);
      sum = __fadd_rn(sum, s_sc[j]);
    }
    const bool pass = ok_in && ((flags & BP_PNP_RAW_POINTS) || !(mx < 0.3f));
    s_state = pass ? 1 : 0;
    proposal[i] = pass ? __fadd_rn(__fadd_rn(__fdiv_rn(sum, (float)K), det_score ? det_score[i] : 0.f), __fmul_rn(1.25f, mx)) : 0.f;
  }
  __syncthreads();
  if (tid == 0) {
    int c = 0;
    for (int j = 0; j < K; ++j) c += s_sel[j];
    s_nsel = c;
  }
  if (tid < K) {
    selected[(long)i * K + tid] = s_state ? s_sel[tid] : 0;
    s_inl[tid] = 0;
  }
  __syncthreads();

  const bool run = s_state == 1 && s_nsel >= 4 && !(flags & BP_PNP_NMS_ONLY);
  const bool ransac = run && mode == 0 && s_nsel >= 6;

  // ---- stage B: hypotheses (RANSAC: one 5-point EPnP per thread; all-points mode: thread 0 solves all selected)
  if (tid < kMaxHyp) s_cnt[tid] = -1;
  __syncthreads();
  if (run && (ransac ? tid < n_hyp : tid == 0)) {
    int pool[kMaxK];
    int m = 0;
    for (int j = 0; j < K; ++j)
      if (s_sel[j]) pool[m++] = j;
    if (ransac) {
      bp::pnp::sample_subset(pool, m, tid, seed, 5);
      m = 5;
    }
    double R[9], t[3];
    if (bp::pnp::epnp(s_pw, s_uv, pool, m, fx, fy, cx, cy, R, t)) {
      int cnt = m;
      double tot = 0.0;
      if (ransac) bp::pnp::score_hypothesis(R, t, s_pw, s_uv, s_sel, K, fx, fy, cx, cy, thr2, &cnt, &tot);
      s_cnt[tid] = cnt;
      s_tot[tid] = tot;
      for (int k = 0; k < 9; ++k) s_R[tid * 9 + k] = R[k];
      for (int k = 0; k < 3; ++k) s_t[tid * 3 + k] = t[k];
    }
  }
  __syncthreads();

  // ---- stage C + D (warp 0): pick the consensus winner, then alternate {classify points against the current
  // pose, LM refit on the consensus set} until the set is stable (at most BP_PNP_LO_ROUNDS refits)
  if (tid < 32) {
    int bc = -1, bh = 0x7fffffff;
    double bt = INFINITY;
    for (int h = tid; h < kMaxHyp; h += 32) {
      const int c = s_cnt[h];
      if (c < 0) continue;
      const double tt = s_tot[h];
      if (c > bc || (c == bc && (tt < bt || (tt == bt && h < bh)))) {
        bc = c;
        bt = tt;
        bh = h;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      const double ot = __shfl_xor_sync(0xffffffffu, bt, o);
      const int oh = __shfl_xor_sync(0xffffffffu, bh, o);
      if (oc > bc || (oc == bc && (ot < bt || (ot == bt && oh < bh)))) {
        bc = oc;
        bt = ot;
        bh = oh;
      }
    }
    bool ok = run && bc >= 4;
    double R[9], t[3];
    for (int k = 0; k < 9; ++k) R[k] = ok ? s_R[bh * 9 + k] : 0.0;
    for (int k = 0; k < 3; ++k) t[k] = ok ? s_t[bh * 3 + k] : 0.0;
    if (ok) {
      WarpLanes ln;
      for (int round = 0; round < BP_PNP_LO_ROUNDS; ++round) {
        int changed = 0, cnt = 0;
        for (int j = tid; j < K; j += 32) {
          const uint8_t in = s_sel[j] && (!ransac || bp::pnp::within_threshold(R, t, s_pw, s_uv, j, fx, fy, cx, cy, thr2)) ? 1 : 0;
          changed |= in != s_inl[j];
          s_inl[j] = in;
          cnt += in;
        }
        __syncwarp();
        changed = __any_sync(0xffffffffu, changed);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (round > 0 && !changed) break;
        if (cnt < 4) {
          ok = false;
          break;
        }
        bp::pnp::lm_refine(ln, R, t, s_pw, s_uv, s_inl, K, fx, fy, cx, cy, 50);
        if (!ransac) break;
      }
    }
    if (tid == 0) {
      for (int k = 0; k < 9; ++k) R_out[(long)i * 9 + k] = ok ? R[k] : 0.0;
      for (int k = 0; k < 3; ++k) t_out[(long)i * 3 + k] = ok ? t[k] : 0.0;
      status[i] = s_state == 0 ? 0 : ((ok || (flags & BP_PNP_NMS_ONLY)) ? 1 : -1);
    }
    if (!ok)
      for (int j = tid; j < K; j += 32) s_inl[j] = 0;
  }
  __syncthreads();
  if (tid < K) inlier[(long)i * K + tid] = s_inl[tid];
}

}  // namespace

extern "C" int bp_pose_pnp(bp_engine* e, const float* preds_img, const float* maxval, const float* det_score,
                           const uint8_t* valid, int n, int K, const double* kp3d, const int32_t* model_idx,
                           const double* cam, int left_number, int mode, int flags, float reproj_thr, int n_hyp,
                           uint32_t seed, float* keypoints, float* kp_score, float* proposal, uint8_t* selected, double* R, double* t,
                           uint8_t* inlier, int32_t* status, void* stream) {
  if (!e || !preds_img || !kp3d || !cam || n <= 0 || (!(flags & BP_PNP_RAW_POINTS) && (!maxval || !det_score))) return bp_fail(BP_ERR_INVALID, "bp_pose_pnp: bad arguments");
  if (K < 1 || K > kMaxK) return bp_fail(BP_ERR_UNSUPPORTED, "bp_pose_pnp: K must be in [1, 64]");
  if (n_hyp < 1 || n_hyp > kMaxHyp) return bp_fail(BP_ERR_UNSUPPORTED, "bp_pose_pnp: n_hyp must be in [1, 128]");
  if (mode != 0 && mode != 1) return bp_fail(BP_ERR_INVALID, "bp_pose_pnp: mode");
  pose_pnp_kernel<<<n, kMaxHyp, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      preds_img, maxval, det_score, valid, K, kp3d, model_idx, cam[0], cam[1], cam[2], cam[3], left_number, mode, flags,
      (double)reproj_thr * (double)reproj_thr, n_hyp, seed, keypoints, kp_score, proposal, selected, R, t, inlier, status);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? BP_OK : bp_fail(BP_ERR_CUDA, cudaGetErrorString(err));
}
