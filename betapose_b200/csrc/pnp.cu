// a9 (single-proposal pose-NMS) + a10 (key-point selection) + a11 (PnP) in two launches:
//   pnp_hypotheses_serial_kernel : one CTA per image, one thread per five-point EPnP hypothesis (fp64): pose-NMS + selection of
//                           the image once, then every thread solves its hypothesis -- control points, 12 x 12 M^T M
//                           eigen-problem by Householder + QL in local memory, three beta initialisations + Gauss-Newton,
//                           Horn absolute orientation -- scores it against all selected points (12 px) and leaves it in a
//                           global scratch row.
//   pnp_refine_kernel     : one warp per image picks the consensus winner (shuffle arg-max) and alternates
//                           {classify points against the current pose, Levenberg-Marquardt refit, lane-parallel over
//                           points} until the consensus set is stable.
// Reference: pPose_nms.py:24-122 (n = 1 branch), dataloader.py:715-726, utils/utils.py:17-41.
#include <cuda_runtime.h>

#include <cstdlib>

#include "betapose_b200.h"
#include "engine.h"
#include "pnp_math.cuh"

namespace {

constexpr int kMaxK = 64;
constexpr int kMaxHyp = 128;
constexpr int kHypRow = 16;         // doubles per hypothesis in the scratch: count, total, R[9], t[3], pad

// One warp, lanes over points; the 28 sums of an LM step are all-reduced by shuffles (independent chains: they
// pipeline).  A shared-memory variant in which lane l adds up sum number l over the points measured 1.6x slower:
// 50-long dependent fp64 add chains.
struct WarpLanes {
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int count() const { return 32; }
  template <int N>
  __device__ __forceinline__ void allreduce_n(double* v) const {  // N compile-time: v stays in registers
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double x = v[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      v[i] = x;
    }
  }
  __device__ __forceinline__ void accumulate(const double* R, const double* t, const double* pw, const double* uv,
                                             const uint8_t* mask, int n, double fx, double fy, double cx, double cy,
                                             double* acc) const {
    bp::pnp::lm_accumulate(*this, R, t, pw, uv, mask, n, fx, fy, cx, cy, acc);
  }
};

// ---- a9 + a10 for one image, by the first K threads of the CTA; leaves the selected points in shared memory.
// Returns (to every thread) whether PnP should run; writes the per-image outputs when `write_out`.
struct ImageState {
  double pw[kMaxK * 3];
  double uv[kMaxK * 2];
  float sc[kMaxK];
  uint8_t sel[kMaxK];
  int state;  // 1 = passed pose-NMS
  int nsel;
};

__device__ __forceinline__ void stage_a(ImageState& S, int i, int K, const float* preds_img, const float* maxval,
                                        const float* det_score, const uint8_t* valid, const double* kp3d,
                                        const int32_t* model_idx, int left_number, int flags, bool write_out, float* keypoints,
                                        float* kp_score, float* proposal, uint8_t* selected) {
  const int tid = threadIdx.x;
  const bool ok_in = !valid || valid[i];
  const bool raw = flags & BP_PNP_RAW_POINTS;  // inputs are final key-points: no pose-NMS arithmetic
  for (int k = tid; k < K; k += blockDim.x) {
    float sc = ok_in ? (raw ? 1.f : maxval[(long)i * K + k]) : 0.f;
    if (sc == 0.f) sc = 1e-5f;
    S.sc[k] = sc;
    const double* mp = kp3d + ((long)(model_idx ? model_idx[i] : 0) * K + k) * 3;
    S.pw[3 * k] = mp[0];
    S.pw[3 * k + 1] = mp[1];
    S.pw[3 * k + 2] = mp[2];
    const float shift = raw ? 0.f : 0.3f;
    const float kx = ok_in ? __fsub_rn(preds_img[((long)i * K + k) * 2], shift) : 0.f;
    const float ky = ok_in ? __fsub_rn(preds_img[((long)i * K + k) * 2 + 1], shift) : 0.f;
    S.uv[2 * k] = (double)kx;
    S.uv[2 * k + 1] = (double)ky;
    if (write_out) {
      keypoints[((long)i * K + k) * 2] = kx;
      keypoints[((long)i * K + k) * 2 + 1] = ky;
      kp_score[(long)i * K + k] = sc;
    }
  }
  __syncthreads();
  for (int k = tid; k < K; k += blockDim.x) {
    // delete arg-min (first on ties) until left_number remain  <=>  drop the (K - left) lowest by (score, index)
    int rank = 0;
    const float me = S.sc[k];
    for (int j = 0; j < K; ++j) rank += (S.sc[j] < me || (S.sc[j] == me && j < k)) ? 1 : 0;
    const int drop = K > left_number ? K - left_number : 0;
    S.sel[k] = rank >= drop ? 1 : 0;
  }
  if (tid == 0) {
    float mx = S.sc[0], sum = 0.f;
    for (int j = 0; j < K; ++j) {
      mx = fmaxf(mx, S.sc[j]);
      sum = __fadd_rn(sum, S.sc[j]);
    }
    const bool pass = ok_in && (raw || !(mx < 0.3f));
    S.state = pass ? 1 : 0;
    if (write_out)
      proposal[i] = pass ? __fadd_rn(__fadd_rn(__fdiv_rn(sum, (float)K), det_score ? det_score[i] : 0.f), __fmul_rn(1.25f, mx)) : 0.f;
  }
  __syncthreads();
  if (tid == 0) {
    int c = 0;
    for (int j = 0; j < K; ++j) c += S.sel[j];
    S.nsel = c;
  }
  if (write_out)
    for (int k = tid; k < K; k += blockDim.x) selected[(long)i * K + k] = S.state ? S.sel[k] : 0;
  __syncthreads();
}

// ONE thread per hypothesis, one CTA per image.  The 12 x 12 eigen-problem runs as a Householder + QL pass in the thread's
// own local memory (pnp_math.cuh: sym_eig_ql, ~5x fewer flops than Jacobi sweeps), no shared-memory matrices, no shuffles:
// 128 warps of independent fp64 work per 64-frame batch, pose-NMS + selection once per image.  (Round 1 ran 8 / 16
// cooperating lanes per hypothesis around a parallel-order Jacobi in shared memory: 1024 warps in lock step, 40 KB of
// shared memory per CTA, 0.55 ms per batch of 64 against 0.3 ms now; one algorithm for every batch size also keeps the
// poses bit-identical whatever batch a frame arrives in.)
__global__ void __launch_bounds__(128)
pnp_hypotheses_serial_kernel(const float* __restrict__ preds_img, const float* __restrict__ maxval, const float* __restrict__ det_score,
                             const uint8_t* __restrict__ valid, int K, const double* __restrict__ kp3d,
                             const int32_t* __restrict__ model_idx, double fx, double fy, double cx, double cy, int left_number,
                             int mode, int flags, double thr2, int n_hyp, uint32_t seed, float* __restrict__ keypoints,
                             float* __restrict__ kp_score, float* __restrict__ proposal, uint8_t* __restrict__ selected,
                             double* __restrict__ hyp /* [images][kMaxHyp][kHypRow] */) {
  __shared__ ImageState S;
  const int i = blockIdx.x;
  const int h = threadIdx.x;  // hypothesis index
  stage_a(S, i, K, preds_img, maxval, det_score, valid, kp3d, model_idx, left_number, flags, true, keypoints, kp_score, proposal,
          selected);
  const bool run = S.state == 1 && S.nsel >= 4 && !(flags & BP_PNP_NMS_ONLY);
  const bool ransac = run && mode == 0 && S.nsel >= 6;
  if (h >= kMaxHyp) return;
  double* row = hyp + ((long)i * kMaxHyp + h) * kHypRow;
  bool solved = false;
  int cnt = 0;
  double tot = 0.0;
  double R[9], t[3];
  if (run && (ransac ? h < n_hyp : h == 0)) {
    int pool[kMaxK];
    int m = 0;
    for (int j = 0; j < K; ++j)
      if (S.sel[j]) pool[m++] = j;
    if (ransac) {
      bp::pnp::sample_subset(pool, m, h, seed, 5);
      m = 5;
    }
    bp::pnp::QlEig12 eig;
    if (bp::pnp::epnp(eig, S.pw, S.uv, pool, m, fx, fy, cx, cy, R, t)) {
      solved = true;
      cnt = m;
      if (ransac) bp::pnp::score_hypothesis(R, t, S.pw, S.uv, S.sel, K, fx, fy, cx, cy, thr2, &cnt, &tot);
    }
  }
  row[0] = solved ? (double)cnt : -1.0;
  row[1] = tot;
  if (solved) {
#pragma unroll
    for (int k = 0; k < 9; ++k) row[2 + k] = R[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) row[11 + k] = t[k];
  }
}

__global__ void __launch_bounds__(32)
pnp_refine_kernel(const float* __restrict__ preds_img, const float* __restrict__ maxval, const uint8_t* __restrict__ valid, int K,
                  const double* __restrict__ kp3d, const int32_t* __restrict__ model_idx, double fx, double fy, double cx, double cy,
                  int left_number, int mode, int flags, double thr2, int n_hyp, const double* __restrict__ hyp,
                  double* __restrict__ R_out, double* __restrict__ t_out, uint8_t* __restrict__ inlier, int32_t* __restrict__ status) {
  __shared__ ImageState S;
  __shared__ uint8_t s_inl[kMaxK];
  const int i = blockIdx.x;
  const int tid = threadIdx.x;
  stage_a(S, i, K, preds_img, maxval, nullptr, valid, kp3d, model_idx, left_number, flags, false, nullptr, nullptr, nullptr, nullptr);
  for (int k = tid; k < K; k += 32) s_inl[k] = 0;
  __syncwarp();

  const bool run = S.state == 1 && S.nsel >= 4 && !(flags & BP_PNP_NMS_ONLY);
  const bool ransac = run && mode == 0 && S.nsel >= 6;
  const int nh = run ? (ransac ? n_hyp : 1) : 0;

  // consensus winner: most inliers, then smallest summed squared error, then lowest hypothesis index
  int bc = -1, bh = 0x7fffffff;
  double bt = INFINITY;
  const double* rows = hyp + (long)i * kMaxHyp * kHypRow;
  for (int h = tid; h < nh; h += 32) {
    const int c = (int)rows[h * kHypRow];
    if (c < 0) continue;
    const double tt = rows[h * kHypRow + 1];
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
  // `have`: some hypothesis was solved.  A survivor of pose-NMS whose PnP fails (status -1) still gets the solver's last
  // estimate written out: the reference appends cam_R / cam_t for every survivor (dataloader.py:722-727 ignores
  // solvePnP's return flag), so the frame must stay in the result list and score as a miss, not vanish.
  const bool have = run && bc >= 0;
  double R[9], t[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = have ? rows[bh * kHypRow + 2 + k] : 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) t[k] = have ? rows[bh * kHypRow + 11 + k] : 0.0;
  if (ok) {
    // alternate {classify points against the current pose, LM refit on the consensus set} until the set is stable
    WarpLanes ln;
    for (int round = 0; round < BP_PNP_LO_ROUNDS; ++round) {
      int changed = 0, cnt = 0;
      for (int j = tid; j < K; j += 32) {
        const uint8_t in = S.sel[j] && (!ransac || bp::pnp::within_threshold(R, t, S.pw, S.uv, j, fx, fy, cx, cy, thr2)) ? 1 : 0;
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
      bp::pnp::lm_refine(ln, R, t, S.pw, S.uv, s_inl, K, fx, fy, cx, cy, BP_PNP_LM_ITERS);
      if (!ransac) break;
    }
  }
  if (tid == 0) {
    for (int k = 0; k < 9; ++k) R_out[(long)i * 9 + k] = have ? R[k] : 0.0;
    for (int k = 0; k < 3; ++k) t_out[(long)i * 3 + k] = have ? t[k] : 0.0;
    status[i] = S.state == 0 ? 0 : ((ok || (flags & BP_PNP_NMS_ONLY)) ? 1 : -1);
  }
  __syncwarp();
  for (int j = tid; j < K; j += 32) inlier[(long)i * K + j] = ok ? s_inl[j] : 0;
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
  // hypothesis scratch (grow-only; (re)allocated at most once per batch size, outside steady state)
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t need = (size_t)n * kMaxHyp * kHypRow * sizeof(double);
  bp_engine::StreamScratch& sc = e->scratch_for(st);
  if (!e->grow(reinterpret_cast<void**>(&sc.pnp), &sc.pnp_bytes, need)) return bp_fail(BP_ERR_CUDA, "bp_pose_pnp: scratch allocation failed");
  const double thr2 = (double)reproj_thr * (double)reproj_thr;
  {
    const int threads = mode == 0 ? ((n_hyp + 31) / 32) * 32 : 32;
    pnp_hypotheses_serial_kernel<<<n, threads, 0, st>>>(preds_img, maxval, det_score, valid, K, kp3d, model_idx, cam[0], cam[1], cam[2],
                                                        cam[3], left_number, mode, flags, thr2, n_hyp, seed, keypoints, kp_score,
                                                        proposal, selected, sc.pnp);
  }
  pnp_refine_kernel<<<n, 32, 0, st>>>(preds_img, maxval, valid, K, kp3d, model_idx, cam[0], cam[1], cam[2], cam[3], left_number, mode,
                                      flags, thr2, n_hyp, sc.pnp, R, t, inlier, status);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? BP_OK : bp_fail(BP_ERR_CUDA, cudaGetErrorString(err));
}
