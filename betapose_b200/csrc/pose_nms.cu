// General parametric pose-NMS for n >= 1 proposals of one image: greedy pick of the best mean
// score, suppression by key-point similarity, score-weighted merge of every suppressed cluster.
// Reference: 3_6Dpose_estimator/pPose_nms.py:24-122 (pose_nms), :204-240 (p_merge_fast), :243-267
// (get_parametric_distance), :270-281 (PCK_match); constants :12-20.
// One CTA per image; a warp owns a proposal, lanes run over the key-points; the pick loop is sequential (<= n rounds).
#include <cuda_runtime.h>
#include <math.h>

#include "betapose_b200.h"
#include "engine.h"

namespace {

constexpr int kMaxN = 64;  // proposals per image
constexpr int kMaxK = 64;  // key-points
constexpr float kDelta1 = 1.f, kMu = 1.7f, kDelta2 = 2.65f, kGamma = 22.48f, kScoreThr = 0.3f, kAlpha = 0.1f;
constexpr int kMatchThr = 5;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
pose_nms_kernel(const float* __restrict__ bboxes, const float* __restrict__ bbox_scores, const float* __restrict__ pose_preds,
                const float* __restrict__ pose_scores, const int32_t* __restrict__ first, const int32_t* __restrict__ count, int K,
                int32_t* __restrict__ out_count, int32_t* __restrict__ out_pick, float* __restrict__ out_kp,
                float* __restrict__ out_score, float* __restrict__ out_prop) {
  __shared__ float s_human[kMaxN];
  __shared__ float s_ref[kMaxN];
  __shared__ unsigned char s_alive[kMaxN];
  __shared__ unsigned char s_del[kMaxN];
  __shared__ int s_pick[kMaxN];
  __shared__ unsigned long long s_merge[kMaxN];
  __shared__ int s_npick, s_cur, s_nout;
  const int img = blockIdx.x;
  const int base = first ? first[img] : 0;
  const int n = count[img];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* P = pose_preds + (long)base * K * 2;
  const float* S = pose_scores + (long)base * K;
  auto score = [&](int i, int k) {  // zero scores count as 1e-5 (pPose_nms.py:34)
    const float v = S[(long)i * K + k];
    return v == 0.f ? 1e-5f : v;
  };
  for (int i = warp; i < n; i += nw) {
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += score(i, k);
    s = warp_sum(s);
    if (lane == 0) {
      s_human[i] = s / (float)K;
      const float* b = bboxes + (long)(base + i) * 4;
      s_ref[i] = kAlpha * fmaxf(b[2] - b[0], b[3] - b[1]);
      s_alive[i] = 1;
    }
  }
  if (threadIdx.x == 0) s_npick = 0;
  __syncthreads();
  // ---- greedy pick / suppress
  for (int round = 0; round < n; ++round) {
    if (threadIdx.x == 0) {
      int best = -1;
      for (int i = 0; i < n; ++i)
        if (s_alive[i] && (best < 0 || s_human[i] > s_human[best])) best = i;  // first maximum in proposal order
      s_cur = best;
    }
    __syncthreads();
    const int pid = s_cur;
    if (pid < 0) break;
    const float ref7 = fminf(s_ref[pid], 7.f);
    for (int i = warp; i < n; i += nw) {
      if (!s_alive[i]) continue;
      float sd = 0.f, pd = 0.f;
      int nm = 0;
      for (int k = lane; k < K; k += 32) {
        const float dx = P[((long)pid * K + k) * 2] - P[((long)i * K + k) * 2];
        const float dy = P[((long)pid * K + k) * 2 + 1] - P[((long)i * K + k) * 2 + 1];
        const float d = sqrtf(dx * dx + dy * dy);
        if (d <= 1.f) sd += tanhf(score(pid, k) / kDelta1) * tanhf(score(i, k) / kDelta1);
        pd += expf(-1.f * d / kDelta2);
        nm += (d / ref7 <= 1.f) ? 1 : 0;
      }
      sd = warp_sum(sd);
      pd = warp_sum(pd);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nm += __shfl_xor_sync(0xffffffffu, nm, o);
      if (lane == 0) s_del[i] = (sd + kMu * pd > kGamma || nm >= kMatchThr) ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long m = 0ull;
      for (int i = 0; i < n; ++i)
        if (s_alive[i] && s_del[i]) m |= 1ull << i;
      if (m == 0ull) m = 1ull << pid;  // nothing matched: only the pick leaves (pPose_nms.py:73-74)
      for (int i = 0; i < n; ++i)
        if ((m >> i) & 1ull) s_alive[i] = 0;
      s_pick[s_npick] = pid;
      s_merge[s_npick] = m;
      ++s_npick;
    }
    __syncthreads();
  }
  // ---- merge every cluster, filter, emit (sequential over picks to keep the reference's output order)
  if (threadIdx.x == 0) s_nout = 0;
  __syncthreads();
  __shared__ float s_mx, s_my, s_ms;  // scratch for reductions by warp 0
  for (int j = 0; j < s_npick; ++j) {
    const int pk = s_pick[j];
    const unsigned long long m = s_merge[j];
    if (warp == 0) {
      float mx = -INFINITY;
      for (int k = lane; k < K; k += 32) mx = fmaxf(mx, score(pk, k));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      bool keep = !(mx < kScoreThr);
      const float ref15 = fminf(s_ref[pk], 15.f);
      const int slot = s_nout;
      float msum = 0.f, mmax = -INFINITY;
      float kx[2], ky[2], ks[2];
      for (int q = 0, k = lane; k < K; k += 32, ++q) {
        // p_merge_fast: weights = score * [dist <= ref] normalised over the cluster, per key-point
        float wsum = 0.f;
        for (int i = 0; i < n; ++i) {
          if (!((m >> i) & 1ull)) continue;
          const float dx = P[((long)pk * K + k) * 2] - P[((long)i * K + k) * 2];
          const float dy = P[((long)pk * K + k) * 2 + 1] - P[((long)i * K + k) * 2 + 1];
          if (sqrtf(dx * dx + dy * dy) <= ref15) wsum += score(i, k);
        }
        float px = 0.f, py = 0.f, ps = 0.f;
        for (int i = 0; i < n; ++i) {
          if (!((m >> i) & 1ull)) continue;
          const float dx = P[((long)pk * K + k) * 2] - P[((long)i * K + k) * 2];
          const float dy = P[((long)pk * K + k) * 2 + 1] - P[((long)i * K + k) * 2 + 1];
          const float ms = (sqrtf(dx * dx + dy * dy) <= ref15) ? score(i, k) : 0.f;
          const float w = ms / wsum;
          px += P[((long)i * K + k) * 2] * w;
          py += P[((long)i * K + k) * 2 + 1] * w;
          ps += ms * w;
        }
        kx[q] = px; ky[q] = py; ks[q] = ps;
        msum += ps;
        mmax = fmaxf(mmax, ps);
      }
      msum = warp_sum(msum);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mmax = fmaxf(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
      keep = keep && !(mmax < kScoreThr);  // (areaThres is 0 in the reference: the area test never rejects)
      if (keep) {
        for (int q = 0, k = lane; k < K; k += 32, ++q) {
          out_kp[(((long)base + slot) * K + k) * 2] = kx[q] - 0.3f;
          out_kp[(((long)base + slot) * K + k) * 2 + 1] = ky[q] - 0.3f;
          out_score[((long)base + slot) * K + k] = ks[q];
        }
        if (lane == 0) {
          out_pick[base + slot] = pk;
          out_prop[base + slot] = msum / (float)K + bbox_scores[base + pk] + 1.25f * mmax;
          s_nout = slot + 1;
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out_count[img] = s_nout;
  (void)s_mx; (void)s_my; (void)s_ms;
}

}  // namespace

extern "C" int bp_pose_nms(bp_engine* e, int n_images, const int32_t* first, const int32_t* count, int max_count, int K,
                           const float* bboxes, const float* bbox_scores, const float* pose_preds, const float* pose_scores,
                           int32_t* out_count, int32_t* out_pick, float* out_keypoints, float* out_kp_score, float* out_proposal,
                           void* stream) {
  if (!e || n_images <= 0 || !count || !bboxes || !bbox_scores || !pose_preds || !pose_scores || !out_count || !out_pick ||
      !out_keypoints || !out_kp_score || !out_proposal)
    return bp_fail(BP_ERR_INVALID, "bp_pose_nms: bad arguments");
  if (K < 1 || K > kMaxK || max_count < 1 || max_count > kMaxN)
    return bp_fail(BP_ERR_UNSUPPORTED, "bp_pose_nms: at most 64 proposals per image and 64 key-points");
  pose_nms_kernel<<<n_images, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(bboxes, bbox_scores, pose_preds, pose_scores, first,
                                                                              count, K, out_count, out_pick, out_keypoints,
                                                                              out_kp_score, out_proposal);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? BP_OK : bp_fail(BP_ERR_CUDA, cudaGetErrorString(err));
}
