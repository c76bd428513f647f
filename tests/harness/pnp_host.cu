// TEST HARNESS: the PnP math of betapose_b200/csrc/pnp_math.cuh compiled for the HOST so the CPU test-suite can
// check it against oracle/pnp.py without a GPU.  Mirrors the control flow of pose_pnp_kernel (stages B-D) with a
// sequential loop over hypotheses.  Not part of libbetapose_b200.so and never used by the product.
#include <cmath>
#include <cstring>

#include "../../betapose_b200/csrc/pnp_math.cuh"

extern "C" int bp_host_pnp(const double* pw, const double* uv, const unsigned char* sel, int K, const double* cam, int mode,
                           double thr, int n_hyp, unsigned seed, double* R_out, double* t_out, unsigned char* inl,
                           int* best_h) {
  using namespace bp::pnp;
  const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
  int nsel = 0;
  for (int j = 0; j < K; ++j) nsel += sel[j] ? 1 : 0;
  memset(inl, 0, K);
  *best_h = -1;
  if (nsel < 4) return -1;
  const bool ransac = mode == 0 && nsel >= 6;
  double bR[9], bt[3];
  int bc = -1;
  double btot = INFINITY;
  int pool[64];
  if (ransac) {
    for (int h = 0; h < n_hyp; ++h) {
      int m = 0;
      for (int j = 0; j < K; ++j)
        if (sel[j]) pool[m++] = j;
      sample_subset(pool, m, h, seed, 5);
      double R[9], t[3];
      if (!epnp(pw, uv, pool, 5, fx, fy, cx, cy, R, t)) continue;
      int cnt;
      double tot;
      score_hypothesis(R, t, pw, uv, sel, K, fx, fy, cx, cy, thr * thr, &cnt, &tot);
      if (cnt > bc || (cnt == bc && tot < btot)) {
        bc = cnt; btot = tot; *best_h = h;
        memcpy(bR, R, sizeof bR); memcpy(bt, t, sizeof bt);
      }
    }
    if (bc < 4) {  // no consensus: failure, with the best hypothesis as the last estimate (as the kernel and the oracle do)
      if (bc >= 0) { memcpy(R_out, bR, sizeof bR); memcpy(t_out, bt, sizeof bt); }
      return -1;
    }
  } else {
    int m = 0;
    for (int j = 0; j < K; ++j)
      if (sel[j]) pool[m++] = j;
    if (!epnp(pw, uv, pool, m, fx, fy, cx, cy, bR, bt)) return -1;
    *best_h = 0;
  }
  SingleLane ln;
  for (int round = 0; round < BP_PNP_LO_ROUNDS; ++round) {
    int changed = 0, cnt = 0;
    for (int j = 0; j < K; ++j) {
      const unsigned char in = sel[j] && (!ransac || within_threshold(bR, bt, pw, uv, j, fx, fy, cx, cy, thr * thr)) ? 1 : 0;
      changed |= in != inl[j];
      inl[j] = in;
      cnt += in;
    }
    if (round > 0 && !changed) break;
    if (cnt < 4) {
      memset(inl, 0, K);
      memcpy(R_out, bR, sizeof bR);
      memcpy(t_out, bt, sizeof bt);
      return -1;
    }
    lm_refine(ln, bR, bt, pw, uv, inl, K, fx, fy, cx, cy, BP_PNP_LM_ITERS);
    if (!ransac) break;
  }
  memcpy(R_out, bR, sizeof bR);
  memcpy(t_out, bt, sizeof bt);
  return 0;
}

// EPnP alone (all selected points) for unit tests
extern "C" int bp_host_epnp(const double* pw, const double* uv, int K, const double* cam, double* R_out, double* t_out) {
  int pool[64];
  for (int j = 0; j < K; ++j) pool[j] = j;
  return bp::pnp::epnp(pw, uv, pool, K, cam[0], cam[1], cam[2], cam[3], R_out, t_out) ? 0 : -1;
}

// EPnP with the cyclic-Jacobi policy (what the device's small-batch path does cooperatively) for comparison
extern "C" int bp_host_epnp_jacobi(const double* pw, const double* uv, int K, const double* cam, double* R_out, double* t_out) {
  int pool[64];
  for (int j = 0; j < K; ++j) pool[j] = j;
  bp::pnp::SerialEig12 eig;
  return bp::pnp::epnp(eig, pw, uv, pool, K, cam[0], cam[1], cam[2], cam[3], R_out, t_out) ? 0 : -1;
}

// 12 x 12 symmetric eigen-solvers alone: which = 0 Householder + QL, 1 cyclic Jacobi.  a[144] in; vectors (columns) and w[12] out
extern "C" void bp_host_sym_eig12(const double* a, int which, double* vec, double* w) {
  double A[144];
  for (int i = 0; i < 144; ++i) A[i] = a[i];
  if (which == 0) {
    bp::pnp::sym_eig_ql<12>(A, w);
    for (int i = 0; i < 144; ++i) vec[i] = A[i];
  } else {
    bp::pnp::jacobi_eig<12>(A, vec, w);
  }
}
