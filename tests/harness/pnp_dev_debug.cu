// bring-up aid: 5-point EPnP hypotheses on device vs host for one planted pose
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../betapose_b200/csrc/pnp_math.cuh"
using namespace bp::pnp;
#ifndef NTHREADS
#define NTHREADS 64
#endif

__global__ void k(const double* pw, const double* uv, const unsigned char* sel, int K, double fx, double fy, double cx, double cy, unsigned seed, int* cnt, double* Rt, int* okf) {
  int h = threadIdx.x;
  int pool[64];
  for (int j = 0; j < K; ++j) pool[j] = j;
  sample_subset(pool, K, h, seed, 5);
  double R[9], t[3];
  bool ok = epnp(pw, uv, pool, 5, fx, fy, cx, cy, R, t);
  okf[h] = ok;
  int c = -1; double tot = 0;
  if (ok) score_hypothesis(R, t, pw, uv, sel, K, fx, fy, cx, cy, 144.0, &c, &tot);
  cnt[h] = c;
  for (int i = 0; i < 9; ++i) Rt[h * 12 + i] = R[i];
  for (int i = 0; i < 3; ++i) Rt[h * 12 + 9 + i] = t[i];
}

int main() {
  const int K = 50;
  std::vector<double> pw(K * 3), uv(K * 2);
  std::vector<unsigned char> sel(K, 1);
  unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) / 16777216.0; };
  for (int i = 0; i < K * 3; ++i) pw[i] = (rnd() - 0.5) * 0.08;
  double th = 0.7, R[9] = {cos(th), -sin(th), 0, sin(th), cos(th), 0, 0, 0, 1}, t[3] = {0.05, -0.03, 0.9};
  const double fx = 572.4114, fy = 573.57043, cx = 325.2611, cy = 242.04899;
  for (int i = 0; i < K; ++i) { double u, v, z; project(R, t, &pw[3 * i], fx, fy, cx, cy, &u, &v, &z); uv[2 * i] = (float)u; uv[2 * i + 1] = (float)v; }
  double *dpw, *duv, *dRt; unsigned char* dsel; int *dcnt, *dok;
  cudaMalloc(&dpw, K * 24); cudaMalloc(&duv, K * 16); cudaMalloc(&dsel, K); cudaMalloc(&dcnt, 64 * 4); cudaMalloc(&dok, 64 * 4); cudaMalloc(&dRt, 64 * 12 * 8);
  cudaMemcpy(dpw, pw.data(), K * 24, cudaMemcpyHostToDevice); cudaMemcpy(duv, uv.data(), K * 16, cudaMemcpyHostToDevice); cudaMemcpy(dsel, sel.data(), K, cudaMemcpyHostToDevice);
  k<<<1, NTHREADS>>>(dpw, duv, dsel, K, fx, fy, cx, cy, 11, dcnt, dRt, dok);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  int cnt[64], okf[64]; double Rt[64 * 12];
  cudaMemcpy(cnt, dcnt, sizeof cnt, cudaMemcpyDeviceToHost); cudaMemcpy(okf, dok, sizeof okf, cudaMemcpyDeviceToHost); cudaMemcpy(Rt, dRt, sizeof Rt, cudaMemcpyDeviceToHost);
  for (int h = 0; h < 8; ++h) {
    int pool[64]; for (int j = 0; j < K; ++j) pool[j] = j;
    sample_subset(pool, K, h, 11, 5);
    double Rh[9], thh[3]; bool ok = epnp(pw.data(), uv.data(), pool, 5, fx, fy, cx, cy, Rh, thh);
    int c = -1; double tot = 0; if (ok) score_hypothesis(Rh, thh, pw.data(), uv.data(), sel.data(), K, fx, fy, cx, cy, 144.0, &c, &tot);
    printf("h%d host ok=%d cnt=%d R0=%.4f t=(%.4f %.4f %.4f) | dev ok=%d cnt=%d R0=%.4f t=(%.4f %.4f %.4f)\n", h, ok, c, Rh[0], thh[0], thh[1], thh[2], okf[h], cnt[h], Rt[h * 12], Rt[h * 12 + 9], Rt[h * 12 + 10], Rt[h * 12 + 11]);
  }
  return 0;
}
