// Stand-alone bring-up harness for conv_umma_kernel (no torch, no python): checks the tcgen05/TMA conv against a
// naive fp32 GPU reference over the shape/epilogue matrix the two networks need, probes the im2col TMA
// semantics, and times a few production shapes.  Build: see betapose_b200/csrc/Makefile (target conv_harness).
//   ./conv_harness            -> correctness matrix + timings
//   ./conv_harness probe      -> im2col probe only
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../betapose_b200/csrc/conv_plan.cuh"

using namespace bp;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

// ------------------------------------------------------------------ naive reference
__global__ void ref_conv_kernel(const __half* x, int N, int H, int W, int C, int xp, const __half* w, int wp,
                                const float* bias, int R, int S, int stride, int pad, int P, int Q, int Cout, int act,
                                const __half* res, int rp, int res_mode, float* out) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)N * P * Q * Cout;
  if (idx >= total) return;
  const int co = idx % Cout;
  const long m = idx / Cout;
  const int q = m % Q;
  const int p = (m / Q) % P;
  const int n = m / ((long)P * Q);
  float acc = 0.f;
  for (int r = 0; r < R; ++r) {
    const int h = p * stride - pad + r;
    if (h < 0 || h >= H) continue;
    for (int s = 0; s < S; ++s) {
      const int ww = q * stride - pad + s;
      if (ww < 0 || ww >= W) continue;
      const __half* xr = x + (((long)n * H + h) * W + ww) * xp;
      const __half* wr = w + (long)co * wp + (r * S + s) * C;
      for (int c = 0; c < C; ++c) acc += __half2float(xr[c]) * __half2float(wr[c]);
    }
  }
  float v = acc + bias[co];
  float rv = res ? __half2float(res[m * rp + co]) : 0.f;
  if (res && res_mode == RES_BEFORE_ACT) v += rv;
  v = act == ACT_LEAKY ? (v > 0 ? v : 0.1f * v) : act == ACT_RELU ? fmaxf(v, 0.f) : act == ACT_SIGMOID ? 1.f / (1.f + expf(-v)) : v;
  if (res && res_mode == RES_AFTER_ACT) v += rv;
  out[idx] = v;
}

// ------------------------------------------------------------------ im2col probe
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, int c, int w, int h, int n, int offw, int offh,
                             __half* out, int bytes) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  uint8_t* s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, bytes);
    tma_load_im2col_4d(&tm, &bar, s, c, w, h, n, (uint16_t)offw, (uint16_t)offh);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<__half*>(s)[i];
}

struct Rng {  // xorshift32: fast enough to fill the batch-64 tensors
  uint32_t s;
  explicit Rng(unsigned seed) : s(seed ? seed : 1u) {}
  float uni(float a, float b) {
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    return a + (b - a) * ((s >> 8) * (1.0f / 16777216.0f));
  }
};

static TmapApi g_api;
static bool g_trace = false;  // `trace` mode: per-CTA clock stamps of the production shapes

struct Case {
  const char* name;
  int N, H, W, C, Cout, R, S, stride, pad;
  int act = ACT_LEAKY, res_mode = RES_NONE, store_mode = STORE_PLAIN, out_f32 = 0;
  int x_pitch_extra = 0, out_pitch_extra = 0, out_coff = 0;
  int force_bn = 0, force_st = 0;
  int iters = 0;  // >0: also time it
  int gather = 0;  // unused (kept so the positional initialisers below stay valid)
  int cg = 0;      // 0 = planner's choice, 1 = single CTA, 2 = CTA pairs (cta_group::2)
  int mt = 0;      // 0 = planner's choice, 1 = 128-pixel tiles, 2 = 256-pixel tiles
};

static int run_case(const Case& cs) {
  Rng rng(1234);
  const int xp = cs.C + cs.x_pitch_extra;
  const int K = cs.R * cs.S * cs.C;
  const int wp = (K + 7) / 8 * 8;
  const int Cout_pad = (cs.Cout + 255) / 256 * 256;
  const int P = (cs.H + 2 * cs.pad - cs.R) / cs.stride + 1, Q = (cs.W + 2 * cs.pad - cs.S) / cs.stride + 1;
  const long M = (long)cs.N * P * Q;
  std::vector<__half> hx((size_t)cs.N * cs.H * cs.W * xp), hw((size_t)Cout_pad * wp, __float2half(0.f));
  std::vector<float> hb(Cout_pad, 0.f);
  for (auto& v : hx) v = __float2half(rng.uni(-1.f, 1.f));
  const float ws = 1.0f / sqrtf((float)K);
  for (int co = 0; co < cs.Cout; ++co)
    for (int k = 0; k < K; ++k) hw[(size_t)co * wp + k] = __float2half(rng.uni(-1.f, 1.f) * ws * 1.7f);
  for (int co = 0; co < cs.Cout; ++co) hb[co] = rng.uni(-0.5f, 0.5f);
  std::vector<__half> hres;
  if (cs.res_mode != RES_NONE) {
    hres.resize((size_t)M * cs.Cout);
    for (auto& v : hres) v = __float2half(rng.uni(-1.f, 1.f));
  }
  // output geometry
  const int up = cs.store_mode != STORE_PLAIN ? 2 : 1;
  const int oc = cs.store_mode == STORE_PIXSHUF2 ? cs.Cout / 4 : cs.Cout;
  const int opitch = oc + cs.out_coff + cs.out_pitch_extra;
  const size_t out_elems = (size_t)cs.N * P * up * Q * up * opitch;
  const size_t esz = cs.out_f32 ? 4 : 2;

  __half *dx, *dw, *dres = nullptr;
  float *db, *dref;
  void* dout;
  CK(cudaMalloc(&dx, hx.size() * 2));
  CK(cudaMalloc(&dw, hw.size() * 2));
  CK(cudaMalloc(&db, hb.size() * 4));
  CK(cudaMalloc(&dref, (size_t)M * cs.Cout * 4));
  CK(cudaMalloc(&dout, out_elems * esz));
  CK(cudaMemset(dout, 0xFF, out_elems * esz));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  if (!hres.empty()) {
    CK(cudaMalloc(&dres, hres.size() * 2));
    CK(cudaMemcpy(dres, hres.data(), hres.size() * 2, cudaMemcpyHostToDevice));
  }

  ConvDesc d;
  d.x = dx; d.N = cs.N; d.H = cs.H; d.W = cs.W; d.C = cs.C; d.x_pitch = xp;
  d.w = dw; d.bias = db; d.w_pitch = wp; d.Cout = cs.Cout; d.Cout_pad = Cout_pad;
  d.R = cs.R; d.S = cs.S; d.stride = cs.stride; d.pad = cs.pad; d.act = cs.act;
  d.res = dres; d.res_pitch = cs.Cout; d.res_mode = cs.res_mode;
  d.out = dout; d.out_pitch = opitch; d.out_coff = cs.out_coff; d.out_f32 = cs.out_f32; d.store_mode = cs.store_mode;
  d.force_block_n = cs.force_bn; d.force_stages = cs.force_st; d.force_cg = cs.cg; d.force_mt = cs.mt;
  ConvPlan pl;
  std::string err;
  if (!conv_plan_build(g_api, &pl, d, &err)) {
    printf("[%-28s] PLAN FAILED: %s\n", cs.name, err.c_str());
    return 1;
  }
  cudaError_t le = conv_plan_launch(pl, 0);
  cudaError_t se = cudaDeviceSynchronize();
  if (le != cudaSuccess || se != cudaSuccess) {
    printf("[%-28s] LAUNCH/RUN FAILED: %s / %s (bn=%d bk=%d st=%d grid=%d)\n", cs.name, cudaGetErrorString(le),
           cudaGetErrorString(se), pl.block_n, pl.block_k, pl.stages, pl.grid);
    exit(3);  // context is likely poisoned
  }
  {
    const long total = M * cs.Cout;
    ref_conv_kernel<<<(unsigned)((total + 255) / 256), 256>>>(dx, cs.N, cs.H, cs.W, cs.C, xp, dw, wp, db, cs.R, cs.S,
                                                             cs.stride, cs.pad, P, Q, cs.Cout, cs.act, dres, cs.Cout,
                                                             cs.res_mode, dref);
    CK(cudaDeviceSynchronize());
  }
  std::vector<float> href((size_t)M * cs.Cout);
  CK(cudaMemcpy(href.data(), dref, href.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<uint8_t> hout(out_elems * esz);
  CK(cudaMemcpy(hout.data(), dout, hout.size(), cudaMemcpyDeviceToHost));
  auto get = [&](size_t i) -> float {
    return cs.out_f32 ? reinterpret_cast<float*>(hout.data())[i] : __half2float(reinterpret_cast<__half*>(hout.data())[i]);
  };
  double max_err = 0, max_ref = 0;
  long bad = 0, first_bad = -1;
  for (long m = 0; m < M; ++m) {
    const int q = m % Q, p = (m / Q) % P, n = m / ((long)P * Q);
    for (int co = 0; co < cs.Cout; ++co) {
      const float r = href[(size_t)m * cs.Cout + co];
      int ndst = 1;
      size_t dst[4];
      if (cs.store_mode == STORE_PLAIN) {
        dst[0] = (size_t)m * opitch + cs.out_coff + co;
      } else if (cs.store_mode == STORE_UPSAMPLE2) {
        ndst = 4;
        for (int k = 0; k < 4; ++k)
          dst[k] = (((size_t)n * 2 * P + 2 * p + (k >> 1)) * 2 * Q + 2 * q + (k & 1)) * opitch + cs.out_coff + co;
      } else {
        // harness weights are NOT permuted: channel index co is already "o' = sub*C4 + c"
        const int c4 = cs.Cout / 4, sub = co / c4, cc = co % c4;
        dst[0] = (((size_t)n * 2 * P + 2 * p + (sub >> 1)) * 2 * Q + 2 * q + (sub & 1)) * opitch + cs.out_coff + cc;
      }
      for (int k = 0; k < ndst; ++k) {
        const float g = get(dst[k]);
        const double e = fabs((double)g - r);
        const double tol = 2e-2 + 4e-3 * fabs(r);
        if (!(e <= tol)) {
          if (first_bad < 0) first_bad = m * cs.Cout + co;
          ++bad;
        }
        if (e > max_err || e != e) max_err = e;
        if (fabs(r) > max_ref) max_ref = fabs(r);
      }
    }
  }
  double ms = 0;
  if (cs.iters > 0 && bad == 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) conv_plan_launch(pl, 0);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < cs.iters; ++i) conv_plan_launch(pl, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float t;
    CK(cudaEventElapsedTime(&t, e0, e1));
    ms = t / cs.iters;
  }
  printf("[%-28s] bn=%3d bk=%2d st=%2d cg=%d mt=%d grid=%6d M=%7ld K=%5d  max_err=%.4g (max|ref|=%.3g) bad=%ld%s", cs.name,
         pl.block_n, pl.block_k, pl.stages, pl.cg, pl.mt, pl.grid, M, K, max_err, max_ref, bad, bad ? "  <-- FAIL" : "  ok");
  if (bad) printf(" first_bad: m=%ld co=%ld", first_bad / cs.Cout, first_bad % cs.Cout);
  if (ms > 0) printf("  %.3f ms  %.1f TFLOP/s", ms, pl.flops / ms * 1e-9);
  printf("\n");
  if (g_trace && bad == 0) {
    // one launch with per-CTA clock stamps (ConvArgs::trace): where a CTA's time goes, tile by tile
    unsigned long long* dtr;
    const size_t nslots = (size_t)pl.grid * kTraceSlots;
    CK(cudaMalloc(&dtr, nslots * 8));
    CK(cudaMemset(dtr, 0, nslots * 8));
    ConvPlan pt = pl;
    pt.args.trace = dtr;
    conv_plan_launch(pt, 0);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> tr(nslots);
    CK(cudaMemcpy(tr.data(), dtr, nslots * 8, cudaMemcpyDeviceToHost));
    cudaFree(dtr);
    double tot = 0, totmax = 0, fill = 0, main_ = 0, epi = 0, period = 0, idle = 0, first = 0;
    long nt = 0, np = 0, nc = 0, maxtiles = 0;
    for (int c = 0; c < pl.grid; ++c) {
      const unsigned long long* t = tr.data() + (size_t)c * kTraceSlots;
      if (!t[1] || !t[0]) continue;
      if (pl.cg == 2 && (c & 1)) {  // the peer CTA of a pair has no MMA stamps: epilogue only
        continue;
      }
      ++nc;
      tot += double(t[1] - t[0]);
      totmax = std::max(totmax, double(t[1] - t[0]));
      long tiles = 0;
      for (int i = 0; i < kTraceTiles; ++i) {
        const unsigned long long *q = t + 2 + 4 * i;
        if (!q[2] || !q[3]) break;
        ++tiles;
        if (i == 0) first += double(q[1] - t[0]);   // start -> first k-block landed
        fill += double(q[1] - q[0]);               // accumulator free -> first k-block of the tile landed
        main_ += double(q[2] - q[1]);              // first k-block landed -> accumulator complete
        epi += double(q[3] - q[2]);                // accumulator complete -> epilogue done
        ++nt;
        if (i > 0) {
          period += double(q[2] - (q - 4)[2]);     // accumulator-complete to accumulator-complete
          idle += q[0] > (q - 4)[2] ? double(q[0] - (q - 4)[2]) : 0.0;  // MMA warp waited for the epilogue to free an accumulator
          ++np;
        }
      }
      maxtiles = std::max(maxtiles, tiles);
    }
    {  // stamps inside the first chunks of the first tile: an epilogue thread (thread 64) and the store warp
      for (int k = 0; k < 5; ++k) {
        double d[8] = {};
        long cnt = 0;
        for (int c = 0; c < pl.grid; ++c) {
          const unsigned long long* t = tr.data() + (size_t)c * kTraceSlots + kTraceFine;
          const unsigned long long *q = t + 5 * k, *w = t + 25 + 3 * k;
          if (!q[0] || !q[4] || !w[0] || !w[2]) continue;
          ++cnt;
          for (int e = 0; e < 4; ++e) d[e] += double(q[e + 1] - q[e]);
          if (k > 0 && (q - 5)[4]) d[4] += double(q[0] - (q - 5)[4]);
          d[5] += double((long long)(w[0] - q[4]));  // arrive -> store warp released
          d[6] += double(w[1] - w[0]);
          d[7] += double(w[2] - w[1]);
        }
        if (cnt)
          printf("    chunk %d of tile 0 (%ld CTAs): epilogue thread: loop top %.0f | TMEM wait %.0f | next load + buffer wait %.0f | math + st.shared %.0f | fence + arrive %.0f"
                 " || store warp: released %.0f after that arrive | store + commit %.0f | wait_group.read + hand-on %.0f\n",
                 k, cnt, d[4] / cnt, d[0] / cnt, d[1] / cnt, d[2] / cnt, d[3] / cnt, d[5] / cnt, d[6] / cnt, d[7] / cnt);
      }
    }
    if (nc && nt)
      printf("    trace: CTAs %ld, tiles/CTA max %ld | CTA active cycles mean %.0f max %.0f | start->first k-block %.0f | per tile: wait for first k-block %.0f, "
             "main loop %.0f (%.0f per k-block), epilogue %.0f | tile period %.0f | MMA warp blocked on epilogue %.0f\n",
             nc, maxtiles, tot / nc, totmax, first / nc, fill / nt, main_ / nt, main_ / nt / (double)pl.args.num_kb, epi / nt, np ? period / np : 0.0,
             np ? idle / np : 0.0);
  }
  fflush(stdout);
  cudaFree(dx); cudaFree(dw); cudaFree(db); cudaFree(dref); cudaFree(dout);
  if (dres) cudaFree(dres);
  return bad ? 1 : 0;
}

// Stem convolution exactly as betapose_b200/csrc/net.cu builds it: input fp16 [N, H, W + 8, 8] (data from column 3,
// channels 3..7 and pad columns zero), a k x 1 convolution over "virtual pixels" of Cv channels whose pixel stride
// (16 B) is smaller than their extent; weights [Cout][r][q*8 + c].  Reference: the naive kernel on a dense C = 8 copy.
static int run_stem_case(const char* name, int N, int H, int W, int k, int stride, int pad, int Cout, int act, int iters, int mt = 0) {
  Rng rng(4321);
  const int PADL = 3, PADC = 8, Wp = W + PADC;
  const int Cv = k * 8 <= 32 ? 32 : 64;
  const int K = k * Cv, wp = K;
  const int Cout_pad = (Cout + 255) / 256 * 256;
  const int P = (H + 2 * pad - k) / stride + 1, Q = (W + 2 * pad - k) / stride + 1;
  const long M = (long)N * P * Q;
  std::vector<__half> hx((size_t)N * H * Wp * 8, __float2half(0.f)), hd((size_t)N * H * W * 8, __float2half(0.f));
  for (int n = 0; n < N; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w)
        for (int c = 0; c < 3; ++c) {
          const __half v = __float2half((float)(int)rng.uni(0.f, 255.99f));
          hx[(((size_t)n * H + h) * Wp + PADL + w) * 8 + c] = v;
          hd[(((size_t)n * H + h) * W + w) * 8 + c] = v;
        }
  std::vector<__half> hw((size_t)Cout_pad * wp, __float2half(0.f)), hwd((size_t)Cout_pad * k * k * 8, __float2half(0.f));
  std::vector<float> hb(Cout_pad, 0.f);
  const float ws = 1.0f / (255.f * sqrtf((float)(k * k * 3)));
  for (int co = 0; co < Cout; ++co) {
    for (int r = 0; r < k; ++r)
      for (int q = 0; q < k; ++q)
        for (int c = 0; c < 3; ++c) {
          const __half v = __float2half(rng.uni(-1.f, 1.f) * ws * 3.4f);
          hw[(size_t)co * wp + r * Cv + q * 8 + c] = v;
          hwd[(size_t)co * k * k * 8 + (r * k + q) * 8 + c] = v;
        }
    hb[co] = rng.uni(-0.5f, 0.5f);
  }
  __half *dx, *dd, *dw, *dwd, *dout;
  float *db, *dref;
  CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&dd, hd.size() * 2));
  CK(cudaMalloc(&dw, hw.size() * 2)); CK(cudaMalloc(&dwd, hwd.size() * 2));
  CK(cudaMalloc(&db, hb.size() * 4)); CK(cudaMalloc(&dref, (size_t)M * Cout * 4)); CK(cudaMalloc(&dout, (size_t)M * Cout * 2));
  CK(cudaMemset(dout, 0xFF, (size_t)M * Cout * 2));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dd, hd.data(), hd.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dwd, hwd.data(), hwd.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));

  ConvDesc d;
  d.x = dx + (size_t)(PADL - pad) * 8;
  d.N = N; d.H = H; d.W = (Q - 1) * stride + 1; d.C = Cv; d.x_pitch = 8;
  d.x_row_pitch = (long)Wp * 8; d.x_img_pitch = (long)H * Wp * 8;
  d.R = k; d.S = 1; d.stride = stride; d.pad = pad; d.pad_w = 0; d.real_k = k * k * 3;
  d.w = dw; d.bias = db; d.w_pitch = wp; d.Cout = Cout; d.Cout_pad = Cout_pad; d.act = act;
  d.out = dout; d.out_pitch = Cout; d.force_mt = mt;
  ConvPlan pl;
  std::string err;
  if (!conv_plan_build(g_api, &pl, d, &err)) {
    printf("[%-28s] PLAN FAILED: %s\n", name, err.c_str());
    return 1;
  }
  cudaError_t le = conv_plan_launch(pl, 0);
  cudaError_t se = cudaDeviceSynchronize();
  if (le != cudaSuccess || se != cudaSuccess) {
    printf("[%-28s] LAUNCH/RUN FAILED: %s / %s\n", name, cudaGetErrorString(le), cudaGetErrorString(se));
    exit(3);
  }
  const long total = M * Cout;
  ref_conv_kernel<<<(unsigned)((total + 255) / 256), 256>>>(dd, N, H, W, 8, 8, dwd, k * k * 8, db, k, k, stride, pad, P, Q, Cout,
                                                           act, nullptr, 0, RES_NONE, dref);
  CK(cudaDeviceSynchronize());
  std::vector<float> href((size_t)total);
  std::vector<__half> hout((size_t)total);
  CK(cudaMemcpy(href.data(), dref, href.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hout.data(), dout, hout.size() * 2, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0;
  long bad = 0, first_bad = -1;
  for (long i = 0; i < total; ++i) {
    const double r = href[i], g = __half2float(hout[i]), e = fabs(g - r);
    if (!(e <= 2e-2 + 4e-3 * fabs(r))) {
      if (first_bad < 0) first_bad = i;
      ++bad;
    }
    if (e > max_err || e != e) max_err = e;
    if (fabs(r) > max_ref) max_ref = fabs(r);
  }
  double ms = 0;
  if (iters > 0 && bad == 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) conv_plan_launch(pl, 0);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) conv_plan_launch(pl, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float t;
    CK(cudaEventElapsedTime(&t, e0, e1));
    ms = t / iters;
  }
  printf("[%-28s] bn=%3d bk=%2d st=%2d cg=%d mt=%d grid=%6d M=%7ld K=%5d  max_err=%.4g (max|ref|=%.3g) bad=%ld%s", name, pl.block_n,
         pl.block_k, pl.stages, pl.cg, pl.mt, pl.grid, M, K, max_err, max_ref, bad, bad ? "  <-- FAIL" : "  ok");
  if (bad) printf(" first_bad: m=%ld co=%ld", first_bad / Cout, first_bad % Cout);
  if (ms > 0) printf("  %.3f ms  %.1f GB/s out", ms, (double)total * 2 / ms * 1e-6);
  printf("\n");
  fflush(stdout);
  cudaFree(dx); cudaFree(dd); cudaFree(dw); cudaFree(dwd); cudaFree(db); cudaFree(dref); cudaFree(dout);
  return bad ? 1 : 0;
}

// Dump what one im2col TMA load actually fetches, against the expected gather, to pin the coordinate semantics.
static int run_probe(int N, int H, int W, int C, int R, int S, int stride, int pad, int m0, int fr, int fs) {
  const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
  std::vector<__half> hx((size_t)N * H * W * C);
  // value encodes (n,h,w) exactly in fp16 range: v = n*512 + h*20 + w  (small dims), channel adds 0
  for (int n = 0; n < N; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w)
        for (int c = 0; c < C; ++c) hx[(((size_t)n * H + h) * W + w) * C + c] = __float2half(float(n * 512 + h * 20 + w + 1));
  __half *dx, *dout;
  CK(cudaMalloc(&dx, hx.size() * 2));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  const int bk = 64;
  const int bytes = 128 * bk * 2;
  CK(cudaMalloc(&dout, bytes));
  alignas(64) CUtensorMap tm;
  std::string err;
  if (!make_tmap_im2col(g_api, &tm, dx, N, H, W, C, C, R, S, stride, pad, pad, bk, &err)) {
    printf("probe: tmap failed %s\n", err.c_str());
    return 1;
  }
  const int img = m0 / (P * Q), rem = m0 % (P * Q), op = rem / Q, oq = rem % Q;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 1024));
  probe_kernel<<<1, 128, bytes + 1024>>>(tm, 0, oq * stride - pad, op * stride - pad, img, fs, fr, dout, bytes);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("probe: kernel failed: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  std::vector<__half> ho(bytes / 2);
  CK(cudaMemcpy(ho.data(), dout, bytes, cudaMemcpyDeviceToHost));
  int bad = 0;
  printf("probe N=%d H=%d W=%d R=%d S=%d stride=%d pad=%d P=%d Q=%d m0=%d tap=(%d,%d)\n", N, H, W, R, S, stride, pad,
         P, Q, m0, fr, fs);
  for (int i = 0; i < 128; ++i) {
    const int m = m0 + i;
    float expect = 0.f;
    if (m < N * P * Q) {
      const int n = m / (P * Q), r2 = m % (P * Q), p = r2 / Q, q = r2 % Q;
      const int h = p * stride - pad + fr, w = q * stride - pad + fs;
      if (h >= 0 && h < H && w >= 0 && w < W) expect = float(n * 512 + h * 20 + w + 1);
    }
    // 128B swizzle: 16B chunk index XOR (row & 7); read chunk 0 of the logical row
    const int chunk = 0 ^ (i & 7);
    const float got = __half2float(ho[(size_t)i * 64 + chunk * 8]);
    if (got != expect) {
      if (bad < 24) printf("  row %3d: got %6.0f expect %6.0f\n", i, got, expect);
      ++bad;
    }
  }
  printf("probe result: %d/128 rows mismatched\n", bad);
  cudaFree(dx);
  cudaFree(dout);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  std::string err;
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  if (!g_api.load(&err)) {
    printf("tmap api: %s\n", err.c_str());
    return 2;
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs, driver %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount,
         g_api.driver_version);
  if (argc > 1 && !strcmp(argv[1], "trace")) {
    // batch-64 production shapes with per-tile clock stamps: which of {operand arrival, main loop, epilogue} a tile waits for
    g_trace = true;
    std::vector<Case> tc = {
        {"K conv1 1x1 1024->256 @20x16", 64, 20, 16, 1024, 256, 1, 1, 1, 0, ACT_RELU, 0, 0, 0, 0, 0, 0, 0, 0, 10},
        {"K conv2 3x3 256->256 @20x16", 64, 20, 16, 256, 256, 3, 3, 1, 1, ACT_RELU, 0, 0, 0, 0, 0, 0, 0, 0, 10},
        {"K conv3 1x1 256->1024 +res", 64, 20, 16, 256, 1024, 1, 1, 1, 0, ACT_RELU, RES_BEFORE_ACT, 0, 0, 0, 0, 0, 0, 0, 10},
        {"K conv3 1x1 256->1024 no res", 64, 20, 16, 256, 1024, 1, 1, 1, 0, ACT_RELU, 0, 0, 0, 0, 0, 0, 0, 0, 10},
        {"Y 1x1 512->256 @26", 64, 26, 26, 512, 256, 1, 1, 1, 0, ACT_LEAKY, 0, 0, 0, 0, 0, 0, 0, 0, 10},
        {"Y 3x3 256->512 @26 +res", 64, 26, 26, 256, 512, 3, 3, 1, 1, ACT_LEAKY, RES_AFTER_ACT, 0, 0, 0, 0, 0, 0, 0, 10},
        {"Y 3x3 128->256 @52 +res", 64, 52, 52, 128, 256, 3, 3, 1, 1, ACT_LEAKY, RES_AFTER_ACT, 0, 0, 0, 0, 0, 0, 0, 10},
        {"Y 1x1 1024->512 @13", 64, 13, 13, 1024, 512, 1, 1, 1, 0, ACT_LEAKY, 0, 0, 0, 0, 0, 0, 0, 0, 10},
        {"Y 3x3 512->1024 @13 +res", 64, 13, 13, 512, 1024, 3, 3, 1, 1, ACT_LEAKY, RES_AFTER_ACT, 0, 0, 0, 0, 0, 0, 0, 10},
        {"Y 1x1 256->128 @52", 64, 52, 52, 256, 128, 1, 1, 1, 0, ACT_LEAKY, 0, 0, 0, 0, 0, 0, 0, 0, 10},
        {"Y head 1x1 256->18 f32 @52", 64, 52, 52, 256, 18, 1, 1, 1, 0, ACT_NONE, 0, 0, 1, 0, 2, 0, 0, 0, 10},
        {"Y 1x1 512->256 @13 up2", 64, 13, 13, 512, 256, 1, 1, 1, 0, ACT_LEAKY, 0, STORE_UPSAMPLE2, 0, 0, 512, 0, 0, 0, 10},
        {"Y 1x1 256->128 @26 up2", 64, 26, 26, 256, 128, 1, 1, 1, 0, ACT_LEAKY, 0, STORE_UPSAMPLE2, 0, 0, 256, 0, 0, 0, 10},
        {"K 3x3 128->50 f32 @80x64", 64, 80, 64, 128, 50, 3, 3, 1, 1, ACT_NONE, 0, 0, 1, 0, 2, 0, 0, 0, 10},
        {"K 3x3 256->512 @40x32 ps2", 64, 40, 32, 256, 512, 3, 3, 1, 1, ACT_RELU, 0, STORE_PIXSHUF2, 0, 0, 0, 0, 0, 0, 10},
    };
    int f = 0;
    for (auto& c : tc) f += run_case(c);
    printf("TOTAL FAILURES: %d\n", f);
    return f ? 1 : 0;
  }
  const bool probe_only = argc > 1 && !strcmp(argv[1], "probe");
  const bool quick = argc > 1 && !strcmp(argv[1], "quick");
  int fails = 0;
  if (argc > 1 && !strcmp(argv[1], "small")) {
    // the small-tile, bandwidth-class layers only (profiling target)
    fails += run_stem_case("stem 3x3 s1 3->32 @416 B64", 64, 416, 416, 3, 1, 1, 32, ACT_LEAKY, 3);
    Case c1 = {"Y 3x3 32->64 s2 @416 bn64", 64, 416, 416, 32, 64, 3, 3, 2, 1, ACT_LEAKY, 0, 0, 0, 0, 0, 0, 64, 0, 3};
    Case c2 = {"Y 1x1 64->32 @208", 64, 208, 208, 64, 32, 1, 1, 1, 0, ACT_LEAKY, 0, 0, 0, 0, 0, 0, 0, 0, 3};
    fails += run_case(c1);
    fails += run_case(c2);
    return fails ? 1 : 0;
  }

  if (!probe_only) {
    // 1) plain GEMM path (2-D TMA only) first: isolates descriptor/idesc/TMEM plumbing from im2col
    std::vector<Case> gemm = {
        {"gemm 1x1 C64->128 bn128", 1, 1, 1000, 64, 128, 1, 1, 1, 0},
        {"gemm 1x1 C256->256 bn256", 2, 20, 16, 256, 256, 1, 1, 1, 0},
        {"gemm 1x1 C512->64", 1, 13, 13, 512, 64, 1, 1, 1, 0},
        {"gemm 1x1 C1024->18 f32", 2, 13, 13, 1024, 18, 1, 1, 1, 0, ACT_NONE, RES_NONE, STORE_PLAIN, 1, 0, 14},
        {"gemm K=27 ragged bk32", 1, 1, 3000, 27, 32, 1, 1, 1, 0, ACT_LEAKY, RES_NONE, STORE_PLAIN, 0, 5},
        {"gemm K=147 ragged bk64", 1, 1, 3000, 147, 64, 1, 1, 1, 0, ACT_RELU, RES_NONE, STORE_PLAIN, 0, 5},
        {"gemm 1x1 C32->64 bk32", 1, 16, 16, 32, 64, 1, 1, 1, 0},
        {"gemm pitch/coff", 1, 26, 26, 128, 256, 1, 1, 1, 0, ACT_LEAKY, RES_NONE, STORE_PLAIN, 0, 64, 128, 64},
        {"gemm res after-act", 1, 26, 26, 128, 256, 1, 1, 1, 0, ACT_LEAKY, RES_AFTER_ACT},
        {"gemm res before-act relu", 1, 20, 16, 256, 1024, 1, 1, 1, 0, ACT_RELU, RES_BEFORE_ACT},
        {"gemm sigmoid (SE fc)", 1, 1, 64, 256, 256, 1, 1, 1, 0, ACT_SIGMOID},
        {"gemm upsample2 store", 2, 13, 13, 512, 256, 1, 1, 1, 0, ACT_LEAKY, RES_NONE, STORE_UPSAMPLE2, 0, 0, 512, 0},
        {"gemm bn256 st2", 2, 20, 16, 256, 512, 1, 1, 1, 0, ACT_LEAKY, RES_NONE, STORE_PLAIN, 0, 0, 0, 0, 256, 2},
        {"gemm bn128 st6", 2, 20, 16, 256, 512, 1, 1, 1, 0, ACT_LEAKY, RES_NONE, STORE_PLAIN, 0, 0, 0, 0, 128, 6},
        {"gemm Cout32 chunk32 res", 3, 31, 29, 64, 32, 1, 1, 1, 0, ACT_LEAKY, RES_AFTER_ACT},
        {"gemm Cout96 ragged N res", 3, 31, 29, 64, 96, 1, 1, 1, 0, ACT_RELU, RES_BEFORE_ACT},
        {"gemm many tiles persistent", 40, 33, 31, 64, 256, 1, 1, 1, 0, ACT_LEAKY, RES_AFTER_ACT, STORE_PLAIN, 0, 0, 0, 0, 64},
        {"gemm many tiles bn256", 40, 33, 31, 128, 512, 1, 1, 1, 0, ACT_LEAKY, RES_BEFORE_ACT},
        {"gemm coff+res pitch", 2, 26, 26, 128, 256, 1, 1, 1, 0, ACT_LEAKY, RES_AFTER_ACT, STORE_PLAIN, 0, 64, 128, 64},
    };
    for (auto& c : gemm) fails += run_case(c);
  }

  // 2) im2col probes (cheap, very informative if the conv cases below fail)
  int pf = 0;
  pf += run_probe(2, 13, 13, 64, 3, 3, 1, 1, 0, 0, 0);
  pf += run_probe(2, 13, 13, 64, 3, 3, 1, 1, 128, 1, 2);
  pf += run_probe(2, 13, 13, 64, 3, 3, 1, 1, 256, 2, 2);
  pf += run_probe(2, 16, 16, 64, 3, 3, 2, 1, 0, 0, 0);
  pf += run_probe(2, 16, 16, 64, 3, 3, 2, 1, 0, 2, 1);
  pf += run_probe(3, 15, 11, 64, 1, 1, 2, 0, 0, 0, 0);
  printf("probe failures: %d\n", pf);
  if (probe_only) return pf ? 1 : 0;

  if (!probe_only) {
    // packed stem convolutions over the padded network-input layout [N, H, W + 8, 8] (overlapping im2col map)
    fails += run_stem_case("stem 3x3 s1 3->32 40x36", 3, 40, 36, 3, 1, 1, 32, ACT_LEAKY, 0);
    fails += run_stem_case("stem 7x7 s2 3->64 64x48", 3, 64, 48, 7, 2, 3, 64, ACT_RELU, 0);
    fails += run_stem_case("stem 3x3 s1 3->32 53x47", 2, 53, 47, 3, 1, 1, 32, ACT_LEAKY, 0);
    fails += run_stem_case("stem 3x3 s1 3->32 @416 B64", 64, 416, 416, 3, 1, 1, 32, ACT_LEAKY, 5);
    fails += run_stem_case("stem 7x7 s2 3->64 @320x256 B64", 64, 320, 256, 7, 2, 3, 64, ACT_RELU, 5);
    fails += run_stem_case("stem 3x3 3->32 53x47 mt4", 2, 53, 47, 3, 1, 1, 32, ACT_LEAKY, 0, 4);
    fails += run_stem_case("stem 3x3 3->32 @416 B64 mt4", 64, 416, 416, 3, 1, 1, 32, ACT_LEAKY, 5, 4);
  }

  std::vector<Case> conv = {
      {"3x3 s1 C64->128 13x13", 2, 13, 13, 64, 128, 3, 3, 1, 1},
      {"3x3 s1 C128->256 52x52", 1, 52, 52, 128, 256, 3, 3, 1, 1},
      {"3x3 s2 C128->256 26x26", 2, 26, 26, 128, 256, 3, 3, 2, 1},
      {"3x3 s2 C32->64 bk32", 1, 32, 32, 32, 64, 3, 3, 2, 1},
      {"3x3 s1 C32->64 bk32", 1, 24, 24, 32, 64, 3, 3, 1, 1},
      {"1x1 s2 C256->512 40x32", 2, 40, 32, 256, 512, 1, 1, 2, 0, ACT_NONE},
      {"3x3 s2 C128->128 80x64", 1, 80, 64, 128, 128, 3, 3, 2, 1, ACT_RELU},
      {"3x3 pixshuf C512->1024", 1, 20, 16, 512, 1024, 3, 3, 1, 1, ACT_RELU, RES_NONE, STORE_PIXSHUF2},
      {"3x3 C128->50 f32 head", 1, 80, 64, 128, 50, 3, 3, 1, 1, ACT_NONE, RES_NONE, STORE_PLAIN, 1, 0, 14},
      {"3x3 s1 x_pitch 768", 1, 26, 26, 512, 256, 3, 3, 1, 1, ACT_LEAKY, RES_NONE, STORE_PLAIN, 0, 256},
      {"3x3 s1 C512->1024 res", 2, 13, 13, 512, 1024, 3, 3, 1, 1, ACT_LEAKY, RES_AFTER_ACT},
  };
  for (auto& c : conv) fails += run_case(c);

  // CTA pairs (cta_group::2): odd m-tile counts, several n-tiles, residual, fused stores, strided / padded layouts
  std::vector<Case> pairs = {
      {"pair gemm 256->256 1 mtile", 1, 8, 16, 256, 256, 1, 1, 1, 0, ACT_LEAKY, RES_NONE, STORE_PLAIN, 0, 0, 0, 0, 256, 0, 0, 0, 2},
      {"pair gemm 256->512 res", 2, 20, 16, 256, 512, 1, 1, 1, 0, ACT_RELU, RES_BEFORE_ACT, STORE_PLAIN, 0, 0, 0, 0, 256, 0, 0, 0, 2},
      {"pair gemm bn128 odd tiles", 3, 31, 29, 64, 128, 1, 1, 1, 0, ACT_LEAKY, RES_AFTER_ACT, STORE_PLAIN, 0, 0, 0, 0, 128, 0, 0, 0, 2},
      {"pair gemm many tiles", 40, 33, 31, 128, 512, 1, 1, 1, 0, ACT_LEAKY, RES_BEFORE_ACT, STORE_PLAIN, 0, 0, 0, 0, 256, 0, 0, 0, 2},
      {"pair gemm coff+res pitch", 2, 26, 26, 128, 256, 1, 1, 1, 0, ACT_LEAKY, RES_AFTER_ACT, STORE_PLAIN, 0, 64, 128, 64, 256, 0, 0, 0, 2},
      {"pair gemm upsample2 store", 2, 13, 13, 512, 256, 1, 1, 1, 0, ACT_LEAKY, RES_NONE, STORE_UPSAMPLE2, 0, 0, 512, 0, 256, 0, 0, 0, 2},
      {"pair 3x3 s1 C128->256 52x52", 3, 52, 52, 128, 256, 3, 3, 1, 1, ACT_LEAKY, RES_AFTER_ACT, STORE_PLAIN, 0, 0, 0, 0, 256, 0, 0, 0, 2},
      {"pair 3x3 s2 C128->256 26x26", 2, 26, 26, 128, 256, 3, 3, 2, 1, ACT_LEAKY, RES_NONE, STORE_PLAIN, 0, 0, 0, 0, 256, 0, 0, 0, 2},
      {"pair 3x3 s1 C64->128 13x13", 5, 13, 13, 64, 128, 3, 3, 1, 1, ACT_LEAKY, RES_NONE, STORE_PLAIN, 0, 0, 0, 0, 128, 0, 0, 0, 2},
      {"pair 3x3 pixshuf C512->1024", 1, 20, 16, 512, 1024, 3, 3, 1, 1, ACT_RELU, RES_NONE, STORE_PIXSHUF2, 0, 0, 0, 0, 256, 0, 0, 0, 2},
      {"pair 3x3 x_pitch 768", 1, 26, 26, 512, 256, 3, 3, 1, 1, ACT_LEAKY, RES_NONE, STORE_PLAIN, 0, 256, 0, 0, 256, 0, 0, 0, 2},
      {"pair 1x1 s2 C256->512 40x32", 2, 40, 32, 256, 512, 1, 1, 2, 0, ACT_NONE, RES_NONE, STORE_PLAIN, 0, 0, 0, 0, 256, 0, 0, 0, 2},
      {"pair 3x3 C512->1024 res 13", 2, 13, 13, 512, 1024, 3, 3, 1, 1, ACT_LEAKY, RES_AFTER_ACT, STORE_PLAIN, 0, 0, 0, 0, 256, 0, 0, 0, 2},
  };
  for (auto& c : pairs) fails += run_case(c);

  // 256-pixel tiles (MT = 2): ragged last tile (second half empty / partial), residual, BK = 32 and 64, strided
  std::vector<Case> wide = {
      {"mt2 gemm Cout32 res", 3, 31, 29, 64, 32, 1, 1, 1, 0, ACT_LEAKY, RES_AFTER_ACT, STORE_PLAIN, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2},
      {"mt2 gemm Cout64 129 rows", 1, 1, 129, 64, 64, 1, 1, 1, 0, ACT_RELU, RES_NONE, STORE_PLAIN, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2},
      {"mt2 gemm C32->64 bk32 res", 2, 33, 31, 32, 64, 1, 1, 1, 0, ACT_LEAKY, RES_BEFORE_ACT, STORE_PLAIN, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2},
      {"mt2 3x3 s2 C32->64 bk32", 2, 48, 40, 32, 64, 3, 3, 2, 1, ACT_LEAKY, RES_NONE, STORE_PLAIN, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2},
      {"mt2 3x3 s1 C32->64 res", 3, 24, 27, 32, 64, 3, 3, 1, 1, ACT_LEAKY, RES_AFTER_ACT, STORE_PLAIN, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2},
      {"mt2 3x3 s1 C64->64", 2, 80, 64, 64, 64, 3, 3, 1, 1, ACT_RELU, RES_NONE, STORE_PLAIN, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2},
      {"mt2 1x1 C128->64 coff", 2, 52, 52, 128, 64, 1, 1, 1, 0, ACT_LEAKY, RES_NONE, STORE_PLAIN, 0, 0, 64, 64, 0, 0, 0, 0, 0, 2},
      {"mt2 many tiles C64->32", 40, 33, 31, 64, 32, 1, 1, 1, 0, ACT_LEAKY, RES_AFTER_ACT, STORE_PLAIN, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2},
      {"mt4 gemm C64->32 res", 3, 31, 29, 64, 32, 1, 1, 1, 0, ACT_LEAKY, RES_AFTER_ACT, STORE_PLAIN, 0, 0, 0, 0, 32, 0, 0, 0, 0, 4},
      {"mt4 gemm C32->64 bk32 res", 2, 33, 31, 32, 64, 1, 1, 1, 0, ACT_LEAKY, RES_BEFORE_ACT, STORE_PLAIN, 0, 0, 0, 0, 64, 0, 0, 0, 0, 4},
      {"mt4 3x3 s2 C32->64 bk32", 2, 48, 40, 32, 64, 3, 3, 2, 1, ACT_LEAKY, RES_NONE, STORE_PLAIN, 0, 0, 0, 0, 64, 0, 0, 0, 0, 4},
      {"mt4 3x3 s1 C32->64 res", 3, 24, 27, 32, 64, 3, 3, 1, 1, ACT_LEAKY, RES_AFTER_ACT, STORE_PLAIN, 0, 0, 0, 0, 64, 0, 0, 0, 0, 4},
      {"mt4 3x3 s1 C32->32", 5, 24, 27, 32, 32, 3, 3, 1, 1, ACT_LEAKY, RES_NONE, STORE_PLAIN, 0, 0, 0, 0, 32, 0, 0, 0, 0, 4},
      {"mt4 many tiles C64->32", 40, 33, 31, 64, 32, 1, 1, 1, 0, ACT_LEAKY, RES_AFTER_ACT, STORE_PLAIN, 0, 0, 0, 0, 32, 0, 0, 0, 0, 4},
      {"mt2 bn128 3x3 C64->128 res", 3, 52, 50, 64, 128, 3, 3, 1, 1, ACT_LEAKY, RES_AFTER_ACT, STORE_PLAIN, 0, 0, 0, 0, 128, 0, 0, 0, 1, 2},
      {"mt2 bn128 gemm 256->384", 3, 31, 29, 256, 384, 1, 1, 1, 0, ACT_RELU, RES_BEFORE_ACT, STORE_PLAIN, 0, 0, 0, 0, 128, 0, 0, 0, 1, 2},
  };
  for (auto& c : wide) fails += run_case(c);

  if (!quick && fails == 0) {
    printf("---- timing (batch 64 production shapes) ----\n");
    std::vector<Case> perf = {
        {"Y 3x3 32->64 s2 @416 mt2", 64, 416, 416, 32, 64, 3, 3, 2, 1, ACT_LEAKY, 0, 0, 0, 0, 0, 0, 64, 0, 5, 0, 0, 2},
        {"Y 3x3 32->64 s2 @416 mt4", 64, 416, 416, 32, 64, 3, 3, 2, 1, ACT_LEAKY, 0, 0, 0, 0, 0, 0, 64, 0, 5, 0, 0, 4},
        {"Y 3x3 32->64 s1 @208 +res mt2", 64, 208, 208, 32, 64, 3, 3, 1, 1, ACT_LEAKY, RES_AFTER_ACT, 0, 0, 0, 0, 0, 64, 0, 5, 0, 0, 2},
        {"Y 3x3 32->64 s1 @208 +res mt4", 64, 208, 208, 32, 64, 3, 3, 1, 1, ACT_LEAKY, RES_AFTER_ACT, 0, 0, 0, 0, 0, 64, 0, 5, 0, 0, 4},
        {"Y 1x1 64->32 @208 mt2", 64, 208, 208, 64, 32, 1, 1, 1, 0, ACT_LEAKY, 0, 0, 0, 0, 0, 0, 32, 0, 5, 0, 0, 2},
        {"Y 1x1 64->32 @208 mt4", 64, 208, 208, 64, 32, 1, 1, 1, 0, ACT_LEAKY, 0, 0, 0, 0, 0, 0, 32, 0, 5, 0, 0, 4},
    };
    for (auto& c : perf) fails += run_case(c);
  }
  printf("TOTAL FAILURES: %d\n", fails);
  return fails ? 1 : 0;
}
