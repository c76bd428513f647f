// Warp-/block-level kernels for the non-network stages of the evaluate path (C-ABI in betapose_b200.h):
// PIL-exact bicubic resize, fused YOLO head decode + arg-max, crop/pad/bilinear, heat-map peak decode,
// result packing.  All HBM-bound integer/byte or light fp32 work: coalesced, vectorised where the layout allows.
#include <cuda_fp16.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "betapose_b200.h"
#include "engine.h"

// ===================================================================================================
// a1  Pillow ImagingResample, bicubic, 8 bits per channel (dataloader.py:94-99,162)
// ===================================================================================================
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// precompute_coeffs + normalize_coeffs_8bpc restated (double precision, same operation order as Pillow)
void build_resize_tables(int in_size, int out_size, std::vector<int32_t>& bounds, std::vector<int32_t>& kk, int& ksize) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  ksize = (int)std::ceil(support) * 2 + 1;
  bounds.assign((size_t)out_size * 2, 0);
  kk.assign((size_t)out_size * ksize, 0);
  const double ss = 1.0 / filterscale;
  std::vector<double> w(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      double v = w[x];
      if (ww != 0.0) v /= ww;
      kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
}

const ResizeTables* get_tables(bp_engine* e, int in_size, int out_size) {
  auto key = std::make_pair(in_size, out_size);
  auto it = e->resize_tables.find(key);
  if (it != e->resize_tables.end()) return &it->second;
  std::vector<int32_t> b, k;
  ResizeTables t;
  t.in_size = in_size;
  t.out_size = out_size;
  build_resize_tables(in_size, out_size, b, k, t.ksize);
  if (cudaMalloc(&t.bounds, b.size() * 4) != cudaSuccess || cudaMalloc(&t.coeffs, k.size() * 4) != cudaSuccess) return nullptr;
  cudaMemcpy(t.bounds, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(t.coeffs, k.data(), k.size() * 4, cudaMemcpyHostToDevice);
  e->owned.push_back(t.bounds);
  e->owned.push_back(t.coeffs);
  t.host_bounds = b;
  return &(e->resize_tables[key] = t);
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: [B,H,W,3] u8 -> [B,H,ow,3] u8 ; one thread per output pixel
__global__ void resize_h_kernel(const uint8_t* __restrict__ in, int rows, int W, int ow, const int32_t* __restrict__ bounds,
                                const int32_t* __restrict__ kk, int ksize, uint8_t* __restrict__ out) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= (long)rows * ow) return;
  const int xx = idx % ow;
  const long y = idx / ow;
  const int x0 = bounds[2 * xx], n = bounds[2 * xx + 1];
  const int32_t* k = kk + (long)xx * ksize;
  const uint8_t* src = in + (y * W + x0) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < n; ++x) {
    const int c = k[x];
    s0 += src[3 * x] * c;
    s1 += src[3 * x + 1] * c;
    s2 += src[3 * x + 2] * c;
  }
  uint8_t* dst = out + idx * 3;
  dst[0] = (uint8_t)clip8(s0);
  dst[1] = (uint8_t)clip8(s1);
  dst[2] = (uint8_t)clip8(s2);
}

// vertical pass: [B,H,ow,3] u8 -> network input fp16 [B,oh,ow+BP_IN_PAD_COLS,8] (raw 0..255, exact in fp16) and/or
// fp32 [B,3,oh,ow] (value/255)
__global__ void resize_v_kernel(const uint8_t* __restrict__ in, int B, int H, int oh, int ow, const int32_t* __restrict__ bounds,
                                const int32_t* __restrict__ kk, int ksize, __half* __restrict__ out_net,
                                float* __restrict__ out_f32) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= (long)B * oh * ow) return;
  const int xx = idx % ow;
  const int yy = (idx / ow) % oh;
  const int b = idx / ((long)ow * oh);
  const int y0 = bounds[2 * yy], n = bounds[2 * yy + 1];
  const int32_t* k = kk + (long)yy * ksize;
  const uint8_t* src = in + (((long)b * H + y0) * ow + xx) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int y = 0; y < n; ++y) {
    const int c = k[y];
    const uint8_t* p = src + (long)y * ow * 3;
    s0 += p[0] * c;
    s1 += p[1] * c;
    s2 += p[2] * c;
  }
  const int r = clip8(s0), g = clip8(s1), bl = clip8(s2);
  if (out_net) {
    uint4 pk = make_uint4(0u, 0u, 0u, 0u);
    __half2* h = reinterpret_cast<__half2*>(&pk);
    h[0] = __floats2half2_rn((float)r, (float)g);
    h[1] = __floats2half2_rn((float)bl, 0.f);
    reinterpret_cast<uint4*>(out_net)[((long)b * oh + yy) * (ow + BP_IN_PAD_COLS) + BP_IN_PAD_LEFT + xx] = pk;
  }
  if (out_f32) {
    const long plane = (long)oh * ow;
    float* o = out_f32 + (long)b * 3 * plane + (long)yy * ow + xx;
    o[0] = __fdiv_rn((float)r, 255.f);
    o[plane] = __fdiv_rn((float)g, 255.f);
    o[2 * plane] = __fdiv_rn((float)bl, 255.f);
  }
}

// Both passes in one kernel: a block owns TH output rows of one image.  It stages the input rows those need in shared
// memory (16-byte loads) together with the coefficient tables, runs Pillow's horizontal pass on them into a uint8
// intermediate -- the rounding to 8 bits between the passes is part of the bit-exact semantics --, then the vertical pass,
// and writes the network-input layout.  No global intermediate, coalesced traffic only; the horizontal pass is redone for
// the few rows neighbouring tiles share.  One thread per output COLUMN in both passes: its horizontal taps (<= kMaxTaps
// 22-bit coefficients) live in registers for all rows, the vertical taps are a shared-memory broadcast, and no index
// arithmetic (division / modulo) is left in the inner loops -- the kernel is bound by its integer multiply-adds.
constexpr int kMaxTaps = 12;

__global__ void __launch_bounds__(448)
resize_fused_kernel(const uint8_t* __restrict__ in, int H, int W, int oh, int ow, const int32_t* __restrict__ hb,
                    const int32_t* __restrict__ hk, int hks, const int32_t* __restrict__ vb, const int32_t* __restrict__ vk, int vks,
                    int TH, int in_pitch, int mid_pitch, int max_rows, __half* __restrict__ out_net, float* __restrict__ out_f32) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  uint8_t* s_in = rs_smem;                                  // [max_rows][in_pitch]
  uint8_t* s_mid = rs_smem + (size_t)max_rows * in_pitch;   // [max_rows][mid_pitch]
  int32_t* s_vk = reinterpret_cast<int32_t*>(s_mid + (size_t)max_rows * mid_pitch);  // [TH][vks] vertical taps of this tile
  int32_t* s_vb = s_vk + TH * vks;                                                   // [TH][2]
  const int b = blockIdx.y;
  const int y0 = blockIdx.x * TH, y1 = min(oh, y0 + TH);
  const int r0 = vb[2 * y0];
  const int r1 = vb[2 * (y1 - 1)] + vb[2 * (y1 - 1) + 1];
  const int nrows = r1 - r0;
  const int row_bytes = W * 3;
  const uint8_t* src = in + ((long)b * H + r0) * row_bytes;
  for (int e = threadIdx.x; e < (y1 - y0) * vks; e += blockDim.x) s_vk[e] = __ldg(vk + (long)y0 * vks + e);
  for (int e = threadIdx.x; e < (y1 - y0) * 2; e += blockDim.x) s_vb[e] = __ldg(vb + 2 * y0 + e);
  if ((row_bytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(in) & 15) == 0)) {
    const int vec = row_bytes >> 4;
    for (int r = 0; r < nrows; ++r)
      for (int c = threadIdx.x; c < vec; c += blockDim.x)
        reinterpret_cast<uint4*>(s_in + (size_t)r * in_pitch)[c] = __ldg(reinterpret_cast<const uint4*>(src + (long)r * row_bytes) + c);
  } else {
    for (int r = 0; r < nrows; ++r)
      for (int c = threadIdx.x; c < row_bytes; c += blockDim.x) s_in[(size_t)r * in_pitch + c] = src[(long)r * row_bytes + c];
  }
  __syncthreads();
  for (int xx = threadIdx.x; xx < ow; xx += blockDim.x) {
    // ---- horizontal pass of column xx over the staged rows
    const int x0 = __ldg(hb + 2 * xx), n = __ldg(hb + 2 * xx + 1);
    int k[kMaxTaps];
#pragma unroll
    for (int t = 0; t < kMaxTaps; ++t) k[t] = t < n ? __ldg(hk + (long)xx * hks + t) : 0;
    const uint8_t* p = s_in + x0 * 3;
    uint8_t* d = s_mid + xx * 3;
    for (int r = 0; r < nrows; ++r, p += in_pitch, d += mid_pitch) {
      int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
#pragma unroll
      for (int t = 0; t < kMaxTaps; ++t) {
        if (t < n) {
          s0 += p[3 * t] * k[t];
          s1 += p[3 * t + 1] * k[t];
          s2 += p[3 * t + 2] * k[t];
        }
      }
      d[0] = (uint8_t)clip8(s0);
      d[1] = (uint8_t)clip8(s1);
      d[2] = (uint8_t)clip8(s2);
    }
  }
  __syncthreads();
  for (int xx = threadIdx.x; xx < ow; xx += blockDim.x) {
    // ---- vertical pass of column xx for the TH output rows
    for (int yl = 0; yl < y1 - y0; ++yl) {
      const int yy = y0 + yl;
      const int yb = s_vb[2 * yl] - r0, n = s_vb[2 * yl + 1];
      const int32_t* k = s_vk + yl * vks;
      const uint8_t* q = s_mid + (size_t)yb * mid_pitch + xx * 3;
      int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
      for (int y = 0; y < n; ++y, q += mid_pitch) {
        const int c = k[y];
        s0 += q[0] * c;
        s1 += q[1] * c;
        s2 += q[2] * c;
      }
      const int r = clip8(s0), g = clip8(s1), bl = clip8(s2);
      if (out_net) {
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        __half2* h = reinterpret_cast<__half2*>(&pk);
        h[0] = __floats2half2_rn((float)r, (float)g);
        h[1] = __floats2half2_rn((float)bl, 0.f);
        reinterpret_cast<uint4*>(out_net)[((long)b * oh + yy) * (ow + BP_IN_PAD_COLS) + BP_IN_PAD_LEFT + xx] = pk;
      }
      if (out_f32) {
        const long plane = (long)oh * ow;
        float* o = out_f32 + (long)b * 3 * plane + (long)yy * ow + xx;
        o[0] = __fdiv_rn((float)r, 255.f);
        o[plane] = __fdiv_rn((float)g, 255.f);
        o[2 * plane] = __fdiv_rn((float)bl, 255.f);
      }
    }
  }
}

}  // namespace

extern "C" int bp_resize_bicubic(bp_engine* e, const uint8_t* frames, int B, int H, int W, int oh, int ow,
                                 void* out_net, float* out_f32_chw, void* stream) {
  if (!e || !frames || B <= 0 || (!out_net && !out_f32_chw)) return bp_fail(BP_ERR_INVALID, "bp_resize_bicubic: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const ResizeTables* th = get_tables(e, W, ow);
  const ResizeTables* tv = get_tables(e, H, oh);
  if (!th || !tv) return bp_fail(BP_ERR_CUDA, "bp_resize_bicubic: table upload failed");
  {
    // fused single-kernel path whenever a row tile fits in shared memory: largest TH (output rows per block) whose
    // input rows + uint8 intermediate stay below ~100 KB (two blocks per SM)
    const int in_pitch = (W * 3 + 15) / 16 * 16, mid_pitch = (ow * 3 + 15) / 16 * 16;
    for (int TH = 32; TH >= 1; TH >>= 1) {
      int max_rows = 0;
      for (int y0 = 0; y0 < oh; y0 += TH) {
        const int yl = std::min(oh, y0 + TH) - 1;
        max_rows = std::max(max_rows, tv->host_bounds[2 * yl] + tv->host_bounds[2 * yl + 1] - tv->host_bounds[2 * y0]);
      }
      if (th->ksize > kMaxTaps) break;  // (very strong down-scaling: the two-kernel path below has no tap limit)
      const size_t smem = (size_t)max_rows * (in_pitch + mid_pitch) + (size_t)TH * (tv->ksize + 2) * sizeof(int32_t);
      if (smem > 100 * 1024) continue;
      if (smem > 48 * 1024) {
        static bool attr_done[64] = {};  // per device
        if (bp_attr_once(attr_done)) {
          if (cudaFuncSetAttribute(resize_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) break;
        }
      }
      dim3 grid((oh + TH - 1) / TH, B);
      const int threads = std::min(448, (ow + 31) / 32 * 32);  // one thread per output column (416 for the detector input)
      resize_fused_kernel<<<grid, threads, smem, st>>>(frames, H, W, oh, ow, th->bounds, th->coeffs, th->ksize, tv->bounds, tv->coeffs,
                                                  tv->ksize, TH, in_pitch, mid_pitch, max_rows, (__half*)out_net, out_f32_chw);
      cudaError_t err = cudaGetLastError();
      return err == cudaSuccess ? BP_OK : bp_fail(BP_ERR_CUDA, cudaGetErrorString(err));
    }
  }
  // fall-back for extreme sizes: two passes through a global uint8 intermediate
  const size_t need = (size_t)B * H * ow * 3;
  bp_engine::StreamScratch& sc = e->scratch_for(st);
  // per-stream, grow-only scratch; (re)allocation happens at most once per batch size, outside steady state
  if (!e->grow(reinterpret_cast<void**>(&sc.resize_tmp), &sc.resize_tmp_bytes, need))
    return bp_fail(BP_ERR_CUDA, "bp_resize_bicubic: scratch allocation failed");
  const long n1 = (long)B * H * ow;
  resize_h_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(frames, B * H, W, ow, th->bounds, th->coeffs, th->ksize,
                                                              sc.resize_tmp);
  const long n2 = (long)B * oh * ow;
  resize_v_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(sc.resize_tmp, B, H, oh, ow, tv->bounds, tv->coeffs,
                                                              tv->ksize, (__half*)out_net, out_f32_chw);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? BP_OK : bp_fail(BP_ERR_CUDA, cudaGetErrorString(err));
}

// ===================================================================================================
// a3 + a4 + a5  YOLO head decode + per-image arg-max objectness + box rescale
// ===================================================================================================
namespace {

struct HeadDesc {
  const float* ptr[3];
  int grid[3];
  int pitch[3];
  int rows_before[3];  // flat row offset of each head
  float anchor[3][6];  // (w,h) x 3 anchors, pixels
  int n_heads;
};

__device__ __forceinline__ float sigmoidf_ref(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// one block per image; 256 threads scan all candidate rows (obj logit only), block arg-max with lowest-index
// tie-break, then thread 0 decodes the winner.  Optionally every row is decoded into `decoded` [B,R,n_attr].
__global__ void yolo_decode_argmax_kernel(HeadDesc hd, int n_attr, int total_rows, int reso, float conf, float wr, float hr,
                                          float* __restrict__ det, float* __restrict__ box, float* __restrict__ score,
                                          int32_t* __restrict__ row_out, uint8_t* __restrict__ valid,
                                          float* __restrict__ decoded) {
  const int b = blockIdx.x;
  __shared__ float s_val[256];
  __shared__ int s_idx[256];
  float best = -1.f;
  int best_row = 0x7fffffff;
  for (int r = threadIdx.x; r < total_rows; r += blockDim.x) {
    int h = 0;
    while (h + 1 < hd.n_heads && r >= hd.rows_before[h + 1]) ++h;
    const int g = hd.grid[h];
    const int local = r - hd.rows_before[h];
    const int a = local / (g * g);
    const int cell = local - a * g * g;
    const float* px = hd.ptr[h] + ((long)b * g * g + cell) * hd.pitch[h] + a * n_attr;
    const float obj = sigmoidf_ref(px[4]);
    bool cand = obj > conf;
    if (n_attr > 6 && cand) {  // multi-class cfg: the reference keeps class 0 only (yolo/util.py:166-167)
      int am = 0;
      float mv = px[5];
      for (int c = 1; c < n_attr - 5; ++c)
        if (px[5 + c] > mv) { mv = px[5 + c]; am = c; }
      cand = am == 0;
    }
    if (decoded) {
      const int cy = cell / g, cx = cell - cy * g;
      const float stride = (float)(reso / g);
      float* o = decoded + ((long)b * total_rows + r) * n_attr;
      o[0] = __fmul_rn(__fadd_rn(sigmoidf_ref(px[0]), (float)cx), stride);
      o[1] = __fmul_rn(__fadd_rn(sigmoidf_ref(px[1]), (float)cy), stride);
      o[2] = __fmul_rn(__fmul_rn(expf(px[2]), __fdiv_rn(hd.anchor[h][2 * a], stride)), stride);
      o[3] = __fmul_rn(__fmul_rn(expf(px[3]), __fdiv_rn(hd.anchor[h][2 * a + 1], stride)), stride);
      o[4] = obj;
      for (int c = 5; c < n_attr; ++c) o[c] = sigmoidf_ref(px[c]);
    }
    if (cand && (obj > best || (obj == best && r < best_row))) {
      best = obj;
      best_row = r;
    }
  }
  s_val[threadIdx.x] = best;
  s_idx[threadIdx.x] = best_row;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const float ov = s_val[threadIdx.x + s];
      const int oi = s_idx[threadIdx.x + s];
      if (ov > s_val[threadIdx.x] || (ov == s_val[threadIdx.x] && oi < s_idx[threadIdx.x])) {
        s_val[threadIdx.x] = ov;
        s_idx[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int r = s_idx[0];
    const bool ok = s_val[0] >= 0.f && r != 0x7fffffff;
    valid[b] = ok ? 1 : 0;
    row_out[b] = ok ? r : -1;
    float d[8] = {(float)b, 0, 0, 0, 0, 0, 0, 0};
    float bx[4] = {0, 0, 0, 0};
    if (ok) {
      int h = 0;
      while (h + 1 < hd.n_heads && r >= hd.rows_before[h + 1]) ++h;
      const int g = hd.grid[h];
      const int local = r - hd.rows_before[h];
      const int a = local / (g * g);
      const int cell = local - a * g * g;
      const int cy = cell / g, cx = cell - cy * g;
      const float* px = hd.ptr[h] + ((long)b * g * g + cell) * hd.pitch[h] + a * n_attr;
      const float stride = (float)(reso / g);
      const float bxc = __fmul_rn(__fadd_rn(sigmoidf_ref(px[0]), (float)cx), stride);
      const float byc = __fmul_rn(__fadd_rn(sigmoidf_ref(px[1]), (float)cy), stride);
      const float bw = __fmul_rn(__fmul_rn(expf(px[2]), __fdiv_rn(hd.anchor[h][2 * a], stride)), stride);
      const float bh = __fmul_rn(__fmul_rn(expf(px[3]), __fdiv_rn(hd.anchor[h][2 * a + 1], stride)), stride);
      d[1] = __fsub_rn(bxc, __fdiv_rn(bw, 2.f));
      d[2] = __fsub_rn(byc, __fdiv_rn(bh, 2.f));
      d[3] = __fadd_rn(bxc, __fdiv_rn(bw, 2.f));
      d[4] = __fadd_rn(byc, __fdiv_rn(bh, 2.f));
      d[5] = s_val[0];
      d[6] = sigmoidf_ref(px[5]);
      d[7] = 0.f;
      bx[0] = __fmul_rn(d[1], wr);
      bx[1] = __fmul_rn(d[2], hr);
      bx[2] = __fmul_rn(d[3], wr);
      bx[3] = __fmul_rn(d[4], hr);
    }
    for (int i = 0; i < 8; ++i) det[b * 8 + i] = d[i];
    if (score) score[b] = d[5];
    for (int i = 0; i < 4; ++i) box[b * 4 + i] = bx[i];
  }
}

}  // namespace

extern "C" int bp_yolo_decode_argmax(bp_engine* e, const float* const* heads, const int* grids, const int* pitches,
                                     int n_heads, const float* anchors, int n_attr, int B, int reso, float conf,
                                     int frame_w, int frame_h, float* det, float* box, float* score, int32_t* row,
                                     uint8_t* valid, float* decoded, void* stream) {
  if (!e || !heads || n_heads < 1 || n_heads > 3 || B <= 0 || n_attr < 6) return bp_fail(BP_ERR_INVALID, "bp_yolo_decode_argmax: bad arguments");
  HeadDesc hd;
  int total = 0;
  for (int i = 0; i < n_heads; ++i) {
    hd.ptr[i] = heads[i];
    hd.grid[i] = grids[i];
    hd.pitch[i] = pitches[i];
    hd.rows_before[i] = total;
    total += 3 * grids[i] * grids[i];
    for (int k = 0; k < 6; ++k) hd.anchor[i][k] = anchors[i * 6 + k];
  }
  hd.n_heads = n_heads;
  const float wr = (float)frame_w / (float)reso, hr = (float)frame_h / (float)reso;
  yolo_decode_argmax_kernel<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(hd, n_attr, total, reso, conf, wr, hr, det,
                                                                                  box, score, row, valid, decoded);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? BP_OK : bp_fail(BP_ERR_CUDA, cudaGetErrorString(err));
}

namespace {
// one block per image over decoded rows (cx, cy, w, h, obj, cls...): arg-max objectness above `conf`, lowest row on ties
__global__ void write_results_kernel(const float* __restrict__ pred, int R, int n_attr, float conf, float* __restrict__ det,
                                     int32_t* __restrict__ row_out, uint8_t* __restrict__ valid) {
  const int b = blockIdx.x;
  __shared__ float s_val[256];
  __shared__ int s_idx[256];
  const float* p = pred + (long)b * R * n_attr;
  float best = -1.f;
  int best_row = 0x7fffffff;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float obj = p[(long)r * n_attr + 4];
    bool cand = obj > conf;
    if (cand && n_attr > 6) {
      int am = 0;
      float mv = p[(long)r * n_attr + 5];
      for (int c = 1; c < n_attr - 5; ++c)
        if (p[(long)r * n_attr + 5 + c] > mv) { mv = p[(long)r * n_attr + 5 + c]; am = c; }
      cand = am == 0;
    }
    if (cand && (obj > best || (obj == best && r < best_row))) { best = obj; best_row = r; }
  }
  s_val[threadIdx.x] = best;
  s_idx[threadIdx.x] = best_row;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const float ov = s_val[threadIdx.x + s];
      const int oi = s_idx[threadIdx.x + s];
      if (ov > s_val[threadIdx.x] || (ov == s_val[threadIdx.x] && oi < s_idx[threadIdx.x])) { s_val[threadIdx.x] = ov; s_idx[threadIdx.x] = oi; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int r = s_idx[0];
    const bool ok = s_val[0] >= 0.f && r != 0x7fffffff;
    valid[b] = ok ? 1 : 0;
    row_out[b] = ok ? r : -1;
    float d[8] = {(float)b, 0, 0, 0, 0, 0, 0, 0};
    if (ok) {
      const float* q = p + (long)r * n_attr;
      d[1] = __fsub_rn(q[0], __fdiv_rn(q[2], 2.f));
      d[2] = __fsub_rn(q[1], __fdiv_rn(q[3], 2.f));
      d[3] = __fadd_rn(q[0], __fdiv_rn(q[2], 2.f));
      d[4] = __fadd_rn(q[1], __fdiv_rn(q[3], 2.f));
      d[5] = q[4];
      d[6] = q[5];
    }
    for (int i = 0; i < 8; ++i) det[b * 8 + i] = d[i];
  }
}
}  // namespace

extern "C" int bp_write_results(bp_engine* e, const float* pred, int B, int R, int n_attr, float conf, float* det, int32_t* row,
                                uint8_t* valid, void* stream) {
  if (!e || !pred || B <= 0 || R <= 0 || n_attr < 6 || !det || !row || !valid) return bp_fail(BP_ERR_INVALID, "bp_write_results: bad arguments");
  write_results_kernel<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, R, n_attr, conf, det, row, valid);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? BP_OK : bp_fail(BP_ERR_CUDA, cudaGetErrorString(err));
}

// ===================================================================================================
// a6  crop_from_dets + cropBox (dataloader.py:794-835, KPD/src/utils/img.py:242-262)
// ===================================================================================================
namespace {

struct CropGeom {
  float pt1x, pt1y, pt2x, pt2y;
  int ulx, uly, hS, wS, Hp, Wp, top, left;
};

__device__ CropGeom crop_geometry(const float* bx, int W, int H, int rh, int rw) {
  CropGeom g;
  const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
  const float ht = __fsub_rn(y2, y1), width = __fsub_rn(x2, x1);
  const float rate = width > 100.f ? 0.2f : 0.3f;
  const float dx = __fdiv_rn(__fmul_rn(width, rate), 2.f), dy = __fdiv_rn(__fmul_rn(ht, rate), 2.f);
  g.pt1x = fmaxf(0.f, __fsub_rn(x1, dx));
  g.pt1y = fmaxf(0.f, __fsub_rn(y1, dy));
  g.pt2x = fmaxf(fminf((float)(W - 1), __fadd_rn(x2, dx)), __fadd_rn(g.pt1x, 5.f));
  g.pt2y = fmaxf(fminf((float)(H - 1), __fadd_rn(y2, dy)), __fadd_rn(g.pt1y, 5.f));
  g.ulx = (int)g.pt1x;
  g.uly = (int)g.pt1y;
  const int brx = (int)g.pt2x, bry = (int)g.pt2y;
  const int hb = bry - g.uly, wb = brx - g.ulx;
  // torch >= 1.5 true-division semantics of cropBox
  const float cand = __fdiv_rn((float)(wb * rh), (float)rw);
  int Hp, Wp;
  if (cand > (float)hb) {
    Hp = (int)cand;
    Wp = (int)__fdiv_rn(__fmul_rn(cand, (float)rw), (float)rh);
  } else {
    Hp = hb;
    Wp = (int)__fdiv_rn((float)(hb * rw), (float)rh);
  }
  g.hS = max(0, min(hb, H - g.uly));
  g.wS = max(0, min(wb, W - g.ulx));
  g.top = (max(Hp - g.hS, 0) + 1) / 2;   // ceil(diff / 2)
  g.left = (max(Wp - g.wS, 0) + 1) / 2;
  g.Hp = max(Hp, g.hS);
  g.Wp = max(Wp, g.wS);
  return g;
}

// (u / 255) - mean[c] of im_to_torch + the mean subtraction (dataloader.py:802-804, BGR means applied in RGB order), one
// exact fp32 division per table entry; built once per engine instead of three divisions per output pixel
__global__ void crop_lut_kernel(float* lut) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 768) return;
  const float mean[3] = {0.406f, 0.457f, 0.480f};
  lut[e] = __fadd_rn(__fdiv_rn((float)(e & 255), 255.f), -mean[e >> 8]);
}

constexpr int kCropPxPerThread = 4;

// 4-tap bilinear (align_corners=True) over the zero-padded, mean-subtracted patch; a block covers 1024 output pixels of one
// crop (4 per thread, so the per-block set-up -- box geometry, scale factors, the 3 KB value table -- is amortised); coalesced
// fp16x8 NHWC stores (+ optional fp32 NCHW for the drop-in seam).  Same fp32 operation order as the reference throughout.
__global__ void __launch_bounds__(256)
crop_resize_kernel(const uint8_t* __restrict__ frames, int H, int W, const float* __restrict__ box,
                   const int32_t* __restrict__ img_idx, const uint8_t* __restrict__ valid, int rh, int rw,
                   const float* __restrict__ lut, __half* __restrict__ out16, float* __restrict__ out32, float* __restrict__ pt1,
                   float* __restrict__ pt2) {
  __shared__ CropGeom s_g;
  __shared__ float s_scale[2];
  __shared__ float s_lut[3][256];
  const int i = blockIdx.y;
  const bool ok = !valid || valid[i];
  if (ok) {
    if (threadIdx.x == 0) {
      const CropGeom g = crop_geometry(box + 4 * i, W, H, rh, rw);
      s_g = g;
      s_scale[0] = rh > 1 ? __fdiv_rn((float)(g.Hp - 1), (float)(rh - 1)) : 0.f;
      s_scale[1] = rw > 1 ? __fdiv_rn((float)(g.Wp - 1), (float)(rw - 1)) : 0.f;
    }
    for (int e = threadIdx.x; e < 768; e += blockDim.x) s_lut[e >> 8][e & 255] = __ldg(lut + e);
  }
  __syncthreads();
  const int npix = rh * rw;
  const CropGeom g = s_g;
  const float rhs = s_scale[0], rws = s_scale[1];
  const uint8_t* fr = ok ? frames + (long)img_idx[i] * H * W * 3 : frames;
  int pix = blockIdx.x * (blockDim.x * kCropPxPerThread) + threadIdx.x;
  int oy = pix / rw, ox = pix - oy * rw;
#pragma unroll
  for (int it = 0; it < kCropPxPerThread; ++it, pix += blockDim.x) {
    if (pix >= npix) break;
    if (it) {  // advance (oy, ox) by blockDim.x pixels without a division
      ox += blockDim.x;
      while (ox >= rw) {
        ox -= rw;
        ++oy;
      }
    }
    float v[3] = {0.f, 0.f, 0.f};
    if (ok) {
      const float ys = __fmul_rn(rhs, (float)oy), xs = __fmul_rn(rws, (float)ox);
      const int y0 = (int)ys, x0 = (int)xs;
      const int y1 = min(y0 + 1, g.Hp - 1), x1 = min(x0 + 1, g.Wp - 1);
      const float ly = __fsub_rn(ys, (float)y0), lx = __fsub_rn(xs, (float)x0);
      const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
      // positions inside the source patch; a tap outside it reads the zero padding
      const int sy0 = y0 - g.top, sy1 = y1 - g.top, sx0 = x0 - g.left, sx1 = x1 - g.left;
      const bool vy0 = (unsigned)sy0 < (unsigned)g.hS, vy1 = (unsigned)sy1 < (unsigned)g.hS;
      const bool vx0 = (unsigned)sx0 < (unsigned)g.wS, vx1 = (unsigned)sx1 < (unsigned)g.wS;
      const uint8_t* r0 = fr + ((long)(g.uly + sy0) * W + g.ulx) * 3;
      const uint8_t* r1 = fr + ((long)(g.uly + sy1) * W + g.ulx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float p00 = (vy0 && vx0) ? s_lut[c][r0[sx0 * 3 + c]] : 0.f;
        const float p01 = (vy0 && vx1) ? s_lut[c][r0[sx1 * 3 + c]] : 0.f;
        const float p10 = (vy1 && vx0) ? s_lut[c][r1[sx0 * 3 + c]] : 0.f;
        const float p11 = (vy1 && vx1) ? s_lut[c][r1[sx1 * 3 + c]] : 0.f;
        const float t = __fadd_rn(__fmul_rn(hx, p00), __fmul_rn(lx, p01));
        const float bt = __fadd_rn(__fmul_rn(hx, p10), __fmul_rn(lx, p11));
        v[c] = __fadd_rn(__fmul_rn(hy, t), __fmul_rn(ly, bt));
      }
    }
    if (out16) {  // network input layout: [n, rh, rw + BP_IN_PAD_COLS, 8]
      uint4 pk = make_uint4(0u, 0u, 0u, 0u);
      __half2* h = reinterpret_cast<__half2*>(&pk);
      h[0] = __floats2half2_rn(v[0], v[1]);
      h[1] = __floats2half2_rn(v[2], 0.f);
      reinterpret_cast<uint4*>(out16)[((long)i * rh + oy) * (rw + BP_IN_PAD_COLS) + BP_IN_PAD_LEFT + ox] = pk;
    }
    if (out32) {
      const long plane = (long)rh * rw;
      float* o = out32 + (long)i * 3 * plane + pix;
      o[0] = v[0];
      o[plane] = v[1];
      o[2 * plane] = v[2];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (ok) {
      pt1[2 * i] = g.pt1x; pt1[2 * i + 1] = g.pt1y;
      pt2[2 * i] = g.pt2x; pt2[2 * i + 1] = g.pt2y;
    } else {
      pt1[2 * i] = pt1[2 * i + 1] = pt2[2 * i] = pt2[2 * i + 1] = 0.f;
    }
  }
}

}  // namespace

extern "C" int bp_crop_resize(bp_engine* e, const uint8_t* frames, int H, int W, const float* box, const int32_t* img_idx,
                              const uint8_t* valid, int n, int rh, int rw, void* out_net, float* out_f32_chw, float* pt1,
                              float* pt2, void* stream) {
  if (!e || !frames || !box || !img_idx || !pt1 || !pt2 || n <= 0 || (!out_net && !out_f32_chw))
    return bp_fail(BP_ERR_INVALID, "bp_crop_resize: bad arguments");
  if (!e->crop_lut) {  // first use (engine set-up; not inside a stream capture): the value table
    float* lut = nullptr;
    if (cudaMalloc(&lut, 768 * sizeof(float)) != cudaSuccess) return bp_fail(BP_ERR_CUDA, "bp_crop_resize: table allocation failed");
    crop_lut_kernel<<<3, 256>>>(lut);
    if (cudaDeviceSynchronize() != cudaSuccess) return bp_fail(BP_ERR_CUDA, "bp_crop_resize: table kernel failed (first use inside a stream capture?)");
    std::lock_guard<std::mutex> g(e->mu);
    e->owned.push_back(lut);
    e->crop_lut = lut;
  }
  const int per_block = 256 * kCropPxPerThread;
  dim3 grid((rh * rw + per_block - 1) / per_block, n);
  crop_resize_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(frames, H, W, box, img_idx, valid, rh, rw, e->crop_lut,
                                                                              (__half*)out_net, out_f32_chw, pt1, pt2);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? BP_OK : bp_fail(BP_ERR_CUDA, cudaGetErrorString(err));
}

// ===================================================================================================
// a8  getPrediction + transformBoxInvert_batch (KPD/src/utils/eval.py:113-147, img.py:216-239)
// ===================================================================================================
namespace {

__device__ __forceinline__ void argmax_merge(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) {
    v = ov;
    i = oi;
  }
}

// grid (images, S), 1024 threads.  NHWC heat-maps (k_stride == 1): thread = (position lane, map) so a warp reads 32
// consecutive maps of one position (coalesced); the positions of an image are cut into S slices (one block each, so a
// 64-image batch fills the machine), every block leaves its per-map partial (max, lowest index) in `part`, and the last
// block of an image to finish merges the S partials in slice order and finalises.  NCHW (pos_stride == 1, the
// getPrediction seam): thread = (map lane, position), S = 1.
__global__ void heatmap_decode_kernel(const float* __restrict__ hm, long img_stride, long k_stride, long pos_stride, int K,
                                      int res_h, int res_w, float inp_ratio_hw, float inp_ratio_wh, const float* __restrict__ pt1,
                                      const float* __restrict__ pt2, float* __restrict__ preds_hm, float* __restrict__ preds_img,
                                      float* __restrict__ maxval, int32_t* __restrict__ idx_out, float* __restrict__ part_val,
                                      int* __restrict__ part_idx, unsigned* __restrict__ counters) {
  extern __shared__ unsigned char sm_raw[];
  __shared__ bool s_last;
  const int n = blockIdx.x;
  const int S = gridDim.y, sl = blockIdx.y;
  const int npos = res_h * res_w;
  const float* base = hm + (long)n * img_stride;
  const int T = blockDim.x;
  float* s_val = reinterpret_cast<float*>(sm_raw);
  int* s_idx = reinterpret_cast<int*>(s_val + T);
  // lanes-per-map layout
  int kslots, lanes;
  if (k_stride == 1) {
    kslots = 1;
    while (kslots < K) kslots <<= 1;  // maps padded to a power of two (64 for K=50)
    lanes = T / kslots;
  } else {
    lanes = 32;                        // one warp per map, looping over maps
    kslots = T / 32;
  }
  if (k_stride == 1) {
    const int k = threadIdx.x % kslots, l = threadIdx.x / kslots;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    const int chunk = (npos + S - 1) / S;
    const int p0 = sl * chunk, p1 = min(npos, p0 + chunk);
    if (k < K)
      for (int p = p0 + l; p < p1; p += lanes) argmax_merge(bv, bi, base[(long)p * pos_stride + k], p);
    s_val[threadIdx.x] = bv;
    s_idx[threadIdx.x] = bi;
    __syncthreads();
    for (int s = lanes / 2; s > 0; s >>= 1) {
      if (l < s) {
        float v = s_val[threadIdx.x];
        int i = s_idx[threadIdx.x];
        argmax_merge(v, i, s_val[threadIdx.x + s * kslots], s_idx[threadIdx.x + s * kslots]);
        s_val[threadIdx.x] = v;
        s_idx[threadIdx.x] = i;
      }
      __syncthreads();
    }
  } else {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < K; k += kslots) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      const float* mp = base + (long)k * k_stride;
      for (int p = lane; p < npos; p += 32) argmax_merge(bv, bi, mp[(long)p * pos_stride], p);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        argmax_merge(bv, bi, ov, oi);
      }
      if (lane == 0) {
        s_val[k] = bv;
        s_idx[k] = bi;
      }
    }
    __syncthreads();
  }
  if (S > 1) {
    // partials -> global, last block of the image merges them (slice order: lowest index wins ties, as in one block)
    const int kslots_g = 64;
    if (threadIdx.x < K) {
      part_val[((long)n * S + sl) * kslots_g + threadIdx.x] = s_val[threadIdx.x];
      part_idx[((long)n * S + sl) * kslots_g + threadIdx.x] = s_idx[threadIdx.x];
      __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned old = atomicAdd(counters + n, 1u);
      s_last = old == (unsigned)S - 1u;
      if (s_last) counters[n] = 0u;  // ready for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < K) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int q = 0; q < S; ++q)
        argmax_merge(bv, bi, __ldcg(part_val + ((long)n * S + q) * kslots_g + threadIdx.x),
                     __ldcg(part_idx + ((long)n * S + q) * kslots_g + threadIdx.x));
      s_val[threadIdx.x] = bv;
      s_idx[threadIdx.x] = bi;
    }
    __syncthreads();
  }
  // finalise: thread k < K refines and maps back its key-point
  const int k = threadIdx.x;
  if (k >= K) return;
  const float mv = s_val[k];
  const int id = s_idx[k];
  const float* mp = base + (long)k * k_stride;
  float px = 0.f, py = 0.f;
  if (mv > 0.f) {
    px = (float)(id % res_w);
    py = floorf(__fdiv_rn((float)id, (float)res_w));
  }
  const int ix = (int)px, iy = (int)py;
  if (ix > 0 && ix < res_w - 1 && iy > 0 && iy < res_h - 1) {
    const float dx = __fsub_rn(mp[(long)(iy * res_w + ix + 1) * pos_stride], mp[(long)(iy * res_w + ix - 1) * pos_stride]);
    const float dy = __fsub_rn(mp[(long)((iy + 1) * res_w + ix) * pos_stride], mp[(long)((iy - 1) * res_w + ix) * pos_stride]);
    const float sx = dx > 0.f ? 1.f : (dx < 0.f ? -1.f : 0.f);
    const float sy = dy > 0.f ? 1.f : (dy < 0.f ? -1.f : 0.f);
    px = __fadd_rn(px, __fmul_rn(sx, 0.25f));
    py = __fadd_rn(py, __fmul_rn(sy, 0.25f));
  }
  px = __fadd_rn(px, 0.2f);
  py = __fadd_rn(py, 0.2f);
  const long o = (long)n * K + k;
  preds_hm[2 * o] = px;
  preds_hm[2 * o + 1] = py;
  maxval[o] = mv;
  idx_out[o] = id;
  // transformBoxInvert_batch
  const float ulx = pt1[2 * n], uly = pt1[2 * n + 1], brx = pt2[2 * n], bry = pt2[2 * n + 1];
  const float cx = __fdiv_rn(__fsub_rn(__fsub_rn(brx, 1.f), ulx), 2.f);
  const float cy = __fdiv_rn(__fsub_rn(__fsub_rn(bry, 1.f), uly), 2.f);
  const float sxz = __fmul_rn(__fsub_rn(brx, ulx), inp_ratio_hw);
  const float syz = __fsub_rn(bry, uly);
  const float lenH = fmaxf(sxz, syz);
  const float lenW = __fmul_rn(lenH, inp_ratio_wh);
  float tx = __fdiv_rn(__fmul_rn(px, lenH), (float)res_h);
  float ty = __fdiv_rn(__fmul_rn(py, lenH), (float)res_h);
  tx = __fsub_rn(tx, fmaxf(__fsub_rn(__fdiv_rn(__fsub_rn(lenW, 1.f), 2.f), cx), 0.f));
  ty = __fsub_rn(ty, fmaxf(__fsub_rn(__fdiv_rn(__fsub_rn(lenH, 1.f), 2.f), cy), 0.f));
  preds_img[2 * o] = __fadd_rn(tx, ulx);
  preds_img[2 * o + 1] = __fadd_rn(ty, uly);
}

}  // namespace

extern "C" int bp_heatmap_decode(bp_engine* e, const float* hm, long img_stride, long k_stride, long pos_stride, int n, int K,
                                 int res_h, int res_w, int inp_h, int inp_w, const float* pt1, const float* pt2,
                                 float* preds_hm, float* preds_img, float* maxval, int32_t* idx, void* stream) {
  if (!e || !hm || n <= 0 || K <= 0 || K > 1024 || !pt1 || !pt2 || !preds_hm || !preds_img || !maxval || !idx)
    return bp_fail(BP_ERR_INVALID, "bp_heatmap_decode: bad arguments");
  if (k_stride != 1 && pos_stride != 1) return bp_fail(BP_ERR_UNSUPPORTED, "bp_heatmap_decode: need NHWC (k_stride 1) or NCHW (pos_stride 1)");
  const int T = 1024;
  const float r_hw = (float)((double)inp_h / (double)inp_w), r_wh = (float)((double)inp_w / (double)inp_h);
  // NHWC maps (the engine's layout): slice every image's positions over S blocks; scratch is grow-only
  int S = 1;
  if (k_stride == 1 && K <= 64) {
    S = (4 * e->num_sms + n - 1) / n;
    S = std::max(1, std::min(S, std::min(16, res_h * res_w / 256)));
  }
  bp_engine::StreamScratch& sc = e->scratch_for(reinterpret_cast<cudaStream_t>(stream));
  if (S > 1) {
    const size_t need = (size_t)n * 16 * 64 * 8 + (size_t)n * 4;
    if (sc.hm_bytes < need) {
      // per-stream and grow-only: the block this replaces stays allocated (graphs captured earlier point into it)
      if (!e->grow(&sc.hm, &sc.hm_bytes, need)) return bp_fail(BP_ERR_CUDA, "bp_heatmap_decode: scratch allocation failed");
      // arrival counters start at zero (a blocking memset: allocation is not steady state and must not be captured)
      if (cudaMemset(sc.hm, 0, need) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
        return bp_fail(BP_ERR_CUDA, "bp_heatmap_decode: scratch memset failed (first use of a batch size inside a stream capture?)");
      sc.hm_n = n;
    }
  }
  // layout of the scratch: [n_alloc][16][64] float | [n_alloc][16][64] int | [n_alloc] counters
  const int na = sc.hm_n;
  float* pv = reinterpret_cast<float*>(sc.hm);
  int* pi = reinterpret_cast<int*>(pv + (size_t)na * 16 * 64);
  unsigned* ctr = reinterpret_cast<unsigned*>(pi + (size_t)na * 16 * 64);
  heatmap_decode_kernel<<<dim3(n, S), T, T * 8, reinterpret_cast<cudaStream_t>(stream)>>>(hm, img_stride, k_stride, pos_stride, K, res_h,
                                                                                         res_w, r_hw, r_wh, pt1, pt2, preds_hm,
                                                                                         preds_img, maxval, idx, pv, pi, ctr);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? BP_OK : bp_fail(BP_ERR_CUDA, cudaGetErrorString(err));
}

// ===================================================================================================
// a12  result records
// ===================================================================================================
namespace {
__global__ void pack_records_kernel(int n, int K, int image_index0, const float* box, const float* det_score,
                                    const float* keypoints, const float* kp_score, const float* proposal, const double* R,
                                    const double* t, const int32_t* status, bp_record* out) {
  const int i = blockIdx.x;
  if (i >= n) return;
  bp_record* r = out + i;
  if (threadIdx.x == 0) {
    r->image_index = image_index0 + i;
    r->status = status[i];
    for (int j = 0; j < 4; ++j) r->box[j] = box[4 * i + j];
    r->det_score = det_score[i];
    r->proposal_score = proposal[i];
    for (int j = 0; j < 9; ++j) r->R[j] = R[9 * i + j];
    for (int j = 0; j < 3; ++j) r->t[j] = t[3 * i + j];
  }
  for (int k = threadIdx.x; k < 50; k += blockDim.x) {
    const bool live = k < K;
    r->keypoints[3 * k] = live ? keypoints[((long)i * K + k) * 2] : 0.f;
    r->keypoints[3 * k + 1] = live ? keypoints[((long)i * K + k) * 2 + 1] : 0.f;
    r->keypoints[3 * k + 2] = live ? kp_score[(long)i * K + k] : 0.f;
  }
}
}  // namespace

extern "C" int bp_pack_records(bp_engine* e, int n, int K, int image_index0, const float* box, const float* det_score,
                               const float* keypoints, const float* kp_score, const float* proposal, const double* R,
                               const double* t, const int32_t* status, bp_record* out, void* stream) {
  if (!e || n <= 0 || K <= 0 || K > 50 || !out) return bp_fail(BP_ERR_INVALID, "bp_pack_records: bad arguments");
  pack_records_kernel<<<n, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(n, K, image_index0, box, det_score, keypoints,
                                                                           kp_score, proposal, R, t, status, out);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? BP_OK : bp_fail(BP_ERR_CUDA, cudaGetErrorString(err));
}
