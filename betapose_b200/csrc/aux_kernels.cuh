// Bandwidth-bound helper kernels around the tcgen05 conv: NHWC fp16, 16-byte (8-channel) vector accesses,
// one thread per 8-channel group, grid-stride free (grids are sized exactly, a few waves of 148 SMs).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace bp {

struct TView {  // NHWC view: element (n,h,w,c) at ptr[((n*H+h)*W+w)*pitch + c]   (ptr already includes coff)
  void* ptr;
  int H, W, C, pitch;
};

__device__ __forceinline__ uint4 ldg16(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(__half* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }

__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
  uint4 r;
  __half2* rr = reinterpret_cast<__half2*>(&r);
  const __half2* aa = reinterpret_cast<const __half2*>(&a);
  const __half2* bb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) rr[i] = __hmax2(aa[i], bb[i]);
  return r;
}

// MaxPool2d(kernel 3, stride 2, pad 1)  -- KPD/src/models/layers/SE_Resnet.py:59
__global__ void maxpool3x3s2_kernel(TView in, TView out, int N) {
  const int c8 = out.C >> 3;
  const long total = (long)N * out.H * out.W * c8;
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = idx % c8;
  long pix = idx / c8;
  const int q = pix % out.W;
  pix /= out.W;
  const int p = pix % out.H;
  const int n = pix / out.H;
  const __half* ip = reinterpret_cast<const __half*>(in.ptr);
  uint4 m;
  bool first = true;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int h = 2 * p - 1 + dy;
    if (h < 0 || h >= in.H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int w = 2 * q - 1 + dx;
      if (w < 0 || w >= in.W) continue;
      const uint4 v = ldg16(ip + (((long)n * in.H + h) * in.W + w) * in.pitch + cg * 8);
      m = first ? v : hmax8(m, v);
      first = false;
    }
  }
  stg16(reinterpret_cast<__half*>(out.ptr) + (((long)n * out.H + p) * out.W + q) * out.pitch + cg * 8, m);
}

// AdaptiveAvgPool2d(1): mean over H*W per (n, c), fp32 accumulate, fp16 result [N, C]  -- SE_module.py:7,16
// grid (C/256, N, S), block 256 = 32 channel-groups x 8 pixel lanes.  The H*W pixels of an image are cut into S
// slices so that even a 64-image batch fills the machine; every block leaves its partial sums in `scratch`
// ([N][S][C] fp32) and the last block to finish an (image, channel-block) adds the S partials in a fixed order
// (bit-reproducible: no floating-point atomics) and writes the fp16 mean.
// The ORDER of the summation is a property of the tensor shape alone: the H*W positions are always cut into S0 "virtual"
// slices (S0 from H*W only); the launch uses S <= S0 blocks per image, block sl taking the virtual slices sl, sl + S, ...,
// and the S0 partials are added in slice order.  So a frame's result does not depend on the batch it arrives in.
__global__ void global_avgpool_kernel(TView in, __half* out, int out_pitch, float* scratch, unsigned* counters, int S, int S0) {
  __shared__ float red[8][32][8];
  __shared__ bool s_last;
  const int n = blockIdx.y;
  const int cg = blockIdx.x * 32 + (threadIdx.x & 31);
  const int pl = threadIdx.x >> 5;
  const int HW = in.H * in.W;
  const int chunk = (HW + S0 - 1) / S0;
  const bool live = cg * 8 < in.C;
  float* part = scratch + ((long)n * S0) * in.C + cg * 8;
  for (int vs = blockIdx.z; vs < S0; vs += S) {
    const int px0 = vs * chunk, px1 = min(HW, px0 + chunk);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (live) {
      const __half* ip = reinterpret_cast<const __half*>(in.ptr) + (long)n * HW * in.pitch + cg * 8;
      for (int px = px0 + pl; px < px1; px += 8) {
        const uint4 v = ldg16(ip + (long)px * in.pitch);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __half22float2(h[i]);
          acc[2 * i] += f.x;
          acc[2 * i + 1] += f.y;
        }
      }
    }
    __syncthreads();  // the previous virtual slice's partials have been read out of `red`
#pragma unroll
    for (int i = 0; i < 8; ++i) red[pl][threadIdx.x & 31][i] = acc[i];
    __syncthreads();
    if (pl == 0 && live) {
      float4 lo, hi;
      float* f = &lo.x;
      float* g = &hi.x;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
          a += red[l][threadIdx.x & 31][i];
          b += red[l][threadIdx.x & 31][4 + i];
        }
        f[i] = a;
        g[i] = b;
      }
      float* dst = part + (long)vs * in.C;
      *reinterpret_cast<float4*>(dst) = lo;
      *reinterpret_cast<float4*>(dst + 4) = hi;
    }
  }
  if (pl == 0 && live) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* ctr = counters + (long)n * gridDim.x + blockIdx.x;
    const unsigned old = atomicAdd(ctr, 1u);
    s_last = old == (unsigned)S - 1u;
    if (s_last) *ctr = 0u;  // ready for the next launch
  }
  __syncthreads();
  if (s_last && pl == 0 && live) {
    __threadfence();
    float sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < S0; ++k) {
      const float4 lo = __ldcg(reinterpret_cast<const float4*>(part + (long)k * in.C));
      const float4 hi = __ldcg(reinterpret_cast<const float4*>(part + (long)k * in.C + 4));
      sum[0] += lo.x; sum[1] += lo.y; sum[2] += lo.z; sum[3] += lo.w;
      sum[4] += hi.x; sum[5] += hi.y; sum[6] += hi.z; sum[7] += hi.w;
    }
    uint4 o;
    __half* oh = reinterpret_cast<__half*>(&o);
    const float inv = 1.f / (float)HW;
#pragma unroll
    for (int i = 0; i < 8; ++i) oh[i] = __float2half_rn(sum[i] * inv);
    stg16(out + (long)n * out_pitch + cg * 8, o);
  }
}

// relu(y * gate[n,c] + skip)  -- SE_module.py:19 + SE_Resnet.py:38-40
__global__ void scale_add_relu_kernel(TView y, const __half* gates, int gate_pitch, TView skip, TView out, int N) {
  const int c8 = y.C >> 3;
  const long total = (long)N * y.H * y.W * c8;
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = idx % c8;
  const long pix = idx / c8;
  const int n = pix / ((long)y.H * y.W);
  const uint4 a = ldg16(reinterpret_cast<const __half*>(y.ptr) + pix * y.pitch + cg * 8);
  const uint4 s = ldg16(gates + (long)n * gate_pitch + cg * 8);
  const uint4 k = ldg16(reinterpret_cast<const __half*>(skip.ptr) + pix * skip.pitch + cg * 8);
  uint4 o;
  const __half2* ah = reinterpret_cast<const __half2*>(&a);
  const __half2* sh = reinterpret_cast<const __half2*>(&s);
  const __half2* kh = reinterpret_cast<const __half2*>(&k);
  __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __half22float2(ah[i]), fs = __half22float2(sh[i]), fk = __half22float2(kh[i]);
    oh[i] = __floats2half2_rn(fmaxf(fa.x * fs.x + fk.x, 0.f), fmaxf(fa.y * fs.y + fk.y, 0.f));
  }
  stg16(reinterpret_cast<__half*>(out.ptr) + pix * out.pitch + cg * 8, o);
}

// PixelShuffle(2): out[n, 2h+i, 2w+j, c] = in[n, h, w, 4c + 2i + j]  -- FastPose.py:21,30
// one thread = one input pixel x 32 input channels -> four 8-channel output vectors.
__global__ void pixel_shuffle2_kernel(TView in, TView out, int N) {
  const int c32 = in.C >> 5;
  const long total = (long)N * in.H * in.W * c32;
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = idx % c32;
  long pix = idx / c32;
  const int w = pix % in.W;
  pix /= in.W;
  const int h = pix % in.H;
  const int n = pix / in.H;
  const __half* ip = reinterpret_cast<const __half*>(in.ptr) + (((long)n * in.H + h) * in.W + w) * in.pitch + cg * 32;
  __half v[32];
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(&v[i * 8]) = ldg16(ip + i * 8);
  __half* op = reinterpret_cast<__half*>(out.ptr);
#pragma unroll
  for (int sub = 0; sub < 4; ++sub) {
    uint4 o;
    __half* oh = reinterpret_cast<__half*>(&o);
#pragma unroll
    for (int c = 0; c < 8; ++c) oh[c] = v[4 * c + sub];
    const long opix = ((long)n * out.H + 2 * h + (sub >> 1)) * out.W + 2 * w + (sub & 1);
    stg16(op + opix * out.pitch + cg * 8, o);
  }
}

// fall-backs: nearest x2 into a (possibly wider) destination, channel copy, element-wise add
__global__ void upsample2_kernel(TView in, TView out, int N) {
  const int c8 = in.C >> 3;
  const long total = (long)N * out.H * out.W * c8;
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = idx % c8;
  long pix = idx / c8;
  const int q = pix % out.W;
  pix /= out.W;
  const int p = pix % out.H;
  const int n = pix / out.H;
  const uint4 v = ldg16(reinterpret_cast<const __half*>(in.ptr) + (((long)n * in.H + (p >> 1)) * in.W + (q >> 1)) * in.pitch + cg * 8);
  stg16(reinterpret_cast<__half*>(out.ptr) + (((long)n * out.H + p) * out.W + q) * out.pitch + cg * 8, v);
}
__global__ void copy_channels_kernel(TView in, TView out, int N) {
  const int c8 = in.C >> 3;
  const long total = (long)N * in.H * in.W * c8;
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = idx % c8;
  const long pix = idx / c8;
  stg16(reinterpret_cast<__half*>(out.ptr) + pix * out.pitch + cg * 8,
        ldg16(reinterpret_cast<const __half*>(in.ptr) + pix * in.pitch + cg * 8));
}
__global__ void add_kernel(TView a, TView b, TView out, int N) {
  const int c8 = a.C >> 3;
  const long total = (long)N * a.H * a.W * c8;
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = idx % c8;
  const long pix = idx / c8;
  const uint4 x = ldg16(reinterpret_cast<const __half*>(a.ptr) + pix * a.pitch + cg * 8);
  const uint4 y = ldg16(reinterpret_cast<const __half*>(b.ptr) + pix * b.pitch + cg * 8);
  uint4 o;
  const __half2* xh = reinterpret_cast<const __half2*>(&x);
  const __half2* yh = reinterpret_cast<const __half2*>(&y);
  __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fx = __half22float2(xh[i]), fy = __half22float2(yh[i]);
    oh[i] = __floats2half2_rn(fx.x + fy.x, fx.y + fy.y);
  }
  stg16(reinterpret_cast<__half*>(out.ptr) + pix * out.pitch + cg * 8, o);
}

}  // namespace bp
