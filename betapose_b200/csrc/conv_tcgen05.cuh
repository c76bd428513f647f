// Implicit-GEMM convolution for sm_100a: D[M = B*P*Q pixels, N = Cout] = im2col(X)[M, K] * W[N, K]^T
//   A (activations, NHWC fp16)  : TMA im2col loads (3x3 / strided) or TMA 2-D tiles (1x1 s1, explicit matrices)
//   B (weights, [Cout][R][S][Cin] fp16, BN folded) : TMA 2-D tiles
//   both land in 128B- (or 64B-) swizzled shared memory, K-major, and feed tcgen05.mma (M=128, N=BLOCK_N, K=16)
//   accumulators live in TMEM; the epilogue reads them back with tcgen05.ld and fuses
//   bias + activation + residual add + {plain | nearest-x2 | pixel-shuffle} store.
// One CTA = one 128 x BLOCK_N output tile, 4 warps: warp0 = TMA producer, warp1 = MMA issuer (+TMEM owner),
// all four warps drain TMEM (warp w owns lanes 32w..32w+31). Several CTAs co-reside per SM so one CTA's
// epilogue overlaps another's main loop.
//
// Replaces, for the reference, every nn.Conv2d+BatchNorm2d+activation (+shortcut) block of
// 3_6Dpose_estimator/yolo/darknet.py:252-259,333-340 and KPD/src/models/layers/SE_Resnet.py:11-40, DUC.py:12-22.
#pragma once
#include "bp_ptx.cuh"

namespace bp {

enum : int { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2, ACT_SIGMOID = 3 };
enum : int { RES_NONE = 0, RES_AFTER_ACT = 1, RES_BEFORE_ACT = 2 };
enum : int { STORE_PLAIN = 0, STORE_UPSAMPLE2 = 1, STORE_PIXSHUF2 = 2 };

struct ConvArgs {
  int M;         // output pixels (B*P*Q)
  int n_tiles;   // Cout_pad / BLOCK_N
  int num_kb;    // K / BLOCK_K
  int a_im2col;  // 1: A through the im2col tensor map over NHWC; 0: A is a row-major [M, K] matrix
  int P, Q;      // output height / width
  int stride, pad;
  int S;         // filter width
  int cblocks;   // Cin / BLOCK_K
  int Cout;      // real output channels
  int act;
  int res_mode;
  int store_mode;
  int out_f32;
  int out_pitch;  // elements between consecutive output pixels
  int out_coff;   // channel offset inside the output pixel
  int res_pitch;
  const float* bias;   // [n_tiles * BLOCK_N], zero padded
  const __half* res;   // residual, same pixel order as the output, res_pitch elements per pixel
  void* out;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_LEAKY: return v > 0.f ? v : 0.1f * v;
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

template <int BLOCK_N, int BLOCK_K, int STAGES>
struct ConvCfg {
  static constexpr int BLOCK_M = 128;
  static constexpr int SWZ = BLOCK_K * 2;  // bytes per smem row == swizzle span
  static constexpr int A_BYTES = BLOCK_M * SWZ;
  static constexpr int B_BYTES = BLOCK_N * SWZ;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + slack for 1024B alignment
};

template <int BLOCK_N, int BLOCK_K, int STAGES>
__global__ void __launch_bounds__(128)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvArgs p) {
  using Cfg = ConvCfg<BLOCK_N, BLOCK_K, STAGES>;
  static_assert(BLOCK_N == 32 || BLOCK_N == 64 || BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");
  static_assert(BLOCK_K == 64 || BLOCK_K == 32, "BLOCK_K");

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_bias[BLOCK_N];

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int tile = blockIdx.x;
  const int n_tile = tile % p.n_tiles;
  const int m_tile = tile / p.n_tiles;
  const int m0 = m_tile * Cfg::BLOCK_M;
  const int n0 = n_tile * BLOCK_N;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<BLOCK_N>(&tmem_base_slot);
  for (int i = threadIdx.x; i < BLOCK_N; i += 128) s_bias[i] = p.bias[n0 + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int w0 = 0, h0 = 0, img = 0;
    if (p.a_im2col) {
      const int pq = p.P * p.Q;
      img = m0 / pq;
      const int rem = m0 - img * pq;
      const int op = rem / p.Q;
      const int oq = rem - op * p.Q;
      w0 = oq * p.stride - p.pad;
      h0 = op * p.stride - p.pad;
    }
    for (int kb = 0; kb < p.num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (lane == 0) {
        uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        if (p.a_im2col) {
          const int tap = kb / p.cblocks;
          const int cb = kb - tap * p.cblocks;
          const int fr = tap / p.S;
          const int fs = tap - fr * p.S;
          tma_load_im2col_4d(&tmA, &full_bar[s], sa, cb * BLOCK_K, w0, h0, img, (uint16_t)fs, (uint16_t)fr);
        } else {
          tma_load_2d(&tmA, &full_bar[s], sa, kb * BLOCK_K, m0);
        }
        tma_load_2d(&tmB, &full_bar[s], sb, kb * BLOCK_K, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(BLOCK_N);
    for (int kb = 0; kb < p.num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::A_BYTES;
        const uint64_t da = umma_smem_desc<Cfg::SWZ>(sa);
        const uint64_t db = umma_smem_desc<Cfg::SWZ>(sb);
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          // advance 16 elements (32 B) along K inside the swizzle span: +2 in the (addr>>4) field
          umma_f16(tmem_acc, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above have read it
        if (kb == p.num_kb - 1) umma_commit(&tmem_full_bar);
      }
      __syncwarp();
    }
  }

  // ---------------------------------------------------------------- epilogue (all 4 warps)
  mbar_wait(&tmem_full_bar, 0);
  tc_fence_after();

  const int row = m0 + warp * 32 + lane;
  const bool row_ok = row < p.M;
  int img = 0, op = 0, oq = 0;
  if (p.store_mode != STORE_PLAIN) {
    const int pq = p.P * p.Q;
    img = row / pq;
    const int rem = row - img * pq;
    op = rem / p.Q;
    oq = rem - op * p.Q;
  }

#pragma unroll 1
  for (int c = 0; c < BLOCK_N / 32; ++c) {
    const int ch0 = n0 + c * 32;
    if (ch0 >= p.Cout) break;  // warp-uniform
    uint32_t acc[32];
    tmem_ld_32x32(tmem_acc + ((warp * 32u) << 16) + uint32_t(c * 32), acc);
    tmem_ld_wait();
    if (!row_ok) continue;

    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) + s_bias[c * 32 + j];

    if (p.res_mode != RES_NONE) {
      const __half* rp = p.res + (size_t)row * p.res_pitch + ch0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (ch0 + g * 8 + 8 <= p.Cout) {
          const uint4 rv = __ldg(reinterpret_cast<const uint4*>(rp + g * 8));
          const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(rh[j]);
            if (p.res_mode == RES_BEFORE_ACT) {
              v[g * 8 + 2 * j] = apply_act(v[g * 8 + 2 * j] + f.x, p.act);
              v[g * 8 + 2 * j + 1] = apply_act(v[g * 8 + 2 * j + 1] + f.y, p.act);
            } else {
              v[g * 8 + 2 * j] = apply_act(v[g * 8 + 2 * j], p.act) + f.x;
              v[g * 8 + 2 * j + 1] = apply_act(v[g * 8 + 2 * j + 1], p.act) + f.y;
            }
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
    }

    if (p.out_f32) {
      // network heads: fp32, plain store only
      float* op32 = reinterpret_cast<float*>(p.out) + (size_t)row * p.out_pitch + p.out_coff + ch0;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (ch0 + g * 4 + 4 <= p.Cout) {
          *reinterpret_cast<float4*>(op32 + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (ch0 + g * 4 + j < p.Cout) op32[g * 4 + j] = v[g * 4 + j];
        }
      }
      continue;
    }

    __half* o16 = reinterpret_cast<__half*>(p.out);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int ch = ch0 + g * 8;
      if (ch >= p.Cout) break;
      uint4 pk;
      __half2* ph2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j) ph2[j] = __floats2half2_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
      if (p.store_mode == STORE_PLAIN) {
        __half* dst = o16 + (size_t)row * p.out_pitch + p.out_coff + ch;
        if (ch + 8 <= p.Cout) {
          *reinterpret_cast<uint4*>(dst) = pk;
        } else {
          const __half* ph1 = reinterpret_cast<const __half*>(&pk);
          for (int j = 0; j < 8 && ch + j < p.Cout; ++j) dst[j] = ph1[j];
        }
      } else if (p.store_mode == STORE_UPSAMPLE2) {
        // nearest x2: this pixel lands on a 2x2 block of the [N, 2P, 2Q, *] destination
        const size_t base = ((size_t)img * (2 * p.P) + 2 * op) * (2 * p.Q) + 2 * oq;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx)
            *reinterpret_cast<uint4*>(o16 + (base + (size_t)dy * (2 * p.Q) + dx) * p.out_pitch + p.out_coff + ch) =
                pk;
      } else {
        // PixelShuffle(2): weight rows were pre-permuted to o' = sub*(Cout/4) + c, sub = 2*i + j
        const int c4 = p.Cout >> 2;
        const int sub = ch / c4;
        const int cc = ch - sub * c4;
        const size_t pix = ((size_t)img * (2 * p.P) + 2 * op + (sub >> 1)) * (2 * p.Q) + 2 * oq + (sub & 1);
        *reinterpret_cast<uint4*>(o16 + pix * p.out_pitch + p.out_coff + cc) = pk;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BLOCK_N>(tmem_acc);
}

}  // namespace bp
