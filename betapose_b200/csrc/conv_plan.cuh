// Host-side planning + launch of one conv_umma_kernel instance: picks the tile configuration, encodes the
// TMA descriptors once (buffers are owned by the engine, so addresses are stable) and replays the launch.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <string>

#include "conv_tcgen05.cuh"
#include "tmap.cuh"

namespace bp {

struct ConvDesc {
  // input activation, NHWC fp16 (for matrix mode: N=1, H=1, W=rows, C=K)
  const __half* x = nullptr;
  int N = 0, H = 0, W = 0, C = 0, x_pitch = 0;
  long x_row_pitch = 0, x_img_pitch = 0;  // elements between input rows / images, 0 = dense (W * x_pitch, H * row)
  // weights [Cout_pad][R][S][C] fp16 (row pitch w_pitch elements, >= R*S*C, multiple of 8), bias [Cout_pad] fp32
  const __half* w = nullptr;
  const float* bias = nullptr;
  int w_pitch = 0;
  int Cout = 0, Cout_pad = 0;
  int R = 1, S = 1, stride = 1, pad = 0;
  int pad_w = -1;  // horizontal padding, -1 = same as `pad` (the packed stem convolutions use 0, see net.cu)
  int stride_w = 0;  // horizontal stride, 0 = same as `stride` (the grouped stem steps 4 virtual pixels per GEMM row)
  int act = ACT_NONE;
  const __half* res = nullptr;
  int res_pitch = 0, res_mode = RES_NONE;
  void* out = nullptr;
  int out_pitch = 0, out_coff = 0, out_f32 = 0, store_mode = STORE_PLAIN;
  int force_block_n = 0;  // 0 = heuristic
  int force_cg = 0;       // 0 = heuristic, 1 = single CTA, 2 = CTA pairs (cta_group::2)
  int force_mt = 0;       // 0 = heuristic, 1 = 128-pixel tiles, 2 = 256-pixel tiles (narrow layers)
  int force_stages = 0;   // kept for the harness; the stage count follows from the tile configuration
  int num_sms = 148;
  bool pdl = true;        // launch with programmatic stream serialisation (off for nets that share the GPU, bp_net_set_share)
  float* sk_ws = nullptr;        // stream-K workspace (>= num_sms slots of 128 x 256 fp32) and flags (>= num_sms), owned by the net;
  unsigned* sk_flags = nullptr;  // null = stream-K off
  int force_sk = 0;              // 0 = heuristic, 1 = off, 2 = on (where supported)
  double real_k = 0;      // reduction length that counts as work (0 = R*S*C); the stems pad K with zero weights
  unsigned long long* trace = nullptr;  // bring-up harness: per-CTA clock stamps (ConvArgs::trace)
};

struct ConvPlan {
  alignas(64) CUtensorMap tmA;
  alignas(64) CUtensorMap tmB;
  alignas(64) CUtensorMap tmOut;  // valid when args.tma_store
  alignas(64) CUtensorMap tmRes;  // valid when args.tma_store && residual
  ConvArgs args;
  int block_n = 0, block_k = 0, stages = 0, grid = 0, cg = 1, mt = 1;
  bool pdl = true;
  bool streamk = false;
  int P = 0, Q = 0;
  double flops = 0;
};

template <int BN, int BK, int ST, int CG = 1, int NB = 4, int MT = 1, int SK = 0>
inline cudaError_t launch_cfg(const ConvPlan& pl, cudaStream_t st) {
  using Cfg = ConvCfg<BN, BK, ST, CG, NB, MT, SK>;
  static bool attr_done[64] = {};  // per instantiation and per device (the attribute belongs to the device's context)
  {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!attr_done[dev]) {
      cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<BN, BK, ST, CG, NB, MT, SK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::SMEM_BYTES);
      if (e != cudaSuccess) return e;
      attr_done[dev] = true;
    }
  }
  // programmatic dependent launch: this kernel's prologue overlaps the tail of the previous kernel in the stream
  // (conv_umma_kernel calls griddepcontrol.wait before it touches global memory)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.grid);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  static const bool no_pdl = getenv("BP_NO_PDL") != nullptr;  // experiment switch
  if (!no_pdl && pl.pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if constexpr (CG == 2) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = 2;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, conv_umma_kernel<BN, BK, ST, CG, NB, MT, SK>, pl.tmA, pl.tmB, pl.tmOut, pl.tmRes, pl.args);
}

// is there a kernel instantiation for this plan?  (mirrors the dispatch in conv_plan_launch)
// stream-K instantiations: the two 256-wide kernels (single CTA and CTA pairs) and the 128-wide single-CTA kernel
inline bool conv_plan_streamk_supported(int bn, int bk, int cg, int mt) {
  return mt == 1 && bk == 64 && ((bn == 256 && (cg == 1 || cg == 2)) || (bn == 128 && cg == 1));
}

inline bool conv_plan_supported(const ConvPlan& pl) {
  const int bn = pl.block_n, bk = pl.block_k;
  if (pl.streamk && !conv_plan_streamk_supported(bn, bk, pl.cg, pl.mt)) return false;
  if (pl.cg == 2) return bk == 64 && (bn == 256 || bn == 128);
  if (pl.mt == 4) return (bn == 64 && bk == 32) || (bn == 32 && bk == 32) || (bn == 32 && bk == 64);
  if (pl.mt == 2) return (bk == 64 && (bn == 128 || bn == 64 || bn == 32)) || (bk == 32 && (bn == 64 || bn == 32));
  if (bk == 64) return (bn == 256 && pl.stages == 3) || (bn == 128 && pl.stages == 5) || (bn == 64 && pl.stages == 6) || (bn == 32 && pl.stages == 9);
  return (bn == 64 && pl.stages == 13) || (bn == 32 && pl.stages == 16);
}

// Four configurations would fill the SM's shared memory to the last kilobyte with a four-buffer epilogue ring.  They run
// with three buffers instead, which leaves >= 17 KB per SM free: the small stage kernels of the OTHER batch in flight (PnP
// hypotheses / refit, decode: <= 5 KB per CTA, long-running fp64 chains) can then co-reside with any convolution CTA instead of
// holding an SM that a persistent, statically scheduled convolution grid has to wait for.
#ifndef BP_NB_TIGHT
#define BP_NB_TIGHT 3
#endif
constexpr int kNbTight = BP_NB_TIGHT;

inline cudaError_t conv_plan_launch(const ConvPlan& pl, cudaStream_t st) {
  if (pl.streamk) {
    if (pl.cg == 2 && pl.block_n == 256 && pl.block_k == 64) return launch_cfg<256, 64, 5, 2, kNbTight, 1, 1>(pl, st);
    if (pl.cg == 1 && pl.mt == 1 && pl.block_n == 256 && pl.block_k == 64 && pl.stages == 3) return launch_cfg<256, 64, 3, 1, 4, 1, 1>(pl, st);
    if (pl.cg == 1 && pl.mt == 1 && pl.block_n == 128 && pl.block_k == 64 && pl.stages == 5) return launch_cfg<128, 64, 5, 1, kNbTight, 1, 1>(pl, st);
    return cudaErrorInvalidConfiguration;
  }
  if (pl.cg == 2) {
    if (pl.block_n == 256 && pl.block_k == 64) return launch_cfg<256, 64, 5, 2, kNbTight>(pl, st);
    if (pl.block_n == 128 && pl.block_k == 64) return launch_cfg<128, 64, 6, 2>(pl, st);
    return cudaErrorInvalidConfiguration;
  }
  if (pl.mt == 4) {  // 512-pixel tiles (the narrowest, tile-count-bound layers)
    if (pl.block_n == 64 && pl.block_k == 32) return launch_cfg<64, 32, 4, 1, 4, 4>(pl, st);
    if (pl.block_n == 32 && pl.block_k == 32) return launch_cfg<32, 32, 5, 1, 4, 4>(pl, st);
    if (pl.block_n == 32 && pl.block_k == 64) return launch_cfg<32, 64, 2, 1, 4, 4>(pl, st);
    return cudaErrorInvalidConfiguration;
  }
  if (pl.mt == 2) {  // 256-pixel tiles (narrow layers)
    if (pl.block_n == 128 && pl.block_k == 64) return launch_cfg<128, 64, 3, 1, 4, 2>(pl, st);
    if (pl.block_n == 64 && pl.block_k == 64) return launch_cfg<64, 64, 4, 1, kNbTight, 2>(pl, st);
    if (pl.block_n == 32 && pl.block_k == 64) return launch_cfg<32, 64, 5, 1, 4, 2>(pl, st);
    if (pl.block_n == 64 && pl.block_k == 32) return launch_cfg<64, 32, 8, 1, kNbTight, 2>(pl, st);
    if (pl.block_n == 32 && pl.block_k == 32) return launch_cfg<32, 32, 10, 1, 4, 2>(pl, st);
    return cudaErrorInvalidConfiguration;
  }
#define BP_CASE(BN, BK, ST) \
  if (pl.block_n == BN && pl.block_k == BK && pl.stages == ST) return launch_cfg<BN, BK, ST>(pl, st);
  // as many stages as fit beside the epilogue ring: the loop TMA issue -> data lands -> MMA -> commit -> slot free
  // takes ~3000 cycles under load, and a CTA sustains (bytes in flight) / (that latency)
  BP_CASE(256, 64, 3)
  if (pl.block_n == 128 && pl.block_k == 64 && pl.stages == 5) return launch_cfg<128, 64, 5, 1, kNbTight>(pl, st);
  BP_CASE(64, 64, 6)
  BP_CASE(32, 64, 9)
  BP_CASE(64, 32, 13)
  BP_CASE(32, 32, 16)
#undef BP_CASE
  return cudaErrorInvalidConfiguration;
}

inline bool conv_plan_build(TmapApi& api, ConvPlan* pl, const ConvDesc& d, std::string* err) {
  const int pad_w = d.pad_w < 0 ? d.pad : d.pad_w;
  const int stride_w = d.stride_w > 0 ? d.stride_w : d.stride;
  const int P = (d.H + 2 * d.pad - d.R) / d.stride + 1;
  const int Q = (d.W + 2 * pad_w - d.S) / stride_w + 1;
  const int M = d.N * P * Q;
  const bool matrix = d.R == 1 && d.S == 1 && d.stride == 1 && stride_w == 1 && d.pad == 0 && pad_w == 0 && d.x_row_pitch == 0;
  // matrix mode may have a ragged K: TMA zero-fills the tail
  const int block_k = matrix ? (d.C >= 64 ? 64 : 32) : ((d.C % 64 == 0) ? 64 : 32);
  if (!matrix && d.C % 32 != 0) {
    if (err) *err = "im2col conv needs Cin % 32 == 0";
    return false;
  }
  const int K = d.R * d.S * d.C;
  const int num_kb = (K + block_k - 1) / block_k;
  const int m_tiles = (M + 127) / 128;

  // Tile configuration by a small cost model fitted to B200 measurements (tests/harness/conv_harness.cu, batch 64):
  // a persistent grid walks the tiles round-robin, so a launch takes ceil(tiles / CTAs) rounds of one tile each, and a
  // k-block costs ~500 + (BLOCK_K / 16) * bn / 2 cycles (bn = 64 / 128 / 256: 628 / 758 / 1000-1040 measured; the
  // second term is the tensor-pipe time, the first is the TMA -> MMA -> commit loop latency spread over the stages that
  // fit in shared memory).  Wider tiles therefore win unless they leave most of the machine idle.
  // CTA pairs (cta_group::2, M = 256): each CTA stages only half of B, so five 32 KB stages fit instead of three
  // 48 KB ones; measured +6-8 % on every 256-wide 3x3 layer (K >= 1152: 1260 vs 1168, 1411 vs 1334, 1404 vs 1319
  // TFLOP/s) and -5-20 % on the short-K 1x1 layers, where the cluster launch and cross-CTA hand-offs do not amortise.
  int bn = 0, cg = 0;
  {
    int cap = 32;
    while (cap < d.Cout && cap < 256) cap *= 2;
    if (block_k == 32) cap = std::min(cap, 64);
    double best = 0;
    for (int c = cap; c >= 32; c /= 2) {
      if (d.force_block_n && c != d.force_block_n) continue;
      if (d.Cout_pad % c) continue;
      const int g = d.force_cg ? d.force_cg : ((c == 256 && block_k == 64 && num_kb >= 18 && m_tiles >= d.num_sms) ? 2 : 1);
      if (g == 2 && (block_k != 64 || c < 128 || d.num_sms < 2)) continue;
      const long tiles = (long)((m_tiles + g - 1) / g) * ((d.Cout + c - 1) / c);
      const long units = g == 2 ? d.num_sms / 2 : d.num_sms;
      const long rounds = (tiles + units - 1) / units;
      const double per_kb = 500.0 + (block_k / 16) * (c / 2.0);
      const double cost = rounds * (num_kb * per_kb + 400.0 + 200.0 * ((c + 63) / 64));
      if (!bn || cost < 0.97 * best) {  // a narrower tile must win by 3 %: wider tiles read A fewer times
        bn = c;
        cg = g;
        best = cost;
      }
    }
    if (!bn) {
      if (err) *err = "no tile configuration fits (forced block_n / cg?)";
      return false;
    }
  }
  if (d.Cout_pad % bn != 0) {
    if (err) *err = "Cout_pad must be a multiple of BLOCK_N";
    return false;
  }
  // 256-pixel tiles for the narrow, TMA-stored layers with plenty of tiles: they are bound by the issue overhead per
  // tile and per k-block of the producer / MMA warps, which a 256-row tile halves per pixel
  const bool tma_store_ok = !d.out_f32 && d.store_mode == STORE_PLAIN && d.Cout % 8 == 0 && d.out_pitch % 8 == 0 &&
                            d.out_coff % 8 == 0 && (!d.res || d.res_pitch % 8 == 0);
  // (measured: stem 0.54 -> 0.41 ms, 3x3 32->64 0.35 -> 0.21 ms, 3x3 64->64 0.050 -> 0.035 ms, 3x3 64->128 @104 0.125 -> 0.113 ms;
  // no gain on 128-wide 1x1 layers)
  int mt = d.force_mt ? d.force_mt
                      : ((cg == 1 && tma_store_ok && m_tiles >= 4 * d.num_sms && (bn <= 64 || (bn == 128 && block_k == 64 && num_kb >= 9))) ? 2 : 1);
  // 512-pixel tiles for the very large, very narrow layers (measured: stem 0.41 -> 0.39 ms, 3x3 32->64 0.23 / 0.25 -> 0.19 /
  // 0.20 ms, 1x1 64->32 @208 0.097 -> 0.088 ms)
  if (!d.force_mt && mt == 2 && m_tiles >= 16 * d.num_sms && ((bn <= 64 && block_k == 32) || bn == 32)) mt = 4;
  if (mt == 2 && (cg != 1 || bn > 128 || (bn == 128 && block_k != 64) || !tma_store_ok)) mt = 1;
  if (mt == 4 && (cg != 1 || bn > 64 || (bn == 64 && block_k != 32) || !tma_store_ok)) mt = 1;
  const int st = mt == 4 ? (block_k == 32 ? (bn == 64 ? 4 : 5) : 2)
                 : mt == 2 ? (block_k == 64 ? (bn == 128 ? 3 : bn == 64 ? 4 : 5) : (bn == 64 ? 8 : 10))
                 : cg == 2 ? (bn == 256 ? 5 : 6)
                           : (block_k == 64 ? (bn == 256 ? 3 : bn == 128 ? 5 : bn == 64 ? 6 : 9) : (bn == 64 ? 13 : 16));
  const int m_tiles_cta = (M + 128 * mt - 1) / (128 * mt);  // tiles as the kernel walks them

  pl->block_n = bn;
  pl->pdl = d.pdl;
  pl->block_k = block_k;
  pl->stages = st;
  pl->cg = cg;
  pl->mt = mt;
  pl->P = P;
  pl->Q = Q;
  // tiles whose channels are all padding are never launched
  const int n_tiles_live = (d.Cout + bn - 1) / bn;
  // persistent: one CTA per SM (pairs: one cluster of 2 per TPC)
  pl->grid = cg == 2 ? 2 * std::min(((m_tiles + 1) / 2) * n_tiles_live, d.num_sms / 2) : std::min(m_tiles_cta * n_tiles_live, d.num_sms);
  pl->flops = 2.0 * M * (double)d.Cout * (d.real_k > 0 ? d.real_k : (double)K);
  // Stream-K (conv_tcgen05.cuh: SkPlan): when the tiles do not fill the last round of the persistent grid, cut the (tile,
  // k-block) units into one contiguous range per cluster.  Worth it when the k-blocks saved per cluster outweigh the
  // hand-over of one partial accumulator (128 x BLOCK_N fp32 written by one CTA and read by its neighbour, ~2 x 128 KB
  // through L2 for BLOCK_N = 256, i.e. about 5 (single CTA) to 8 (pairs) k-blocks' worth of staging traffic).
  pl->streamk = false;
  if (d.sk_ws && d.sk_flags && d.force_sk != 1 && conv_plan_streamk_supported(bn, block_k, cg, mt)) {
    const long units = cg == 2 ? d.num_sms / 2 : d.num_sms;
    const long tiles = (long)((m_tiles + cg - 1) / cg) * n_tiles_live;
    if (tiles >= units && units > 1) {
      const long rounds = (tiles + units - 1) / units;
      const double saved = (double)rounds * num_kb - (double)tiles * num_kb / units;  // k-blocks per cluster
      const double handover = 2.0 * 128 * bn * 4 / (double)((128 + bn / cg) * block_k * 2);
      if (d.force_sk == 2 || saved >= 1.5 * handover + 2.0) pl->streamk = true;
    }
  }

  ConvArgs& a = pl->args;
  a.M = M;
  a.n_tiles = n_tiles_live;
  a.m_tiles = m_tiles_cta;
  a.num_kb = num_kb;
  a.a_im2col = matrix ? 0 : 1;
  a.P = P;
  a.Q = Q;
  a.stride = d.stride;
  a.stride_w = stride_w;
  a.pad = d.pad;
  a.pad_w = pad_w;
  a.C = d.C;
  a.S = d.S;
  a.cblocks = d.C / block_k;
  a.Cout = d.Cout;
  a.act = d.act;
  a.res_mode = d.res ? d.res_mode : RES_NONE;
  a.store_mode = d.store_mode;
  a.out_f32 = d.out_f32;
  a.tma_store = (!d.out_f32 && d.store_mode == STORE_PLAIN && d.Cout % 8 == 0 && d.out_pitch % 8 == 0 && d.out_coff % 8 == 0 &&
                 (!d.res || d.res_pitch % 8 == 0))
                    ? 1
                    : 0;
  static const bool no_staged = getenv("BP_NO_STAGED_STORE") != nullptr;  // experiment switch: per-thread stores as in round 1
  a.staged_store = (!a.tma_store && !no_staged && !d.out_f32 && !d.res && d.store_mode != STORE_PLAIN && d.Cout % 8 == 0 && d.out_pitch % 8 == 0 &&
                    d.out_coff % 8 == 0 && (d.store_mode != STORE_PIXSHUF2 || d.Cout % 32 == 0))
                       ? 1
                       : 0;
  a.out_pitch = d.out_pitch;
  a.out_coff = d.out_coff;
  a.res_pitch = d.res_pitch;
  if (M >= (1 << 26) || P * Q >= (1 << 18) || (long)m_tiles * n_tiles_live >= (1 << 26)) {
    if (err) *err = "problem too large for the kernel's tile arithmetic (M < 2^26 pixels, P*Q < 2^18)";
    return false;
  }
  auto magic = [](int dv) { return (unsigned long long)(((1ull << 44) + (unsigned long long)dv - 1) / (unsigned long long)dv); };
  a.mul_nt = magic(n_tiles_live);
  a.mul_pq = magic(P * Q);
  a.mul_q = magic(Q);
  a.bias = d.bias;
  a.res = d.res;
  a.out = d.out;
  a.sk_ws = d.sk_ws;
  a.sk_flags = d.sk_flags;
  a.trace = d.trace;
  static const int sk_align = getenv("BP_SK_ALIGN") ? atoi(getenv("BP_SK_ALIGN")) : 0;
  a.sk_align = sk_align;

  if (matrix) {
    if (!make_tmap_2d(api, &pl->tmA, d.x, (uint64_t)M, (uint64_t)d.C, (uint64_t)d.x_pitch, mt == 4 ? 256 : 128 * mt, block_k, err))
      return false;
  } else {
    if (!make_tmap_im2col(api, &pl->tmA, d.x, d.N, d.H, d.W, d.C, d.x_pitch, d.R, d.S, d.stride, d.pad, pad_w, block_k, err,
                          d.x_row_pitch, d.x_img_pitch, mt == 4 ? 256 : 128 * mt, stride_w))
      return false;
  }
  if (!make_tmap_2d(api, &pl->tmB, d.w, (uint64_t)d.Cout_pad, (uint64_t)K, (uint64_t)d.w_pitch, bn / cg, block_k, err))
    return false;
  pl->tmOut = pl->tmB;  // placeholders keep the kernel parameters well-formed when the TMA epilogue is off
  pl->tmRes = pl->tmB;
  if (a.tma_store) {
    // dense [M pixels, Cout channels] fp16 boxes of 128 x CHUNK; rows >= M and channels >= Cout are clipped by TMA
    const int chunk = bn >= 64 ? 64 : 32;
    const __half* obase = reinterpret_cast<const __half*>(d.out) + d.out_coff;
    if (!make_tmap_2d(api, &pl->tmOut, obase, (uint64_t)M, (uint64_t)d.Cout, (uint64_t)d.out_pitch, 128, chunk, err)) return false;
    if (d.res && !make_tmap_2d(api, &pl->tmRes, d.res, (uint64_t)M, (uint64_t)d.Cout, (uint64_t)d.res_pitch, 128, chunk, err))
      return false;
  }
  return true;
}

}  // namespace bp
