// Implicit-GEMM convolution for sm_100a: D[M = B*P*Q pixels, N = Cout] = im2col(X)[M, K] * W[N, K]^T
//   A (activations, NHWC fp16)  : TMA im2col loads (3x3 / strided) or TMA 2-D tiles (1x1 s1, explicit matrices)
//   B (weights, [Cout][R][S][Cin] fp16, BN folded) : TMA 2-D tiles
//   both land in 128B- (or 64B-) swizzled shared memory, K-major, and feed tcgen05.mma (M=128, N=BLOCK_N, K=16);
//   accumulators live in TMEM, double buffered (2 x BLOCK_N columns).
//
// Persistent, warp-specialised: one CTA per SM loops over output tiles (n-tile fastest, so the CTAs running at the
// same time share A rows in L2 and all share the weights).
//   warp 0    : TMA producer (A + B per k-block into a STAGES-deep mbarrier ring; warp-uniform loop, elect.sync issue)
//   warp 1    : MMA issuer (warp-uniform loop, elect.sync issue) + TMEM owner
//   warps 2-9 : epilogue.  tcgen05.ld (issued one chunk ahead of its use) -> bias + activation (+ residual) -> fp16 ->
//               swizzled smem staging, 64 output channels at a time, then an ARRIVE on the chunk's named barrier -- the
//               epilogue warps never wait for a store.  The epilogue of tile i overlaps the main loop of tile i+1 through
//               the second TMEM accumulator.
//   warp 10   : store warp.  Waits on the chunk's named barrier, issues the TMA store (cp.async.bulk.tensor), waits until
//               the store has read the ring buffer and hands the buffer on at once: by loading the residual tile of the
//               chunk that will use it next (TMA, NBUF chunks ahead across tile boundaries), or, without a residual, by a
//               plain arrive on the same mbarrier.
//   Stores that are not a dense [pixels, channels] box go out as per-thread stores (fp32 heads) or, for the fused
//   nearest-x2 upsample / PixelShuffle, staged through the same ring and written row-wise coalesced.
// Variants (template parameters): CG = 2 runs CTA pairs (cluster of 2, tcgen05 cta_group::2, M = 256, each CTA stages
// half of B); MT = 2 gives a CTA 256-pixel tiles (two M = 128 MMAs per K step against one B tile) for narrow layers.
// Launched with programmatic stream serialisation: griddepcontrol.wait sits after the set-up.
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
  int n_tiles;   // live N tiles (ceil(Cout / BLOCK_N))
  int m_tiles;   // ceil(M / 128)
  int num_kb;    // K / BLOCK_K
  int a_im2col;  // 1: A through the im2col tensor map over NHWC; 0: A is a row-major [M, K] matrix
  int P, Q;      // output height / width
  int stride, pad;  // vertical stride / padding
  int stride_w;     // horizontal stride (== stride except for the grouped stem convolution)
  int pad_w;        // horizontal padding (== pad except for the packed stem convolutions)
  int C;            // input channels per filter tap as seen by the im2col map (multiple of BLOCK_K)
  int S;         // filter width
  int cblocks;   // Cin / BLOCK_K
  int Cout;      // real output channels
  int act;
  int res_mode;
  int store_mode;
  int out_f32;
  int tma_store;  // 1: dense fp16 box -> staged TMA store (+ TMA residual); 0: per-thread stores
  int staged_store;  // 1 (only when !tma_store): fp16 result staged in the epilogue ring, written out row-wise coalesced (fused
                     // nearest-x2 upsample / PixelShuffle); 0: every thread stores its own pixel's channels
  int out_pitch;  // elements between consecutive output pixels
  int out_coff;   // channel offset inside the output pixel
  int res_pitch;
  // exact division of x < 2^26 by a constant d < 2^18 (quotient < 2^20) as (x * ceil(2^44 / d)) >> 44 (host: conv_plan_build); the tile
  // walk decodes (tile -> m-tile, n-tile) and (pixel -> image, row, column) once per tile, and on the 1-3 k-block
  // layers a hardware-less integer division (~25 instructions) per decode is a visible share of the producer's time
  unsigned long long mul_nt, mul_pq, mul_q;
  const float* bias;   // [n_tiles * BLOCK_N], zero padded
  const __half* res;   // residual, same pixel order as the output, res_pitch elements per pixel
  void* out;
  // stream-K (SK kernels): fp32 partial accumulators [clusters * CG][128][BLOCK_N] and one hand-over flag per slot
  float* sk_ws;
  unsigned* sk_flags;
  int sk_align;  // experiment switch (BP_SK_ALIGN): whole-tile ranges
  // bring-up harness only (null in the library): per CTA kTraceSlots clock64() stamps -- [0] start (after griddepcontrol.wait),
  // [1] end, then per walked tile i < kTraceTiles: [2+4i] MMA warp may start (accumulator free), [3+4i] its first k-block has
  // landed, [4+4i] epilogue sees the accumulator complete, [5+4i] epilogue done with the tile
  unsigned long long* trace;
};
constexpr int kTraceTiles = 24, kTraceFine = 2 + 4 * kTraceTiles, kTraceSlots = kTraceFine + 5 * 5 + 5 * 3;  // + stamps inside the first 5 chunks

// ---- stream-K work split (SK kernels).  The (tile, k-block) units of a launch are cut into one CONTIGUOUS range per cluster
// instead of whole tiles dealt round-robin, so a launch of 160 tiles on 148 SMs takes 1.08 tile times instead of 2.  A range
// covers the tail [k_first, Kb) of its first tile, whole tiles, and the head [0, k_last_end) of its last tile; a tile is
// split between at most two neighbouring clusters (the planner only enables this when tiles >= clusters).  Order inside a
// cluster: the HEAD of the last tile first (its raw fp32 accumulator is stored to the workspace and a flag released), then
// the whole tiles, and last the TAIL of the first tile -- its accumulator is first LOADED with the neighbour's partial
// (tcgen05.st), so the MMAs of the remaining k-blocks add onto it exactly as if one CTA had run the whole reduction: the
// result is bit-identical to the unsplit kernel, and the partial it needs was the first thing the neighbour produced.
enum : int { SK_FULL = 0, SK_STORE = 1, SK_LOAD = 2 };
struct SkSeg {
  int tile, kb0, kb1, mode;
};
struct SkPlan {
  int nA, tf0, nfull, nZ, t_first, k_first, t_last, k_last_end, Kb;
  __device__ __forceinline__ SkPlan(int c, int G, int T, int Kb_, int align_tiles = 0) : Kb(Kb_) {
    const long U = (long)T * Kb_;
    long u0 = (long)c * U / G, u1 = (long)(c + 1) * U / G;
    if (align_tiles) {  // experiment: contiguous ranges of WHOLE tiles (no split), to separate the effect of the tile order
      u0 = ((long)c * T / G) * Kb_;
      u1 = ((long)(c + 1) * T / G) * Kb_;
    }
    t_first = (int)(u0 / Kb_);
    k_first = (int)(u0 - (long)t_first * Kb_);
    t_last = (int)((u1 - 1) / Kb_);
    k_last_end = (int)(u1 - (long)t_last * Kb_);
    nZ = k_first > 0 ? 1 : 0;
    nA = k_last_end < Kb_ ? 1 : 0;
    tf0 = t_first + nZ;
    nfull = (t_last - nA) - tf0 + 1;
    if (u1 <= u0) nA = nZ = nfull = 0;
  }
  __device__ __forceinline__ int count() const { return nA + nfull + nZ; }
  __device__ __forceinline__ SkSeg seg(int i) const {
    if (i < nA) return SkSeg{t_last, 0, k_last_end, SK_STORE};
    i -= nA;
    if (i < nfull) return SkSeg{tf0 + i, 0, Kb, SK_FULL};
    return SkSeg{t_first, k_first, Kb, SK_LOAD};
  }
};

__device__ __forceinline__ int fast_div(int x, unsigned long long mul) {
  return (int)(((unsigned long long)(unsigned)x * mul) >> 44);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_LEAKY: return v > 0.f ? v : 0.1f * v;
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

// CG = 2: the kernel runs as CTA pairs (cluster of 2, tcgen05 cta_group::2).  A pair computes a 256 x BLOCK_N tile:
// each CTA stages its own 128 rows of A and HALF of the B tile, the leader issues M = 256 MMAs that read both CTAs'
// shared memory, and each CTA's TMEM receives its own 128 x BLOCK_N accumulator (so the epilogue is unchanged).
// Per CTA and k-block that is 128*SWZ + BLOCK_N/2*SWZ bytes from L2 instead of 128*SWZ + BLOCK_N*SWZ: the L2->SM
// fabric (~43 B/clk/SM on B200) is what bounds the single-CTA kernel on every compute-heavy layer.
// MT = 2 / 4: a CTA tile is 256 / 512 pixels = MT M = 128 MMAs per K step against the same B (MT accumulators); a TMA
// load stages 256 A rows (two loads for MT = 4).
// MT = 2: a CTA tile is 256 pixels = two M = 128 MMAs per K step against the same B (two accumulators); one TMA load
// stages 256 A rows.  Halves the tile count -- and with it the producer / MMA warps' per-tile and per-k-block
// instruction overhead per pixel -- on the narrow (BLOCK_N <= 64) layers, which are bound by exactly that.
template <int BLOCK_N, int BLOCK_K, int STAGES, int CG = 1, int NB = 4, int MT = 1, int SK = 0>
struct ConvCfg {
  static constexpr int BLOCK_M = 128 * MT;
  static constexpr int SWZ = BLOCK_K * 2;  // bytes per smem row == swizzle span
  static constexpr int A_BYTES = BLOCK_M * SWZ;
  static constexpr int A_SUB_BYTES = 128 * SWZ;  // one M = 128 operand
  static constexpr int B_BYTES = (BLOCK_N / CG) * SWZ;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int CHUNK = BLOCK_N >= 64 ? 64 : 32;  // output channels per epilogue chunk / TMA store box
  static constexpr int OUT_SWZ = CHUNK * 2;
  static constexpr int CHUNK_BYTES = 128 * OUT_SWZ;
  static constexpr int N_CHUNKS = BLOCK_N / CHUNK;
  static constexpr int NBUF = NB;                      // ring of chunk buffers: residual lands in it, result leaves from it
  static constexpr int EPI_BYTES = NBUF * CHUNK_BYTES;
  // The kernel has NO static shared memory, so the dynamic window starts at the (1024-byte aligned) base of the CTA's
  // shared memory and the swizzled tiles need no alignment slack: [stages][epilogue ring][bias x2][mbarriers].
  static constexpr int BIAS_BYTES = 2 * BLOCK_N * 4;
  static constexpr int BAR_BYTES = (2 * STAGES + 4 + NBUF + SK) * 8 + 16;  // SK: + the "accumulator loaded" barrier
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BIAS_BYTES + BAR_BYTES;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
  static constexpr int TMEM_COLS = 2 * MT * BLOCK_N;
  // EPI_SPLIT warps share a TMEM lane quarter, each takes EPI_COLS columns of every chunk: two per quarter (8 epilogue warps,
  // 32 columns each).  Four per quarter (-DBP_EPI_WARPS=16) shorten an isolated tile's epilogue by ~10 % but cost the
  // whole step 3 %: a 608-thread CTA holds 58 K registers and the other lane's PnP CTAs (16 K) no longer fit beside it.
#ifndef BP_EPI_WARPS
#define BP_EPI_WARPS 8
#endif
  static constexpr int EPI_WARPS = BLOCK_N >= 64 ? BP_EPI_WARPS : 8;
  static constexpr int EPI_SPLIT = EPI_WARPS / 4;
  static constexpr int EPI_COLS = CHUNK / EPI_SPLIT;
  static_assert(EPI_COLS == 16 || EPI_COLS == 32, "epilogue threads take 16 or 32 columns of a chunk");
  static constexpr int EPI_THREADS = EPI_WARPS * 32;
  static constexpr int STORE_WARP = 2 + EPI_WARPS;      // issues the TMA stores / residual loads of the staged chunks
  static constexpr int THREADS = 64 + EPI_THREADS + 32;
  // Register budget: the CTA shares its SM with the other lane's PnP CTAs (pnp_hypotheses_serial_kernel: 64 threads x 255
  // registers = 16 K of the SM's 64 K).  11 warps x 136 registers leave exactly that; the launch bound is therefore stated
  // for 480 threads (65 536 / 480 = 136), not for the 352 the kernel runs with (the compiler then uses 128 instead of 139, 8
  // bytes of spill; throughput-neutral on B200: 7 522 / 7 570 vs 7 537 / 7 500 images/s interleaved on one box).
  static constexpr int LAUNCH_BOUND = THREADS <= 480 ? 480 : THREADS;
};

// byte offset of 16-byte unit j of row r inside a swizzled [128][ROW_BYTES] tile (TMA SWIZZLE_128B / SWIZZLE_64B)
template <int ROW_BYTES>
__device__ __forceinline__ uint32_t swz_off(int r, int j) {
  if constexpr (ROW_BYTES == 128) return uint32_t(r * 128 + ((j ^ (r & 7)) << 4));
  else return uint32_t(r * 64 + ((j ^ ((r >> 1) & 3)) << 4));
}

// Epilogue math for one thread's NV consecutive channels of one output pixel: fp32 accumulator + fp32 bias, then
// packed fp16: activation, residual.  `buf` points at this thread's first 16-byte unit slot; units are addressed
// through swz_off so the same code reads the TMA-loaded residual and writes the TMA-stored result in place.
template <int ACT, int RESMODE, int ROW_BYTES, int NV>
__device__ __forceinline__ void epi_math_store(const uint32_t* a, const float* bias, uint8_t* tile, int row_l, int unit0) {
  // Instruction budget matters here: a chunk's epilogue is bounded by issue slots once the TMEM read is overlapped, so the
  // fp32 work is done two lanes at a time (add / mul .f32x2, sm_100) -- the same IEEE operations, half the instructions.
#pragma unroll
  for (int j = 0; j < NV / 8; ++j) {
    const float4 b0 = *reinterpret_cast<const float4*>(bias + j * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(bias + j * 8 + 4);
    float2 v[4];
    v[0] = __fadd2_rn(make_float2(__uint_as_float(a[j * 8 + 0]), __uint_as_float(a[j * 8 + 1])), make_float2(b0.x, b0.y));
    v[1] = __fadd2_rn(make_float2(__uint_as_float(a[j * 8 + 2]), __uint_as_float(a[j * 8 + 3])), make_float2(b0.z, b0.w));
    v[2] = __fadd2_rn(make_float2(__uint_as_float(a[j * 8 + 4]), __uint_as_float(a[j * 8 + 5])), make_float2(b1.x, b1.y));
    v[3] = __fadd2_rn(make_float2(__uint_as_float(a[j * 8 + 6]), __uint_as_float(a[j * 8 + 7])), make_float2(b1.z, b1.w));
    uint8_t* slot = tile + swz_off<ROW_BYTES>(row_l, unit0 + j);
    uint4 pk;
    __half2* h = reinterpret_cast<__half2*>(&pk);
    const __half2 zero = __float2half2_rn(0.f);
    if constexpr (RESMODE != RES_NONE) {
      // residual layers: add in fp32 and round ONCE (the residual chains are 23 / 33 blocks deep; rounding the conv result
      // to fp16 before the add would round every block twice)
      const uint4 rv = *reinterpret_cast<const uint4*>(slot);
      const __half2* r = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 rf = __half22float2(r[e]);
        float2 x = v[e];
        if constexpr (RESMODE == RES_BEFORE_ACT) x = __fadd2_rn(x, rf);
        if constexpr (ACT == ACT_LEAKY) {
          const float2 s = __fmul2_rn(x, make_float2(0.1f, 0.1f));
          x.x = fmaxf(x.x, s.x);
          x.y = fmaxf(x.y, s.y);
        }
        if constexpr (ACT == ACT_RELU && RESMODE == RES_AFTER_ACT) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); }
        if constexpr (ACT == ACT_SIGMOID) { x.x = 1.f / (1.f + __expf(-x.x)); x.y = 1.f / (1.f + __expf(-x.y)); }
        if constexpr (RESMODE == RES_AFTER_ACT) x = __fadd2_rn(x, rf);
        h[e] = __floats2half2_rn(x.x, x.y);
        // ReLU after the rounding: rounding is monotonic and keeps the sign, so max(round(x), 0) == round(max(x, 0))
        if constexpr (ACT == ACT_RELU && RESMODE == RES_BEFORE_ACT) h[e] = __hmax2(h[e], zero);
      }
    } else if constexpr (ACT == ACT_SIGMOID) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        h[e] = __floats2half2_rn(1.f / (1.f + __expf(-v[e].x)), 1.f / (1.f + __expf(-v[e].y)));
    } else {
      const __half2 slope = __float2half2_rn(0.1f);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __half2 x = __floats2half2_rn(v[e].x, v[e].y);
        if constexpr (ACT == ACT_LEAKY) x = __hmax2(x, __hmul2(x, slope));  // slope < 1: max(x, 0.1 x)
        if constexpr (ACT == ACT_RELU) x = __hmax2(x, zero);
        h[e] = x;
      }
    }
    *reinterpret_cast<uint4*>(slot) = pk;
  }
}

template <int ROW_BYTES, int NV>
__device__ __forceinline__ void epi_dispatch(int act, int res_mode, const uint32_t* a, const float* bias, uint8_t* tile,
                                             int row_l, int unit0) {
#define BP_EPI(A, R) epi_math_store<A, R, ROW_BYTES, NV>(a, bias, tile, row_l, unit0)
  if (res_mode == RES_NONE) {
    if (act == ACT_LEAKY) BP_EPI(ACT_LEAKY, RES_NONE);
    else if (act == ACT_RELU) BP_EPI(ACT_RELU, RES_NONE);
    else if (act == ACT_SIGMOID) BP_EPI(ACT_SIGMOID, RES_NONE);
    else BP_EPI(ACT_NONE, RES_NONE);
  } else if (res_mode == RES_AFTER_ACT) {
    if (act == ACT_LEAKY) BP_EPI(ACT_LEAKY, RES_AFTER_ACT);
    else if (act == ACT_RELU) BP_EPI(ACT_RELU, RES_AFTER_ACT);
    else BP_EPI(ACT_NONE, RES_AFTER_ACT);
  } else {
    if (act == ACT_RELU) BP_EPI(ACT_RELU, RES_BEFORE_ACT);
    else if (act == ACT_LEAKY) BP_EPI(ACT_LEAKY, RES_BEFORE_ACT);
    else BP_EPI(ACT_NONE, RES_BEFORE_ACT);
  }
#undef BP_EPI
}

template <int BLOCK_N, int BLOCK_K, int STAGES, int CG = 1, int NB = 4, int MT = 1, int SK = 0>
__global__ void __launch_bounds__(ConvCfg<BLOCK_N, BLOCK_K, STAGES, CG, NB, MT, SK>::LAUNCH_BOUND, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                 const ConvArgs p) {
  using Cfg = ConvCfg<BLOCK_N, BLOCK_K, STAGES, CG, NB, MT, SK>;
  static_assert(SK == 0 || MT == 1, "stream-K: 128-pixel tiles only");
  static_assert(MT == 1 || ((MT == 2 || MT == 4) && CG == 1 && 2 * MT * BLOCK_N <= 512), "256 / 512-pixel tiles: single-CTA layers, 2 x MT accumulators in TMEM");
  static_assert(CG == 1 || (CG == 2 && BLOCK_N >= 64), "pairs need BLOCK_N >= 64");
  static_assert(BLOCK_N == 32 || BLOCK_N == 64 || BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");
  static_assert(BLOCK_K == 64 || BLOCK_K == 32, "BLOCK_K");
  constexpr int CHUNK = Cfg::CHUNK;
  constexpr int N_CHUNKS = Cfg::N_CHUNKS;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem + STAGES * Cfg::STAGE_BYTES;  // NBUF x CHUNK_BYTES
  float(*s_bias)[BLOCK_N] = reinterpret_cast<float(*)[BLOCK_N]>(ring + Cfg::EPI_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + Cfg::EPI_BYTES + Cfg::BIAS_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* res_full_bar = tmem_empty_bar + 2;
  uint64_t* acc_init_bar = res_full_bar + Cfg::NBUF;  // SK only: the epilogue has put a partial accumulator into TMEM
  uint32_t* tmem_base_slot_p = reinterpret_cast<uint32_t*>(res_full_bar + Cfg::NBUF + SK);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) {
    printf("bp: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }

  // tile walk: cluster c of the grid takes (pair-)tiles c, c + num_clusters, ...; inside a pair CTA `rank` owns the
  // rows of m-tile 2 * pair_m_tile + rank
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int cl_id = CG == 2 ? int(blockIdx.x >> 1) : int(blockIdx.x);
  const int cl_num = CG == 2 ? int(gridDim.x >> 1) : int(gridDim.x);
  const int total_tiles = ((p.m_tiles + CG - 1) / CG) * p.n_tiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) {
      tma_prefetch_desc(&tmOut);
      if (p.res_mode != RES_NONE) tma_prefetch_desc(&tmRes);
    }
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], Cfg::EPI_WARPS * CG);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
#pragma unroll
    for (int s = 0; s < Cfg::NBUF; ++s) mbar_init(&res_full_bar[s], 1);
    if constexpr (SK) mbar_init(acc_init_bar, Cfg::EPI_WARPS * CG);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_cg2<Cfg::TMEM_COLS>(tmem_base_slot_p);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_base_slot_p);
  }
  griddep_launch_dependents();  // the next kernel of the stream may start its own set-up as soon as an SM frees up
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  griddep_wait();  // from here on the previous kernel's outputs (our activations / residual) are complete and visible
  const uint32_t tmem_base = *tmem_base_slot_p;
  unsigned long long* const tr = p.trace ? p.trace + (size_t)blockIdx.x * kTraceSlots : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = clock64();
  // tile walk: round-robin whole tiles, or (SK) the segments of this cluster's contiguous unit range
  const SkPlan skp(cl_id, cl_num, total_tiles, p.num_kb, p.sk_align);
  const int n_walk = SK ? skp.count() : (cl_id < total_tiles ? (total_tiles - cl_id + cl_num - 1) / cl_num : 0);
  auto walk = [&](int i) -> SkSeg {
    if constexpr (SK) return skp.seg(i);
    else return SkSeg{cl_id + i * cl_num, 0, p.num_kb, SK_FULL};
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (one thread: the loop is pure issue
    // overhead, so it is kept to a handful of instructions per k-block; the other 31 lanes idle at the final barrier)
    {  // the whole warp walks the loop (uniform control flow); one elected lane issues the copies
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      uint32_t s = 0, ph = 0;
      for (int wi = 0; wi < n_walk; ++wi) {
        const SkSeg sg = walk(wi);
        const int tile = sg.tile;
        const int pm_tile = fast_div(tile, p.mul_nt);
        const int n0 = (tile - pm_tile * p.n_tiles) * BLOCK_N + int(rank) * (BLOCK_N / CG);  // this CTA's share of B
        const int m0 = (pm_tile * CG + int(rank)) * Cfg::BLOCK_M;
        int w0 = 0, h0 = 0, img = 0;
        if (p.a_im2col == 1) {
          const int pq = p.P * p.Q;
          img = fast_div(m0, p.mul_pq);
          const int rem = m0 - img * pq;
          const int op = fast_div(rem, p.mul_q);
          const int oq = rem - op * p.Q;
          w0 = oq * p.stride_w - p.pad_w;
          h0 = op * p.stride - p.pad;
        }
        int w0b = 0, h0b = 0, imgb = 0;  // MT == 4: window origin of the tile's second 256 pixels
        if constexpr (MT == 4) {
          if (p.a_im2col == 1) {
            const int pq = p.P * p.Q, mb = m0 + 256;
            imgb = fast_div(mb, p.mul_pq);
            const int rem = mb - imgb * pq;
            const int op = fast_div(rem, p.mul_q);
            w0b = (rem - op * p.Q) * p.stride_w - p.pad_w;
            h0b = op * p.stride - p.pad;
          }
        }
        int cb = 0, fr = 0, fs = 0, kcol = 0;
        if constexpr (SK) {
          if (sg.kb0) {  // the reduction of this segment starts at k-block kb0: (filter row, filter column, channel block) there
            kcol = sg.kb0 * BLOCK_K;
            const int tap = sg.kb0 / p.cblocks;
            cb = (sg.kb0 - tap * p.cblocks) * BLOCK_K;
            fr = tap / p.S;
            fs = tap - fr * p.S;
          }
        }
        for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
          mbar_wait_a(empty0 + 8u * s, ph ^ 1u);
          const uint32_t sa = smem_base + s * uint32_t(Cfg::STAGE_BYTES);
          const uint32_t fb = CG == 2 ? leader_addr(full0 + 8u * s) : full0 + 8u * s;
          if (elect_one()) {
          if constexpr (CG == 2) {
            // the leader's barrier collects the bytes of both CTAs
            if (rank == 0) mbar_expect_tx_a(fb, 2 * Cfg::STAGE_BYTES);
            if (p.a_im2col == 1) {
              tma_load_im2col_4d_cg2(&tmA, fb, sa, cb, w0, h0, img, (uint16_t)fs, (uint16_t)fr);
            } else {
              tma_load_2d_cg2(&tmA, fb, sa, kcol, m0);
            }
            tma_load_2d_cg2(&tmB, fb, sa + uint32_t(Cfg::A_BYTES), kcol, n0);
          } else {
            mbar_expect_tx_a(fb, Cfg::STAGE_BYTES);
            if (p.a_im2col == 1) {
              tma_load_im2col_4d_a(&tmA, fb, sa, cb, w0, h0, img, (uint16_t)fs, (uint16_t)fr);
              if constexpr (MT == 4)
                tma_load_im2col_4d_a(&tmA, fb, sa + uint32_t(Cfg::A_BYTES / 2), cb, w0b, h0b, imgb, (uint16_t)fs, (uint16_t)fr);
            } else {
              tma_load_2d_a(&tmA, fb, sa, kcol, m0);
              if constexpr (MT == 4) tma_load_2d_a(&tmA, fb, sa + uint32_t(Cfg::A_BYTES / 2), kcol, m0 + 256);
            }
            tma_load_2d_a(&tmB, fb, sa + uint32_t(Cfg::A_BYTES), kcol, n0);
          }
          }
          __syncwarp();
          kcol += BLOCK_K;
          // advance (tap, channel block) without divisions
          cb += BLOCK_K;
          if (cb == p.C) {
            cb = 0;
            if (++fs == p.S) {
              fs = 0;
              ++fr;
            }
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread; pairs: the leader CTA's)
    if (rank == 0) {  // the whole warp walks the loop (uniform control flow); one elected lane issues
      constexpr uint32_t idesc = umma_idesc_f16(BLOCK_N, 128 * CG);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const uint32_t tfull0 = smem_u32(&tmem_full_bar[0]), tempty0 = smem_u32(&tmem_empty_bar[0]);
      // descriptors differ between stages only in the 14-bit start-address field (smem < 256 KB: no carry out of it)
      const uint64_t da0 = umma_smem_desc<Cfg::SWZ>(smem_u32(smem));
      uint32_t s = 0, ph = 0, it = 0;
      for (int wi = 0; wi < n_walk; ++wi, ++it) {
        const SkSeg sg = walk(wi);
        const uint32_t acc = it & 1u;
        mbar_wait_a(tempty0 + 8u * acc, ((it >> 1) & 1u) ^ 1u);  // epilogue(s) drained this accumulator
        if constexpr (SK) {
          if (sg.mode == SK_LOAD) mbar_wait_a(smem_u32(acc_init_bar), 0u);  // ... and loaded it with the neighbour's partial (once per launch)
        }
        tc_fence_after();
        if (tr && wi < kTraceTiles && lane == 0) tr[2 + 4 * wi] = clock64();
        const uint32_t tmem_acc = tmem_base + acc * uint32_t(MT * BLOCK_N);
        const int kb_first = (SK && sg.mode == SK_LOAD) ? -1 : sg.kb0;  // k-block whose first MMA overwrites the accumulator
        for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
          mbar_wait_a(full0 + 8u * s, ph);
          tc_fence_after();
          if (tr && wi < kTraceTiles && kb == sg.kb0 && lane == 0) tr[3 + 4 * wi] = clock64();
          const uint64_t da = da0 + uint64_t(s * uint32_t(Cfg::STAGE_BYTES >> 4));
          const uint64_t db = da + uint64_t(Cfg::A_BYTES >> 4);
          // advance 16 elements (32 B) along K inside the swizzle span: +2 in the (addr>>4) field
          if (elect_one()) {
            if constexpr (CG == 2) {
              umma_f16_cg2(tmem_acc, da, db, idesc, kb != kb_first ? 1u : 0u);
#pragma unroll
              for (int k = 1; k < BLOCK_K / 16; ++k)
                umma_f16_cg2(tmem_acc, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, 1u);
              umma_commit_cg2(empty0 + 8u * s);  // frees this stage in BOTH CTAs
            } else {
#pragma unroll
              for (int k = 0; k < BLOCK_K / 16; ++k) {
#pragma unroll
                for (int sub = 0; sub < MT; ++sub)
                  umma_f16(tmem_acc + uint32_t(sub * BLOCK_N), da + uint64_t(2 * k + sub * (Cfg::A_SUB_BYTES >> 4)),
                           db + uint64_t(2 * k), idesc, (kb != kb_first || k) ? 1u : 0u);
              }
              umma_commit_a(empty0 + 8u * s);  // frees this smem stage once the MMAs above have read it
            }
          }
          __syncwarp();
          if (++s == STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
        if (elect_one()) {
          if constexpr (CG == 2) umma_commit_cg2(tfull0 + 8u * acc);  // accumulators complete -> both epilogues
          else umma_commit_a(tfull0 + 8u * acc);
        }
        __syncwarp();
      }
    }
  } else if (warp == Cfg::STORE_WARP) {
    // ------------------------------------------------------------ store warp (TMA-stored layers).  The epilogue warps stage
    // a chunk in ring buffer g % NBUF and ARRIVE on named barrier 2 + g % NBUF without waiting; this warp waits there,
    // stores the chunk (cp.async.bulk.tensor, one thread: bulk groups are per thread) and, as soon as that store
    // has read its buffer, hands the buffer on: with a residual by loading the residual tile of the chunk that will use
    // it next (TMA, completes res_full_bar), without one by a plain arrive on the same mbarrier ("buffer free").  So the
    // store issue, the commit and wait_group.read (~250-450 cycles per chunk when an epilogue thread did them between
    // two barriers of all epilogue warps: harness trace) are off the epilogue's critical path.
    if (p.tma_store) {
      const bool use_res = p.res_mode != RES_NONE;
      int pf_wi = 0, pf_sub = 0, pf_c = 0, pf_n0 = 0, pf_m0 = 0, pf_live = 0;  // residual prefetch cursor over the walk
      uint32_t pf_g = 0;  // global index of the next chunk to prefetch
      auto pf_decode = [&]() {
        if constexpr (SK) {
          while (pf_wi < n_walk && walk(pf_wi).mode == SK_STORE) ++pf_wi;  // a stored partial has no epilogue output, no residual
        }
        if (pf_wi < n_walk) {
          const int pf_tile = walk(pf_wi).tile;
          const int pm = fast_div(pf_tile, p.mul_nt);
          pf_n0 = (pf_tile - pm * p.n_tiles) * BLOCK_N;
          pf_m0 = (pm * CG + int(rank)) * Cfg::BLOCK_M;
          pf_live = min(N_CHUNKS, (p.Cout - pf_n0 + CHUNK - 1) / CHUNK);
        }
      };
      auto pf_issue = [&]() {  // load the residual of chunk pf_g (if there is one) and advance the cursor
        if (pf_wi >= n_walk) return;
        const uint32_t rb = pf_g % Cfg::NBUF;
        if (lane == 0) {
          mbar_expect_tx(&res_full_bar[rb], Cfg::CHUNK_BYTES);
          tma_load_2d(&tmRes, &res_full_bar[rb], ring + rb * Cfg::CHUNK_BYTES, pf_n0 + pf_c * CHUNK, pf_m0 + pf_sub * 128);
        }
        ++pf_g;
        if (++pf_c == pf_live) {
          pf_c = 0;
          if (++pf_sub == MT) {
            pf_sub = 0;
            ++pf_wi;
            pf_decode();
          }
        }
      };
      if (use_res) {
        // the prefetch cursor runs NBUF chunks ahead of the chunk being stored, across tile boundaries
        pf_decode();
        for (int i = 0; i < Cfg::NBUF; ++i) pf_issue();
      } else if (lane == 0) {
        for (int b = 0; b < Cfg::NBUF; ++b) mbar_arrive(&res_full_bar[b]);  // every buffer starts out free
      }
      uint32_t g = 0;
      for (int wi = 0; wi < n_walk; ++wi) {
        const SkSeg sg = walk(wi);
        if constexpr (SK) {
          if (sg.mode == SK_STORE) continue;
        }
        const int pm_tile = fast_div(sg.tile, p.mul_nt);
        const int m0 = (pm_tile * CG + int(rank)) * Cfg::BLOCK_M;
        const int n0 = (sg.tile - pm_tile * p.n_tiles) * BLOCK_N;
        const int live = min(N_CHUNKS, (p.Cout - n0 + CHUNK - 1) / CHUNK);
        for (int sub = 0; sub < MT; ++sub) {
          for (int c = 0; c < live; ++c, ++g) {
            const uint32_t bsel = g % Cfg::NBUF;
            unsigned long long* const ss = (tr && g < 5 && lane == 0) ? tr + kTraceFine + 25 + 3 * g : nullptr;
            bar_sync_named(2 + int(bsel), Cfg::EPI_THREADS + 32);  // chunk g is staged (its writers fenced the async proxy)
            if (ss) ss[0] = clock64();
            if (lane == 0) {
              tma_store_2d(&tmOut, ring + bsel * Cfg::CHUNK_BYTES, n0 + c * CHUNK, m0 + sub * 128);
              bulk_commit_group();
              if (ss) ss[1] = clock64();
              // This warp has nothing else to do: it waits until the store has read the buffer (a few hundred cycles) and
              // hands the SAME buffer on at once -- the residual of chunk g + NBUF is then under way a whole chunk time
              // earlier than if the hand-over waited for the next store's commit (a buffer is busy from the issue of its
              // residual load to the end of its store's read, ~3 000 cycles: with 3-4 buffers every cycle of lead counts).
              bulk_wait_group_read<0>();
            }
            __syncwarp();
            if (use_res) {
              pf_issue();  // chunk g + NBUF into buffer g % NBUF
            } else if (lane == 0) {
              mbar_arrive(&res_full_bar[bsel]);
            }
            if (ss) ss[2] = clock64();
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2 .. 2 + EPI_WARPS - 1)
    constexpr int HALF = Cfg::EPI_COLS;      // channels of a chunk handled by one thread (16)
    constexpr int ESPLIT = Cfg::EPI_SPLIT;   // warps sharing a lane quarter
    const int q4 = warp & 3;                 // TMEM lane quarter this warp may read
    const int hsel = (warp - 2) >> 2;        // which share of each chunk's columns
    const int row_l = q4 * 32 + lane;        // row inside the tile
    const bool leader = threadIdx.x == 64;   // issues TMA stores / residual loads (bulk groups are per thread)
    const int et = threadIdx.x - 64;         // 0..255
    const bool use_res = p.res_mode != RES_NONE;
    int it = 0;
    uint32_t chunk_ctr = 0;                  // running chunk index: ring buffer = chunk_ctr % NBUF
    float bias_next = (n_walk > 0 && et < BLOCK_N) ? __ldg(p.bias + (walk(0).tile % p.n_tiles) * BLOCK_N + et) : 0.f;
    const uint32_t tfull_wait0 = smem_u32(&tmem_full_bar[0]);
    const uint32_t tempty_arrive0 = CG == 2 ? leader_addr(smem_u32(&tmem_empty_bar[0])) : smem_u32(&tmem_empty_bar[0]);
    // Residual tiles arrive by TMA in the ring buffers the results later leave from (store warp above); res_full_bar[b]
    // completes when buffer b holds the residual of its next chunk -- or, without a residual, when it is free again.
    for (int wi = 0; wi < n_walk; ++wi, ++it) {
      const SkSeg sg = walk(wi);
      const int tile = sg.tile;
      const int pm_tile = fast_div(tile, p.mul_nt);
      const int n_tile = tile - pm_tile * p.n_tiles;
      const int m0 = (pm_tile * CG + int(rank)) * Cfg::BLOCK_M;
      const int n0 = n_tile * BLOCK_N;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const int live = min(N_CHUNKS, (p.Cout - n0 + CHUNK - 1) / CHUNK);  // chunks holding real channels
      const uint32_t tmem_acc0 = tmem_base + uint32_t(acc * MT * BLOCK_N) + ((uint32_t(q4) * 32u) << 16);

      if constexpr (SK) {
        // workspace slot layout [32-column group][float4 index j][row][4]: every warp-wide 16-byte access is 512 contiguous bytes
        if (sg.mode == SK_LOAD) {
          // the tail of a split tile: put the neighbour's partial accumulator (k-blocks [0, kb0)) into TMEM, then let the MMA
          // warp add the remaining k-blocks onto it.  This accumulator buffer was drained two segments ago by this epilogue.
          const int src = (cl_id - 1) * CG + int(rank);
          if (leader) {
            while (ld_acquire_gpu(p.sk_flags + src) == 0u) __nanosleep(64);
            p.sk_flags[src] = 0u;  // consumed: ready for the next launch
          }
          bar_sync_named(1, Cfg::EPI_THREADS);
          __threadfence();
          const float4* wsrc = reinterpret_cast<const float4*>(p.sk_ws) + (size_t)src * (BLOCK_N / 32) * 8 * 128;
#pragma unroll 1
          for (int c = hsel; c < BLOCK_N / 32; c += ESPLIT) {
            uint32_t a[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 v = __ldcg(wsrc + ((size_t)c * 8 + j) * 128 + row_l);
              a[4 * j] = __float_as_uint(v.x); a[4 * j + 1] = __float_as_uint(v.y);
              a[4 * j + 2] = __float_as_uint(v.z); a[4 * j + 3] = __float_as_uint(v.w);
            }
            tmem_st_32x32(tmem_acc0 + uint32_t(c * 32), a);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_cluster_a(leader_addr(smem_u32(acc_init_bar)));
            else mbar_arrive(acc_init_bar);
          }
        }
      }

      float* bias_s = s_bias[it & 1];
      if (et < BLOCK_N) bias_s[et] = bias_next;
      {  // bias of the next tile: in flight during this tile's epilogue
        if (wi + 1 < n_walk && et < BLOCK_N) {
          const int nt = walk(wi + 1).tile;
          bias_next = __ldg(p.bias + (nt - fast_div(nt, p.mul_nt) * p.n_tiles) * BLOCK_N + et);
        }
      }
      bar_sync_named(1, Cfg::EPI_THREADS);  // bias visible

      mbar_wait_a(tfull_wait0 + 8u * acc, acc_ph);
      tc_fence_after();
      if (tr && wi < kTraceTiles && leader) tr[4 + 4 * wi] = clock64();

      if constexpr (SK) {
        if (sg.mode == SK_STORE) {
          // the head of a split tile: leave the raw fp32 accumulator (k-blocks [0, kb1)) in this CTA's workspace slot
          const int dst = cl_id * CG + int(rank);
          float4* wdst = reinterpret_cast<float4*>(p.sk_ws) + (size_t)dst * (BLOCK_N / 32) * 8 * 128;
#pragma unroll 1
          for (int c = hsel; c < BLOCK_N / 32; c += ESPLIT) {
            uint32_t a[32];
            tmem_ld_32x32(tmem_acc0 + uint32_t(c * 32), a);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              __stcg(wdst + ((size_t)c * 8 + j) * 128 + row_l,
                     make_float4(__uint_as_float(a[4 * j]), __uint_as_float(a[4 * j + 1]), __uint_as_float(a[4 * j + 2]), __uint_as_float(a[4 * j + 3])));
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_cluster_a(tempty_arrive0 + 8u * acc);
            else mbar_arrive(&tmem_empty_bar[acc]);
          }
          __threadfence();
          bar_sync_named(1, Cfg::EPI_THREADS);
          if (leader) st_release_gpu(p.sk_flags + dst, 1u);
          continue;
        }
      }

      if (p.tma_store) {
        // The chunks of a tile (MT sub-tiles x `live` chunks) are walked with the TMEM load issued ONE CHUNK AHEAD of its use, so
        // the wait at the top of a chunk finds its data there (harness trace: ~25 cycles).  A chunk costs ~1 000 cycles of a
        // warp's time whatever is moved around in it -- short dependent chains (bias -> add -> pack -> st.shared -> fence ->
        // arrive) on two warps per scheduler -- so a 128 x 256 tile drains in ~4 000 cycles: as long as 8 k-blocks of its
        // main loop.  Layers with K <= 512 are therefore bounded by this epilogue, and every launch ends with one exposed.
        const int nchunk = MT * live;
        uint32_t nx[32];
        if constexpr (HALF == 32) tmem_ld_32x32(tmem_acc0 + uint32_t(hsel * HALF), nx);
        else tmem_ld_32x16(tmem_acc0 + uint32_t(hsel * HALF), nx);
        int sub = 0, c = 0;
#pragma unroll 1
        for (int ci = 0; ci < nchunk; ++ci, ++chunk_ctr) {
          const uint32_t bsel = chunk_ctr % Cfg::NBUF;
          uint32_t a[32];
          // harness trace: the leader's stamps inside the first 5 chunks of the CTA's first tile
          unsigned long long* const fs = (tr && wi == 0 && ci < 5 && leader) ? tr + kTraceFine + 5 * ci : nullptr;
          if (fs) fs[0] = clock64();
          if constexpr (HALF == 32) tmem_ld_wait_dep32(nx);
          else tmem_ld_wait_dep16(nx);
          if (fs) fs[1] = clock64();
#pragma unroll
          for (int i = 0; i < HALF; ++i) a[i] = nx[i];
          int c2 = c + 1, sub2 = sub;
          if (c2 == live) {
            c2 = 0;
            ++sub2;
          }
          if (ci + 1 < nchunk) {
            if constexpr (HALF == 32) tmem_ld_32x32(tmem_acc0 + uint32_t(sub2 * BLOCK_N + c2 * CHUNK + hsel * HALF), nx);
            else tmem_ld_32x16(tmem_acc0 + uint32_t(sub2 * BLOCK_N + c2 * CHUNK + hsel * HALF), nx);
          } else {  // accumulators fully read (the wait above covered the last load): hand them back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CG == 2) mbar_arrive_cluster_a(tempty_arrive0 + 8u * acc);
              else mbar_arrive(&tmem_empty_bar[acc]);
            }
          }
          uint8_t* buf = ring + bsel * Cfg::CHUNK_BYTES;
          mbar_wait(&res_full_bar[bsel], (chunk_ctr / Cfg::NBUF) & 1);  // residual landed / buffer free (store warp)
          if (fs) fs[2] = clock64();
          epi_dispatch<Cfg::OUT_SWZ, HALF>(p.act, p.res_mode, a, bias_s + c * CHUNK + hsel * HALF, buf, row_l, hsel * (HALF / 8));
          if (fs) fs[3] = clock64();
          fence_proxy_async_smem();
          bar_arrive_named(2 + int(bsel), Cfg::EPI_THREADS + 32);  // no waiting: the store warp takes it from here
          if (fs) fs[4] = clock64();
          c = c2;
          sub = sub2;
        }
      } else if (p.staged_store) {
        if constexpr (MT != 1) __trap();  // 256-pixel tiles are planned for TMA-store layers only
        // -------- fused nearest-x2 upsample / PixelShuffle(2) of an fp16 result: the destination of a 128-pixel tile is not a
        // TMA box, and with every thread storing its own pixel's channels a warp-wide 16-byte store touches 32 different
        // 128-byte lines (x4 replicas for the upsample: 1x1 512->256 @13x13 ran at 9x its HBM roofline).  So the chunk is
        // staged in the epilogue ring exactly like a TMA-stored chunk and then written out row-wise: 8 consecutive threads
        // write the 128 contiguous bytes of one pixel's 64 channels, a warp 4 pixels.  One barrier per chunk suffices: the
        // ring has >= 3 buffers, and a thread that is writing chunk g + 1 has passed barrier g, which every thread reaches
        // only after its copy-out of chunk g - 1.
        constexpr int UNITS = CHUNK / 8;  // 16-byte units per staged row
        const uint32_t tmem_acc = tmem_acc0;
        for (int c = 0; c < live; ++c, ++chunk_ctr) {
          const uint32_t bsel = chunk_ctr % Cfg::NBUF;
          uint32_t a[32];
          if constexpr (HALF == 32) tmem_ld_32x32(tmem_acc + uint32_t(c * CHUNK + hsel * HALF), a);
          else tmem_ld_32x16(tmem_acc + uint32_t(c * CHUNK + hsel * HALF), a);
          tmem_ld_wait();
          if (c == live - 1) {  // accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CG == 2) mbar_arrive_cluster_a(tempty_arrive0 + 8u * acc);
              else mbar_arrive(&tmem_empty_bar[acc]);
            }
          }
          uint8_t* buf = ring + bsel * Cfg::CHUNK_BYTES;
          const float* bs = bias_s + c * CHUNK + hsel * HALF;
#pragma unroll
          for (int j = 0; j < HALF / 8; ++j) {  // same arithmetic as the per-thread path: fp32 bias + activation, one rounding
            uint4 pk;
            __half2* h2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              h2[e] = __floats2half2_rn(apply_act(__uint_as_float(a[j * 8 + 2 * e]) + bs[j * 8 + 2 * e], p.act),
                                        apply_act(__uint_as_float(a[j * 8 + 2 * e + 1]) + bs[j * 8 + 2 * e + 1], p.act));
            *reinterpret_cast<uint4*>(buf + swz_off<Cfg::OUT_SWZ>(row_l, hsel * (HALF / 8) + j)) = pk;
          }
          bar_sync_named(1, Cfg::EPI_THREADS);
          __half* o16 = reinterpret_cast<__half*>(p.out);
          const int pq = p.P * p.Q;
#pragma unroll
          for (int i2 = 0; i2 < (128 * UNITS) / Cfg::EPI_THREADS; ++i2) {
            const int idx = i2 * Cfg::EPI_THREADS + et;
            const int r = idx / UNITS, u = idx % UNITS;
            const int row = m0 + r;
            const int ch = n0 + c * CHUNK + u * 8;
            if (row >= p.M || ch >= p.Cout) continue;
            const uint4 pk = *reinterpret_cast<const uint4*>(buf + swz_off<Cfg::OUT_SWZ>(r, u));
            const int img = fast_div(row, p.mul_pq);
            const int rem = row - img * pq;
            const int op = fast_div(rem, p.mul_q);
            const int oq = rem - op * p.Q;
            if (p.store_mode == STORE_UPSAMPLE2) {
              const size_t base = ((size_t)img * (2 * p.P) + 2 * op) * (2 * p.Q) + 2 * oq;
#pragma unroll
              for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx)
                  *reinterpret_cast<uint4*>(o16 + (base + (size_t)dy * (2 * p.Q) + dx) * p.out_pitch + p.out_coff + ch) = pk;
            } else {  // PixelShuffle(2): weight rows were pre-permuted to o' = sub*(Cout/4) + c, sub = 2*i + j
              const int c4 = p.Cout >> 2;
              const int sub = ch / c4;
              const int cc = ch - sub * c4;
              const size_t pix = ((size_t)img * (2 * p.P) + 2 * op + (sub >> 1)) * (2 * p.Q) + 2 * oq + (sub & 1);
              *reinterpret_cast<uint4*>(o16 + pix * p.out_pitch + p.out_coff + cc) = pk;
            }
          }
        }
      } else {
        if constexpr (MT != 1) __trap();  // 256-pixel tiles are planned for TMA-store layers only
        const uint32_t tmem_acc = tmem_acc0;
        // -------- per-thread stores: fp32 heads (and any fp16 result the staged paths do not take)
        const int row = m0 + row_l;
        const bool row_ok = row < p.M;
        int img = 0, op = 0, oq = 0;
        if (p.store_mode != STORE_PLAIN) {
          const int pq = p.P * p.Q;
          img = fast_div(row, p.mul_pq);
          const int rem = row - img * pq;
          op = fast_div(rem, p.mul_q);
          oq = rem - op * p.Q;
        }
        const int live32 = min(BLOCK_N / 32, (p.Cout - n0 + 31) / 32);
        if (hsel >= live32) {  // nothing to read for this warp in this tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
              if constexpr (CG == 2) mbar_arrive_cluster_a(tempty_arrive0 + 8u * acc);
              else mbar_arrive(&tmem_empty_bar[acc]);
            }
        }
#pragma unroll 1
        for (int c = hsel; c < live32; c += ESPLIT) {
          const int ch0 = n0 + c * 32;
          uint32_t a[32];
          tmem_ld_32x32(tmem_acc + uint32_t(c * 32), a);
          tmem_ld_wait();
          if (c + ESPLIT >= live32) {  // this warp's last chunk of the tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CG == 2) mbar_arrive_cluster_a(tempty_arrive0 + 8u * acc);
              else mbar_arrive(&tmem_empty_bar[acc]);
            }
          }
          if (!row_ok) continue;

          if (p.out_f32 && !use_res) {
            // fp32 outputs (detector heads: 18 of 32 channels live, linear; heat-maps: 50 of 64): only the live channels
            // are touched, bias comes as float4, and the activation switch is one uniform branch per group.  (The generic
            // code below ran ~1 300 dependent instructions per tile on ONE warp per scheduler -- the 1x1 256->18 head at 52x52
            // took 4x its HBM time, ncu r03a_head52.)
            const int nl = min(32, p.Cout - ch0);
            float* op32 = reinterpret_cast<float*>(p.out) + (size_t)row * p.out_pitch + p.out_coff + ch0;
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + c * 32);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (g * 4 < nl) {
                const float4 b = b4[g];
                float x0 = __uint_as_float(a[g * 4]) + b.x, x1 = __uint_as_float(a[g * 4 + 1]) + b.y;
                float x2 = __uint_as_float(a[g * 4 + 2]) + b.z, x3 = __uint_as_float(a[g * 4 + 3]) + b.w;
                if (p.act != ACT_NONE) {
                  x0 = apply_act(x0, p.act); x1 = apply_act(x1, p.act);
                  x2 = apply_act(x2, p.act); x3 = apply_act(x3, p.act);
                }
                if (g * 4 + 4 <= nl) {
                  *reinterpret_cast<float4*>(op32 + g * 4) = make_float4(x0, x1, x2, x3);
                } else {
                  op32[g * 4] = x0;
                  if (g * 4 + 1 < nl) op32[g * 4 + 1] = x1;
                  if (g * 4 + 2 < nl) op32[g * 4 + 2] = x2;
                }
              }
            }
            continue;
          }

          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(a[j]) + bias_s[c * 32 + j];

          if (use_res) {
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
                  *reinterpret_cast<uint4*>(o16 + (base + (size_t)dy * (2 * p.Q) + dx) * p.out_pitch + p.out_coff + ch) = pk;
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
      }
      if (tr && wi < kTraceTiles && leader) tr[5 + 4 * wi] = clock64();
    }
  }

  tc_fence_before();
  if (tr && threadIdx.x == 64) tr[1] = clock64();
  if constexpr (CG == 2) cluster_sync_all();  // the peer may still be reading our shared memory / signalling our barriers
  else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    if constexpr (CG == 2) tmem_dealloc_cg2<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace bp
