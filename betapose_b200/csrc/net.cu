// bp_net: ordered list of fused layer ops over NHWC fp16 tensors (C-ABI in include/betapose_b200.h).
// Build time: BN folding (fp64), weight packing to [Cout_pad][R][S][Cin] fp16, buffer allocation, TMA
// descriptor encoding for max_batch.  Run time: one kernel launch per op, no host sync, graph-capturable.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "aux_kernels.cuh"
#include "betapose_b200.h"
#include "conv_plan.cuh"
#include "engine.h"

using namespace bp;

namespace {

enum OpKind { OP_CONV, OP_MAXPOOL, OP_AVGPOOL, OP_SCALE_ADD_RELU, OP_PIXSHUF, OP_UPSAMPLE, OP_COPYC, OP_ADD };

constexpr int kAvgSplitMax = 32;

struct Tensor {
  void* ptr = nullptr;  // includes the channel offset
  int H = 0, W = 0, C = 0, pitch = 0, coff = 0;
  bool f32 = false;
  int in_kind = -1;  // >= 0 for the network input
  int row_px = 0;    // network input only: pixels per buffer row (W + BP_IN_PAD_COLS); data starts at column BP_IN_PAD_LEFT
};

struct Op {
  OpKind kind;
  // OP_CONV: the launch plan (tile configuration + TMA descriptors) depends on the batch; built lazily per batch
  ConvDesc cdesc;
  bool stem = false;
  std::map<int, ConvPlan> plans;
  struct TileCfg { int block_n, cg, mt; };
  std::map<int, TileCfg> tuned;  // per batch size: measured-best tile configuration (bp_net_set_op_config)
  int pq = 0;     // output pixels per image (conv) for batch scaling
  int a = -1, b = -1, c = -1, dst = -1;  // tensor ids for aux ops
  float* scratch = nullptr;              // OP_AVGPOOL: [max_batch][kAvgSplitMax][C] partial sums
  unsigned* counters = nullptr;          // OP_AVGPOOL: [max_batch][C/256] arrival counters (zero between launches)
  double flops = 0, bytes = 0;  // per image
  std::string desc;
};

}  // namespace

struct bp_net {
  bp_engine* eng = nullptr;
  bp_net* share = nullptr;
  int max_batch = 0;
  int in_kind = 0;
  std::vector<Tensor> tensors;
  std::vector<Op> ops;
  std::vector<void*> owned;       // device allocations owned by this net
  std::vector<void*> act_buffers; // activation buffers in allocation order (for sharing)
  std::vector<size_t> act_sizes;
  size_t act_cursor = 0;
  double flops = 0;
  int share_batch = 0;  // > 0: runs concurrently with other nets, see bp_net_set_share
  float* sk_ws = nullptr;        // stream-K workspace: one 128 x 256 fp32 partial accumulator per SM (conv_plan.cuh)
  unsigned* sk_flags = nullptr;  // ... and its hand-over flags (zero between launches)
};

static void* net_alloc_weights(bp_net* n, size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  n->owned.push_back(p);
  return p;
}

// activation buffers: nets with the same topology (one per LineMod object) alias each other's buffers
static void* net_alloc_act(bp_net* n, size_t bytes) {
  const size_t i = n->act_cursor++;
  if (n->share && i < n->share->act_buffers.size() && n->share->act_sizes[i] >= bytes) {
    n->act_buffers.push_back(n->share->act_buffers[i]);
    n->act_sizes.push_back(n->share->act_sizes[i]);
    return n->share->act_buffers[i];
  }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  cudaMemset(p, 0, bytes);
  n->owned.push_back(p);
  n->act_buffers.push_back(p);
  n->act_sizes.push_back(bytes);
  return p;
}

static int new_tensor(bp_net* n, int H, int W, int C, bool f32) {
  Tensor t;
  t.H = H; t.W = W; t.C = C; t.f32 = f32;
  t.pitch = f32 ? (C + 3) / 4 * 4 : (C + 7) / 8 * 8;
  const size_t bytes = (size_t)n->max_batch * H * W * t.pitch * (f32 ? 4 : 2) + 256;
  t.ptr = net_alloc_act(n, bytes);
  if (!t.ptr) return -1;
  n->tensors.push_back(t);
  return (int)n->tensors.size() - 1;
}

static TView view_of(const Tensor& t) { return TView{t.ptr, t.H, t.W, t.C, t.pitch}; }

extern "C" {

int bp_net_create(bp_engine* e, int max_batch, int in_h, int in_w, int in_kind, bp_net* share, bp_net** out) {
  if (!e || !out || max_batch <= 0) return bp_fail(BP_ERR_INVALID, "bp_net_create: bad arguments");
  if (in_kind != BP_IN_RAW255 && in_kind != BP_IN_F16) return bp_fail(BP_ERR_INVALID, "bp_net_create: in_kind");
  if (share && share->max_batch < max_batch) return bp_fail(BP_ERR_INVALID, "bp_net_create: shared net is smaller");
  cudaSetDevice(e->device);
  bp_net* n = new bp_net();
  n->eng = e;
  n->share = share;
  n->max_batch = max_batch;
  n->in_kind = in_kind;
  Tensor t;
  // network input: fp16 [N, H, W + BP_IN_PAD_COLS, 8]; the pad columns and channels 3..7 stay zero for ever (the
  // buffer is zeroed at allocation and the stage kernels write only data pixels, channels 3..7 as zeros), which is
  // what lets the first convolution read whole filter rows as one "virtual pixel" (see bp_net_conv)
  t.H = in_h; t.W = in_w; t.C = 3; t.pitch = 8; t.in_kind = in_kind; t.row_px = in_w + BP_IN_PAD_COLS;
  t.ptr = net_alloc_act(n, (size_t)max_batch * in_h * t.row_px * 8 * 2 + 256);
  if (!t.ptr) {
    delete n;
    return bp_fail(BP_ERR_CUDA, "bp_net_create: cudaMalloc failed");
  }
  n->tensors.push_back(t);
  // Stream-K (conv_tcgen05.cuh: SkPlan) is implemented, bit-exact and OFF by default: measured on B200 it does not pay
  // (DESIGN.md 9).  BP_STREAMK=1 turns the planner's heuristic on for nets created afterwards.
  if (getenv("BP_STREAMK") && atoi(getenv("BP_STREAMK")) > 0) {
    const size_t slots = (size_t)e->num_sms;
    n->sk_ws = (float*)net_alloc_weights(n, slots * 128 * 256 * sizeof(float));
    n->sk_flags = (unsigned*)net_alloc_weights(n, slots * sizeof(unsigned));
    if (!n->sk_ws || !n->sk_flags || cudaMemset(n->sk_flags, 0, slots * sizeof(unsigned)) != cudaSuccess) {
      bp_net_destroy(n);
      return bp_fail(BP_ERR_CUDA, "bp_net_create: cudaMalloc (stream-K workspace) failed");
    }
  }
  *out = n;
  return BP_OK;
}

void bp_net_destroy(bp_net* n) {
  if (!n) return;
  for (void* p : n->owned) cudaFree(p);
  delete n;
}

void* bp_net_input_ptr(bp_net* n) { return n ? n->tensors[0].ptr : nullptr; }

int bp_net_alloc_tensor(bp_net* n, int h, int w, int c) {
  if (!n) return bp_fail(BP_ERR_INVALID, "null net");
  const int id = new_tensor(n, h, w, c, false);
  return id < 0 ? bp_fail(BP_ERR_CUDA, "bp_net_alloc_tensor: cudaMalloc failed") : id;
}

int bp_net_view(bp_net* n, int tensor, int coff, int c) {
  if (!n || tensor < 0 || tensor >= (int)n->tensors.size()) return bp_fail(BP_ERR_INVALID, "bp_net_view: tensor id");
  Tensor t = n->tensors[tensor];
  if (t.f32 || coff % 8 || c % 8 || coff + c > t.C) return bp_fail(BP_ERR_INVALID, "bp_net_view: window");
  t.ptr = reinterpret_cast<__half*>(t.ptr) + coff;
  t.coff += coff;
  t.C = c;
  n->tensors.push_back(t);
  return (int)n->tensors.size() - 1;
}

// Fold BN (fp64) into the weights and pack them for the implicit-GEMM kernel: [Cout_pad][K] fp16 rows + fp32 bias.
// Host code (no CUDA call): shared by bp_net_conv and bp_pack_conv_weights (the packed-weight cache).
//   ordinary conv: K = k*k*Cin ordered (r, q, ci) -- NHWC im2col order;
//   stem (3-channel network input, stored [N, H, W + 8, 8] fp16 with zero pad columns): one filter ROW of k pixels x 8
//   channels is contiguous in memory, so the convolution is run as a k x 1 convolution over "virtual pixels" of
//   Cv = 32 (k <= 4) or 64 (k <= 8) channels = 4 or 8 neighbouring real pixels, whose pixel stride (16 B) is smaller than
//   their extent (overlapping TMA im2col map).  K = k * Cv, ordered (row, pixel-in-row, 8 channels); entries beyond the
//   k real pixels / 3 real channels are zero weights.  For RAW255 inputs ToTensor's 1/255 (dataloader.py:94-99) is
//   folded into the weights, so the 0..255 pixel values are exact in fp16.
struct PackedDims {
  int Cv, K, wpitch, Cout_pad;
};
static PackedDims packed_dims(int Cin, int k, int Cout, bool stem) {
  PackedDims d;
  d.Cv = stem ? (k * 8 <= 32 ? 32 : 64) : 0;
  d.K = stem ? k * d.Cv : k * k * Cin;
  d.wpitch = (d.K + 7) / 8 * 8;
  d.Cout_pad = (Cout + 255) / 256 * 256;
  return d;
}
static void pack_conv_weights(const bp_conv_spec* s, int Cin, int in_kind /* -1: not a stem */, __half* hw, float* hb) {
  const bool stem = in_kind >= 0;
  const int k = s->ksize, Cout = s->cout;
  const PackedDims pd = packed_dims(Cin, k, Cout, stem);
  for (size_t i = 0; i < (size_t)pd.Cout_pad * pd.wpitch; ++i) hw[i] = __float2half(0.f);
  for (int i = 0; i < pd.Cout_pad; ++i) hb[i] = 0.f;
  const int c4 = Cout / 4;
  const double in_scale = (stem && in_kind == BP_IN_RAW255) ? 1.0 / 255.0 : 1.0;
  for (int o = 0; o < Cout; ++o) {
    double scale = 1.0, shift = s->bias ? (double)s->bias[o] : 0.0;
    if (s->bn_gamma) {
      const double inv = (double)s->bn_gamma[o] / std::sqrt((double)s->bn_var[o] + (double)s->bn_eps);
      scale = inv;
      shift = (double)s->bn_beta[o] - (double)s->bn_mean[o] * inv + shift * inv;
    }
    // PixelShuffle(2) fused store: kernel row o' = sub*(Cout/4) + c holds PyTorch channel o = 4c + sub
    const int row = s->store_mode == BP_STORE_PIXSHUF2 ? (o % 4) * c4 + o / 4 : o;
    hb[row] = (float)shift;
    const float* wsrc = s->weight + (size_t)o * Cin * k * k;
    __half* wdst = hw + (size_t)row * pd.wpitch;
    for (int ci = 0; ci < Cin; ++ci)
      for (int r = 0; r < k; ++r)
        for (int q = 0; q < k; ++q) {
          const size_t kidx = stem ? (size_t)r * pd.Cv + q * 8 + ci : (size_t)(r * k + q) * Cin + ci;
          wdst[kidx] = __float2half_rn((float)((double)wsrc[(ci * k + r) * k + q] * scale * in_scale));
        }
  }
}

int bp_pack_conv_weights(const bp_conv_spec* s, int cin, int in_kind, void* w_out, float* b_out, size_t* w_elems, size_t* b_elems) {
  if (!s || cin <= 0 || s->cout <= 0 || s->ksize <= 0 || in_kind > BP_IN_F16) return bp_fail(BP_ERR_INVALID, "bp_pack_conv_weights: bad arguments");
  const bool stem = in_kind >= 0;
  if (stem && (s->ksize > 8 || cin > 8)) return bp_fail(BP_ERR_UNSUPPORTED, "bp_pack_conv_weights: stem convolution geometry (k <= 8, cin <= 8)");
  if (!stem && cin % 32 != 0) return bp_fail(BP_ERR_UNSUPPORTED, "bp_pack_conv_weights: Cin must be a multiple of 32");
  const PackedDims pd = packed_dims(cin, s->ksize, s->cout, stem);
  if (w_elems) *w_elems = (size_t)pd.Cout_pad * pd.wpitch;
  if (b_elems) *b_elems = (size_t)pd.Cout_pad;
  if (!w_out && !b_out) return BP_OK;  // size query
  if (!w_out || !b_out || !s->weight) return bp_fail(BP_ERR_INVALID, "bp_pack_conv_weights: null buffer");
  pack_conv_weights(s, cin, in_kind, reinterpret_cast<__half*>(w_out), b_out);
  return BP_OK;
}

int bp_net_conv(bp_net* n, const bp_conv_spec* s) {
  if (!n || !s || (!s->weight && !s->packed_w)) return bp_fail(BP_ERR_INVALID, "bp_net_conv: null argument");
  if (s->src < 0 || s->src >= (int)n->tensors.size()) return bp_fail(BP_ERR_INVALID, "bp_net_conv: src id");
  cudaSetDevice(n->eng->device);
  const Tensor src = n->tensors[s->src];
  if (src.f32) return bp_fail(BP_ERR_INVALID, "bp_net_conv: fp32 tensors cannot feed a conv");
  const bool stem = src.in_kind >= 0;
  const int Cin = src.C, k = s->ksize, Cout = s->cout;
  const int P = (src.H + 2 * s->pad - k) / s->stride + 1, Q = (src.W + 2 * s->pad - k) / s->stride + 1;
  if (!stem && Cin % 32 != 0) return bp_fail(BP_ERR_UNSUPPORTED, "bp_net_conv: Cin must be a multiple of 32");
  if (s->store_mode == BP_STORE_PIXSHUF2 && (Cout % 32 != 0)) return bp_fail(BP_ERR_UNSUPPORTED, "pixel shuffle needs Cout % 32 == 0");
  if (s->store_mode != BP_STORE_PLAIN && s->out_f32) return bp_fail(BP_ERR_UNSUPPORTED, "fp32 output supports plain stores only");
  if (s->store_mode != BP_STORE_PLAIN && Cout % 8) return bp_fail(BP_ERR_UNSUPPORTED, "fused stores need Cout % 8 == 0");

  // ---- packed weights: folded + packed here (pack_conv_weights above), or taken as they are from a packed-weight cache
  const PackedDims pd = packed_dims(Cin, k, Cout, stem);
  const int Cv = pd.Cv, K = pd.K, wpitch = pd.wpitch, Cout_pad = pd.Cout_pad;
  if (stem && (k > 8 || s->pad > BP_IN_PAD_LEFT || (Q - 1) * s->stride + Cv / 8 > src.row_px - (BP_IN_PAD_LEFT - s->pad)))
    return bp_fail(BP_ERR_UNSUPPORTED, "bp_net_conv: stem convolution geometry (k <= 8, pad <= 3)");
  const size_t n_w = (size_t)Cout_pad * wpitch, n_b = (size_t)Cout_pad;
  std::vector<__half> hw;
  std::vector<float> hb;
  const void* w_host = s->packed_w;
  const float* b_host = s->packed_b;
  if (s->packed_w) {
    if (!s->packed_b || s->packed_w_elems != n_w || s->packed_b_elems != n_b)
      return bp_fail(BP_ERR_INVALID, "bp_net_conv: packed weights do not have this convolution's packed size (stale cache?)");
  } else {
    hw.resize(n_w);
    hb.resize(n_b);
    pack_conv_weights(s, Cin, stem ? src.in_kind : -1, hw.data(), hb.data());
    w_host = hw.data();
    b_host = hb.data();
  }
  // ---- grouped stem: a 3x3 / 1 convolution of the 3-channel network input with few output channels is run with FOUR
  // horizontally adjacent output pixels per GEMM row.  A row's operand is one 128-byte line -- 8 input pixels x 8 channels
  // of one filter row, the window of the four outputs (6 pixels) plus two that meet zero weights -- so a k-block is staged
  // as 128 rows of 128 B instead of 512 rows of 64 B for the same 512 outputs (the ungrouped stem is bound by the TMA row
  // rate), and the result row is 4 pixels x Cout channels = 256 contiguous bytes of the NHWC output, a dense TMA store.
  // The GEMM grows to N = 4 Cout, K = 3 x 64 (mostly zeros: 7x the useful MACs, still far below the layer's HBM time).
  // The weights are derived here from the ordinary packed stem layout, so weight files and the packed cache do not change.
  const bool grouped = stem && k == 3 && s->stride == 1 && s->pad == 1 && Cin <= 8 && Cout % 8 == 0 && Cout * 4 <= 128 && Q % 4 == 0 &&
                       s->store_mode == BP_STORE_PLAIN && !s->out_f32 && s->res < 0 && s->dst < 0 && Cv == 32 && !getenv("BP_NO_GROUPED_STEM");
  std::vector<__half> gw;
  std::vector<float> gb;
  int wpitch_dev = wpitch;
  size_t n_w_dev = n_w, n_b_dev = n_b;
  if (grouped) {
    const int G = 4, Kg = 3 * 64;
    gw.assign((size_t)Cout_pad * Kg, __float2half(0.f));
    gb.assign((size_t)Cout_pad, 0.f);
    const __half* hw_src = reinterpret_cast<const __half*>(w_host);
    for (int j = 0; j < G; ++j)
      for (int co = 0; co < Cout; ++co) {
        gb[(size_t)j * Cout + co] = b_host[co];
        for (int r = 0; r < 3; ++r)
          for (int px = j; px < j + 3; ++px)  // window pixel px holds filter column px - j of output j
            for (int ci = 0; ci < 8; ++ci)
              gw[((size_t)j * Cout + co) * Kg + r * 64 + px * 8 + ci] = hw_src[(size_t)co * wpitch + r * Cv + (px - j) * 8 + ci];
      }
    w_host = gw.data();
    b_host = gb.data();
    wpitch_dev = Kg;
    n_w_dev = gw.size();
    n_b_dev = gb.size();
  }
  __half* dw = (__half*)net_alloc_weights(n, n_w_dev * 2);
  float* db = (float*)net_alloc_weights(n, n_b_dev * 4);
  if (!dw || !db) return bp_fail(BP_ERR_CUDA, "bp_net_conv: cudaMalloc (weights) failed");
  if (cudaMemcpy(dw, w_host, n_w_dev * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(db, b_host, n_b_dev * 4, cudaMemcpyHostToDevice) != cudaSuccess)
    return bp_fail(BP_ERR_CUDA, "bp_net_conv: weight upload failed");

  // ---- destination tensor
  const int up = s->store_mode == BP_STORE_PLAIN ? 1 : 2;
  const int oc = s->store_mode == BP_STORE_PIXSHUF2 ? Cout / 4 : Cout;
  int dst = s->dst;
  int coff = 0;
  if (dst < 0) {
    dst = new_tensor(n, P * up, Q * up, oc, s->out_f32 != 0);
    if (dst < 0) return bp_fail(BP_ERR_CUDA, "bp_net_conv: cudaMalloc (activation) failed");
  } else {
    if (dst >= (int)n->tensors.size()) return bp_fail(BP_ERR_INVALID, "bp_net_conv: dst id");
    const Tensor& d = n->tensors[dst];
    coff = s->dst_coff;
    if (d.H != P * up || d.W != Q * up || coff + oc > d.C || coff % 8 || d.f32 != (s->out_f32 != 0))
      return bp_fail(BP_ERR_INVALID, "bp_net_conv: dst tensor does not fit the output");
  }
  const Tensor& dt = n->tensors[dst];

  ConvDesc d;
  if (stem) {
    // virtual pixel w' = buffer column (BP_IN_PAD_LEFT - pad) + w'; the window of output column q starts at w' = q*stride
    d.x = (const __half*)src.ptr + (size_t)(BP_IN_PAD_LEFT - s->pad) * 8;
    d.N = n->max_batch; d.H = src.H; d.W = (Q - 1) * s->stride + 1; d.C = Cv; d.x_pitch = 8;
    d.x_row_pitch = (long)src.row_px * 8; d.x_img_pitch = (long)src.H * src.row_px * 8;
    d.R = k; d.S = 1; d.stride = s->stride; d.pad = s->pad; d.pad_w = 0;
    d.real_k = (double)k * k * Cin;
    if (grouped) {  // virtual pixels of 8 real pixels, one GEMM row per 4 output columns
      d.W = (Q / 4 - 1) * 4 + 1;
      d.C = 64;
      d.stride_w = 4;
      d.force_mt = 2;
    }
  } else {
    d.x = (const __half*)src.ptr; d.N = n->max_batch; d.H = src.H; d.W = src.W; d.C = Cin; d.x_pitch = src.pitch;
    d.R = k; d.S = k; d.stride = s->stride; d.pad = s->pad;
  }
  d.w = dw; d.bias = db; d.w_pitch = wpitch_dev; d.Cout = grouped ? 4 * Cout : Cout; d.Cout_pad = Cout_pad;
  d.act = s->act;
  if (s->res >= 0) {
    if (s->res >= (int)n->tensors.size()) return bp_fail(BP_ERR_INVALID, "bp_net_conv: res id");
    const Tensor& r = n->tensors[s->res];
    if (r.f32 || r.H != P || r.W != Q || r.C != Cout || s->store_mode != BP_STORE_PLAIN)
      return bp_fail(BP_ERR_INVALID, "bp_net_conv: residual tensor shape mismatch");
    d.res = (const __half*)r.ptr; d.res_pitch = r.pitch; d.res_mode = s->res_mode;
  }
  d.out = dt.f32 ? (void*)((float*)dt.ptr) : (void*)((__half*)dt.ptr);
  d.out_pitch = dt.pitch; d.out_coff = coff; d.out_f32 = s->out_f32; d.store_mode = s->store_mode;
  if (grouped) {
    if (dt.pitch != Cout || coff != 0) return bp_fail(BP_ERR_UNSUPPORTED, "bp_net_conv: grouped stem needs a dense output tensor");
    d.out_pitch = 4 * Cout;  // one GEMM row = 4 pixels of the dense NHWC output
  }
  if (n->eng->force_block_n) d.force_block_n = n->eng->force_block_n;
  if (n->eng->force_stages) d.force_stages = n->eng->force_stages;

  d.num_sms = n->eng->num_sms;
  d.sk_ws = n->sk_ws;
  d.sk_flags = n->sk_flags;
  Op op;
  op.kind = OP_CONV;
  op.cdesc = d;
  op.stem = stem;
  op.pq = P * Q;
  std::string err;
  ConvPlan& plan0 = op.plans[n->max_batch];
  if (!conv_plan_build(n->eng->tmap, &plan0, d, &err)) return bp_fail(BP_ERR_CUDA, ("bp_net_conv: " + err).c_str());
  op.dst = dst;
  op.flops = 2.0 * P * Q * (double)Cout * (k * k * Cin);
  op.bytes = (stem ? (double)src.H * src.row_px * 16 : (double)src.H * src.W * Cin * 2) +
             (double)K * Cout * 2 / n->max_batch +
             (double)P * Q * Cout * (s->out_f32 ? 4 : 2) * (s->store_mode == BP_STORE_UPSAMPLE2 ? 4 : 1) +
             (s->res >= 0 ? (double)P * Q * Cout * 2 : 0.0);
  char buf[160];
  char tile[32] = "";
  if (plan0.cg == 2) snprintf(tile, sizeof tile, " cg2");
  else if (plan0.mt > 1) snprintf(tile, sizeof tile, " mt%d", plan0.mt);
  if (grouped) snprintf(tile + strlen(tile), sizeof tile - strlen(tile), " g4");
  if (plan0.streamk) snprintf(tile + strlen(tile), sizeof tile - strlen(tile), " sk");
  snprintf(buf, sizeof buf, "conv %dx%d/%d %d->%d @%dx%d bn%d bk%d st%d%s%s%s%s", k, k, s->stride, Cin, Cout, P, Q,
           plan0.block_n, plan0.block_k, plan0.stages, tile, s->res >= 0 ? " +res" : "",
           s->store_mode == BP_STORE_UPSAMPLE2 ? " up2" : (s->store_mode == BP_STORE_PIXSHUF2 ? " ps2" : ""),
           s->out_f32 ? " f32" : "");
  op.desc = buf;
  n->ops.push_back(op);
  n->flops += op.flops;
  // the tensor id a consumer reads: for dst given by the caller, a view restricted to our channel window
  if (s->dst >= 0 && (coff != 0 || oc != dt.C)) return bp_net_view(n, dst, coff, oc);
  return dst;
}

static int push_aux(bp_net* n, OpKind kind, int a, int b, int c, int dst, double bytes, const char* name) {
  Op op;
  op.kind = kind;
  op.a = a; op.b = b; op.c = c; op.dst = dst;
  op.bytes = bytes;
  const Tensor& t = n->tensors[dst];
  char buf[128];
  snprintf(buf, sizeof buf, "%s -> %dx%dx%d", name, t.H, t.W, t.C);
  op.desc = buf;
  n->ops.push_back(op);
  return dst;
}

static bool bad_id(bp_net* n, int id) { return !n || id < 0 || id >= (int)n->tensors.size() || n->tensors[id].f32; }

int bp_net_maxpool3x3s2(bp_net* n, int src) {
  if (bad_id(n, src) || n->tensors[src].C % 8) return bp_fail(BP_ERR_INVALID, "bp_net_maxpool3x3s2: src");
  const Tensor s = n->tensors[src];
  const int dst = new_tensor(n, (s.H + 2 - 3) / 2 + 1, (s.W + 2 - 3) / 2 + 1, s.C, false);
  if (dst < 0) return bp_fail(BP_ERR_CUDA, "cudaMalloc failed");
  const Tensor& d = n->tensors[dst];
  return push_aux(n, OP_MAXPOOL, src, -1, -1, dst, (double)s.H * s.W * s.C * 2 + (double)d.H * d.W * d.C * 2, "maxpool3x3/2");
}

int bp_net_global_avgpool(bp_net* n, int src) {
  if (bad_id(n, src) || n->tensors[src].C % 8) return bp_fail(BP_ERR_INVALID, "bp_net_global_avgpool: src");
  const Tensor s = n->tensors[src];
  const int dst = new_tensor(n, 1, 1, s.C, false);
  if (dst < 0) return bp_fail(BP_ERR_CUDA, "cudaMalloc failed");
  const int cblks = (s.C + 255) / 256;
  float* scratch = (float*)net_alloc_weights(n, (size_t)n->max_batch * kAvgSplitMax * s.C * sizeof(float));
  unsigned* counters = (unsigned*)net_alloc_weights(n, (size_t)n->max_batch * cblks * sizeof(unsigned));
  if (!scratch || !counters || cudaMemset(counters, 0, (size_t)n->max_batch * cblks * sizeof(unsigned)) != cudaSuccess)
    return bp_fail(BP_ERR_CUDA, "bp_net_global_avgpool: cudaMalloc (scratch) failed");
  const int r = push_aux(n, OP_AVGPOOL, src, -1, -1, dst, (double)s.H * s.W * s.C * 2, "global_avgpool");
  n->ops.back().scratch = scratch;
  n->ops.back().counters = counters;
  return r;
}

int bp_net_scale_add_relu(bp_net* n, int y, int gates, int skip) {
  if (bad_id(n, y) || bad_id(n, gates) || bad_id(n, skip)) return bp_fail(BP_ERR_INVALID, "bp_net_scale_add_relu: ids");
  const Tensor ty = n->tensors[y], tg = n->tensors[gates], ts = n->tensors[skip];
  if (tg.C != ty.C || ts.C != ty.C || ts.H != ty.H || ts.W != ty.W || ty.C % 8)
    return bp_fail(BP_ERR_INVALID, "bp_net_scale_add_relu: shapes");
  const int dst = new_tensor(n, ty.H, ty.W, ty.C, false);
  if (dst < 0) return bp_fail(BP_ERR_CUDA, "cudaMalloc failed");
  return push_aux(n, OP_SCALE_ADD_RELU, y, gates, skip, dst, 3.0 * ty.H * ty.W * ty.C * 2, "scale_add_relu");
}

int bp_net_pixel_shuffle2(bp_net* n, int src) {
  if (bad_id(n, src) || n->tensors[src].C % 32) return bp_fail(BP_ERR_INVALID, "bp_net_pixel_shuffle2: src");
  const Tensor s = n->tensors[src];
  const int dst = new_tensor(n, s.H * 2, s.W * 2, s.C / 4, false);
  if (dst < 0) return bp_fail(BP_ERR_CUDA, "cudaMalloc failed");
  return push_aux(n, OP_PIXSHUF, src, -1, -1, dst, 2.0 * s.H * s.W * s.C * 2, "pixel_shuffle2");
}

int bp_net_upsample2(bp_net* n, int src, int dst, int dst_coff) {
  if (bad_id(n, src)) return bp_fail(BP_ERR_INVALID, "bp_net_upsample2: src");
  const Tensor s = n->tensors[src];
  int d = dst;
  if (d < 0) {
    d = new_tensor(n, s.H * 2, s.W * 2, s.C, false);
    if (d < 0) return bp_fail(BP_ERR_CUDA, "cudaMalloc failed");
  } else {
    d = bp_net_view(n, dst, dst_coff, s.C);
    if (d < 0) return d;
    if (n->tensors[d].H != 2 * s.H || n->tensors[d].W != 2 * s.W) return bp_fail(BP_ERR_INVALID, "bp_net_upsample2: dst");
  }
  return push_aux(n, OP_UPSAMPLE, src, -1, -1, d, 5.0 * s.H * s.W * s.C * 2, "upsample2");
}

int bp_net_copy_channels(bp_net* n, int src, int dst, int dst_coff) {
  if (bad_id(n, src) || bad_id(n, dst)) return bp_fail(BP_ERR_INVALID, "bp_net_copy_channels: ids");
  const Tensor s = n->tensors[src];
  const int d = bp_net_view(n, dst, dst_coff, s.C);
  if (d < 0) return d;
  if (n->tensors[d].H != s.H || n->tensors[d].W != s.W) return bp_fail(BP_ERR_INVALID, "bp_net_copy_channels: dst");
  return push_aux(n, OP_COPYC, src, -1, -1, d, 2.0 * s.H * s.W * s.C * 2, "copy_channels");
}

int bp_net_add(bp_net* n, int a, int b) {
  if (bad_id(n, a) || bad_id(n, b)) return bp_fail(BP_ERR_INVALID, "bp_net_add: ids");
  const Tensor ta = n->tensors[a], tb = n->tensors[b];
  if (ta.H != tb.H || ta.W != tb.W || ta.C != tb.C || ta.C % 8) return bp_fail(BP_ERR_INVALID, "bp_net_add: shapes");
  const int dst = new_tensor(n, ta.H, ta.W, ta.C, false);
  if (dst < 0) return bp_fail(BP_ERR_CUDA, "cudaMalloc failed");
  return push_aux(n, OP_ADD, a, b, -1, dst, 3.0 * ta.H * ta.W * ta.C * 2, "add");
}

int bp_net_tensor_info(bp_net* n, int tensor, int* dims, void** ptr) {
  if (!n || tensor < 0 || tensor >= (int)n->tensors.size()) return bp_fail(BP_ERR_INVALID, "bp_net_tensor_info: id");
  const Tensor& t = n->tensors[tensor];
  if (dims) {
    dims[0] = t.H; dims[1] = t.W; dims[2] = t.C; dims[3] = t.pitch; dims[4] = t.f32 ? 1 : 0; dims[5] = t.coff;
    dims[6] = t.row_px ? t.row_px : t.W;      // pixels per buffer row
    dims[7] = t.row_px ? BP_IN_PAD_LEFT : 0;  // first data column
  }
  if (ptr) *ptr = t.ptr;
  return BP_OK;
}

int bp_net_num_ops(bp_net* n) { return n ? (int)n->ops.size() : 0; }

int bp_net_set_op_config(bp_net* n, int op, int batch, int block_n, int cg, int mt) {
  if (!n || op < 0 || op >= (int)n->ops.size() || batch <= 0 || batch > n->max_batch) return bp_fail(BP_ERR_INVALID, "bp_net_set_op_config: op / batch");
  Op& o = n->ops[op];
  if (o.kind != OP_CONV) return bp_fail(BP_ERR_INVALID, "bp_net_set_op_config: not a convolution");
  o.plans.erase(batch);
  if (block_n == 0 && cg == 0 && mt == 0) {
    o.tuned.erase(batch);
    return BP_OK;
  }
  o.tuned[batch] = Op::TileCfg{block_n, cg, mt};
  // build now so that an unsupported combination is reported here, not at the first launch
  ConvDesc d = o.cdesc;
  d.N = batch;
  d.force_block_n = block_n;
  d.force_cg = cg;
  d.force_mt = mt;
  if (n->share_batch > 0) return BP_OK;  // shared nets re-plan with their SM budget at launch
  std::string err;
  ConvPlan pl;
  bool ok = conv_plan_build(n->eng->tmap, &pl, d, &err);
  if (ok && ((block_n && pl.block_n != block_n) || (cg && pl.cg != cg) || (mt && pl.mt != mt))) {
    ok = false;
    err = "the planner does not support this combination for this layer";
  }
  if (ok && !conv_plan_supported(pl)) {
    ok = false;
    err = "no kernel instantiation for this tile configuration";
  }
  if (!ok) {
    o.tuned.erase(batch);
    return bp_fail(BP_ERR_UNSUPPORTED, ("bp_net_set_op_config: " + err).c_str());
  }
  o.plans[batch] = pl;
  return BP_OK;
}

int bp_net_op_config(bp_net* n, int op, int batch, int* cfg) {
  if (!n || op < 0 || op >= (int)n->ops.size() || !cfg) return bp_fail(BP_ERR_INVALID, "bp_net_op_config: op");
  Op& o = n->ops[op];
  cfg[0] = cfg[1] = cfg[2] = cfg[3] = cfg[4] = 0;
  if (o.kind != OP_CONV) return BP_OK;
  auto it = o.plans.find(batch);
  if (it == o.plans.end()) {
    ConvDesc d = o.cdesc;
    d.N = batch;
    auto tu = o.tuned.find(batch);
    if (tu != o.tuned.end()) { d.force_block_n = tu->second.block_n; d.force_cg = tu->second.cg; d.force_mt = tu->second.mt; }
    std::string err;
    ConvPlan pl;
    if (!conv_plan_build(n->eng->tmap, &pl, d, &err)) return bp_fail(BP_ERR_CUDA, err.c_str());
    it = o.plans.emplace(batch, pl).first;
  }
  cfg[0] = it->second.block_n; cfg[1] = it->second.cg; cfg[2] = it->second.mt; cfg[3] = it->second.block_k; cfg[4] = it->second.stages;
  return BP_OK;
}

int bp_net_set_share(bp_net* n, int share_batch) {
  if (!n || share_batch < 0) return bp_fail(BP_ERR_INVALID, "bp_net_set_share: bad arguments");
  if (n->share_batch != share_batch) {
    n->share_batch = share_batch;
    for (Op& op : n->ops) op.plans.clear();  // plans depend on the SM budget
  }
  return BP_OK;
}
int bp_net_num_launches(bp_net* n) { return n ? (int)n->ops.size() : 0; }
double bp_net_flops_per_image(bp_net* n) { return n ? n->flops : 0.0; }

int bp_net_op_desc(bp_net* n, int op, char* buf, int buflen, double* flops, double* bytes) {
  if (!n || op < 0 || op >= (int)n->ops.size()) return bp_fail(BP_ERR_INVALID, "bp_net_op_desc: op");
  if (buf && buflen > 0) snprintf(buf, buflen, "%s", n->ops[op].desc.c_str());
  if (flops) *flops = n->ops[op].flops;
  if (bytes) *bytes = n->ops[op].bytes;
  return BP_OK;
}

static inline unsigned blocks_for(long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

int bp_net_forward_range(bp_net* n, int batch, int first, int last, void* stream_) {
  if (!n || batch <= 0 || batch > n->max_batch) return bp_fail(BP_ERR_INVALID, "bp_net_forward: batch out of range");
  if (first < 0 || last > (int)n->ops.size() || first > last) return bp_fail(BP_ERR_INVALID, "bp_net_forward: op range");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  for (int i = first; i < last; ++i) {
    const Op& op = n->ops[i];
    cudaError_t e = cudaSuccess;
    switch (op.kind) {
      case OP_CONV: {
        Op& mop = n->ops[i];
        auto it = mop.plans.find(batch);
        if (it == mop.plans.end()) {
          // first use of this batch size: tile configuration + TMA descriptors for exactly `batch` images
          ConvDesc d = mop.cdesc;
          d.N = batch;
          auto tu = mop.tuned.find(batch);
          if (tu != mop.tuned.end()) {
            d.force_block_n = tu->second.block_n;
            d.force_cg = tu->second.cg;
            d.force_mt = tu->second.mt;
          }
          if (n->share_batch > 0) {
            // this object's share of the machine, in proportion to its frames (even, >= 2: CTA pairs stay possible)
            int budget = (int)(((long)n->eng->num_sms * batch + n->share_batch - 1) / n->share_batch);
            budget = std::max(2, std::min(n->eng->num_sms, (budget + 1) & ~1));
            d.num_sms = budget;
            d.pdl = false;
          }
          std::string err;
          ConvPlan& np = mop.plans[batch];
          if (!conv_plan_build(n->eng->tmap, &np, d, &err)) {
            mop.plans.erase(batch);
            return bp_fail(BP_ERR_CUDA, ("bp_net_forward: " + err).c_str());
          }
          it = mop.plans.find(batch);
        }
        e = conv_plan_launch(it->second, st);
        break;
      }
      case OP_MAXPOOL: {
        const Tensor &s = n->tensors[op.a], &d = n->tensors[op.dst];
        const long total = (long)batch * d.H * d.W * (d.C / 8);
        maxpool3x3s2_kernel<<<blocks_for(total, 256), 256, 0, st>>>(view_of(s), view_of(d), batch);
        e = cudaGetLastError();
        break;
      }
      case OP_AVGPOOL: {
        const Tensor &s = n->tensors[op.a], &d = n->tensors[op.dst];
        const int cblks = (s.C + 255) / 256;
        // summation order: S0 virtual slices of >= 32 positions, a function of the tensor shape only (batch-independent
        // results); blocks per image: enough for ~4 per SM
        const int S0 = std::max(1, std::min(kAvgSplitMax, (s.H * s.W + 31) / 32));
        int S = (4 * n->eng->num_sms + cblks * batch - 1) / (cblks * batch);
        S = std::max(1, std::min(S, S0));
        dim3 grid(cblks, batch, S);
        global_avgpool_kernel<<<grid, 256, 0, st>>>(view_of(s), (__half*)d.ptr, d.pitch, op.scratch, op.counters, S, S0);
        e = cudaGetLastError();
        break;
      }
      case OP_SCALE_ADD_RELU: {
        const Tensor &y = n->tensors[op.a], &g = n->tensors[op.b], &k = n->tensors[op.c], &d = n->tensors[op.dst];
        const long total = (long)batch * y.H * y.W * (y.C / 8);
        scale_add_relu_kernel<<<blocks_for(total, 256), 256, 0, st>>>(view_of(y), (const __half*)g.ptr, g.pitch, view_of(k),
                                                                     view_of(d), batch);
        e = cudaGetLastError();
        break;
      }
      case OP_PIXSHUF: {
        const Tensor &s = n->tensors[op.a], &d = n->tensors[op.dst];
        const long total = (long)batch * s.H * s.W * (s.C / 32);
        pixel_shuffle2_kernel<<<blocks_for(total, 256), 256, 0, st>>>(view_of(s), view_of(d), batch);
        e = cudaGetLastError();
        break;
      }
      case OP_UPSAMPLE: {
        const Tensor &s = n->tensors[op.a], &d = n->tensors[op.dst];
        const long total = (long)batch * d.H * d.W * (s.C / 8);
        upsample2_kernel<<<blocks_for(total, 256), 256, 0, st>>>(view_of(s), view_of(d), batch);
        e = cudaGetLastError();
        break;
      }
      case OP_COPYC: {
        const Tensor &s = n->tensors[op.a], &d = n->tensors[op.dst];
        const long total = (long)batch * s.H * s.W * (s.C / 8);
        copy_channels_kernel<<<blocks_for(total, 256), 256, 0, st>>>(view_of(s), view_of(d), batch);
        e = cudaGetLastError();
        break;
      }
      case OP_ADD: {
        const Tensor &a = n->tensors[op.a], &b = n->tensors[op.b], &d = n->tensors[op.dst];
        const long total = (long)batch * a.H * a.W * (a.C / 8);
        add_kernel<<<blocks_for(total, 256), 256, 0, st>>>(view_of(a), view_of(b), view_of(d), batch);
        e = cudaGetLastError();
        break;
      }
    }
    if (e != cudaSuccess) {
      char buf[256];
      snprintf(buf, sizeof buf, "bp_net_forward: op %d (%s): %s", i, op.desc.c_str(), cudaGetErrorString(e));
      return bp_fail(BP_ERR_CUDA, buf);
    }
  }
  return BP_OK;
}

int bp_net_forward(bp_net* n, int batch, void* stream) {
  return bp_net_forward_range(n, batch, 0, n ? (int)n->ops.size() : 0, stream);
}

}  // extern "C"
