// Greedy IoU-NMS over the candidates of ONE image, written once for the CUDA block (box_nms.cu) and for a serial host
// build (tests/harness/box_nms_host.cu: the CPU test-suite runs the very same code with a one-thread "block").
//
// Restates the branch of write_results the reference ships switched off (yolo/util.py:182-196 behind `nms = False` at :181, with
// bbox_iou, yolo/bbox.py:51-77): candidates = rows with objectness > conf (and class arg-max 0, util.py:166-167), sorted by
// objectness, descending (ties: lower row first), then repeatedly keep the best remaining box and drop every later one
// whose IoU with it is not < nms_thr.  IoU uses the reference's "+1 pixel" convention, fp32, one rounding per operation.
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define BP_NMS_ADD(a, b) __fadd_rn((a), (b))
#define BP_NMS_SUB(a, b) __fsub_rn((a), (b))
#define BP_NMS_MUL(a, b) __fmul_rn((a), (b))
#define BP_NMS_DIV(a, b) __fdiv_rn((a), (b))
#else  // host build: compiled with -ffp-contract=off, so every operation rounds once as well
#define BP_NMS_ADD(a, b) ((a) + (b))
#define BP_NMS_SUB(a, b) ((a) - (b))
#define BP_NMS_MUL(a, b) ((a) * (b))
#define BP_NMS_DIV(a, b) ((a) / (b))
#endif

namespace bp_nms {

struct Box {
  float x1, y1, x2, y2;
};

__host__ __device__ inline Box corners(const float* row) {
  const float hw = BP_NMS_DIV(row[2], 2.f), hh = BP_NMS_DIV(row[3], 2.f);
  return Box{BP_NMS_SUB(row[0], hw), BP_NMS_SUB(row[1], hh), BP_NMS_ADD(row[0], hw), BP_NMS_ADD(row[1], hh)};
}

__host__ __device__ inline float fmax_(float a, float b) { return a > b ? a : b; }
__host__ __device__ inline float fmin_(float a, float b) { return a < b ? a : b; }

__host__ __device__ inline float iou_plus1(const Box& a, const Box& b) {
  const float ix1 = fmax_(a.x1, b.x1), iy1 = fmax_(a.y1, b.y1), ix2 = fmin_(a.x2, b.x2), iy2 = fmin_(a.y2, b.y2);
  const float iw = fmax_(BP_NMS_ADD(BP_NMS_SUB(ix2, ix1), 1.f), 0.f), ih = fmax_(BP_NMS_ADD(BP_NMS_SUB(iy2, iy1), 1.f), 0.f);
  const float inter = BP_NMS_MUL(iw, ih);
  const float aa = BP_NMS_MUL(BP_NMS_ADD(BP_NMS_SUB(a.x2, a.x1), 1.f), BP_NMS_ADD(BP_NMS_SUB(a.y2, a.y1), 1.f));
  const float ab = BP_NMS_MUL(BP_NMS_ADD(BP_NMS_SUB(b.x2, b.x1), 1.f), BP_NMS_ADD(BP_NMS_SUB(b.y2, b.y1), 1.f));
  return BP_NMS_DIV(inter, BP_NMS_SUB(BP_NMS_ADD(aa, ab), inter));
}

// sort key: objectness in the high word (mapped so that unsigned order = float order), inverted row in the low word, so
// one descending sort gives "objectness descending, lower row first".  Real keys are never 0 (row < 2^32 - 1); 0 pads.
__host__ __device__ inline unsigned long long make_key(float obj, uint32_t row) {
  union {
    float f;
    uint32_t u;
  } c;
  c.f = obj;
  const uint32_t ordered = (c.u & 0x80000000u) ? ~c.u : (c.u | 0x80000000u);
  return ((unsigned long long)ordered << 32) | (0xFFFFFFFFu - row);
}
__host__ __device__ inline uint32_t key_row(unsigned long long k) { return 0xFFFFFFFFu - uint32_t(k & 0xFFFFFFFFull); }

struct HostBlock {  // a block of one thread
  __host__ __device__ int tid() const { return 0; }
  __host__ __device__ int size() const { return 1; }
  __host__ __device__ void sync() const {}
  __host__ __device__ int fetch_add(int* p, int v) const {
    const int o = *p;
    *p += v;
    return o;
  }
};

// pred: the image's rows [R, n_attr] (cx, cy, w, h, obj, cls...).  keys[cap] (cap = power of two >= R), supp[cap] and
// counter[1] are block-shared scratch.  Outputs: out_det[max_det, 8], out_row[max_det], *out_count = min(kept, max_det),
// *out_total = kept (every survivor is counted even past max_det: dynamic_write_results' "> 100 detections" retry needs it).
template <class Block>
__host__ __device__ inline void nms_image(const Block& blk, const float* pred, int R, int n_attr, float conf, float nms_thr, int max_det,
                                          int img, unsigned long long* keys, uint8_t* supp, int* counter, float* out_det,
                                          int32_t* out_row, int32_t* out_count, int32_t* out_total) {
  const int tid = blk.tid(), nt = blk.size();
  if (tid == 0) *counter = 0;
  blk.sync();
  // 1. candidates (any order: they are sorted next)
  for (int r = tid; r < R; r += nt) {
    const float* p = pred + (long)r * n_attr;
    bool cand = p[4] > conf;
    if (cand && n_attr > 6) {
      int am = 0;
      float mv = p[5];
      for (int c = 1; c < n_attr - 5; ++c)
        if (p[5 + c] > mv) {
          mv = p[5 + c];
          am = c;
        }
      cand = am == 0;
    }
    if (cand) keys[blk.fetch_add(counter, 1)] = make_key(p[4], uint32_t(r));
  }
  blk.sync();
  const int count = *counter;
  int n2 = 1;
  while (n2 < count) n2 <<= 1;
  for (int i = count + tid; i < n2; i += nt) keys[i] = 0ull;
  for (int i = tid; i < count; i += nt) supp[i] = 0;
  blk.sync();
  // 2. bitonic sort, descending
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < n2; t += nt) {
        const int u = t ^ j;
        if (u > t) {
          const unsigned long long a = keys[t], b = keys[u];
          const bool descending = (t & k) == 0;
          if (descending ? a < b : a > b) {
            keys[t] = b;
            keys[u] = a;
          }
        }
      }
      blk.sync();
    }
  }
  // 3. greedy suppression; every thread walks the same sequence (supp[i] is final once i is reached)
  int kept = 0;
  for (int i = 0; i < count; ++i) {
    if (supp[i]) continue;
    const uint32_t row = key_row(keys[i]);
    const float* p = pred + (long)row * n_attr;
    const Box bi = corners(p);
    if (tid == 0 && kept < max_det) {
      float* d = out_det + (long)kept * 8;
      d[0] = (float)img;
      d[1] = bi.x1;
      d[2] = bi.y1;
      d[3] = bi.x2;
      d[4] = bi.y2;
      d[5] = p[4];
      d[6] = p[5];
      d[7] = 0.f;
      out_row[kept] = (int32_t)row;
    }
    ++kept;
    for (int j = i + 1 + tid; j < count; j += nt) {
      if (supp[j]) continue;
      const Box bj = corners(pred + (long)key_row(keys[j]) * n_attr);
      if (!(iou_plus1(bi, bj) < nms_thr)) supp[j] = 1;  // NaN compares false and is dropped, like `ious < nms_conf` indexing
    }
    blk.sync();
  }
  if (tid == 0) {
    *out_count = kept < max_det ? kept : max_det;
    *out_total = kept;
  }
}

}  // namespace bp_nms
