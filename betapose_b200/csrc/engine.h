// Internal definition of the opaque engine handle shared by the translation units of libbetapose_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "tmap.cuh"

struct ResizeTables {  // Pillow bicubic coefficient tables for one (in, out) size pair, device resident
  int in_size = 0, out_size = 0, ksize = 0;
  int32_t* bounds = nullptr;  // [out, 2] (first tap, tap count)
  int32_t* coeffs = nullptr;  // [out, ksize] 22-bit fixed point
  std::vector<int32_t> host_bounds;  // host copy of `bounds` (tile planning of the fused resize kernel)
};

struct bp_engine {
  int device = 0;
  bp::TmapApi tmap;
  int num_sms = 148;
  int force_block_n = 0;  // tuning overrides (env BP_FORCE_BLOCK_N / BP_FORCE_STAGES)
  int force_stages = 0;
  std::map<std::pair<int, int>, ResizeTables> resize_tables;
  float* crop_lut = nullptr;  // [3][256] (u / 255) - mean[c], built on first use (bp_crop_resize)
  // Kernel scratch is kept PER STREAM (calls on one stream are ordered, calls on different streams never share a
  // block) and is grow-only: a block that became too small is retired into `owned`, never freed while the engine
  // lives, because CUDA graphs captured earlier have its address baked in (BetaposeEngine captures one graph per
  // batch size; the native engine is a per-device singleton shared by every Python-side engine).
  struct StreamScratch {
    uint8_t* resize_tmp = nullptr;  // horizontal-pass intermediate of the two-kernel resize fall-back
    size_t resize_tmp_bytes = 0;
    void* hm = nullptr;  // heat-map decode: per-slice partial arg-max + arrival counters
    size_t hm_bytes = 0;
    int hm_n = 0;
    double* pnp = nullptr;  // per-hypothesis rows between the two PnP kernels
    size_t pnp_bytes = 0;
  };
  std::map<cudaStream_t, StreamScratch> scratch;
  std::mutex mu;  // guards `scratch`, `owned`, `resize_tables`
  std::vector<void*> owned;

  StreamScratch& scratch_for(cudaStream_t st) {
    std::lock_guard<std::mutex> g(mu);
    return scratch[st];
  }
  // grow-only (re)allocation of one scratch block; the old block is retired, not freed.  Returns false on failure.
  bool grow(void** block, size_t* have, size_t need) {
    if (*have >= need) return true;
    void* p = nullptr;
    if (cudaMalloc(&p, need) != cudaSuccess) return false;
    std::lock_guard<std::mutex> g(mu);
    if (*block) owned.push_back(*block);
    *block = p;
    *have = need;
    return true;
  }
};

// cudaFuncSetAttribute applies to the current device only; an engine exists per device, so "done" is tracked per device
// (up to 64) in a static table of the calling site
inline bool bp_attr_once(bool (&done)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

// records the message for bp_last_error() and returns `code`
int bp_fail(int code, const char* msg);
