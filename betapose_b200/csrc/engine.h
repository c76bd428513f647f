// Internal definition of the opaque engine handle shared by the translation units of libbetapose_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
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
  uint8_t* resize_tmp = nullptr;  // horizontal-pass intermediate
  size_t resize_tmp_bytes = 0;
  void* hm_scratch = nullptr;  // heat-map decode: per-slice partial arg-max + arrival counters
  size_t hm_scratch_bytes = 0;
  int hm_scratch_n = 0;
  double* pnp_scratch = nullptr;  // per-hypothesis rows between the two PnP kernels
  size_t pnp_scratch_bytes = 0;
  std::vector<void*> owned;
};

// records the message for bp_last_error() and returns `code`
int bp_fail(int code, const char* msg);
