// Engine lifetime + error reporting for the C ABI (include/betapose_b200.h).
#include <cstdlib>
#include <cstring>
#include <string>

#include "betapose_b200.h"
#include "engine.h"

static thread_local std::string g_last_error;

int bp_fail(int code, const char* msg) {
  g_last_error = msg ? msg : "";
  return code;
}

extern "C" {

const char* bp_last_error(void) { return g_last_error.c_str(); }
int bp_version(void) { return 100; }

int bp_engine_create(int device, bp_engine** out) {
  if (!out) return bp_fail(BP_ERR_INVALID, "bp_engine_create: null out");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return bp_fail(BP_ERR_CUDA, "bp_engine_create: no CUDA device (this engine has no CPU path)");
  if (device < 0 || device >= count) return bp_fail(BP_ERR_INVALID, "bp_engine_create: device index");
  if (cudaSetDevice(device) != cudaSuccess) return bp_fail(BP_ERR_CUDA, "bp_engine_create: cudaSetDevice failed");
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    std::string m = std::string("bp_engine_create: device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                    ", this library is built for sm_100a only";
    return bp_fail(BP_ERR_UNSUPPORTED, m.c_str());
  }
  cudaFree(0);
  bp_engine* e = new bp_engine();
  e->device = device;
  e->num_sms = prop.multiProcessorCount;
  std::string err;
  if (!e->tmap.load(&err)) {
    delete e;
    return bp_fail(BP_ERR_CUDA, err.c_str());
  }
  if (const char* s = getenv("BP_FORCE_BLOCK_N")) e->force_block_n = atoi(s);
  if (const char* s = getenv("BP_FORCE_STAGES")) e->force_stages = atoi(s);
  *out = e;
  return BP_OK;
}

void bp_engine_destroy(bp_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  for (void* p : e->owned) cudaFree(p);
  for (auto& kv : e->scratch) {
    if (kv.second.resize_tmp) cudaFree(kv.second.resize_tmp);
    if (kv.second.pnp) cudaFree(kv.second.pnp);
    if (kv.second.hm) cudaFree(kv.second.hm);
  }
  delete e;
}

}  // extern "C"
