// bp_write_results_nms: write_results with the IoU-NMS branch on (multi-instance scenes).
// One block per image; candidates, a bitonic sort of (objectness, row) keys and the suppression flags live in shared
// memory, the greedy loop is serial over kept boxes and block-parallel over the boxes each one may suppress
// (box_nms.cuh, shared with the host build the CPU tests run).
#include <cuda_runtime.h>

#include "betapose_b200.h"
#include "box_nms.cuh"
#include "engine.h"

namespace {

struct CudaBlock {
  __device__ int tid() const { return threadIdx.x; }
  __device__ int size() const { return blockDim.x; }
  __device__ void sync() const { __syncthreads(); }
  __device__ int fetch_add(int* p, int v) const { return atomicAdd(p, v); }
};

__global__ void __launch_bounds__(1024)
write_results_nms_kernel(const float* __restrict__ pred, int R, int n_attr, float conf, float nms_thr, int max_det, int cap,
                         float* __restrict__ out_det, int32_t* __restrict__ out_row, int32_t* __restrict__ out_count,
                         int32_t* __restrict__ out_total) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(nms_smem);
  uint8_t* supp = nms_smem + (size_t)cap * 8;
  int* counter = reinterpret_cast<int*>(nms_smem + (size_t)cap * 9);
  const int b = blockIdx.x;
  bp_nms::nms_image(CudaBlock{}, pred + (long)b * R * n_attr, R, n_attr, conf, nms_thr, max_det, b, keys, supp, counter,
                    out_det + (long)b * max_det * 8, out_row + (long)b * max_det, out_count + b, out_total + b);
}

}  // namespace

extern "C" int bp_write_results_nms(bp_engine* e, const float* pred, int B, int R, int n_attr, float conf, float nms_thr, int max_det,
                                    float* out_det, int32_t* out_row, int32_t* out_count, int32_t* out_total, void* stream) {
  if (!e || !pred || !out_det || !out_row || !out_count || !out_total || B <= 0 || R <= 0 || n_attr < 6 || max_det <= 0)
    return bp_fail(BP_ERR_INVALID, "bp_write_results_nms: bad arguments");
  if (R > 16384) return bp_fail(BP_ERR_UNSUPPORTED, "bp_write_results_nms: more than 16384 rows per image (shared-memory sort)");
  int cap = 64;
  while (cap < R) cap <<= 1;
  const size_t smem = (size_t)cap * 9 + 16;
  if (smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(write_results_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return bp_fail(BP_ERR_CUDA, cudaGetErrorString(err));
  }
  const int threads = cap >= 4096 ? 1024 : 256;
  write_results_nms_kernel<<<B, threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(pred, R, n_attr, conf, nms_thr, max_det, cap,
                                                                                        out_det, out_row, out_count, out_total);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? BP_OK : bp_fail(BP_ERR_CUDA, cudaGetErrorString(err));
}
