// Host-side TMA descriptor (CUtensorMap) construction. The driver entry points are resolved at run time
// through cudaGetDriverEntryPoint so the library links without libcuda (the build box has no driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace bp {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
using EncodeIm2colFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmapApi {
  EncodeTiledFn tiled = nullptr;
  EncodeIm2colFn im2col = nullptr;
  int driver_version = 0;
  bool load(std::string* err) {
    if (tiled && im2col) return true;
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
      if (err) *err = "cuTensorMapEncodeTiled not available";
      return false;
    }
    tiled = reinterpret_cast<EncodeTiledFn>(f);
    f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
      if (err) *err = "cuTensorMapEncodeIm2col not available";
      return false;
    }
    im2col = reinterpret_cast<EncodeIm2colFn>(f);
    cudaDriverGetVersion(&driver_version);
    return true;
  }
};

inline CUtensorMapSwizzle swizzle_for(int block_k) {
  return block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
}

// Row-major fp16 matrix [rows, cols] with `pitch` elements between rows; box = [box_rows, block_k].
inline bool make_tmap_2d(TmapApi& api, CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                         uint64_t pitch, uint32_t box_rows, uint32_t block_k, std::string* err) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch * 2};
  cuuint32_t box[2] = {block_k, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = api.tiled(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(block_k), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed: " + std::to_string((int)r);
    return false;
  }
  return true;
}

// NHWC fp16 activation [N, H, W, C] with `pitch` elements per pixel, traversed as the im2col matrix of an
// R x S convolution with the given stride and padding (pad_h rows above/below, pad_w columns left/right).
// One load = `pixels` (128 or 256) output pixels x block_k channels.  row_pitch / img_pitch (elements) default to the dense layout;
// the packed stem convolutions pass a pixel pitch SMALLER than C (overlapping "virtual pixels", see net.cu) and the
// padded row pitch of the network-input buffer.
inline bool make_tmap_im2col(TmapApi& api, CUtensorMap* out, const void* base, int N, int H, int W, int C,
                             int pitch, int R, int S, int stride, int pad_h, int pad_w, uint32_t block_k, std::string* err,
                             long row_pitch = 0, long img_pitch = 0, uint32_t pixels = 128, int stride_w = 0) {
  if (stride_w <= 0) stride_w = stride;
  if (row_pitch <= 0) row_pitch = (long)W * pitch;
  if (img_pitch <= 0) img_pitch = (long)H * row_pitch;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)row_pitch * 2, (cuuint64_t)img_pitch * 2};
  int lower[2] = {-pad_w, -pad_h};
  int upper[2] = {pad_w - (S - 1), pad_h - (R - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride_w, (cuuint32_t)stride, 1};
  CUresult r = api.im2col(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, lower,
                          upper, block_k, pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(block_k),
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeIm2col failed: " + std::to_string((int)r);
    return false;
  }
  // Drivers up to 13.1 mis-encode im2col maps of tensors smaller than 128 KiB; clearing bit 21 of the second
  // descriptor word is the documented-by-practice fix (same guard as CUTLASS' make_im2col_tma_copy_desc).
  const uint64_t bytes = (uint64_t)N * img_pitch * 2;
  if (api.driver_version <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
  return true;
}

}  // namespace bp
