// Host build of the block-level IoU-NMS (betapose_b200/csrc/box_nms.cuh) for the CPU test-suite: the same source the
// CUDA kernel instantiates, run by a one-thread "block".  Test infrastructure, not part of libbetapose_b200.so.
#include <cstdlib>
#include <vector>

#include "../../betapose_b200/csrc/box_nms.cuh"

extern "C" int bp_box_nms_host(const float* pred, int B, int R, int n_attr, float conf, float nms_thr, int max_det, float* out_det,
                               int32_t* out_row, int32_t* out_count, int32_t* out_total) {
  int cap = 64;
  while (cap < R) cap <<= 1;
  std::vector<unsigned long long> keys(cap);
  std::vector<uint8_t> supp(cap);
  int counter = 0;
  for (int b = 0; b < B; ++b)
    bp_nms::nms_image(bp_nms::HostBlock{}, pred + (long)b * R * n_attr, R, n_attr, conf, nms_thr, max_det, b, keys.data(), supp.data(),
                      &counter, out_det + (long)b * max_det * 8, out_row + (long)b * max_det, out_count + b, out_total + b);
  return 0;
}

extern "C" float bp_box_iou_host(const float* a, const float* b) {
  return bp_nms::iou_plus1(bp_nms::Box{a[0], a[1], a[2], a[3]}, bp_nms::Box{b[0], b[1], b[2], b[3]});
}
