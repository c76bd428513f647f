/*
 * betapose_b200 -- C ABI of the B200-native engine for Betapose's per-frame evaluate hot path.
 *
 * The reference (sjtuytc/betapose) has no plugin / FFI boundary on this path: its seams are Python call
 * signatures.  This header is the boundary a maintainer binds instead (ctypes stub in
 * INTEGRATION.md); each entry point names the reference code it replaces.  Conventions:
 *   - plain pointers and sizes only; device pointers are raw CUDA addresses owned by the CALLER unless stated;
 *   - every call enqueues on the given cudaStream_t (passed as void*) and never synchronises, except
 *     bp_net_conv (load-time: host->device weight upload), bp_*_create and the first use of a new batch size or
 *     scratch size (descriptor encoding / a one-off cudaMalloc: outside steady state and before CUDA-graph capture);
 *   - int return: 0 = OK, negative = error, message via bp_last_error() (thread-local);
 *   - a handle is bound to one device and is not thread-safe.
 */
#ifndef BETAPOSE_B200_H
#define BETAPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BP_OK 0
#define BP_ERR_INVALID (-1)
#define BP_ERR_CUDA (-2)
#define BP_ERR_UNSUPPORTED (-3)
#define BP_ERR_IO (-4)

/* activation / residual / store selectors of a fused conv block */
#define BP_ACT_NONE 0
#define BP_ACT_LEAKY 1   /* LeakyReLU(0.1): yolo/darknet.py:259 */
#define BP_ACT_RELU 2    /* SE_Resnet.py:28-30,40 ; DUC.py:21 */
#define BP_ACT_SIGMOID 3 /* SE_module.py:12 */
#define BP_RES_NONE 0
#define BP_RES_AFTER_ACT 1  /* darknet [shortcut]: act(conv) + skip, yolo/darknet.py:338-340 */
#define BP_RES_BEFORE_ACT 2 /* ResNet bottleneck: relu(conv + skip), SE_Resnet.py:38-40 */
#define BP_STORE_PLAIN 0
#define BP_STORE_UPSAMPLE2 1 /* nn.Upsample(2, nearest) fused into the producer, darknet.py:273-276 */
#define BP_STORE_PIXSHUF2 2  /* nn.PixelShuffle(2) fused into the producer, DUC.py:16,22 */

typedef struct bp_engine bp_engine;
typedef struct bp_net bp_net;

const char* bp_last_error(void);
int bp_version(void);

/* ---------------------------------------------------------------------------------------------- engine */
/* One per device.  Owns the TMA driver entry points, the PIL coefficient tables and scratch space. */
int bp_engine_create(int device, bp_engine** out);
void bp_engine_destroy(bp_engine* e);

/* ---------------------------------------------------------------------------------------------- networks
 * A bp_net is an ordered list of fused layer ops over NHWC fp16 tensors, built once (weights uploaded, BN
 * folded, TMA descriptors encoded for max_batch) and replayed per batch.  It replaces
 *   Darknet.build_model / load_weights / forward   (3_6Dpose_estimator/yolo/darknet.py:223-432) and
 *   FastPose / SEResnet / Bottleneck / SELayer / DUC (KPD/src/models/FastPose.py:13-35, layers/{SE_Resnet,SE_module,DUC}.py).
 * Tensors are named by small integer ids returned by the builder calls.
 */
/* The network input is always fp16 [N, H, W + BP_IN_PAD_COLS, 8]: pixel (h, w) at column BP_IN_PAD_LEFT + w, channels
 * 0..2 = R,G,B, channels 3..7 and the pad columns zero (zeroed once at creation; bp_resize_bicubic / bp_crop_resize
 * write only the data pixels).  This is what lets the first convolution fetch a whole filter row per TMA row. */
#define BP_IN_RAW255 0 /* pixel values 0..255; ToTensor's 1/255 is folded into the first conv's weights */
#define BP_IN_F16 1    /* values used as they are (the key-point net's mean-subtracted crop)            */
#define BP_IN_PAD_LEFT 3
#define BP_IN_PAD_COLS 8

int bp_net_create(bp_engine* e, int max_batch, int in_h, int in_w, int in_kind, bp_net* share_buffers_with,
                  bp_net** out);
void bp_net_destroy(bp_net* n);
/* id of the network input tensor (always 0) and its device address (caller writes the input there) */
void* bp_net_input_ptr(bp_net* n);

typedef struct bp_conv_spec {
  int src;      /* input tensor id */
  int cout;     /* output channels */
  int ksize;    /* square kernel: 1, 3 or 7 */
  int stride;   /* 1 or 2 */
  int pad;      /* symmetric zero padding */
  int act;      /* BP_ACT_* */
  int res;      /* residual tensor id, or -1 */
  int res_mode; /* BP_RES_* */
  int dst;      /* tensor id to write into (e.g. a concat buffer), or -1 to allocate a fresh tensor */
  int dst_coff; /* channel offset inside dst */
  int store_mode; /* BP_STORE_* */
  int out_f32;  /* 1: fp32 output (network heads) */
  /* host fp32 parameters, PyTorch layouts; bn_* may be NULL (then `bias` is the conv bias or NULL) */
  const float* weight; /* [cout, cin, k, k] */
  const float* bias;   /* [cout] */
  const float* bn_gamma;
  const float* bn_beta;
  const float* bn_mean;
  const float* bn_var;
  float bn_eps;
  /* optional: this convolution's weights already folded and packed (bp_pack_conv_weights, e.g. read from a packed-weight
   * cache).  When packed_w != NULL it is uploaded as it is and weight / bias / bn_* are ignored (may be NULL). */
  const void* packed_w;  /* fp16 [packed_w_elems] */
  const float* packed_b; /* fp32 [packed_b_elems] */
  size_t packed_w_elems;
  size_t packed_b_elems;
} bp_conv_spec;

/* conv (+folded BN) + bias + activation (+ residual) (+ fused upsample / pixel-shuffle store). Returns the
 * output tensor id (>= 0) or a negative error. */
int bp_net_conv(bp_net* n, const bp_conv_spec* spec);
/* SURVEY 8(f) item 4 -- the packed form bp_net_conv uploads, computed on the HOST (no CUDA call, no bp_net): BN folded in
 * fp64, fp16 rows [cout_pad][K] in the kernel's K order (NHWC im2col; for a stem the "virtual pixel" row order), fp32 bias
 * [cout_pad], PixelShuffle row permutation applied.  cin = input channels; in_kind = BP_IN_RAW255 / BP_IN_F16 when the
 * convolution reads the network input (stem), -1 otherwise.  Only cout, ksize, store_mode and the parameter pointers of
 * `spec` are read.  *w_elems / *b_elems receive the sizes; with w_out == b_out == NULL the call is a size query. */
int bp_pack_conv_weights(const bp_conv_spec* spec, int cin, int in_kind, void* w_out, float* b_out, size_t* w_elems,
                         size_t* b_elems);
/* pre-allocate a tensor several producers write into (darknet [route] with two layers = channel concat) */
int bp_net_alloc_tensor(bp_net* n, int h, int w, int c);
/* channel window [coff, coff+c) of an existing tensor as a tensor of its own (no copy) */
int bp_net_view(bp_net* n, int tensor, int coff, int c);
/* MaxPool2d(3, 2, 1): SE_Resnet.py:59 */
int bp_net_maxpool3x3s2(bp_net* n, int src);
/* AdaptiveAvgPool2d(1) -> [N,1,1,C] fp16: SE_module.py:7,16 */
int bp_net_global_avgpool(bp_net* n, int src);
/* relu(y * s + skip), s = [N,1,1,C] channel gates: SE_module.py:19 + SE_Resnet.py:38-40 */
int bp_net_scale_add_relu(bp_net* n, int y, int gates, int skip);
/* stand-alone PixelShuffle(2): FastPose.py:21,30 */
int bp_net_pixel_shuffle2(bp_net* n, int src);
/* generic fall-backs for cfg files whose route/upsample pattern cannot be fused */
int bp_net_upsample2(bp_net* n, int src, int dst, int dst_coff);
int bp_net_copy_channels(bp_net* n, int src, int dst, int dst_coff);
/* element-wise sum (darknet [shortcut] whose producer could not absorb it) */
int bp_net_add(bp_net* n, int a, int b);

/* query a tensor: dims[0..7] = {H, W, C, pitch (elements per pixel), is_f32, coff, pixels per buffer row, first data
 * column}; *ptr = device address of the buffer */
int bp_net_tensor_info(bp_net* n, int tensor, int* dims, void** ptr);
int bp_net_num_launches(bp_net* n);
double bp_net_flops_per_image(bp_net* n);
/* run all ops for `batch` images on `stream`; input must already be at bp_net_input_ptr() */
int bp_net_forward(bp_net* n, int batch, void* stream);
/* debugging / profiling: run only ops [first, last) */
int bp_net_forward_range(bp_net* n, int batch, int first, int last, void* stream);
int bp_net_num_ops(bp_net* n);
/* Measured tile plans (betapose_b200/tune.py): force the tile configuration of convolution `op` at batch size `batch` --
 * BLOCK_N (32..256), cg (1 = single CTA, 2 = CTA pairs), mt (1 / 2 / 4 = 128- / 256- / 512-pixel tiles); 0 = leave that
 * choice to the planner, all three 0 = remove the override.  BP_ERR_UNSUPPORTED if the layer cannot run that way. */
int bp_net_set_op_config(bp_net* n, int op, int batch, int block_n, int cg, int mt);
/* cfg[0..4] = {BLOCK_N, cg, mt, BLOCK_K, stages} of the plan convolution `op` runs at `batch` (zeros for aux ops) */
int bp_net_op_config(bp_net* n, int op, int batch, int* cfg);
/* Mixed-object batches (BASELINE.json configs[3]; the reference loads one detector + key-point net per object id,
 * KPD/src/main_fast_inference.py:29-36): this net runs CONCURRENTLY with other objects' nets whose batches sum to
 * `share_batch` images.  Its grids are then sized for num_sms * batch / share_batch SMs instead of the whole GPU (the
 * objects split the machine in proportion to their frame counts) and its kernels are launched without programmatic
 * stream serialisation (early-launched dependents would hold SMs another object could use).  0 = the net owns the GPU. */
int bp_net_set_share(bp_net* n, int share_batch);
/* per-op description for profiling tables: writes a short text into buf */
int bp_net_op_desc(bp_net* n, int op, char* buf, int buflen, double* flops_per_image, double* bytes_per_image);

/* ---------------------------------------------------------------------------------------------- stages */
/* a1: transforms.Resize((oh,ow), BICUBIC) + ToTensor  (dataloader.py:94-99,162) -- Pillow's two integer
 * passes, bit-exact.  frames: uint8 [B,H,W,3] RGB.  out_net: a BP_IN_RAW255 network input buffer
 * (fp16 [B,oh,ow+BP_IN_PAD_COLS,8], raw 0..255 values) and/or out_f32_chw: fp32 [B,3,oh,ow] in 0..1 (what the
 * reference feeds Darknet); either may be NULL. */
int bp_resize_bicubic(bp_engine* e, const uint8_t* frames, int B, int H, int W, int oh, int ow, void* out_net,
                      float* out_f32_chw, void* stream);

/* a3+a4+a5: DetectionLayer.forward (yolo/darknet.py:129-169) + write_results (yolo/util.py:118-223, nms off,
 * arg-max objectness) + box rescale (dataloader.py:350-364), fused; never materialises the 10647x6 tensor
 * unless `decoded` != NULL.  heads[i]: fp32 NHWC [B,g_i,g_i,pitch_i] raw head i (stride 32,16,8 order),
 * channel c = a*(5+classes)+attr.  anchors: 2*3 floats per head (pixels).
 * Outputs per image: det[B,8] = (img, x1,y1,x2,y2 in reso-space, obj, cls_conf, cls_idx), box[B,4] rescaled to
 * the frame, score[B] = objectness of the winner (what DetectionLoader hands on as `scores`, dataloader.py:352),
 * row[B] winning flat row (anchor-major, lowest index on ties), valid[B] (0: no candidate > conf).
 * score and decoded may be NULL. */
int bp_yolo_decode_argmax(bp_engine* e, const float* const* heads, const int* grids, const int* pitches,
                          int n_heads, const float* anchors, int n_attr, int B, int reso, float conf, int frame_w,
                          int frame_h, float* det, float* box, float* score, int32_t* row, uint8_t* valid,
                          float* decoded, void* stream);

/* a4 alone (the `dynamic_write_results(prediction, confidence, ...)` seam, yolo/util.py:104-223, nms off): arg-max
 * objectness per image over an already decoded prediction tensor pred[B,R,n_attr].  det[B,8] as above; valid[B]. */
int bp_write_results(bp_engine* e, const float* pred, int B, int R, int n_attr, float conf, float* det, int32_t* row,
                     uint8_t* valid, void* stream);

/* a4 with the IoU-NMS branch the reference ships switched off (yolo/util.py:182-196 behind `nms = False` at :181; bbox_iou,
 * yolo/bbox.py:51-77) -- SURVEY 8(f) item 3, scenes with several instances.  pred[B,R,n_attr] decoded rows as above
 * (R <= 16384).  Per image: candidates (objectness > conf, class arg-max 0) sorted by objectness, descending (ties: lower
 * row first); keep the best remaining box, drop every later one whose IoU with it ("+1 pixel" convention, fp32) is not
 * < nms_thr; repeat.  out_det[B,max_det,8] rows as bp_write_results, best first; out_row[B,max_det]; out_count[B] =
 * min(kept, max_det); out_total[B] = kept before the cap (what dynamic_write_results' "> 100 detections" retry reads). */
int bp_write_results_nms(bp_engine* e, const float* pred, int B, int R, int n_attr, float conf, float nms_thr, int max_det,
                         float* out_det, int32_t* out_row, int32_t* out_count, int32_t* out_total, void* stream);

/* a6: im_to_torch + crop_from_dets + cropBox (KPD/src/utils/img.py:13-18,242-262; dataloader.py:794-835).
 * frames uint8 [F,H,W,3] RGB; box[n,4]; img_idx[n] (frame of each box); valid[n] (may be NULL).
 * Outputs: out_net = a BP_IN_F16 network input buffer (fp16 [n,rh,rw+BP_IN_PAD_COLS,8]) and/or out_f32_chw fp32
 * [n,3,rh,rw]; pt1/pt2 [n,2] fp32 un-truncated expanded corners. */
int bp_crop_resize(bp_engine* e, const uint8_t* frames, int H, int W, const float* box, const int32_t* img_idx,
                   const uint8_t* valid, int n, int rh, int rw, void* out_net, float* out_f32_chw, float* pt1,
                   float* pt2, void* stream);

/* a8: getPrediction + transformBoxInvert_batch (KPD/src/utils/eval.py:113-147; img.py:216-239).
 * hm: fp32 heat-maps addressed as hm[img*img_stride + k*k_stride + pos*pos_stride], pos = y*res_w + x.
 * Outputs: preds_hm[n,K,2], preds_img[n,K,2], maxval[n,K], idx[n,K] (flat arg-max, lowest index on ties). */
int bp_heatmap_decode(bp_engine* e, const float* hm, long img_stride, long k_stride, long pos_stride, int n, int K,
                      int res_h, int res_w, int inp_h, int inp_w, const float* pt1, const float* pt2,
                      float* preds_hm, float* preds_img, float* maxval, int32_t* idx, void* stream);

/* a9 (n = 1 branch) + a10 + a11: pose_nms for a single proposal (pPose_nms.py:24-122), keypoint selection
 * (dataloader.py:715-724) and PnP (utils/utils.py:17-41).  mode 0: RANSAC-EPnP (5-point hypotheses, 12 px
 * consensus, LM refit on the consensus set) mirroring cv2.solvePnPRansac; mode 1: EPnP on all selected points
 * + LM, mirroring the active cv2.solvePnP call.
 * In: preds_img[n,K,2], maxval[n,K], det_score[n], valid[n] (NULL = all), kp3d f64 [n_models,K,3] + model_idx[n]
 * (NULL = model 0), cam f64[4] = (fx,fy,cx,cy).
 * Out: keypoints[n,K,2] (= preds - 0.3), kp_score[n,K], proposal[n], selected[n,K] (u8 mask of the points
 * handed to PnP), R f64[n,9] row-major, t f64[n,3], inlier[n,K] u8, status[n]: 1 = pose, 0 = rejected by
 * pose-NMS (max score < 0.3) or invalid, -1 = PnP failed. */
#define BP_PNP_RAW_POINTS 1 /* preds_img are final key-points (the bare `pnp(points_3D, points_2D, K)` seam): no score
                               floor / threshold, no -0.3 shift; maxval and det_score may be NULL */
#define BP_PNP_NMS_ONLY 2   /* stop after pose-NMS + selection (the bare `pose_nms` seam); R, t are left zero */
int bp_pose_pnp(bp_engine* e, const float* preds_img, const float* maxval, const float* det_score,
                const uint8_t* valid, int n, int K, const double* kp3d, const int32_t* model_idx, const double* cam,
                int left_number, int mode, int flags, float reproj_thr, int n_hyp, uint32_t seed, float* keypoints,
                float* kp_score, float* proposal, uint8_t* selected, double* R, double* t, uint8_t* inlier,
                int32_t* status, void* stream);

/* a12: fixed-size result record per image, the unit the multi-GPU all-gather moves. */
typedef struct bp_record {
  int32_t image_index; /* global image index */
  int32_t status;      /* see bp_pose_pnp */
  float box[4];
  float det_score;
  float proposal_score;
  float keypoints[50 * 3]; /* x, y, score */
  double R[9];
  double t[3];
} bp_record;
int bp_pack_records(bp_engine* e, int n, int K, int image_index0, const float* box, const float* det_score,
                    const float* keypoints, const float* kp_score, const float* proposal, const double* R,
                    const double* t, const int32_t* status, bp_record* out, void* stream);

/* a9 in general: parametric pose-NMS over n >= 1 proposals per image (pPose_nms.py:24-122 with
 * p_merge_fast :204-240, get_parametric_distance :243-267, PCK_match :270-281).  Proposals of all images are
 * concatenated: image i owns rows [first[i], first[i] + count[i]) (first == NULL: a single image starting at row 0);
 * count[i] <= max_count <= 64, K <= 64.  bboxes [N,4] corners, bbox_scores [N], pose_preds [N,K,2], pose_scores [N,K].
 * Outputs, written at the image's first rows in pick order: out_count[i] surviving poses, out_pick (proposal index
 * inside the image), out_keypoints [N,K,2] (merged pose - 0.3), out_kp_score [N,K] (merged scores), out_proposal [N]
 * (mean + bbox score of the pick + 1.25 max).  As in the reference every result carries the image's FIRST box. */
int bp_pose_nms(bp_engine* e, int n_images, const int32_t* first, const int32_t* count, int max_count, int K, const float* bboxes,
                const float* bbox_scores, const float* pose_preds, const float* pose_scores, int32_t* out_count, int32_t* out_pick,
                float* out_keypoints, float* out_kp_score, float* out_proposal, void* stream);

/* scoring (the evaluation loop after the per-frame path, betapose_evaluate.py:203-266): per image
 *   add_err[n]  = mean_v |(R_gt v + t_gt) - (R_est v + t_est)|        (utils/metrics.py:10-22, model units: metres)
 *   proj_err[n] = mean_v |proj(K [R_gt|t_gt] v) - proj(K [R_est|t_est] v)| in pixels (metrics.py:96-127)
 *   iou[n]      = IoU of box_gt and box_est, both (x1, y1, x2, y2)       (metrics.py:77-93)
 *   scored[n]   = 1 where the reference would score ADD / reprojection: the frame has a pose (status == 1; NULL = all)
 *                 and iou >= 0.5.
 * model: f64 [n_models, n_vertices, 3]; model_idx[n] (NULL = model 0); cam f64[4] = (fx, fy, cx, cy). */
int bp_score_poses(bp_engine* e, int n, const double* R_est, const double* t_est, const int32_t* status,
                   const float* box_est, const double* R_gt, const double* t_gt, const float* box_gt, const double* model,
                   const int32_t* model_idx, int n_vertices, const double* cam, double* add_err, double* proj_err, float* iou,
                   uint8_t* scored, void* stream);

/* ---------------------------------------------------------------------------------------------- frame ingest
 * what stands before a1 -- cv2.imread / PIL.Image.open per frame, twice, on one Python thread
 * (ImageLoader.getitem_yolo, dataloader.py:150-179; prep_image, yolo/preprocess.py:34-46).  Host-only (no CUDA call, no
 * bp_engine): a pool of threads decodes PNG files straight into caller memory, normally the pinned buffers the engine
 * uploads from, so decoding of batch i+1.. overlaps the GPU work of batch i.  The result is the 8-bit 3-channel image
 * both reference decoders produce: alpha dropped, grey replicated, palette expanded, 16-bit samples reduced to the high
 * byte.  Adam7-interlaced PNGs and formats other than PNG / PPM / PGM / .npy return BP_ERR_UNSUPPORTED (the caller decodes
 * those elsewhere). */
#define BP_ORDER_RGB 0 /* PIL.Image.open: the detector branch and this engine's frame layout */
#define BP_ORDER_BGR 1 /* cv2.imread: the reference's orig_img */
typedef struct bp_ingest bp_ingest;
/* n_threads <= 0: one per hardware thread */
int bp_ingest_create(int n_threads, bp_ingest** out);
/* waits for queued files (their output buffers must stay valid until this returns) */
void bp_ingest_destroy(bp_ingest* g);
int bp_ingest_num_threads(bp_ingest* g);
/* header fields of an in-memory PNG; channels counts the stored samples per pixel (palette = 3) */
int bp_png_info(const uint8_t* png, size_t len, int* H, int* W, int* channels, int* depth);
/* decode one in-memory PNG on the calling thread into out[H][row_pitch] (row_pitch 0 = 3*W); the file must be HxW */
int bp_png_decode(const uint8_t* png, size_t len, int H, int W, int order, uint8_t* out, size_t row_pitch);
/* what the pool runs per file, on the calling thread: PNG as above, or a container without entropy coding -- binary
 * PPM / PGM ("P6" / "P5", maxval 255) or a NumPy .npy file holding C-ordered uint8 [H,W,3] (or [H,W]) -- told apart by
 * their magic bytes.  A sequence converted once to one of those is delivered at memory speed. */
int bp_frame_decode(const uint8_t* data, size_t len, int H, int W, int order, uint8_t* out, size_t row_pitch);
/* the ingest's inflate on its own: one whole zlib stream (RFC 1950) -> out[0..out_cap); *out_len = bytes produced.
 * BP_ERR_INVALID for corrupt / truncated streams and for streams that hold more than out_cap bytes. */
int bp_zlib_inflate(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len);
/* queue n files; file i is decoded into out + i*frame_pitch (frame_pitch 0 = H*W*3) and status[i] (may be NULL) gets
 * its return code.  Returns a ticket > 0, or a negative error.  May be called again before earlier tickets finished. */
int64_t bp_ingest_submit(bp_ingest* g, const char* const* paths, int n, int H, int W, int order, uint8_t* out, size_t frame_pitch,
                         int32_t* status);
/* blocks until every file of the ticket is done; 0, or the first failing file's code with its path in bp_last_error() */
int bp_ingest_wait(bp_ingest* g, int64_t ticket);

#ifdef __cplusplus
}
#endif
#endif /* BETAPOSE_B200_H */
