"""Tensor-level wrappers of the stage kernels (C ABI: bp_resize_bicubic, bp_yolo_decode_argmax, bp_crop_resize,
bp_heatmap_decode, bp_pose_pnp, bp_pack_records).  Inputs/outputs are CUDA torch tensors; every call enqueues on
the current torch stream and returns immediately.  These are what the drop-in seam functions in
`betapose_b200.compat` and the fused `BetaposeEngine` are built from.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

CAM_K = np.array([[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]], np.float64)
"""LineMod intrinsics hard-coded by the reference (betapose_evaluate.py:59)."""


def _eng(t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.BetaposeError("expected a CUDA tensor (betapose_b200 has no CPU path)")
    return _lib.Engine.get(t.device.index)


def net_input_buffer(n: int, h: int, w: int, device) -> torch.Tensor:
    """A zeroed network-input buffer fp16 [n, h, w + IN_PAD_COLS, 8] (layout of include/betapose_b200.h, BP_IN_*)."""
    return torch.zeros((n, h, w + _lib.IN_PAD_COLS, 8), dtype=torch.float16, device=device)


def net_input_pixels(buf: torch.Tensor) -> torch.Tensor:
    """The data pixels [n, h, w, 3] of a network-input buffer (a view)."""
    w = buf.shape[2] - _lib.IN_PAD_COLS
    return buf[:, :, _lib.IN_PAD_LEFT:_lib.IN_PAD_LEFT + w, :3]


def resize_bicubic(frames_u8: torch.Tensor, out_h: int = 416, out_w: int = 416, want_net: bool = True,
                   want_f32: bool = False, out_net: torch.Tensor | None = None):
    """frames uint8 [B,H,W,3] RGB (cuda) -> (network-input buffer fp16 [B,oh,ow+8,8] holding the raw 0..255 values | None,
    fp32 [B,3,oh,ow] | None); Pillow-exact."""
    assert frames_u8.dtype == torch.uint8 and frames_u8.dim() == 4 and frames_u8.shape[3] == 3 and frames_u8.is_contiguous()
    e = _eng(frames_u8)
    B, H, W, _ = frames_u8.shape
    if want_net and out_net is None:
        out_net = net_input_buffer(B, out_h, out_w, frames_u8.device)
    f32 = torch.empty((B, 3, out_h, out_w), dtype=torch.float32, device=frames_u8.device) if want_f32 else None
    _lib.check(_lib.lib().bp_resize_bicubic(e.handle, _lib.ptr(frames_u8), B, H, W, out_h, out_w,
                                            _lib.ptr(out_net if want_net else None), _lib.ptr(f32), _lib.stream_ptr()),
               "bp_resize_bicubic")
    return (out_net if want_net else None), f32


def yolo_decode_argmax(heads, anchors, B: int, reso: int = 416, conf: float = 0.01, frame_w: int = 640, frame_h: int = 480,
                       n_attr: int = 6, want_decoded: bool = False):
    """heads: list of fp32 NHWC tensors [>=B, g, g, C] (strided views allowed, channel stride 1);
    anchors: list (per head) of 3 (w,h) pairs in pixels.
    -> dict(det [B,8], box [B,4], score [B], row int32 [B], valid uint8 [B], decoded [B,R,n_attr] | None)"""
    h0 = heads[0]
    e = _eng(h0)
    nh = len(heads)
    ptrs = (C.c_void_p * nh)(*[h.data_ptr() for h in heads])
    grids = (C.c_int * nh)(*[int(h.shape[1]) for h in heads])
    pitches = (C.c_int * nh)(*[int(h.stride(2)) for h in heads])
    flat = (C.c_float * (nh * 6))(*[float(v) for hd in anchors for wh in hd for v in wh])
    dev = h0.device
    det = torch.empty((B, 8), dtype=torch.float32, device=dev)
    box = torch.empty((B, 4), dtype=torch.float32, device=dev)
    score = torch.empty((B,), dtype=torch.float32, device=dev)
    row = torch.empty((B,), dtype=torch.int32, device=dev)
    valid = torch.empty((B,), dtype=torch.uint8, device=dev)
    total = sum(3 * int(h.shape[1]) ** 2 for h in heads)
    dec = torch.empty((B, total, n_attr), dtype=torch.float32, device=dev) if want_decoded else None
    _lib.check(_lib.lib().bp_yolo_decode_argmax(e.handle, ptrs, grids, pitches, nh, flat, n_attr, B, reso, float(conf),
                                                frame_w, frame_h, _lib.ptr(det), _lib.ptr(box), _lib.ptr(score),
                                                _lib.ptr(row),
                                                _lib.ptr(valid), _lib.ptr(dec), _lib.stream_ptr()), "bp_yolo_decode_argmax")
    return dict(det=det, box=box, score=score, row=row, valid=valid, decoded=dec)


def write_results(pred: torch.Tensor, conf: float = 0.01):
    """pred fp32 [B,R,n_attr] decoded rows (cuda) -> dict(det [B,8], row int32 [B], valid uint8 [B])."""
    e = _eng(pred)
    pred = pred.contiguous()
    B, R, A = pred.shape
    det = torch.empty((B, 8), dtype=torch.float32, device=pred.device)
    row = torch.empty((B,), dtype=torch.int32, device=pred.device)
    valid = torch.empty((B,), dtype=torch.uint8, device=pred.device)
    _lib.check(_lib.lib().bp_write_results(e.handle, _lib.ptr(pred), B, R, A, float(conf), _lib.ptr(det), _lib.ptr(row),
                                           _lib.ptr(valid), _lib.stream_ptr()), "bp_write_results")
    return dict(det=det, row=row, valid=valid)


def write_results_nms(pred: torch.Tensor, conf: float = 0.01, nms_thr: float = 0.6, max_det: int = 100):
    """write_results with the IoU-NMS branch on (yolo/util.py:182-196, shipped disabled by `nms = False` at :181).
    pred fp32 [B,R,n_attr] decoded rows (cuda) -> dict(det [B,max_det,8], row int32 [B,max_det], count int32 [B] =
    min(kept, max_det), total int32 [B] = kept before the cap), detections best first."""
    e = _eng(pred)
    pred = pred.contiguous()
    B, R, A = pred.shape
    dev = pred.device
    det = torch.zeros((B, max_det, 8), dtype=torch.float32, device=dev)
    row = torch.full((B, max_det), -1, dtype=torch.int32, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    total = torch.empty((B,), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().bp_write_results_nms(e.handle, _lib.ptr(pred), B, R, A, float(conf), float(nms_thr), int(max_det), _lib.ptr(det),
                                               _lib.ptr(row), _lib.ptr(count), _lib.ptr(total), _lib.stream_ptr()), "bp_write_results_nms")
    return dict(det=det, row=row, count=count, total=total)


def crop_resize(frames_u8: torch.Tensor, box: torch.Tensor, img_idx: torch.Tensor, valid: torch.Tensor | None = None,
                res_h: int = 320, res_w: int = 256, out_net: torch.Tensor | None = None, want_f16: bool = True,
                want_f32: bool = False):
    """frames uint8 [F,H,W,3]; box fp32 [n,4]; img_idx int32 [n] -> dict(net = network-input buffer fp16 [n,rh,rw+8,8],
    f32 [n,3,rh,rw], pt1, pt2)."""
    e = _eng(frames_u8)
    n = int(box.shape[0])
    _, H, W, _ = frames_u8.shape
    dev = frames_u8.device
    if want_f16 and out_net is None:
        out_net = net_input_buffer(n, res_h, res_w, dev)
    f32 = torch.empty((n, 3, res_h, res_w), dtype=torch.float32, device=dev) if want_f32 else None
    pt1 = torch.empty((n, 2), dtype=torch.float32, device=dev)
    pt2 = torch.empty((n, 2), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().bp_crop_resize(e.handle, _lib.ptr(frames_u8), H, W, _lib.ptr(box.contiguous()),
                                         _lib.ptr(img_idx.contiguous()), _lib.ptr(valid), n, res_h, res_w,
                                         _lib.ptr(out_net if want_f16 else None), _lib.ptr(f32), _lib.ptr(pt1),
                                         _lib.ptr(pt2), _lib.stream_ptr()), "bp_crop_resize")
    return dict(net=out_net if want_f16 else None, f32=f32, pt1=pt1, pt2=pt2)


def heatmap_decode(hm: torch.Tensor, pt1: torch.Tensor, pt2: torch.Tensor, layout: str = "nchw", inp_h: int = 320,
                   inp_w: int = 256):
    """hm fp32: layout 'nchw' = [n,K,H,W] (the reference's), 'nhwc' = [n,H,W,K] (possibly channel-padded view).
    -> dict(preds_hm [n,K,2], preds_img [n,K,2], maxval [n,K,1], idx int32 [n,K])"""
    e = _eng(hm)
    assert hm.dtype == torch.float32 and hm.dim() == 4
    if layout == "nchw":
        n, K, H, W = hm.shape
        assert hm.stride(3) == 1 and hm.stride(2) == W
        strides = (hm.stride(0), hm.stride(1), 1)
    else:
        n, H, W, K = hm.shape
        assert hm.stride(3) == 1 and hm.stride(1) == W * hm.stride(2)
        strides = (hm.stride(0), 1, hm.stride(2))
    dev = hm.device
    ph = torch.empty((n, K, 2), dtype=torch.float32, device=dev)
    pi = torch.empty((n, K, 2), dtype=torch.float32, device=dev)
    mv = torch.empty((n, K, 1), dtype=torch.float32, device=dev)
    idx = torch.empty((n, K), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().bp_heatmap_decode(e.handle, _lib.ptr(hm), strides[0], strides[1], strides[2], n, K, H, W, inp_h,
                                            inp_w, _lib.ptr(pt1.contiguous()), _lib.ptr(pt2.contiguous()), _lib.ptr(ph),
                                            _lib.ptr(pi), _lib.ptr(mv), _lib.ptr(idx), _lib.stream_ptr()), "bp_heatmap_decode")
    return dict(preds_hm=ph, preds_img=pi, maxval=mv, idx=idx)


def pinhole4(cam_K) -> tuple[float, float, float, float]:
    """(fx, fy, cx, cy) of a 3x3 camera matrix.  The PnP and scoring kernels model a plain pinhole camera; a matrix with
    skew or a non-unit last row (which cv2.solvePnP / utils/metrics.py would honour) is refused instead of being
    silently mis-read."""
    K = np.asarray(cam_K, np.float64)
    if K.shape != (3, 3) or K[0, 1] != 0 or K[1, 0] != 0 or K[2, 0] != 0 or K[2, 1] != 0 or K[2, 2] != 1:
        raise _lib.BetaposeError(f"camera matrix must be [[fx,0,cx],[0,fy,cy],[0,0,1]] (no skew), got {K.tolist()}")
    return float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])


MODE_RANSAC, MODE_ALLPTS = 0, 1
PNP_RAW_POINTS, PNP_NMS_ONLY = 1, 2


def pose_pnp(preds_img: torch.Tensor, maxval: torch.Tensor, det_score: torch.Tensor, kp3d: torch.Tensor,
             valid: torch.Tensor | None = None, model_idx: torch.Tensor | None = None, cam_K=CAM_K, left_number: int = 50,
             mode: int = MODE_RANSAC, reproj_thr: float = 12.0, n_hyp: int = 64, seed: int = 0, flags: int = 0):
    """Single-proposal pose-NMS + key-point selection + PnP for n detections.
    kp3d float64 [K,3] or [n_models,K,3] (cuda).  Returns a dict of cuda tensors (see bp_pose_pnp)."""
    e = _eng(preds_img)
    n, K = int(preds_img.shape[0]), int(preds_img.shape[1])
    dev = preds_img.device
    assert kp3d.dtype == torch.float64 and kp3d.is_cuda
    cam_arr = (C.c_double * 4)(*pinhole4(cam_K))
    out = dict(
        keypoints=torch.empty((n, K, 2), dtype=torch.float32, device=dev),
        kp_score=torch.empty((n, K), dtype=torch.float32, device=dev),
        proposal=torch.empty((n,), dtype=torch.float32, device=dev),
        selected=torch.empty((n, K), dtype=torch.uint8, device=dev),
        R=torch.empty((n, 9), dtype=torch.float64, device=dev),
        t=torch.empty((n, 3), dtype=torch.float64, device=dev),
        inlier=torch.empty((n, K), dtype=torch.uint8, device=dev),
        status=torch.empty((n,), dtype=torch.int32, device=dev),
    )
    _lib.check(_lib.lib().bp_pose_pnp(e.handle, _lib.ptr(preds_img.contiguous()),
                                      _lib.ptr(None if maxval is None else maxval.contiguous()),
                                      _lib.ptr(None if det_score is None else det_score.contiguous()), _lib.ptr(valid), n, K, _lib.ptr(kp3d.contiguous()),
                                      _lib.ptr(model_idx), C.cast(cam_arr, C.c_void_p), int(left_number), int(mode), int(flags),
                                      float(reproj_thr), int(n_hyp), int(seed) & 0xFFFFFFFF, _lib.ptr(out["keypoints"]),
                                      _lib.ptr(out["kp_score"]), _lib.ptr(out["proposal"]), _lib.ptr(out["selected"]),
                                      _lib.ptr(out["R"]), _lib.ptr(out["t"]), _lib.ptr(out["inlier"]),
                                      _lib.ptr(out["status"]), _lib.stream_ptr()), "bp_pose_pnp")
    return out


def pack_records(image_index0: int, box, det_score, pose: dict) -> torch.Tensor:
    """-> uint8 [n, RECORD_BYTES] cuda tensor of bp_record structs."""
    e = _eng(box)
    n, K = pose["kp_score"].shape
    out = torch.empty((n, _lib.RECORD_BYTES), dtype=torch.uint8, device=box.device)
    _lib.check(_lib.lib().bp_pack_records(e.handle, n, K, int(image_index0), _lib.ptr(box.contiguous()),
                                          _lib.ptr(det_score.contiguous()), _lib.ptr(pose["keypoints"]),
                                          _lib.ptr(pose["kp_score"]), _lib.ptr(pose["proposal"]), _lib.ptr(pose["R"]),
                                          _lib.ptr(pose["t"]), _lib.ptr(pose["status"]), _lib.ptr(out), _lib.stream_ptr()),
               "bp_pack_records")
    return out


RECORD_DTYPE = np.dtype([("image_index", "<i4"), ("status", "<i4"), ("box", "<f4", 4), ("det_score", "<f4"),
                         ("proposal_score", "<f4"), ("keypoints", "<f4", 150), ("R", "<f8", 9), ("t", "<f8", 3)], align=True)
assert RECORD_DTYPE.itemsize == _lib.RECORD_BYTES, (RECORD_DTYPE.itemsize, _lib.RECORD_BYTES)


def records_to_numpy(rec_u8: torch.Tensor) -> np.ndarray:
    """host copy of packed records as a numpy structured array."""
    return rec_u8.detach().cpu().numpy().view(RECORD_DTYPE).reshape(-1)


def score_poses(R_est, t_est, box_est, R_gt, t_gt, box_gt, model, cam_K=CAM_K, status=None, model_idx=None):
    """Scoring stage (betapose_evaluate.py:203-266; utils/metrics.py add_err / projection_error_2d / iou), batched.
    R_* f64 [n,3,3] or [n,9], t_* f64 [n,3], box_* fp32 [n,4] corners (x1,y1,x2,y2), model f64 [V,3] or [n_models,V,3]
    (metres), all CUDA tensors.  -> dict(add [n] f64 metres, proj [n] f64 px, iou [n] f32, scored [n] u8)."""
    dev = R_est.device
    e = _eng(R_est)
    n = int(R_est.shape[0])
    f64 = lambda x: x.to(dev, torch.float64).reshape(n, -1).contiguous()  # noqa: E731
    Re, te, Rg, tg = f64(R_est), f64(t_est), f64(R_gt), f64(t_gt)
    be, bg = box_est.to(dev, torch.float32).contiguous(), box_gt.to(dev, torch.float32).contiguous()
    m = model.to(dev, torch.float64).contiguous()
    V = int(m.shape[-2])
    cam = (C.c_double * 4)(*pinhole4(cam_K))
    add = torch.empty(n, dtype=torch.float64, device=dev)
    proj = torch.empty(n, dtype=torch.float64, device=dev)
    iou = torch.empty(n, dtype=torch.float32, device=dev)
    scored = torch.empty(n, dtype=torch.uint8, device=dev)
    st = status.to(dev, torch.int32).contiguous() if status is not None else None
    mi = model_idx.to(dev, torch.int32).contiguous() if model_idx is not None else None
    _lib.check(_lib.lib().bp_score_poses(e.handle, n, _lib.ptr(Re), _lib.ptr(te), _lib.ptr(st), _lib.ptr(be), _lib.ptr(Rg),
                                         _lib.ptr(tg), _lib.ptr(bg), _lib.ptr(m), _lib.ptr(mi), V, C.cast(cam, C.c_void_p),
                                         _lib.ptr(add), _lib.ptr(proj), _lib.ptr(iou), _lib.ptr(scored), _lib.stream_ptr()),
               "bp_score_poses")
    return dict(add=add, proj=proj, iou=iou, scored=scored)


def summarize_scores(add_m, proj_px, iou, scored, diameter_mm: float, pixel_thresh: float = 5.0) -> dict:
    """The reference's summary numbers (betapose_evaluate.py:259-266): ADD accuracy at diameter / 10 (errors in mm),
    2-D reprojection accuracy at 5 px, fraction of frames with IoU > 0.5.  Host arrays / tensors in."""
    add_mm = np.asarray(torch.as_tensor(add_m).cpu(), np.float64) * 1000.0
    proj = np.asarray(torch.as_tensor(proj_px).cpu(), np.float64)
    io = np.asarray(torch.as_tensor(iou).cpu(), np.float64)
    sc = np.asarray(torch.as_tensor(scored).cpu()).astype(bool)
    return dict(mean_add_err_mm=float(add_mm[sc].mean()) if sc.any() else float("nan"),
                add_accuracy=float((add_mm[sc] < diameter_mm / 10.0).mean()) if sc.any() else float("nan"),
                proj2d_accuracy=float((proj[sc] < pixel_thresh).mean()) if sc.any() else float("nan"),
                iou_accuracy=float((io > 0.5).mean()) if len(io) else float("nan"), n_scored=int(sc.sum()))


def pose_nms(bboxes, bbox_scores, pose_preds, pose_scores, counts=None):
    """General parametric pose-NMS (pPose_nms.py:24-122) for the proposals of one image (counts=None) or of several
    images concatenated (counts: list of proposals per image).  CUDA fp32 tensors: bboxes [N,4], bbox_scores [N],
    pose_preds [N,K,2], pose_scores [N,K].  -> dict(count int32 [n_images], pick int32 [N], keypoints [N,K,2],
    kp_score [N,K], proposal [N]); image i's results are rows first[i] .. first[i] + count[i] - 1."""
    e = _eng(pose_preds)
    dev = pose_preds.device
    N, K = int(pose_preds.shape[0]), int(pose_preds.shape[1])
    cl = [N] if counts is None else [int(c) for c in counts]
    assert sum(cl) == N
    first = torch.tensor(np.concatenate([[0], np.cumsum(cl)[:-1]]).astype(np.int32), device=dev)
    cnt = torch.tensor(cl, dtype=torch.int32, device=dev)
    out = dict(count=torch.zeros(len(cl), dtype=torch.int32, device=dev), pick=torch.zeros(N, dtype=torch.int32, device=dev),
               keypoints=torch.zeros((N, K, 2), dtype=torch.float32, device=dev),
               kp_score=torch.zeros((N, K), dtype=torch.float32, device=dev), proposal=torch.zeros(N, dtype=torch.float32, device=dev),
               first=first)
    f32 = lambda x, shape: x.to(dev, torch.float32).reshape(shape).contiguous()  # noqa: E731
    _lib.check(_lib.lib().bp_pose_nms(e.handle, len(cl), _lib.ptr(first), _lib.ptr(cnt), max(cl), K, _lib.ptr(f32(bboxes, (N, 4))),
                                      _lib.ptr(f32(bbox_scores, (N,))), _lib.ptr(f32(pose_preds, (N, K, 2))),
                                      _lib.ptr(f32(pose_scores, (N, K))), _lib.ptr(out["count"]), _lib.ptr(out["pick"]),
                                      _lib.ptr(out["keypoints"]), _lib.ptr(out["kp_score"]), _lib.ptr(out["proposal"]),
                                      _lib.stream_ptr()), "bp_pose_nms")
    return out
