"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, fp32/fp64 as the reference computes) of every
non-network stage of Betapose's per-frame evaluate path.  Imported only by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; the product (betapose_b200/) never imports it.

Each function cites the reference lines it restates (paths relative to
/root/reference/3_6Dpose_estimator/).  Pinned against the reference itself by tests/test_oracle_pinned.py
(runs where /root/reference exists) and against tests/golden/*.npz everywhere else.

Third-party arithmetic the reference delegates to un-vendored libraries is restated from the published
algorithm and pinned to this image's versions: Pillow 12.2.0 (bicubic resize), torch 2.11.0
(bilinear align_corners=True), OpenCV 4.13.0 (solvePnPRansac / solvePnP -> see oracle/pnp.py).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32

# ----------------------------------------------------------------------------------------------------
# a1  ImageLoader.getitem_yolo: transforms.Resize((416,416), interpolation=3) + ToTensor
#     dataloader.py:94-99,162  -> Pillow ImagingResample (two integer passes, 22-bit coefficients)
# ----------------------------------------------------------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def _bicubic_filter(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_resample_coeffs(in_size: int, out_size: int):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter (support 2.0).

    Returns (bounds int32 [out,2] = (xmin, count), coeffs int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            if v < 0:
                kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS))
            else:
                kk[xx, x] = int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def pil_resize_bicubic(img_u8: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """uint8 [H,W,C] -> uint8 [out_h,out_w,C]; horizontal pass (uint8 intermediate) then vertical."""
    H, W, C = img_u8.shape
    src = img_u8.astype(np.int64)
    if W != out_w:
        bx, kx = pil_resample_coeffs(W, out_w)
        tmp = np.empty((H, out_w, C), np.int64)
        for xx in range(out_w):
            x0, n = bx[xx]
            acc = (src[:, x0:x0 + n, :] * kx[xx, :n].astype(np.int64)[None, :, None]).sum(axis=1)
            tmp[:, xx, :] = np.clip((acc + (1 << (PRECISION_BITS - 1))) >> PRECISION_BITS, 0, 255)
        src = tmp
    if H != out_h:
        by, ky = pil_resample_coeffs(H, out_h)
        out = np.empty((out_h, src.shape[1], C), np.int64)
        for yy in range(out_h):
            y0, n = by[yy]
            acc = (src[y0:y0 + n, :, :] * ky[yy, :n].astype(np.int64)[:, None, None]).sum(axis=0)
            out[yy] = np.clip((acc + (1 << (PRECISION_BITS - 1))) >> PRECISION_BITS, 0, 255)
        src = out
    return src.astype(np.uint8)


def yolo_input_from_frame(frame_rgb_u8: np.ndarray, reso: int = 416) -> np.ndarray:
    """-> fp32 [3,reso,reso] RGB in 0..1 (ToTensor divides by 255 in fp32)."""
    r = pil_resize_bicubic(frame_rgb_u8, reso, reso)
    return (r.astype(F32) / F32(255)).transpose(2, 0, 1).copy()


# ----------------------------------------------------------------------------------------------------
# a3  DetectionLayer.forward  yolo/darknet.py:129-169  (anchor-major flatten, heads 32|16|8)
# ----------------------------------------------------------------------------------------------------
YOLO_ANCHORS = {  # yolo/cfg/yolov3-single.cfg:584-585,670-671,757-758 (mask 6,7,8 / 3,4,5 / 0,1,2)
    32: ((116, 90), (156, 198), (373, 326)),
    16: ((30, 61), (62, 45), (59, 119)),
    8: ((10, 13), (16, 30), (33, 23)),
}


def _sigmoid(x):
    x = x.astype(F32)
    return (F32(1) / (F32(1) + np.exp(-x))).astype(F32)


def yolo_decode_head(x: np.ndarray, reso: int = 416, anchors=None) -> np.ndarray:
    """x: [B, 18, g, g] fp32 raw head -> [B, 3*g*g, 6] (cx, cy, w, h, obj, cls0), pixels of the reso^2 input."""
    B, ch, g, _ = x.shape
    stride = reso // g
    anchors = anchors if anchors is not None else YOLO_ANCHORS[stride]
    nA = len(anchors)
    nattr = ch // nA
    x = x.reshape(B, nA, nattr, g, g).transpose(0, 1, 3, 4, 2).astype(F32)
    gx = np.arange(g, dtype=F32)[None, None, None, :]
    gy = np.arange(g, dtype=F32)[None, None, :, None]
    aw = np.array([F32(a[0] / stride) for a in anchors], F32)[None, :, None, None]
    ah = np.array([F32(a[1] / stride) for a in anchors], F32)[None, :, None, None]
    det = np.empty_like(x)
    det[..., 0] = _sigmoid(x[..., 0]) + gx
    det[..., 1] = _sigmoid(x[..., 1]) + gy
    det[..., 2] = np.exp(x[..., 2]).astype(F32) * aw
    det[..., 3] = np.exp(x[..., 3]).astype(F32) * ah
    det[..., :4] *= F32(stride)
    det[..., 4:] = _sigmoid(x[..., 4:])
    return det.reshape(B, -1, nattr)


def yolo_decode(heads, reso: int = 416, anchors=None) -> np.ndarray:
    """heads: list of raw [B,18,g,g] in network order (stride 32, 16, 8) -> [B,10647,6].
    anchors: optional per-head anchor lists (default: the yolov3-single.cfg masks by stride)."""
    return np.concatenate([yolo_decode_head(h, reso, None if anchors is None else anchors[i]) for i, h in enumerate(heads)], axis=1)


# ----------------------------------------------------------------------------------------------------
# a4  write_results  yolo/util.py:118-223  (nms hard-coded off; keep arg-max objectness per image)
# a5  box rescale    dataloader.py:350-364
# ----------------------------------------------------------------------------------------------------
def write_results(pred: np.ndarray, confidence: float = 0.01):
    """pred [B,R,6] -> (dets [D,8] fp32, rows int64 [D]) or (0, None) when no image has a candidate.

    Row = (img_idx, x1, y1, x2, y2, obj, cls_conf, cls_idx).  Tie-break: lowest flat row index
    (the reference's sort is unstable; parity inputs are tie-free)."""
    out, rows = [], []
    for b in range(pred.shape[0]):
        p = pred[b]
        obj = p[:, 4]
        mask = obj > F32(confidence)
        if not mask.any():
            continue
        masked = np.where(mask, obj, F32(-1))
        r = int(np.argmax(masked))
        cx, cy, w, h = p[r, 0], p[r, 1], p[r, 2], p[r, 3]
        half = F32(2)
        x1, y1 = cx - w / half, cy - h / half
        x2, y2 = cx + w / half, cy + h / half
        cls_conf = p[r, 5]  # max over the (single) class column of the masked prediction
        out.append([F32(b), x1, y1, x2, y2, obj[r], cls_conf, F32(0)])
        rows.append(r)
    if not out:
        return 0, None
    return np.array(out, F32), np.array(rows, np.int64)


def bbox_iou_plus1(box: np.ndarray, others: np.ndarray) -> np.ndarray:
    """yolo/bbox.py:51-77 `bbox_iou`: corner boxes, the "+1 pixel" convention, fp32, one rounding per operation."""
    ix1 = np.maximum(box[0], others[:, 0])
    iy1 = np.maximum(box[1], others[:, 1])
    ix2 = np.minimum(box[2], others[:, 2])
    iy2 = np.minimum(box[3], others[:, 3])
    one, zero = F32(1), F32(0)
    inter = np.maximum(ix2 - ix1 + one, zero) * np.maximum(iy2 - iy1 + one, zero)
    a1 = (box[2] - box[0] + one) * (box[3] - box[1] + one)
    a2 = (others[:, 2] - others[:, 0] + one) * (others[:, 3] - others[:, 1] + one)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / (a1 + a2 - inter)).astype(F32)


def write_results_nms(pred: np.ndarray, confidence: float = 0.01, nms_conf: float = 0.6):
    """The IoU-NMS branch the reference ships switched off (yolo/util.py:182-196, `nms = False` at :181), restated for
    multi-instance scenes.  pred [B,R,6] -> (dets [D,8] fp32, rows int64 [D], counts int64 [B]):
    per image the candidates (obj > confidence) sorted by objectness, descending, then greedily: keep the first, drop
    every later box whose IoU with it is not < nms_conf, repeat.  Detections are image-major, best first.
    Tie-break of the sort: lower flat row first (torch.sort is unstable there; parity inputs are tie-free)."""
    out, rows, counts = [], [], []
    half = F32(2)
    for b in range(pred.shape[0]):
        p = pred[b].astype(F32)
        cand = np.flatnonzero(p[:, 4] > F32(confidence))
        order = cand[np.argsort(-p[cand, 4].astype(np.float64), kind="stable")]
        q = p[order]
        boxes = np.stack([q[:, 0] - q[:, 2] / half, q[:, 1] - q[:, 3] / half, q[:, 0] + q[:, 2] / half, q[:, 1] + q[:, 3] / half], 1).astype(F32)
        alive = np.ones(len(order), bool)
        n = 0
        for i in range(len(order)):
            if not alive[i]:
                continue
            out.append([F32(b), *boxes[i], q[i, 4], q[i, 5], F32(0)])
            rows.append(int(order[i]))
            n += 1
            if i + 1 < len(order):
                iou = bbox_iou_plus1(boxes[i], boxes[i + 1:])
                alive[i + 1:] &= iou < F32(nms_conf)  # NaN compares false: dropped, as `image_pred_class[1:][ious < nms_conf]`
        counts.append(n)
    return np.array(out, F32).reshape(-1, 8), np.array(rows, np.int64), np.array(counts, np.int64)


def rescale_boxes(dets: np.ndarray, im_w: int, im_h: int, reso: int = 416):
    """-> boxes [D,4] fp32 in frame pixels, scores [D,1]."""
    wr = F32(im_w) / F32(reso)
    hr = F32(im_h) / F32(reso)
    boxes = dets[:, 1:5].copy()
    boxes[:, 0] *= wr
    boxes[:, 1] *= hr
    boxes[:, 2] *= wr
    boxes[:, 3] *= hr
    return boxes, dets[:, 5:6].copy()


# ----------------------------------------------------------------------------------------------------
# a6  im_to_torch (KPD/src/utils/img.py:13-18) + crop_from_dets (dataloader.py:794-835) + cropBox
#     (img.py:242-262), torch-2.11 division semantics
# ----------------------------------------------------------------------------------------------------
CROP_MEANS = (F32(0.406), F32(0.457), F32(0.480))  # applied to (R,G,B) in that order, dataloader.py:802-804


def expand_box(box, im_w: int, im_h: int):
    """box (x1,y1,x2,y2) fp32 -> (pt1, pt2) fp32 *un-truncated* expanded corners (dataloader.py:806-833)."""
    x1, y1, x2, y2 = (F32(v) for v in box)
    ht = y2 - y1
    width = x2 - x1
    rate = F32(0.2) if width > 100 else F32(0.3)
    ulx = max(F32(0), x1 - width * rate / F32(2))
    uly = max(F32(0), y1 - ht * rate / F32(2))
    brx = max(min(F32(im_w - 1), x2 + width * rate / F32(2)), ulx + F32(5))
    bry = max(min(F32(im_h - 1), y2 + ht * rate / F32(2)), uly + F32(5))
    return np.array([ulx, uly], F32), np.array([brx, bry], F32)


def crop_geometry(pt1, pt2, im_w: int, im_h: int, res_h: int = 320, res_w: int = 256):
    """Integer geometry of cropBox: returns dict(ulx, uly, hb, wb, Hp, Wp, hS, wS, top, left)."""
    ulx, uly = int(pt1[0]), int(pt1[1])  # .int() truncates toward zero (values are >= 0)
    brx, bry = int(pt2[0]), int(pt2[1])
    hb, wb = bry - uly, brx - ulx
    cand = F32(wb * res_h) / F32(res_w)  # int tensor * int -> int; / -> true division (fp32)
    if cand > hb:
        lenH = cand
        lenW = lenH * F32(res_w) / F32(res_h)
    else:
        lenH = hb
        lenW = F32(hb * res_w) / F32(res_h)
    Hp, Wp = int(lenH), int(lenW)
    hS = max(0, min(hb, im_h - uly))
    wS = max(0, min(wb, im_w - ulx))
    top = int(math.ceil(max(Hp - hS, 0) / 2.0))
    left = int(math.ceil(max(Wp - wS, 0) / 2.0))
    return dict(ulx=ulx, uly=uly, hb=hb, wb=wb, Hp=max(Hp, hS), Wp=max(Wp, wS), hS=hS, wS=wS, top=top, left=left)


def crop_box(frame_rgb_u8: np.ndarray, pt1, pt2, res_h: int = 320, res_w: int = 256) -> np.ndarray:
    """-> fp32 [3,res_h,res_w]: (u8/255 - mean) patch, centred zero pad to the 320:256 aspect, bilinear
    align_corners=True (torch CPU upsample_bilinear2d arithmetic, fp32)."""
    H, W, _ = frame_rgb_u8.shape
    g = crop_geometry(pt1, pt2, W, H, res_h, res_w)
    img = frame_rgb_u8.astype(F32).transpose(2, 0, 1) / F32(255)
    for c in range(3):
        img[c] = img[c] + (-CROP_MEANS[c])
    Hp, Wp = g["Hp"], g["Wp"]
    pad = np.zeros((3, Hp, Wp), F32)
    patch = img[:, g["uly"]:g["uly"] + g["hS"], g["ulx"]:g["ulx"] + g["wS"]]
    pad[:, g["top"]:g["top"] + g["hS"], g["left"]:g["left"] + g["wS"]] = patch
    # bilinear, align_corners=True
    rh = F32(Hp - 1) / F32(res_h - 1) if res_h > 1 else F32(0)
    rw = F32(Wp - 1) / F32(res_w - 1) if res_w > 1 else F32(0)
    ys = (rh * np.arange(res_h, dtype=F32)).astype(F32)
    xs = (rw * np.arange(res_w, dtype=F32)).astype(F32)
    y0 = ys.astype(np.int64)
    x0 = xs.astype(np.int64)
    y1 = np.minimum(y0 + 1, Hp - 1)
    x1 = np.minimum(x0 + 1, Wp - 1)
    ly = (ys - y0.astype(F32)).astype(F32)
    lx = (xs - x0.astype(F32)).astype(F32)
    hy = F32(1) - ly
    hx = F32(1) - lx
    p00 = pad[:, y0][:, :, x0]
    p01 = pad[:, y0][:, :, x1]
    p10 = pad[:, y1][:, :, x0]
    p11 = pad[:, y1][:, :, x1]
    top_ = hx[None, None, :] * p00 + lx[None, None, :] * p01
    bot_ = hx[None, None, :] * p10 + lx[None, None, :] * p11
    return (hy[None, :, None] * top_ + ly[None, :, None] * bot_).astype(F32)


# ----------------------------------------------------------------------------------------------------
# a8  getPrediction (KPD/src/utils/eval.py:113-147) + transformBoxInvert_batch (img.py:216-239)
# ----------------------------------------------------------------------------------------------------
def get_prediction(hms: np.ndarray, pt1: np.ndarray, pt2: np.ndarray, inp_h=320, inp_w=256, res_h=80, res_w=64):
    """hms [n,K,res_h,res_w] fp32; pt1/pt2 [n,2] -> (preds_hm [n,K,2], preds_img [n,K,2], maxval [n,K,1],
    idx int64 [n,K], sign int8 [n,K,2])."""
    n, K, H, W = hms.shape
    flat = hms.reshape(n, K, -1)
    idx = flat.argmax(axis=2)  # first (lowest) index on ties
    maxval = np.take_along_axis(flat, idx[..., None], axis=2).astype(F32)
    preds = np.stack([(idx % W).astype(F32), np.floor(idx.astype(F32) / F32(W))], axis=2)
    preds *= (maxval > 0).astype(F32)
    sign = np.zeros((n, K, 2), np.int8)
    for i in range(n):
        for j in range(K):
            px, py = int(round(float(preds[i, j, 0]))), int(round(float(preds[i, j, 1])))
            if 0 < px < W - 1 and 0 < py < H - 1:
                hm = hms[i, j]
                dx = np.sign(hm[py, px + 1] - hm[py, px - 1])
                dy = np.sign(hm[py + 1, px] - hm[py - 1, px])
                sign[i, j] = (dx, dy)
                preds[i, j, 0] += F32(dx) * F32(0.25)
                preds[i, j, 1] += F32(dy) * F32(0.25)
    preds = (preds + F32(0.2)).astype(F32)
    return preds, transform_box_invert(preds, pt1, pt2, inp_h, inp_w, res_h, res_w), maxval, idx, sign


def transform_box_invert(pt: np.ndarray, ul: np.ndarray, br: np.ndarray, inp_h=320, inp_w=256, res_h=80, res_w=64):
    ul = ul.astype(F32)
    br = br.astype(F32)
    center = (br - F32(1) - ul) / F32(2)
    size = br - ul
    size[:, 0] = size[:, 0] * F32(inp_h / inp_w)
    lenH = size.max(axis=1)
    lenW = lenH * F32(inp_w / inp_h)
    _pt = (pt.astype(F32) * lenH[:, None, None]) / F32(res_h)
    _pt[:, :, 0] = _pt[:, :, 0] - np.maximum((lenW[:, None] - F32(1)) / F32(2) - center[:, 0:1], F32(0))
    _pt[:, :, 1] = _pt[:, :, 1] - np.maximum((lenH[:, None] - F32(1)) / F32(2) - center[:, 1:2], F32(0))
    out = np.zeros_like(_pt)
    out[:, :, 0] = _pt[:, :, 0] + ul[:, 0:1]
    out[:, :, 1] = _pt[:, :, 1] + ul[:, 1:2]
    return out.astype(F32)


# ----------------------------------------------------------------------------------------------------
# a9  pose_nms (pPose_nms.py:24-122), general n >= 1; on the evaluate path n == 1 always
# ----------------------------------------------------------------------------------------------------
NMS_DELTA1, NMS_MU, NMS_DELTA2, NMS_GAMMA = 1, 1.7, 2.65, 22.48
NMS_SCORE_THR, NMS_MATCH_THR, NMS_AREA_THR, NMS_ALPHA = 0.3, 5, 0, 0.1


def _pairwise_kp_dist(pick: np.ndarray, allp: np.ndarray) -> np.ndarray:
    return np.sqrt(((pick[None] - allp) ** 2).sum(axis=2, dtype=F32)).astype(F32)


def pose_nms(bboxes, bbox_scores, pose_preds, pose_scores):
    """bboxes [n,4], bbox_scores [n,1], pose_preds [n,K,2], pose_scores [n,K,1] (all fp32) ->
    list of dict(bbox, keypoints [K,2], kp_score [K,1], proposal_score [1])."""
    pose_scores = pose_scores.astype(F32).copy()
    pose_scores[pose_scores == 0] = F32(1e-5)
    pose_preds = pose_preds.astype(F32)
    ori_preds, ori_scores, ori_bs = pose_preds.copy(), pose_scores.copy(), bbox_scores.astype(F32).copy()
    widths = bboxes[:, 2] - bboxes[:, 0]
    heights = bboxes[:, 3] - bboxes[:, 1]
    ref_dists = (F32(NMS_ALPHA) * np.maximum(widths, heights)).astype(F32)
    n = bboxes.shape[0]
    human_scores = pose_scores.mean(axis=1, dtype=F32)[:, 0]
    ids = np.arange(n)
    preds, scores = pose_preds, pose_scores
    pick, merge_ids = [], []
    while human_scores.shape[0] != 0:
        pid = int(np.argmax(human_scores))
        pick.append(ids[pid])
        ref = ref_dists[ids[pid]]
        dist = _pairwise_kp_dist(preds[pid], preds)  # [m,K]
        # get_parametric_distance (pPose_nms.py:243-267)
        mask = dist <= 1
        sd = np.zeros_like(dist)
        ps = np.broadcast_to(scores[pid, :, 0][None], dist.shape)
        ks = scores[:, :, 0]
        sd[mask] = np.tanh(ps[mask] / F32(NMS_DELTA1)) * np.tanh(ks[mask] / F32(NMS_DELTA1))
        point_dist = np.exp(F32(-1) * dist / F32(NMS_DELTA2))
        simi = sd.sum(axis=1) + F32(NMS_MU) * point_dist.sum(axis=1)
        # PCK_match (pPose_nms.py:270-281)
        nmatch = (dist / min(ref, F32(7)) <= 1).sum(axis=1)
        dele = np.nonzero((simi > NMS_GAMMA) | (nmatch >= NMS_MATCH_THR))[0]
        if dele.shape[0] == 0:
            dele = np.array([pid])
        merge_ids.append(ids[dele])
        preds = np.delete(preds, dele, axis=0)
        scores = np.delete(scores, dele, axis=0)
        ids = np.delete(ids, dele)
        human_scores = np.delete(human_scores, dele)
    out = []
    for j, pk in enumerate(pick):
        if ori_scores[pk, :, 0].max() < NMS_SCORE_THR:
            continue
        mid = merge_ids[j]
        # p_merge_fast (pPose_nms.py:204-240)
        cp, csc = ori_preds[mid], ori_scores[mid]
        dist = _pairwise_kp_dist(ori_preds[pk], cp)
        m = (dist <= min(ref_dists[pk], F32(15))).astype(F32)[..., None]
        masked = csc * m
        normed = masked / masked.sum(axis=0)
        merge_pose = (cp * np.repeat(normed, 2, axis=2)).sum(axis=0).astype(F32)
        merge_score = (masked * normed).sum(axis=0).astype(F32)  # [K,1]
        mx = merge_score.max()
        if mx < NMS_SCORE_THR:
            continue
        xs, ys = merge_pose[:, 0], merge_pose[:, 1]
        if 1.5 ** 2 * (xs.max() - xs.min()) * (ys.max() - ys.min()) < NMS_AREA_THR:
            continue
        out.append({
            "bbox": bboxes[0],
            "keypoints": (merge_pose - F32(0.3)).astype(F32),
            "kp_score": merge_score,
            "proposal_score": (merge_score.mean(dtype=F32) + ori_bs[pk] + F32(1.25) * mx).astype(F32).reshape(1),
        })
    return out


def pose_nms_single(det_score: float, preds_img: np.ndarray, maxval: np.ndarray):
    """n == 1 closed form: returns None if rejected, else (keypoints [K,2], kp_score [K],
    proposal_score)."""
    sc = maxval.reshape(-1).astype(F32).copy()
    sc[sc == 0] = F32(1e-5)
    mx = sc.max()
    if mx < F32(NMS_SCORE_THR):
        return None
    return (preds_img.astype(F32) - F32(0.3)).astype(F32), sc, F32(sc.mean(dtype=F32) + F32(det_score) + F32(1.25) * mx)


# ----------------------------------------------------------------------------------------------------
# a10 keypoint selection  dataloader.py:715-724
# ----------------------------------------------------------------------------------------------------
def select_keypoints(kp_score: np.ndarray, left_number: int) -> np.ndarray:
    """Indices (ascending, original order preserved) of the keypoints that survive repeatedly deleting the
    arg-min score (first on ties) until left_number remain."""
    keep = list(range(len(kp_score)))
    sc = list(np.asarray(kp_score, F32))
    while len(keep) > left_number:
        d = int(np.argmin(np.array(sc, F32)))
        del keep[d]
        del sc[d]
    return np.array(keep, np.int64)


# ----------------------------------------------------------------------------------------------------
# a13 Model3D.load / refine  utils/model.py:29-46,79-85 ; load_sixd_models betapose_evaluate.py:53-84
# ----------------------------------------------------------------------------------------------------
CAM_K = np.array([[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]], np.float64)


def load_ply_vertices(path: str, scale: float = 0.001) -> np.ndarray:
    """ASCII PLY -> [N,3] float64 * scale (plyfile is un-vendored; the files shipped are ASCII)."""
    with open(path) as f:
        lines = f.read().split("\n")
    nv, i = 0, 0
    while lines[i].strip() != "end_header":
        t = lines[i].split()
        if len(t) == 3 and t[0] == "element" and t[1] == "vertex":
            nv = int(t[2])
        i += 1
    body = lines[i + 1:i + 1 + nv]
    v = np.array([[float(x) for x in ln.split()[:3]] for ln in body], np.float64)
    return v * scale


def refine_vertices(v: np.ndarray, n_keep: int) -> np.ndarray:
    """Greedy closest-pair deletion until n_keep points remain (utils/model.py:29-46)."""
    v = v.copy()
    while v.shape[0] > n_keep:
        d = np.sqrt(np.sum(np.square(v[:, None] - v[None]), axis=2))
        d[np.arange(len(v)), np.arange(len(v))] = np.inf
        i, _ = np.unravel_index(np.argmin(d), d.shape)
        if not d[i, _] < 100.0:  # reference starts from min_dist = 100.0 and keeps a stale index otherwise
            i = 0
        v = np.delete(v, i, axis=0)
    return v


# ---------------------------------------------------------------------------------------------------------------------
# scoring
def add_err(gt_pose, est_pose, model):
    """3_6Dpose_estimator/utils/metrics.py:10-22: mean_v |(R_gt v + t_gt) - (R_est v + t_est)|, 4x4 poses, model [V,3]."""
    gt_pose, est_pose, model = np.asarray(gt_pose, np.float64), np.asarray(est_pose, np.float64), np.asarray(model, np.float64)
    a = model @ gt_pose[:3, :3].T + gt_pose[:3, 3]
    b = model @ est_pose[:3, :3].T + est_pose[:3, 3]
    return float(np.mean(np.linalg.norm(a - b, axis=1)))


def projection_error_2d(gt_pose, est_pose, model, cam):
    """utils/metrics.py:96-127: mean pixel distance between the vertices projected by cam @ pose[:3]."""
    model = np.asarray(model, np.float64)
    h = np.concatenate([model, np.ones((len(model), 1))], 1)
    g = (np.asarray(cam, np.float64) @ np.asarray(gt_pose, np.float64)[:3]) @ h.T
    e = (np.asarray(cam, np.float64) @ np.asarray(est_pose, np.float64)[:3]) @ h.T
    g, e = (g / g[2])[:2].T, (e / e[2])[:2].T
    return float(np.mean(np.linalg.norm(g - e, axis=1)))


def box_iou(gt_box, est_box):
    """utils/metrics.py:77-93 (corners x1, y1, x2, y2)."""
    xA, yA = max(gt_box[0], est_box[0]), max(gt_box[1], est_box[1])
    xB, yB = min(gt_box[2], est_box[2]), min(gt_box[3], est_box[3])
    if xB <= xA or yB <= yA:
        return 0.0
    inter = (xB - xA) * (yB - yA)
    return float(inter / float((gt_box[2] - gt_box[0]) * (gt_box[3] - gt_box[1]) + (est_box[2] - est_box[0]) * (est_box[3] - est_box[1]) - inter))
