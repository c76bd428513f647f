"""Drop-in seams: the Python call signatures at which the reference's stage objects reach into its models and
post-processing, re-pointed at the CUDA engine.  Same names, argument meaning, return types and
sentinels, so `dataloader.py` / `betapose_evaluate.py` can import these instead of their own:

    from betapose_b200.compat import Darknet, dynamic_write_results      # yolo/darknet.py, yolo/util.py
    from betapose_b200.compat import crop_from_dets                       # dataloader.py:794-835
    from betapose_b200.compat import InferenNet_fast                      # KPD/src/main_fast_inference.py:26-46
    from betapose_b200.compat import getPrediction, pose_nms, pnp, write_json

Tensors may live on the CPU (as in the reference's threads) or on the GPU; compute always runs on the GPU and the
result comes back on the caller's device.  There is no CPU fallback: without CUDA these raise BetaposeError.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import _lib, net as _net, stages, yolo_cfg
from .opt import opt

_CROP_MEANS = (0.406, 0.457, 0.480)  # dataloader.py:802-804 (BGR means applied in RGB order; kept)


def _dev():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


# ---------------------------------------------------------------------------------------------- detector
class Darknet:
    """yolo/darknet.py:209-432.  `__call__(img[B,3,reso,reso] fp32 RGB 0..1) -> [B, rows, 5+classes]` (anchor-major
    rows, heads in network order), on the caller's device."""

    def __init__(self, cfgfile: str | None = None, reso: int = 416, max_batch: int = 8):
        text = open(cfgfile).read() if cfgfile else yolo_cfg.default_cfg_text()
        self.blocks = yolo_cfg.parse_cfg_text(text)
        self.net_info = self.blocks[0]  # the cfg has no [net] section (darknet.py:231); dataloader.py:296 writes 'height' here
        self.reso = int(reso)
        self.max_batch = int(max_batch)
        self._net = None
        self._params = None
        self.header = None
        self.seen = 0

    def load_weights(self, path, cutoff=None):
        """16-byte header {major, minor, revision, seen} then the fp32 stream (darknet.py:377-432)."""
        with open(path, "rb") as f:
            self.header = np.fromfile(f, dtype=np.int32, count=4)
            stream = np.fromfile(f, dtype=np.float32)
        self.seen = int(self.header[3])
        self.load_stream(stream)

    def load_stream(self, stream: np.ndarray):
        self._params, used = _net.split_darknet_stream(self.blocks, np.asarray(stream, np.float32))
        self._net = None

    def eval(self):
        return self

    def cuda(self, device=None):
        return self

    def _build(self, batch: int):
        if self._params is None:
            raise _lib.BetaposeError("Darknet: load_weights() first")
        if self._net is None or self._net.max_batch < batch:
            self.max_batch = max(self.max_batch, batch)
            self._net = _net.Net(self.max_batch, self.reso, self.reso, _lib.IN_F16)
            self._heads = _net.build_darknet(self._net, self.blocks, self._params)

    def __call__(self, x: torch.Tensor, CUDA: bool = True) -> torch.Tensor:
        dev = _dev()
        B = int(x.shape[0])
        assert tuple(x.shape[1:]) == (3, self.reso, self.reso), x.shape
        self._build(B)
        inp = self._net.input(B)
        inp.copy_(x.to(dev, non_blocking=True).permute(0, 2, 3, 1))  # data pixels R,G,B of the padded fp16 input buffer
        self._net.forward(B)
        heads = [self._net.tensor(h["tensor"], B) for h in self._heads]
        out = stages.yolo_decode_argmax(heads, [h["anchors"] for h in self._heads], B, reso=self.reso, conf=opt.confidence,
                                        n_attr=5 + self._heads[0]["classes"], want_decoded=True)
        self.last = out
        return out["decoded"].to(x.device)

    forward = __call__


def dynamic_write_results(prediction, confidence, num_classes, nms=True, nms_conf=0.4, box_nms: bool = False):
    """yolo/util.py:104-223: one row (img_idx, x1, y1, x2, y2, obj, cls_conf, cls_idx) per image that has a candidate
    (NMS is hard-coded off there, whatever `nms` says: arg-max objectness), or the int 0 when no image has one.

    box_nms=True (not a reference argument) switches the reference's shipped-but-disabled IoU-NMS branch on, for scenes
    with several instances: every surviving box per image, image-major, best first, with dynamic_write_results' own
    retry -- more than 100 detections in the batch => once more with nms_conf - 0.05 (util.py:111-113)."""
    dev = _dev()
    if box_nms:
        pred = prediction.to(dev, dtype=torch.float32)

        def run(thr):
            r = stages.write_results_nms(pred, float(confidence), float(thr), max_det=int(pred.shape[1]))
            cnt = r["count"].cpu()
            keep = torch.arange(r["det"].shape[1])[None, :] < cnt[:, None]
            return r["det"][keep.to(dev)], int(cnt.sum())

        dets, n = run(nms_conf)
        if n == 0:
            return 0
        if n > 100:
            dets, n = run(nms_conf - 0.05)
        return dets.to(prediction.device)
    res = stages.write_results(prediction.to(dev, dtype=torch.float32), float(confidence))
    valid = res["valid"].bool()
    if not bool(valid.any()):
        return 0
    return res["det"][valid].to(prediction.device)


write_results = dynamic_write_results


# ---------------------------------------------------------------------------------------------- crop
def crop_from_dets(img, boxes, inps, pt1, pt2):
    """dataloader.py:794-835: img [3,H,W] fp32 RGB 0..1 (mean-subtracted IN PLACE, like the reference), boxes [n,4];
    fills and returns the caller's inps [n,3,320,256], pt1, pt2."""
    dev = _dev()
    H, W = int(img.shape[1]), int(img.shape[2])
    frame = (img.to(dev) * 255.0).round().clamp_(0, 255).to(torch.uint8).permute(1, 2, 0).contiguous()[None]
    n = int(boxes.shape[0])
    out = stages.crop_resize(frame, boxes.to(dev, dtype=torch.float32), torch.zeros(n, dtype=torch.int32, device=dev),
                             res_h=int(inps.shape[2]), res_w=int(inps.shape[3]), want_f16=False, want_f32=True)
    inps.copy_(out["f32"])
    pt1.copy_(out["pt1"])
    pt2.copy_(out["pt2"])
    for c in range(3):
        img[c].add_(-_CROP_MEANS[c])
    return inps, pt1, pt2


# ---------------------------------------------------------------------------------------------- key-point network
_MODEL_NAMES = {1: "seq1_model", 2: "seq2_model", 4: "seq4_model", 5: "seq5_model", 6: "seq6_model", 8: "seq8_model",
                9: "seq9_model", 10: "Semmetry_obj10", 11: "seq11_model", 12: "seq12_model", 13: "seq13_model",
                14: "seq14_model", 15: "seq15_model"}  # main_fast_inference.py:29-32 ('NULL' for ids 0, 3, 7)


class InferenNet_fast:
    """KPD/src/main_fast_inference.py:26-46: FastPose + `narrow(1, 0, 50)`; `__call__(x[n,3,320,256]) -> [n,50,80,64]`."""

    def __init__(self, kernel_size=5, obj_id=1, dataset=None, state_dict=None, n_maps: int | None = None, max_batch: int = 80):
        if state_dict is None:
            name = _MODEL_NAMES.get(int(obj_id))
            if name is None:
                raise _lib.BetaposeError(f"no key-point model for object id {obj_id}")
            path = os.path.join("./exp/final_model", name + ".pkl")
            print("Loading pose model from {}".format(path))
            state_dict = torch.load(path, map_location="cpu")
        self.sd = state_dict
        self.n_maps = int(n_maps if n_maps is not None else opt.nClasses)
        self.max_batch = int(max_batch)
        self._net = None

    def cuda(self, device=None):
        return self

    def eval(self):
        return self

    def _build(self, n):
        if self._net is None or self._net.max_batch < n:
            self.max_batch = max(self.max_batch, n)
            self._net = _net.Net(self.max_batch, opt.inputResH, opt.inputResW, _lib.IN_F16)
            self._hm = _net.build_fastpose(self._net, self.sd, self.n_maps)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        dev = _dev()
        n = int(x.shape[0])
        self._build(n)
        inp = self._net.input(n)
        inp.copy_(x.to(dev, non_blocking=True).permute(0, 2, 3, 1))  # data pixels R,G,B of the padded fp16 input buffer
        self._net.forward(n)
        return self._net.tensor(self._hm, n).permute(0, 3, 1, 2).contiguous().to(x.device)

    forward = __call__


# ---------------------------------------------------------------------------------------------- decode / nms / pnp
def getPrediction(hms, pt1, pt2, inpH, inpW, resH, resW):
    """KPD/src/utils/eval.py:113-147 -> (preds_hm [n,K,2], preds_img [n,K,2], maxval [n,K,1]) on hms' device."""
    dev = _dev()
    out = stages.heatmap_decode(hms.to(dev, dtype=torch.float32).contiguous(), pt1.to(dev, dtype=torch.float32),
                                pt2.to(dev, dtype=torch.float32), layout="nchw", inp_h=int(inpH), inp_w=int(inpW))
    return out["preds_hm"].to(hms.device), out["preds_img"].to(hms.device), out["maxval"].to(hms.device)


def pose_nms(bboxes, bbox_scores, pose_preds, pose_scores):
    """pPose_nms.py:24-122: parametric pose-NMS over the n proposals of one image -> list of dicts
    {bbox (always bboxes[0], as in the reference :116), keypoints [K,2] (merged - 0.3), kp_score [K,1], proposal_score
    [1]}.  Zero scores become 1e-5 in place, like the reference.  On the evaluate path n == 1 (yolo/util.py:205-211
    keeps the arg-max row); n > 1 is SURVEY.md 8(f) item 3."""
    dev = _dev()
    n, K = int(pose_preds.shape[0]), int(pose_preds.shape[1])
    pose_scores[pose_scores == 0] = 1e-5
    out = stages.pose_nms(bboxes.to(dev), bbox_scores.to(dev).reshape(n), pose_preds.to(dev), pose_scores.to(dev).reshape(n, K))
    cnt = int(out["count"][0])
    home = pose_preds.device
    return [{"bbox": bboxes[0], "keypoints": out["keypoints"][j].to(home), "kp_score": out["kp_score"][j].reshape(K, 1).to(home),
             "proposal_score": out["proposal"][j].reshape(1).to(home)} for j in range(cnt)]


def pnp(points_3D, points_2D, cameraMatrix, mode: int = stages.MODE_RANSAC):
    """utils/utils.py:17-41 -> (R [3,3] f64, t [3,1] f64).  mode RANSAC mirrors the cv2.solvePnPRansac(12 px) variant,
    MODE_ALLPTS the active cv2.solvePnP call where that converges."""
    assert points_3D.shape[0] == points_2D.shape[0], "points 3D and points 2D must have same number of vertices"
    dist = getattr(pnp, "distCoeffs", None)  # optional function attribute of the reference's pnp (utils.py:18-21); default zeros
    if dist is not None and np.any(np.asarray(dist, np.float64) != 0):
        raise _lib.BetaposeError("pnp: non-zero lens distortion (pnp.distCoeffs) is not supported by the PnP kernel; undistort the points first")
    dev = _dev()
    K = int(points_3D.shape[0])
    p2 = torch.as_tensor(np.ascontiguousarray(np.asarray(points_2D, np.float32)[:, :2])).to(dev).reshape(1, K, 2)
    p3 = torch.as_tensor(np.ascontiguousarray(np.asarray(points_3D, np.float64))).to(dev)
    out = stages.pose_pnp(p2, None, None, p3, cam_K=np.asarray(cameraMatrix, np.float64), left_number=K, mode=mode,
                          flags=stages.PNP_RAW_POINTS)
    if int(out["status"][0]) != 1:
        raise _lib.BetaposeError("pnp: no pose (degenerate correspondences)")
    return out["R"][0].cpu().numpy().reshape(3, 3), out["t"][0].cpu().numpy().reshape(3, 1)


# ---------------------------------------------------------------------------------------------- output
def result_from_record(rec, imgname: str, K: int = 50) -> dict:
    """bp_record -> the dict DataWriter appends (dataloader.py:708-730): {'imgname', 'result', 'cam_R', 'cam_t'}."""
    if int(rec["status"]) == 0:  # no detection, or the pose was rejected by pose-NMS: DataWriter appends result = [] (:727-728)
        return {"imgname": imgname, "result": [], "cam_R": [], "cam_t": []}
    # status 1, and status -1 = a pose-NMS survivor whose PnP found no consensus: the reference appends bbox, key-points
    # and cam_R / cam_t for EVERY survivor (dataloader.py:715-727; solvePnP's return flag is ignored), so the frame stays
    # in Betapose-results.json and in the IoU / ADD / 2-D statistics.  cam_R / cam_t then hold the solver's last estimate
    # (best hypothesis; all zeros if none could be solved), which scores as a miss.
    kp = np.asarray(rec["keypoints"], np.float32).reshape(50, 3)[:K]
    human = {"bbox": np.asarray(rec["box"]), "keypoints": kp[:, :2], "kp_score": kp[:, 2:3], "proposal_score": float(rec["proposal_score"])}
    return {"imgname": imgname, "result": [human], "cam_R": np.asarray(rec["R"]).reshape(3, 3), "cam_t": np.asarray(rec["t"]).reshape(3, 1)}


def write_json(all_results, outputpath, for_eval=False):
    """pPose_nms.py:284-371 (default format): <outputpath>/Betapose-results.json = one JSON array, one entry per pose:
    image_id, cam_R [9] row-major, cam_t [3], keypoints [x, y, score]*K, score.  Images without a pose emit nothing."""
    json_results = []
    for im_res in all_results:
        im_name = im_res["imgname"]
        for human in im_res["result"]:
            result = {}
            if for_eval:
                result["image_id"] = int(im_name.split("/")[-1].split(".")[0].split("_")[-1])
            else:
                result["image_id"] = im_name.split("/")[-1]
            if len(im_res["cam_R"]) > 0:
                result["cam_R"] = np.array(im_res["cam_R"]).reshape((9, 1))[:, 0].tolist()
                result["cam_t"] = np.array(im_res["cam_t"]).reshape((3, 1))[:, 0].tolist()
            kp_preds, kp_scores = human["keypoints"], human["kp_score"]
            keypoints = []
            for n in range(kp_scores.shape[0]):
                keypoints += [float(kp_preds[n, 0]), float(kp_preds[n, 1]), float(np.asarray(kp_scores[n]).reshape(-1)[0])]
            result["keypoints"] = keypoints
            result["score"] = float(human["proposal_score"])
            json_results.append(result)
    os.makedirs(outputpath, exist_ok=True)
    with open(os.path.join(outputpath, "Betapose-results.json"), "w") as json_file:
        json_file.write(json.dumps(json_results))
    return json_results


# ---------------------------------------------------------------------------------------------- scoring (utils/metrics.py)
def _pose_batch(pose):
    p = np.asarray(pose, np.float64)
    return torch.from_numpy(p[None, :3, :3].copy()), torch.from_numpy(p[None, :3, 3].copy())


def _score_one(gt_pose, est_pose, model, cam=None, gt_box=None, est_box=None):
    dev = _dev()
    Rg, tg = _pose_batch(gt_pose)
    Re, te = _pose_batch(est_pose)
    gb = torch.tensor([gt_box if gt_box is not None else [0, 0, 1, 1]], dtype=torch.float32)
    eb = torch.tensor([est_box if est_box is not None else [0, 0, 1, 1]], dtype=torch.float32)
    return stages.score_poses(Re.to(dev), te.to(dev), eb.to(dev), Rg.to(dev), tg.to(dev), gb.to(dev),
                              torch.from_numpy(np.asarray(model, np.float64)).to(dev),
                              cam_K=stages.CAM_K if cam is None else cam)


def add_err(gt_pose, est_pose, model):
    """utils/metrics.py:10-22: mean distance of the model vertices under the two 4x4 poses."""
    return float(_score_one(gt_pose, est_pose, model)["add"][0])


def projection_error_2d(gt_pose, est_pose, model, cam):
    """utils/metrics.py:96-127: mean pixel distance of the projected model vertices."""
    return float(_score_one(gt_pose, est_pose, model, cam=cam)["proj"][0])


def iou(gt_box, est_box):
    """utils/metrics.py:77-93, boxes as corners (x1, y1, x2, y2)."""
    I = np.eye(4)
    return float(_score_one(I, I, np.zeros((1, 3)), gt_box=list(gt_box), est_box=list(est_box))["iou"][0])
