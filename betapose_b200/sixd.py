"""Ground-truth side of the evaluation: the SIXD / LineMod benchmark loader and the scoring loop
that follows the per-frame path in the reference.

Replaces `load_sixd` (3_6Dpose_estimator/utils/sixd.py:60-111), the model part of `load_sixd_models`
(betapose_evaluate.py:53-84) and the loop at betapose_evaluate.py:203-266 (occlusion variant :218-253), with the
ADD / 2-D reprojection / IoU arithmetic batched on the GPU (`bp_score_poses`).  Host code apart from that call."""
from __future__ import annotations

import os

import numpy as np
import torch
import yaml

from . import model3d, stages

SCALE_TO_METERS = 0.001  # gt translations and model vertices are in millimetres (sixd.py:63, betapose_evaluate.py:57)


class Frame:
    def __init__(self, nr: int, path: str, cam: np.ndarray):
        self.nr, self.path, self.cam = nr, path, cam
        self.gt: list = []  # (obj_id, pose [4,4] metres, obj_bb [x, y, w, h])


class Benchmark:
    def __init__(self):
        self.cam = np.identity(3)
        self.frames: list[Frame] = []
        self.diameter: list[float] = []


def _yaml(path: str):
    with open(path) as f:
        return yaml.safe_load(f)


def load_sixd(base_path: str, seq: int | None, nr_frames: int = 0) -> Benchmark:
    """sixd.py:60-111: camera.yml (optional), models/models_info.yml (diameters, list index = object id),
    test/<seq>/{info,gt}.yml."""
    b = Benchmark()
    cam_yml = os.path.join(base_path, "camera.yml")
    if os.path.exists(cam_yml):
        c = _yaml(cam_yml)
        b.cam[0, 0], b.cam[0, 2], b.cam[1, 1], b.cam[1, 2] = c["fx"], c["cx"], c["fy"], c["cy"]
    b.diameter.append(10000.0)  # index 0 unused: object ids start at 1
    for _, val in _yaml(os.path.join(base_path, "models", "models_info.yml")).items():
        b.diameter.append(float(val["diameter"]))
    if seq is None:
        return b
    path = os.path.join(base_path, "test", f"{seq:02d}")
    info, gts = _yaml(os.path.join(path, "info.yml")), _yaml(os.path.join(path, "gt.yml"))
    n = nr_frames if nr_frames > 0 else len(info)
    for i in range(n):
        cam = np.array(info[i]["cam_K"], np.float64).reshape(3, 3) if "cam_K" in info[i] else np.identity(3)
        fr = Frame(i, os.path.join(path, "rgb", f"{i:04d}.png"), cam)
        for gt in gts[i]:
            pose = np.identity(4)
            pose[:3, :3] = np.array(gt["cam_R_m2c"], np.float64).reshape(3, 3)
            pose[:3, 3] = np.array(gt["cam_t_m2c"], np.float64).reshape(3) * SCALE_TO_METERS
            fr.gt.append((int(gt["obj_id"]), pose, [float(v) for v in gt["obj_bb"]]))
        b.frames.append(fr)
    return b


def load_models(base_path: str, obj_id: int, n_kp: int = 50):
    """betapose_evaluate.py:53-84: -> (model vertices [V,3] metres, key-point model [n_kp,3] metres, diameter mm)."""
    info = _yaml(os.path.join(base_path, "models", "models_info.yml"))
    diam = float({int(k): v for k, v in info.items()}[obj_id]["diameter"])
    verts = model3d.load_ply(os.path.join(base_path, "models", f"obj_{obj_id:02d}.ply"), SCALE_TO_METERS)
    kp = model3d.load_kp_model(os.path.join(base_path, "kpmodels", f"obj_{obj_id:02d}.ply"), n_kp)
    return verts, kp, diam


def collect_scoring_pairs(final_result, bench: Benchmark, obj_id: int, occlusion: bool = False):
    """Which (ground truth, estimate) pairs the reference's loop scores (host logic, no GPU):
    betapose_evaluate.py:216-240 -- only the frame's FIRST ground-truth entry counts, and only if it is `obj_id`;
    occlusion_betapose_evaluate.py:216-236 -- every ground-truth entry of the frame that is `obj_id` (the Occlusion
    sequence annotates several objects per frame).  Frames without a pose are skipped either way.
    -> lists (R_gt, t_gt, box_gt corners, R_est, t_est, box_est)."""
    Rg, tg, bg, Re, te, be = [], [], [], [], [], []
    for f in final_result:
        nr = int(os.path.basename(f["imgname"])[0:-4])  # '0123.png' -> 123
        fr = bench.frames[nr]
        assert fr.nr == nr
        for gt_obj, gt_pose, gt_bb in (fr.gt if occlusion else fr.gt[:1]):
            if gt_obj != obj_id:
                continue
            if len(f["result"]) < 1 or len(f["result"][0]) < 1:
                continue
            Rg.append(gt_pose[:3, :3]); tg.append(gt_pose[:3, 3])
            bg.append([gt_bb[0], gt_bb[1], gt_bb[0] + gt_bb[2], gt_bb[1] + gt_bb[3]])  # [xmin, ymin, w, h] -> corners
            Re.append(np.asarray(f["cam_R"], np.float64).reshape(3, 3)); te.append(np.asarray(f["cam_t"], np.float64).reshape(3))
            be.append(np.asarray(f["result"][0]["bbox"], np.float64).reshape(4))
    return Rg, tg, bg, Re, te, be


def evaluate_results(final_result, bench: Benchmark, obj_id: int, model_vertices: np.ndarray, cam=model3d.CAM_K,
                     pixel_thresh: float | None = None, device=None, log=print, occlusion: bool = False,
                     left_keypoints: int | None = None) -> dict:
    """The scoring loop of betapose_evaluate.py:203-266 (occlusion=True: occlusion_betapose_evaluate.py:203-270, which
    scores every matching ground-truth entry, uses a 20 px reprojection threshold instead of 5 and names left_keypoints in
    its summary line).  final_result: list of {'imgname', 'result', 'cam_R', 'cam_t'} (compat.result_from_record).
    ADD and 2-D reprojection only when the boxes overlap by IoU >= 0.5.  All arithmetic in one bp_score_poses launch."""
    if pixel_thresh is None:
        pixel_thresh = 20.0 if occlusion else 5.0
    Rg, tg, bg, Re, te, be = collect_scoring_pairs(final_result, bench, obj_id, occlusion)
    if not Rg:
        return dict(n_frames=0, n_scored=0, add_accuracy=float("nan"), proj2d_accuracy=float("nan"), iou_accuracy=float("nan"))
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    T = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.array(a), dtype=dt)).to(dev)  # noqa: E731
    sc = stages.score_poses(T(Re, np.float64), T(te, np.float64), T(be, np.float32), T(Rg, np.float64), T(tg, np.float64),
                            T(bg, np.float32), T(model_vertices, np.float64), cam_K=np.asarray(cam, np.float64))
    out = stages.summarize_scores(sc["add"], sc["proj"], sc["iou"], sc["scored"], diameter_mm=bench.diameter[obj_id],
                                  pixel_thresh=pixel_thresh)
    out["n_frames"] = len(Rg)
    log("Mean add accuracy for seq %02d is: %.3f" % (obj_id, out["add_accuracy"]))
    if occlusion:
        log("2d reprojection accuracy with leftkeypoints %d for seq %02d is: %.3f" % (int(left_keypoints or 0), obj_id, out["proj2d_accuracy"]))
    else:
        log("2d reprojection accuracy for seq %02d is: %.3f" % (obj_id, out["proj2d_accuracy"]))
    log("Mean IoU for seq %02d is: %.3f" % (obj_id, out["iou_accuracy"]))
    return out
