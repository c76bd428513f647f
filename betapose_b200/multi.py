"""Scenes with several instances of the object: the evaluate path with the two switches the
reference leaves off turned on -- IoU box-NMS after the detector (yolo/util.py:182-196, disabled at :181) and the
multi-proposal branch of the parametric pose-NMS (pPose_nms.py:24-122) -- so that every surviving detection of a frame is
cropped, run through the key-point network, merged / suppressed against its neighbours and solved for its own pose.

    frames -> a1 resize -> a2 detector -> a3 decode (all 10647 rows) -> box NMS (bp_write_results_nms) -> a5 rescale
           -> a6 crop per detection -> a7 key-point net (chunks of the engine's batch) -> a8 heat-map decode
           -> a9 general pose-NMS per image (bp_pose_nms) -> a11 PnP per surviving pose (bp_pose_pnp, raw points)

Built on a BetaposeEngine's networks and buffers (object slot 0).  The number of detections is data-dependent, so unlike
BetaposeEngine.run this path reads two small counts back per batch and is not graph-captured; it is the multi-instance
variant, not the benchmarked one.  With max_det = 1 it reduces to the reference's behaviour (arg-max detection, identity
merge) and returns the poses of BetaposeEngine.run.  Key-point selection (`left_number` < K, the occlusion script) is not
offered here: bp_pose_pnp's raw-point mode carries no scores to select by.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, stages


class MultiInstance:
    def __init__(self, engine, nms_thr: float = 0.6, max_det: int = 8):
        if engine.left_number != engine.K:
            raise _lib.BetaposeError("MultiInstance: key-point selection (left_number < K) is not supported on this path")
        if not 1 <= int(max_det) <= 64:
            raise _lib.BetaposeError("MultiInstance: max_det must be in 1..64 (bp_pose_nms handles at most 64 proposals per image)")
        self.eng, self.nms_thr, self.max_det = engine, float(nms_thr), int(max_det)

    @torch.no_grad()
    def run(self, frames_u8):
        """frames uint8 [n <= engine.B, H, W, 3] RGB (host or cuda) -> list (per frame) of lists of dicts
        {bbox[4] (the frame's first box, as pose_nms reports it), bbox_pick[4] (the box of the picked proposal), det_score,
        keypoints[K,2], kp_score[K,1], proposal_score, cam_R[3,3] f64, cam_t[3,1] f64, status (1 pose / -1 PnP failed)}."""
        eng = self.eng
        fr = torch.as_tensor(frames_u8)
        n = int(fr.shape[0])
        assert n <= eng.B and tuple(fr.shape[1:]) == (eng.frame_h, eng.frame_w, 3) and fr.dtype == torch.uint8
        K = eng.K
        with torch.cuda.device(eng.device):
            L, e = _lib.lib(), eng.yolo[0].engine.handle
            yolo, kpd = eng.yolo[0], eng.kpd[0]
            st = _lib.stream_ptr()
            eng.frames[:n].copy_(fr, non_blocking=True)
            frames = eng.frames[:n]
            # a1 + a2
            _lib.check(L.bp_resize_bicubic(e, _lib.ptr(eng.frames), n, eng.frame_h, eng.frame_w, eng.reso, eng.reso,
                                           C.c_void_p(yolo.tensor_info(0)["ptr"]), None, st), "bp_resize_bicubic")
            yolo.forward(n)
            # a3 (every row decoded) + a4 with box NMS
            heads = [yolo.tensor(h["tensor"], n) for h in eng.heads[0]]
            dec = stages.yolo_decode_argmax(heads, [h["anchors"] for h in eng.heads[0]], n, reso=eng.reso, conf=eng.conf,
                                            frame_w=eng.frame_w, frame_h=eng.frame_h, n_attr=5 + eng.heads[0][0]["classes"],
                                            want_decoded=True)["decoded"]
            nms = stages.write_results_nms(dec, eng.conf, self.nms_thr, self.max_det)
            cnt = nms["count"].cpu().numpy().astype(np.int64)  # first read-back: detections per frame
            out = [[] for _ in range(n)]
            N = int(cnt.sum())
            if N == 0:
                return out
            keep = torch.from_numpy(np.arange(self.max_det)[None, :] < cnt[:, None]).to(eng.device)
            det = nms["det"][keep]  # [N, 8], frame-major, best first
            img_idx = det[:, 0].to(torch.int32).contiguous()
            # a5: boxes to frame pixels (dataloader.py:350-364), one fp32 multiply each, no clamp
            ratio = torch.tensor([eng.frame_w / eng.reso, eng.frame_h / eng.reso] * 2, dtype=torch.float32, device=eng.device)
            boxes = (det[:, 1:5] * ratio).contiguous()
            scores = det[:, 5].contiguous()
            # a6 -> a7 -> a8, in chunks of the key-point net's batch
            preds, maxv = [], []
            hm_id = eng.hm_id[0]
            for c0 in range(0, N, eng.B):
                m = min(eng.B, N - c0)
                crop = stages.crop_resize(frames, boxes[c0:c0 + m], img_idx[c0:c0 + m], res_h=eng.inp_h, res_w=eng.inp_w)
                kpd.input(m).copy_(stages.net_input_pixels(crop["net"]))
                kpd.forward(m)
                d8 = stages.heatmap_decode(kpd.tensor(hm_id, m), crop["pt1"], crop["pt2"], layout="nhwc", inp_h=eng.inp_h, inp_w=eng.inp_w)
                preds.append(d8["preds_img"])
                maxv.append(d8["maxval"].reshape(m, K))
            preds, maxv = torch.cat(preds), torch.cat(maxv)
            maxv[maxv == 0] = 1e-5  # pose_nms does this in place to the caller's scores (pPose_nms.py:31)
            # a9: general pose-NMS, one block per frame that has detections
            has = np.flatnonzero(cnt > 0)
            pn = stages.pose_nms(boxes, scores, preds, maxv, counts=[int(c) for c in cnt[has]])
            n_out = pn["count"].cpu().numpy().astype(np.int64)  # second read-back: surviving poses per frame
            first = np.concatenate([[0], np.cumsum(cnt[has])[:-1]]).astype(np.int64)
            rows = np.concatenate([first[i] + np.arange(n_out[i]) for i in range(len(has))]).astype(np.int64) if n_out.sum() else np.zeros(0, np.int64)
            if rows.size == 0:
                return out
            ridx = torch.from_numpy(rows).to(eng.device)
            kps = pn["keypoints"][ridx].contiguous()  # merged key-points, already shifted by -0.3
            # a10 (no-op: left_number == K) + a11
            pose = stages.pose_pnp(kps, None, None, eng.kp3d[0].contiguous(), cam_K=eng.cam_K, left_number=K, mode=eng.pnp_mode,
                                   reproj_thr=eng.reproj_thr, n_hyp=eng.n_hyp, seed=eng.seed, flags=stages.PNP_RAW_POINTS)
            kps_h = kps.cpu().numpy()
            sc_h = pn["kp_score"][ridx].cpu().numpy()
            prop_h = pn["proposal"][ridx].cpu().numpy()
            pick_h = pn["pick"][ridx].cpu().numpy()
            R_h, t_h, st_h = pose["R"].cpu().numpy(), pose["t"].cpu().numpy(), pose["status"].cpu().numpy()
            boxes_h, scores_h = boxes.cpu().numpy(), scores.cpu().numpy()
            j = 0
            for i, f in enumerate(has):
                for _ in range(int(n_out[i])):
                    src = int(first[i] + pick_h[j])
                    out[int(f)].append({"bbox": boxes_h[first[i]], "bbox_pick": boxes_h[src], "det_score": float(scores_h[src]),
                                        "keypoints": kps_h[j], "kp_score": sc_h[j].reshape(K, 1), "proposal_score": float(prop_h[j]),
                                        "cam_R": R_h[j].reshape(3, 3), "cam_t": t_h[j].reshape(3, 1), "status": int(st_h[j])})
                    j += 1
            return out
