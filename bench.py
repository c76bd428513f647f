#!/usr/bin/env python
"""Benchmark of the Betapose per-frame evaluate hot path on B200 (contract: see the task prompt / DESIGN.md 5).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, weak scaling)

A step = one batch of B synthetic 640x480 RGB frames per GPU through a1..a12 (resize, YOLOv3-416, decode+arg-max,
crop, FastPose, heat-map decode, pose-NMS, PnP, record packing) and, for N > 1, one NCCL all-gather of the packed
result records.  `value` is timed with the frames already in HBM; `e2e` goes through BetaposeEngine.run_stream with
pinned HOST frames (H2D + D2H inside the timed region).  Rank 0 prints ONE JSON line.  The headline workload is
BASELINE.json configs[2]; configs[3] (13 objects mixed, 32 frames per GPU) and configs[4] (Occlusion variant, 16 frames
per GPU) are measured the same way and reported as the sub-records `configs3` / `configs4` of the same line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before CUDA initialises: see betapose_b200/__init__.py

METRIC = "linemod_640x480_images_per_sec"
UNIT = "images/s"
LINEMOD_IDS = (1, 2, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14, 15)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(src="measured", hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sus=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])))
    return dict(src="fallback", hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], 0.0, set(), 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
                pw = max(pw, float(f[3]))
            except Exception:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "power_w_max": pw, "samples": len(sm), "reasons": sorted(reasons)}


# ====================================================================================================== CPU arm
def _reference_tree():
    """The unmodified reference, when this machine has it ($BETAPOSE_REF, <repo>/baseline/_ref, /root/reference).  It never
    exists on the GPU box; there the CPU arm is the oracle port."""
    try:
        from oracle import ref_shim
    except Exception:
        return None
    cand = os.path.join(ROOT, "baseline", "_ref")
    if not os.environ.get("BETAPOSE_REF") and os.path.isfile(os.path.join(cand, "3_6Dpose_estimator", "opt.py")):
        os.environ["BETAPOSE_REF"] = cand
    return ref_shim if ref_shim.available() else None


def cpu_pipeline_factory(threads: int, impl: str = "auto"):
    """The reference's path on host cores, one frame at a time like the reference (--detbatch 1, one detection per frame).
      impl "reference": the UNMODIFIED reference modules driven through oracle/ref_shim.py -- Darknet + dynamic_write_results +
             im_to_torch + crop_from_dets + FastPose + getPrediction + pose_nms + selection + pnp (cv2.solvePnP, the active call)
             -- when the reference tree is on this machine;
      impl "port": the oracle port -- torch-CPU fp32 networks (oracle/nets.py) + numpy stages (oracle/restate.py) + the
             reference's own third-party calls where importable (Pillow resize, cv2.solvePnPRansac);
      "auto" = reference if available, else port.
    Returns run(frames_u8, collect=None) -> number of poses; `collect` (a list) receives per-frame decisions (port only)."""
    import torch

    from betapose_b200 import synth, yolo_cfg
    from oracle import nets as onets
    from oracle import pnp as opnp
    from oracle import restate as R

    try:
        from PIL import Image
    except Exception:
        Image = None
    try:
        import cv2
    except Exception:
        cv2 = None
    torch.set_num_threads(threads)
    shim = _reference_tree() if impl in ("auto", "reference") else None
    if impl == "reference" and shim is None:
        raise RuntimeError("the reference tree is not on this machine")
    kp = synth.synth_kp_model(1, 50)
    stream = synth.cached_yolo_weights(1000)
    sd = synth.cached_kpd_state_dict(2000)

    if shim is not None:
        import tempfile

        import torchvision.transforms as transforms

        ref = shim.load_reference()
        ref_pnp = shim.load_pnp()
        with tempfile.TemporaryDirectory() as td:
            wp = os.path.join(td, "synth.weights")
            synth.write_darknet_weights(wp, stream)
            det_model = ref.Darknet(ref.cfg_path, 416)     # dataloader.py:289-301
            det_model.load_weights(wp)
        det_model.net_info["height"] = "416"
        det_model.eval()
        pose_model = ref.FastPose()                        # KPD/src/main_fast_inference.py:26-40
        pose_model.load_state_dict(sd, strict=False)
        pose_model.eval()
        tf = transforms.Compose([transforms.Resize((416, 416), interpolation=3), transforms.ToTensor()])  # dataloader.py:94-99

        def run(frames, collect=None):
            n_pose = 0
            with torch.no_grad():
                for fr in frames:
                    img = tf(Image.fromarray(fr)).unsqueeze(0)                                   # ImageLoader.getitem_yolo :162
                    pred = det_model(img, False) if _takes_cuda_flag(det_model) else det_model(img)  # DetectionLoader.update :341
                    dets = ref.dynamic_write_results(pred, 0.01, 80, nms=True, nms_conf=0.6)      # :343
                    if isinstance(dets, int):
                        continue
                    dets = dets.cpu()
                    h, w = fr.shape[:2]
                    boxes = dets[:, 1:5] * torch.tensor([w / 416, h / 416, w / 416, h / 416])    # :350-364
                    scores = dets[:, 5:6]
                    inp = ref.im_to_torch(fr)                                                    # DetectionProcessor.update :452 (frame is RGB already)
                    inps, pt1, pt2 = torch.zeros(1, 3, 320, 256), torch.zeros(1, 2), torch.zeros(1, 2)
                    inps, pt1, pt2 = ref.crop_from_dets(inp, boxes, inps, pt1, pt2)              # :453
                    hm = pose_model(inps).narrow(1, 0, 50)                                       # betapose_evaluate.py:168-172
                    _, pi, ps = ref.getPrediction(hm, pt1, pt2, 320, 256, 80, 64)                # DataWriter.update :704
                    result = ref.pose_nms(boxes, scores, pi, ps)                                 # :707
                    if not result:
                        continue
                    kp_score = np.array(result[0]["kp_score"][:, 0])
                    kp_2d, kp_3d = np.array(result[0]["keypoints"]), np.array(kp)
                    while len(kp_2d) > 50:
                        d = np.argmin(kp_score, axis=0)
                        kp_score, kp_2d, kp_3d = np.delete(kp_score, d), np.delete(kp_2d, d, axis=0), np.delete(kp_3d, d, axis=0)
                    ref_pnp(kp_3d, kp_2d, R.CAM_K)                                               # :726 (cv2.solvePnP ITERATIVE)
                    n_pose += 1
            return n_pose

        run.info = {"kind": "reference", "resize": "torchvision/Pillow", "pnp": "cv2.solvePnP (ITERATIVE, the active call)"}
        return run

    blocks = yolo_cfg.parse_cfg_text(yolo_cfg.default_cfg_text())
    yparams, _ = onets.split_darknet_weights(blocks, stream)

    def run(frames, collect=None):
        n_pose = 0
        with torch.no_grad():
            for fr in frames:
                rec = {"row": -1, "status": 0}
                if collect is not None:
                    collect.append(rec)
                if Image is not None:
                    rs = np.asarray(Image.fromarray(fr).resize((416, 416), Image.BICUBIC))
                    x = torch.from_numpy((rs.astype(np.float32) / np.float32(255)).transpose(2, 0, 1).copy())[None]
                else:
                    x = torch.from_numpy(R.yolo_input_from_frame(fr))[None]
                heads = [h.numpy() for h in onets.darknet_forward(blocks, yparams, x)]
                dets, rows = R.write_results(R.yolo_decode(heads), 0.01)
                if rows is None:
                    continue
                boxes, scores = R.rescale_boxes(dets, fr.shape[1], fr.shape[0])
                pt1, pt2 = R.expand_box(boxes[0], fr.shape[1], fr.shape[0])
                crop = torch.from_numpy(R.crop_box(fr, pt1, pt2))[None]
                hm = onets.fastpose_forward(sd, crop).numpy()
                _, pi, mv, idx, _ = R.get_prediction(hm, pt1[None], pt2[None])
                rec.update(row=int(rows[0]), box=boxes[0], hm_idx=idx[0].astype(np.int64), preds_img=pi[0], maxval=mv[0].reshape(-1))
                ref = R.pose_nms_single(float(scores[0, 0]), pi[0], mv[0])
                if ref is None:
                    continue
                keep = R.select_keypoints(ref[1], 50)
                if cv2 is not None:
                    ok, rvec, tvec, _ = cv2.solvePnPRansac(kp[keep], np.ascontiguousarray(ref[0][keep], dtype=np.float32), R.CAM_K,
                                                          np.zeros((8, 1), np.float32), reprojectionError=12.0)
                    rec["status"] = 1 if ok else -1
                    if ok:
                        rec["R"], rec["t"] = cv2.Rodrigues(rvec)[0], tvec.reshape(3)
                    n_pose += int(bool(ok))
                else:
                    sol = opnp.solve_pnp(kp[keep], ref[0][keep], R.CAM_K, mode=0, thr=12.0, n_hyp=64, seed=0)
                    rec["status"] = 1 if sol["ok"] else -1
                    rec["R"], rec["t"] = sol["R"], sol["t"]
                    n_pose += int(sol["ok"])
        return n_pose

    run.info = {"kind": "port", "resize": "Pillow" if Image else "numpy port", "pnp": "cv2.solvePnPRansac" if cv2 else "numpy port"}
    return run


def _takes_cuda_flag(model) -> bool:
    import inspect

    try:
        return len(inspect.signature(model.forward).parameters) >= 2
    except Exception:
        return False


def time_cpu(frames, threads: int, steps: int, warmup: int, impl: str = "auto", collect=None):
    run = cpu_pipeline_factory(threads, impl)
    for _ in range(warmup):
        run(frames[:1])
    t0 = time.perf_counter()
    for s in range(steps):
        run(frames, collect if s == 0 else None)
    dt = time.perf_counter() - t0
    return steps * len(frames) / dt, dt / steps, run.info


def darknet_c_leg(cores: int):
    """The vendored darknet C detector (the code the north_star says the new engine replaces) on host cores: YOLOv3-416
    forward alone, all threads and one thread.  None when oracle/_ref/darknet was not built (no reference tree at build time)."""
    try:
        from betapose_b200 import synth
        from oracle import darknet_cpu

        if not darknet_cpu.available():
            return None
        fr = synth.synth_frames(3, seed=100)
        a = darknet_cpu.time_forward(fr, cores)
        b = darknet_cpu.time_forward(fr[:2], 1)
        return {"what": "vendored darknet C (train_YOLO/src, GPU=0 AVX=1 OPENMP=1 -Ofast), YOLOv3-416 network_predict alone, per frame",
                "ms_per_frame": a["median_ms"], "threads": a["threads"], "frames_timed": a["frames_timed"],
                "ms_per_frame_1thread": b["median_ms"], "frames_timed_1thread": b["frames_timed"]}
    except Exception as ex:  # a timing column, never fatal
        return {"error": str(ex)[:200]}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from betapose_b200 import synth

    cores = os.cpu_count() or 1
    per_step = 2  # frames per step: a bounded sample of the batch-64 workload (the CPU path is ~0.1-0.3 s / frame)
    frames = synth.synth_frames(per_step, seed=100)
    ips, s_per_step, info = time_cpu(frames, cores, args.steps, max(1, min(args.warmup, 2)))
    what = ("the UNMODIFIED reference modules through oracle/ref_shim.py (Darknet, dynamic_write_results, crop_from_dets, FastPose, "
            "getPrediction, pose_nms, pnp)" if info["kind"] == "reference" else
            "CPU oracle port: torch-CPU fp32 YOLOv3 + FastPose, numpy stages, fp64 RANSAC-EPnP+LM (the reference tree is not on this machine)")
    sample = (f"{per_step} frames/step x {args.steps} steps of the same synthetic 640x480 stream, one frame at a time; "
              f"resize={info['resize']}, pnp={info['pnp']}")
    out = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "obj_01 synthetic 640x480 frames (BASELINE.json configs[2] stream), " + what, "frames_per_step": per_step},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": info["kind"], "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ====================================================================================================== GPU arm
def per_op_profile(eng, B: int, reps: int = 3):
    """CUDA-event duration of every op of both networks (one launch per event pair, whole pass repeated `reps` times).
    Returns list of dict(net, i, desc, ms, flops, bytes) for batch B."""
    import torch

    out = []
    for name, net in (("yolo", eng.yolo[0]), ("kpd", eng.kpd[0])):
        n = net.num_ops
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(reps)]
        for r in range(reps):
            ev[r][0].record()
            for i in range(n):
                net.forward(B, i, i + 1)
                ev[r][i + 1].record()
        torch.cuda.synchronize()
        for i in range(n):
            desc, fl, by = net.op_desc(i)
            ms = float(np.median([ev[r][i].elapsed_time(ev[r][i + 1]) for r in range(reps)]))
            out.append(dict(net=name, i=i, desc=desc, ms=ms, flops=fl * B, bytes=by * B))
    return out


def stage_profile(eng, B: int, reps: int = 5):
    """CUDA-event time of every stage of one step (the engine's own per-stage readout, BetaposeEngine.profile_stages):
    the `--profile` numbers of betapose_evaluate.py:132-136,178-186, but measured on the device."""
    return eng.profile_stages(B, reps=reps)


def _planted(rng, kp3d, n, sigma, n_out, cam):
    fx, fy, cx, cy = cam[0, 0], cam[1, 1], cam[0, 2], cam[1, 2]
    K = len(kp3d)
    Rg, tg, uv = [], [], []
    for _ in range(n):
        rv = rng.standard_normal(3)
        rv *= rng.uniform(0, np.pi) / np.linalg.norm(rv)
        th = np.linalg.norm(rv)
        k = rv / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        Rm = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        t = np.array([rng.choice([-0.1, 0.1]), rng.choice([-0.1, 0.1]), rng.uniform(0.6, 1.2)])
        pc = kp3d @ Rm.T + t
        p = np.stack([fx * pc[:, 0] / pc[:, 2] + cx, fy * pc[:, 1] / pc[:, 2] + cy], 1)
        if sigma > 0:
            p = p + rng.normal(0, sigma, (K, 2))
        if n_out:
            out = rng.choice(K, n_out, replace=False)
            p[out] += rng.normal(0, 60.0, (n_out, 2))
        Rg.append(Rm); tg.append(t); uv.append(p.astype(np.float32))
    return np.array(Rg), np.array(tg), np.array(uv)


def pnp_vs_cv2_grid(eng, per_cell: int = 32):
    """PnP stage (bp_pose_pnp, RANSAC mode) against cv2.solvePnPRansac(reprojectionError=12) -- the call the north_star names,
    utils/utils.py:32-36 -- over SURVEY 8(d)'s KAT grid: key-point noise sigma in {0, 0.5, 1, 2} px x {0, 5, 10} gross
    outliers (60 px), `per_cell` planted poses each.  Per cell: the fraction of frames with max|dR| and max|dt| <= 1e-3 (the
    tolerance of record), how many frames end with a consensus set different from OpenCV's, and the worst |dR|, |dt| among
    the frames whose consensus sets coincide (there the two are the same least-squares problem)."""
    import torch

    from betapose_b200 import stages

    try:
        import cv2
    except Exception:
        return None
    dev = eng.device
    kp3d = eng.kp3d[0].cpu().numpy()
    K = len(kp3d)
    cam = eng.cam_K
    rng = np.random.default_rng(4242)
    cells, worst_same = [], 0.0
    for sigma in (0.0, 0.5, 1.0, 2.0):
        for n_out in (0, 5, 10):
            Rg, tg, uv = _planted(rng, kp3d, per_cell, sigma, n_out, cam)
            pose = stages.pose_pnp(torch.from_numpy(uv).to(dev), None, None, eng.kp3d[0].contiguous(), left_number=K, mode=stages.MODE_RANSAC,
                                   reproj_thr=eng.reproj_thr, n_hyp=eng.n_hyp, seed=eng.seed, flags=stages.PNP_RAW_POINTS)
            torch.cuda.synchronize()
            Re, te = pose["R"].cpu().numpy().reshape(-1, 3, 3), pose["t"].cpu().numpy()
            st, inl = pose["status"].cpu().numpy(), pose["inlier"].cpu().numpy().astype(bool)
            within = same = both = ours_closer = 0
            dR_same = dt_same = dR_all = dt_all = 0.0
            for b in range(per_cell):
                ok, rvec, tvec, ci = cv2.solvePnPRansac(kp3d, uv[b], cam, np.zeros((8, 1), np.float32), reprojectionError=float(eng.reproj_thr))
                if not ok or st[b] != 1:
                    continue
                both += 1
                Rc, tc = cv2.Rodrigues(rvec)[0], tvec.reshape(3)
                m = np.zeros(K, bool)
                m[ci.reshape(-1)] = True
                dR, dt = float(np.abs(Re[b] - Rc).max()), float(np.abs(te[b] - tc).max())
                dR_all, dt_all = max(dR_all, dR), max(dt_all, dt)
                within += int(dR <= 1e-3 and dt <= 1e-3)
                if np.array_equal(m, inl[b]):
                    same += 1
                    dR_same, dt_same = max(dR_same, dR), max(dt_same, dt)
                else:  # different consensus sets: which pose is nearer the planted one?
                    eo = max(np.abs(Re[b] - Rg[b]).max(), np.abs(te[b] - tg[b]).max())
                    ec = max(np.abs(Rc - Rg[b]).max(), np.abs(tc - tg[b]).max())
                    ours_closer += int(eo <= ec)
            worst_same = max(worst_same, dR_same, dt_same)
            cells.append({"sigma_px": sigma, "outliers": n_out, "frames": per_cell, "both_found": both,
                          "frac_within_1e-3": within / max(both, 1), "consensus_set_differs": both - same,
                          "ours_nearer_planted_pose_when_sets_differ": ours_closer,
                          "max_dR_same_set": dR_same, "max_dt_same_set": dt_same, "max_dR": dR_all, "max_dt": dt_all})
    differ = sum(c["consensus_set_differs"] for c in cells)
    return {"oracle": "cv2.solvePnPRansac(reprojectionError=12.0)", "tolerance": 1e-3, "cells": cells,
            "frames_with_different_consensus_set": differ,
            "of_those_ours_nearer_the_planted_pose": sum(c["ours_nearer_planted_pose_when_sets_differ"] for c in cells),
            "why_sets_differ": "OpenCV keeps the inliers of its best 5-point hypothesis (its own RNG); this engine re-classifies against "
                               "the refitted pose until the set is stable (DESIGN.md 3.2, INTEGRATION.md)",
            "all_within_tolerance_where_sets_coincide": bool(worst_same <= 1e-3), "worst_same_set": worst_same,
            "min_frac_within_1e-3": min(c["frac_within_1e-3"] for c in cells)}


def add_vs_ref(eng, B):
    """BASELINE.json's second clause, "ADD(-S)@0.1d vs ref", on synthetic data: B planted poses of the key-point model
    (|rotation| < pi, t = (+-0.1, +-0.1, 0.6..1.2) m), their projections with 1 px noise and 5 gross
    outliers (60 px) as key-points -> the engine's PnP stage (bp_pose_pnp) against (i) the CPU oracle and (ii)
    cv2.solvePnPRansac (the reference's third-party solver) on the same points; ADD over the key-point model scored on
    the GPU by bp_score_poses, d = model diameter.  (The key-points a randomly initialised network produces fit no pose
    at all, so the step's own poses cannot be compared between solvers: which local solution wins is arbitrary.)"""
    import torch

    from betapose_b200 import stages
    from oracle import pnp as opnp
    from oracle import restate as R

    try:
        import cv2
    except Exception:
        cv2 = None
    dev = eng.device
    rng = np.random.default_rng(2024)
    kp3d = eng.kp3d[0].cpu().numpy()
    K = len(kp3d)
    diam = float(np.max(np.linalg.norm(kp3d[:, None] - kp3d[None], axis=2)))
    Rg, tg, uv = _planted(rng, kp3d, B, 1.0, 5, R.CAM_K)
    pose = stages.pose_pnp(torch.from_numpy(uv).to(dev), None, None, eng.kp3d[0].contiguous(), left_number=K, mode=stages.MODE_RANSAC,
                           reproj_thr=eng.reproj_thr, n_hyp=eng.n_hyp, seed=eng.seed, flags=1)  # BP_PNP_RAW_POINTS
    torch.cuda.synchronize()
    Re, te, st = pose["R"].cpu().numpy().reshape(B, 3, 3), pose["t"].cpu().numpy(), pose["status"].cpu().numpy()
    Ro, to, Rc, tc = np.zeros_like(Re), np.zeros_like(te), np.zeros_like(Re), np.zeros_like(te)
    ok_o, ok_c = np.zeros(B, bool), np.zeros(B, bool)
    for b in range(B):
        sol = opnp.solve_pnp(kp3d, uv[b], R.CAM_K, mode=0, thr=eng.reproj_thr, n_hyp=eng.n_hyp, seed=eng.seed)
        ok_o[b] = sol["ok"]
        Ro[b], to[b] = sol["R"], sol["t"]
        if cv2 is not None:
            ok, rvec, tvec, _ = cv2.solvePnPRansac(kp3d, uv[b], R.CAM_K, np.zeros((8, 1), np.float32), reprojectionError=float(eng.reproj_thr))
            ok_c[b] = bool(ok)
            if ok:
                Rc[b], tc[b] = cv2.Rodrigues(rvec)[0], tvec.reshape(3)
    box = torch.zeros((B, 4), dtype=torch.float32, device=dev)
    box[:, 2:] = 1
    model = torch.from_numpy(kp3d).to(dev)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731

    def add_mm(Ra, ta, Rb, tb):
        return stages.score_poses(T(Ra), T(ta), box[:len(Ra)], T(Rb), T(tb), box[:len(Ra)], model)["add"].cpu().numpy() * 1000.0

    thr = 100.0 * diam  # 0.1 d in mm
    ours_gt = add_mm(Re, te, Rg, tg)
    res = {"frames": B, "diameter_mm": diam * 1000.0, "noise_px": 1.0, "outliers": 5, "poses_found": int((st == 1).sum()),
           "ours_add_pass_at_0.1d_vs_planted": float((ours_gt[st == 1] < thr).mean()),
           "ours_median_add_mm_vs_planted": float(np.median(ours_gt[st == 1]))}
    m = (st == 1) & ok_o
    a = add_mm(Re[m], te[m], Ro[m], to[m])
    og = add_mm(Ro, to, Rg, tg)
    res["vs_oracle"] = {"frames": int(m.sum()), "add_pass_at_0.1d": float((a < thr).mean()), "max_add_mm": float(a.max()),
                        "pass_decisions_agree": float(((ours_gt < thr) == (og < thr))[m].mean()),
                        "max_abs_dR": float(np.abs(Re[m] - Ro[m]).max()), "max_abs_dt": float(np.abs(te[m] - to[m]).max())}
    if cv2 is not None:
        m = (st == 1) & ok_c
        a = add_mm(Re[m], te[m], Rc[m], tc[m])
        cg = add_mm(Rc, tc, Rg, tg)
        res["vs_cv2_solvePnPRansac"] = {"frames": int(m.sum()), "add_pass_at_0.1d": float((a < thr).mean()), "max_add_mm": float(a.max()),
                                        "pass_decisions_agree": float(((ours_gt < thr) == (cg < thr))[m].mean()),
                                        "cv2_add_pass_at_0.1d_vs_planted": float((cg[ok_c] < thr).mean()),
                                        "max_abs_dR": float(np.abs(Re[m] - Rc[m]).max()), "max_abs_dt": float(np.abs(te[m] - tc[m]).max())}
    return res


def fp16_vs_fp32_flips(eng, frames, cpu_frames_out):
    """End-to-end decision flips of the fp16 engine against the fp32 CPU port on the SAME frames: winning
    detector row, the 50 heat-map arg-max indices per frame, key-point coordinates, pose status; R / t only where both found
    a pose (random networks' key-points fit no rigid pose, so their poses are ill-posed: see add_vs_ref / pnp_vs_cv2 for R, t)."""
    import torch

    n = len(frames)
    rec = eng.run(frames)
    torch.cuda.synchronize()
    row = eng.row[:n].cpu().numpy()
    box = eng.box[:n].cpu().numpy()
    idx = eng.hm_idx[:n].cpu().numpy()
    kp = eng.preds_img[:n].cpu().numpy()
    same_row = np.array([int(row[i]) == c["row"] for i, c in enumerate(cpu_frames_out)])
    live = np.array(["hm_idx" in c for c in cpu_frames_out])
    m = same_row & live
    out = {"frames": n, "detector_row_agree": float(same_row.mean()), "frames_compared_downstream": int(m.sum())}
    if m.any():
        ii = np.nonzero(m)[0]
        out["box_max_abs_diff_px"] = float(max(np.abs(box[i] - cpu_frames_out[i]["box"]).max() for i in ii))
        agree = np.array([(idx[i] == cpu_frames_out[i]["hm_idx"]) for i in ii])
        out["heatmap_argmax_agree"] = float(agree.mean())
        d = np.array([np.abs(kp[i] - cpu_frames_out[i]["preds_img"]).max(axis=1) for i in ii])
        out["keypoint_median_abs_diff_px"] = float(np.median(d))
        out["keypoint_p99_abs_diff_px"] = float(np.quantile(d, 0.99))
        st_c = np.array([cpu_frames_out[i]["status"] for i in ii])
        st_g = rec["status"][ii]
        out["pose_rejected_agree"] = float(((st_c == 0) == (st_g == 0)).mean())
        out["pose_found_gpu"], out["pose_found_cpu"] = int((st_g == 1).sum()), int((st_c == 1).sum())
    # the key-point network alone: fp32 oracle network on the ENGINE's own crops (no upstream box difference)
    try:
        from betapose_b200 import synth as _synth
        from oracle import nets as onets

        kin = eng.kpd[0].input(eng.B)[:n].float().cpu()[..., :3].permute(0, 3, 1, 2).contiguous()
        with torch.no_grad():
            ref_hm = onets.fastpose_forward(_synth.cached_kpd_state_dict(2000), kin)
        ir = ref_hm.reshape(n, 50, -1).argmax(2).numpy()
        top2 = ref_hm.reshape(n, 50, -1).topk(2, dim=2).values
        out["heatmap_argmax_agree_same_crop"] = float((ir == idx).mean())
        out["fp32_top2_gap_median_over_peak"] = float(((top2[..., 0] - top2[..., 1]) / top2[..., 0].abs().clamp_min(1e-6)).median())
        # Can the fp16 noise move an arg-max the fp32 network is sure of?  Per heat-map: e = max |engine - fp32| over the map;
        # a differing arg-max is EXPLAINED by that noise when the fp32 map's value at the engine's position is within 2 e of
        # its own maximum.  `unexplained` must be 0; and among the heat-maps whose fp32 maximum stands out from the
        # runner-up by more than 2 e ("peaked", what a trained network produces) the two arg-maxes must agree everywhere.
        hm_gpu = eng.kpd[0].tensor(eng.hm_id[0], eng.B)[:n].float().cpu().permute(0, 3, 1, 2).reshape(n, 50, -1)
        ref = ref_hm.reshape(n, 50, -1)
        e = (hm_gpu - ref).abs().amax(dim=2)                                   # [n, 50]
        at_gpu = torch.gather(ref, 2, torch.from_numpy(idx.astype(np.int64))[..., None])[..., 0]
        flip = torch.from_numpy(ir != idx)
        explained = (top2[..., 0] - at_gpu) <= 2 * e
        peaked = (top2[..., 0] - top2[..., 1]) > 2 * e
        out["same_crop"] = {"heatmaps": int(flip.numel()), "argmax_differs": int(flip.sum()),
                            "differs_beyond_fp16_noise": int((flip & ~explained).sum()),
                            "fp16_noise_max_over_peak_median": float((e / top2[..., 0].abs().clamp_min(1e-6)).median()),
                            "peaked_heatmaps": int(peaked.sum()), "peaked_argmax_differs": int((flip & peaked).sum())}
        rel_gap = (top2[..., 0] - top2[..., 1]) / top2[..., 0].abs().clamp_min(1e-6)
        for thr in (0.02, 0.05, 0.10):  # agreement among the heat-maps whose fp32 peak leads the runner-up by >= thr of its height
            sel = rel_gap >= thr
            out["same_crop"]["gap_ge_%d%%" % round(thr * 100)] = {"heatmaps": int(sel.sum()), "argmax_differs": int((flip & sel).sum())}
    except Exception as ex:
        out["heatmap_argmax_agree_same_crop"] = f"error: {str(ex)[:120]}"
    out["note"] = ("end-to-end, a 1 px difference of the detector box changes the integer crop window, and a random-init network's "
                   "heat-maps are nearly flat (top-2 gap above), so the arg-max moves; on identical crops see *_same_crop.  "
                   "random-init networks: the 50 key-points fit no rigid pose, so R, t of the two arms are not comparable (the winning "
                   "local solution is arbitrary); R, t agreement is measured on planted poses in add_vs_ref / pnp_vs_cv2")
    return out


def ours_arm(args):
    import torch
    import torch.distributed as dist

    from betapose_b200 import _lib, synth
    from betapose_b200.engine import PipelinedEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, K, W = args.batch, args.steps, args.warmup
    peaks = _peaks()

    class Gather:
        """One NCCL all-gather of the packed records per step, issued asynchronously: the records are copied into one of
        two send buffers on the compute stream, the collective runs on NCCL's stream, and the compute stream never waits
        for it -- so step i + 1 does not wait for the slowest rank's step i (the host waits for gather i - 2 before it
        reuses a buffer)."""

        def __init__(self, n):
            self.send = [torch.empty((n, _lib.RECORD_BYTES), dtype=torch.uint8, device=dev) for _ in range(2)]
            self.recv = [torch.empty((world * n, _lib.RECORD_BYTES), dtype=torch.uint8, device=dev) for _ in range(2)]
            self.work = [None, None]
            self.k = 0

        def __call__(self, rec):
            if world == 1:
                return
            b = self.k & 1
            if self.work[b] is not None:
                self.work[b].wait()
            self.send[b][:rec.shape[0]].copy_(rec, non_blocking=True)
            self.work[b] = dist.all_gather_into_tensor(self.recv[b], self.send[b], async_op=True)
            self.k += 1

        def drain(self):
            for i in range(2):
                if self.work[i] is not None:
                    self.work[i].wait()
                    self.work[i] = None

    def timed(fn, gather, steps=K, warm=W, pipe=None):
        """W warm-up + K timed steps, barrier + synchronize on both sides, CUDA events, max over ranks.  `pipe`: a
        PipelinedEngine whose lane streams are forked from / joined into the timing stream around the timed steps."""
        for i in range(warm):
            fn(i)
        gather.drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        if pipe is not None:
            pipe.fork()
        for i in range(steps):
            fn(warm + i)
        if pipe is not None:
            pipe.join()
        gather.drain()          # the last collectives are inside the timed region
        e1.record()
        torch.cuda.synchronize()
        own = e0.elapsed_time(e1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t1 = time.time()
        ms = torch.tensor([own], dtype=torch.float64, device=dev)
        lo = ms.clone()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        return float(ms.item()), t0, t1, float(lo.item())

    # ------------------------------------------------------------------ configs[2]: obj_01, batch 64 per GPU
    # two lanes: batch i + 1 is in flight on the GPU while batch i finishes (engine.py: PipelinedEngine)
    pipe = PipelinedEngine(args.lanes, B, synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50))
    eng = pipe.lanes[0]
    n_sets = 4  # distinct frame batches rotated through (4 x 59 MB > L2; activations rewritten every step anyway)
    host = [torch.from_numpy(synth.synth_frames(min(B, 16), seed=100 + 17 * rank + s)) for s in range(n_sets)]
    host = [h.repeat((B + h.shape[0] - 1) // h.shape[0], 1, 1, 1)[:B].contiguous().pin_memory() for h in host]
    dev_sets = [h.to(dev) for h in host]
    gather = Gather(B)

    def step_device(i):
        # device->device copy into the lane's frame buffer (inputs already resident in HBM), the step, the all-gather
        pipe.submit_device(i, dev_sets[i % n_sets], graph=bool(args.graph), after_step=gather)

    def timed_e2e():
        """The public streaming API (PipelinedEngine.run_stream) on pinned HOST batches: every step's host->device copy
        (on a side stream, overlapping earlier steps' compute) and the device->host read of its result records are inside
        the timed region.  The timed region is one complete stream of exactly K batches, from an EMPTY pipeline (after a
        synchronize) until the last record is on the host and the device is idle again: pipeline fill and drain are paid
        inside it (with L batches in flight, timing K yields of a longer stream would count only K - L batches of work)."""
        last = None
        for last in pipe.run_stream((host[i % n_sets] for i in range(W)), graph=bool(args.graph), after_step=gather):
            pass
        gather.drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for last in pipe.run_stream((host[(W + i) % n_sets] for i in range(K)), graph=bool(args.graph), after_step=gather):
            pass
        gather.drain()
        torch.cuda.synchronize()
        dt_ms = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([dt_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), last

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    # Order of the two measurements.  Measured on one box, interleaved (gpurun_out/r03t_*): device-resident loop first: value 7 676 /
    # 7 708, e2e 7 428 / 7 491; end-to-end stream first: e2e 7 391 / 7 418, value 7 456 / 7 450.  The end-to-end figure is the same
    # either way (and equals the device-resident loop run warm, scripts/diag_e2e.py); the device-resident loop gains ~3 % when
    # it is the first sustained load after start-up (SM clock before the power cap settles).  Default: value first, as in
    # every earlier round; --e2e-first 1 gives the other order.
    if args.e2e_first:
        step_device(0)  # graphs captured, buffers allocated (not timed)
        torch.cuda.synchronize()
        ms_e2e, last_records = timed_e2e()
        ms_dev, t0, t1, ms_dev_min = timed(step_device, gather, pipe=pipe)
        clocks = sampler.stop(t0, t1) if sampler else None
    else:
        ms_dev, t0, t1, ms_dev_min = timed(step_device, gather, pipe=pipe)
        clocks = sampler.stop(t0, t1) if sampler else None
        ms_e2e, last_records = timed_e2e()
    value = world * B * K / (ms_dev * 1e-3)
    e2e = world * B * K / (ms_e2e * 1e-3)

    scaling_diag = None
    if world > 1:
        # what N > 1 costs: the collective alone (20 back-to-back all-gathers of one step's records) and how far the ranks
        # drift apart over the timed region (max - min of the ranks' own CUDA-event times)
        g2 = Gather(B)
        rec = eng.records[:B]
        for _ in range(3):
            g2(rec)
        g2.drain()
        torch.cuda.synchronize()
        dist.barrier()
        ta = time.perf_counter()
        for _ in range(20):
            g2(rec)
            g2.drain()
        torch.cuda.synchronize()
        scaling_diag = {"allgather_alone_ms": (time.perf_counter() - ta) * 1e3 / 20, "allgather_bytes_per_rank": B * _lib.RECORD_BYTES,
                        "rank_step_ms_max": ms_dev / K, "rank_step_ms_min": ms_dev_min / K,
                        "allgather": "async on NCCL's stream, double-buffered send; the compute stream never waits for it"}

    # ------------------------------------------------------------------ configs[3] / configs[4] sub-records (all ranks)
    sub = {}
    if not args.no_extra:
        sub["configs3"] = bench_configs3(args, world, rank, dev, timed, Gather)
        sub["configs4"] = bench_configs4(args, world, rank, dev, timed, Gather)

    extra = {}
    roof = cpu = None
    if rank == 0:
        # ---- roofline of the dominant kernel family (conv_umma_kernel: every conv / fc launch of both networks).
        # Duration: both networks' forward passes replayed back to back as one CUDA graph (the way they run inside the
        # step), CUDA events on the launching stream.  It includes the 10 small aux kernels (max-pool, SE pooling /
        # scaling, one pixel-shuffle: ~3 % of the time), so `achieved` is a lower bound for the conv kernel alone.
        # (Timing each op between its own pair of events adds ~5 us of launch gap per op -- 1 ms over 200 ops -- and is
        # only used for the per-op table.)
        prof = per_op_profile(eng, B)
        conv = [p for p in prof if p["desc"].startswith("conv")]
        fl = sum(p["flops"] for p in conv)
        g = torch.cuda.CUDAGraph()
        eng.yolo[0].forward(B)
        eng.kpd[0].forward(B)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            eng.yolo[0].forward(B)
            eng.kpd[0].forward(B)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        t_nets_one = e0.elapsed_time(e1) / reps * 1e-3
        # ... and the way the step runs them: every lane replaying its own networks graph on its own stream, lanes in turn
        t_nets = t_nets_one
        if len(pipe.lanes) > 1:
            gs = [g]
            for ln in pipe.lanes[1:]:
                gl = torch.cuda.CUDAGraph()
                cs = torch.cuda.Stream()
                cs.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(cs):
                    ln.yolo[0].forward(B)
                    ln.kpd[0].forward(B)
                cs.synchronize()
                with torch.cuda.graph(gl, stream=cs):
                    ln.yolo[0].forward(B)
                    ln.kpd[0].forward(B)
                gs.append(gl)
            streams = [st_ if st_ is not None else torch.cuda.current_stream() for st_ in pipe.streams]

            def replay_all(n):
                for i in range(n):
                    with torch.cuda.stream(streams[i % len(gs)]):
                        gs[i % len(gs)].replay()

            replay_all(2 * len(gs))
            torch.cuda.synchronize()
            e0.record()
            pipe.fork()
            replay_all(2 * reps)
            pipe.join()
            e1.record()
            torch.cuda.synchronize()
            t_nets = e0.elapsed_time(e1) / (2 * reps) * 1e-3
        t_aux = sum(p["ms"] for p in prof if not p["desc"].startswith("conv")) * 1e-3
        achieved = fl / t_nets / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "conv_dram_traffic.json")  # written from an ncu capture (scripts/ncu_conv_dram.py)
        if os.path.isfile(tp):
            try:
                tj = json.load(open(tp))
                if int(tj.get("batch", 0)) == B:
                    traffic = float(tj["dram_bytes_per_launch"])
            except Exception:
                traffic = None
        roof = {"bound": "tensor", "kernel": "conv_umma_kernel (all %d conv/fc launches of YOLOv3 + FastPose, batch %d)" % (len(conv), B),
                "achieved": achieved, "peak": peaks["tf_sus"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sus"],
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']}); burst {peaks['tf_burst']}",
                "flops_per_launch": fl / len(conv), "avg_launch_ms": t_nets / len(conv) * 1e3, "launches": len(conv),
                "nets_ms": t_nets * 1e3, "nets_ms_one_lane": t_nets_one * 1e3, "frac_one_lane": fl / t_nets_one / 1e12 / peaks["tf_sus"],
                "how": "conv FLOPs of one step / time per step of both networks replayed as CUDA graphs the way the step runs them "
                       f"({len(pipe.lanes)} lanes in flight, CUDA events around {2 * reps if len(pipe.lanes) > 1 else reps} passes); *_one_lane: a single lane alone",
                "aux_ms_per_op_events": t_aux * 1e3, "traffic": traffic,
                "traffic_source": "profiles/conv_dram_traffic.json <- profiles/r03_counters_b64.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per conv launch)" if traffic else None}
        slow = sorted(prof, key=lambda p: -p["ms"])[:8]
        extra["top_ops"] = [{"op": f"{p['net']}[{p['i']}] {p['desc']}", "ms": round(p["ms"], 4),
                             "tflops": round(p["flops"] / (p["ms"] * 1e-3) / 1e12, 1) if p["flops"] else None,
                             "gbs": round(p["bytes"] / (p["ms"] * 1e-3) / 1e9, 1)} for p in slow]
        if args.dump_ops:
            with open(args.dump_ops, "w") as f:
                json.dump(prof, f, indent=1)
        # ---- per-stage device times of one step (the reference's --profile readout, betapose_evaluate.py:132-136,178-186)
        try:
            extra["stage_ms"] = stage_profile(eng, B)
        except Exception as ex:
            extra["stage_ms"] = {"error": str(ex)[:200]}
        # ---- latency at batch 1 (BASELINE.json configs[1]), CUDA graph replay
        eng.frames[:1].copy_(dev_sets[0][:1])
        for _ in range(3):
            eng.run_device(1, graph=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.run_device(1, graph=True)
        e1.record()
        torch.cuda.synchronize()
        extra["latency_batch1_ms"] = e0.elapsed_time(e1) / 20
        # ---- CPU baseline beside it (the reference's path on this box's host cores, bounded sample)
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            fr = synth.synth_frames(6, seed=100)
            cpu_out = []
            ips, _, info = time_cpu(fr, cores, 2, 1, collect=cpu_out)
            cpu = {"value": ips, "unit": UNIT, "cores": cores, "kind": info["kind"],
                   "sample": f"{len(fr) * 2} frames of the same synthetic stream, one at a time: "
                             + ("unmodified reference modules through oracle/ref_shim.py" if info["kind"] == "reference" else "torch-CPU fp32 nets, numpy decode/crop/heat-map stages")
                             + f", resize={info['resize']}, pnp={info['pnp']}"}
            ips1, _, _ = time_cpu(fr[:2], 1, 1, 1)
            cpu["value_1thread"] = ips1
            cpu["sample_1thread"] = "2 frames, torch.set_num_threads(1)"
            cpu["darknet_c"] = darknet_c_leg(cores)
            # ---- end-to-end fp16-vs-fp32 decision flips on the frames the CPU arm just processed
            if info["kind"] == "port" and cpu_out:
                extra["fp16_vs_fp32"] = fp16_vs_fp32_flips(eng, fr, cpu_out)
            # ---- BASELINE.json's second clause, "ADD(-S)@0.1d vs ref": the engine's poses of one batch against the CPU
            # oracle's poses (and cv2.solvePnPRansac's, the reference's third-party call, where importable) computed from
            # the same key-points, scored on the GPU by bp_score_poses (ADD over the key-point model, d = its diameter)
            extra["add_vs_ref"] = add_vs_ref(eng, B)
            extra["pnp_vs_cv2"] = pnp_vs_cv2_grid(eng)

    if rank == 0:
        st = last_records["status"]
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": "obj_01 synthetic 640x480, batch 64 per GPU, 50 keypoints (BASELINE.json configs[2]); "
                                   "1 detection per frame", "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "cuda_graph": bool(args.graph), "lanes": args.lanes, "pnp": "RANSAC-EPnP 64 hyp + LM (fp64)",
                       "l2": f"{n_sets} frame sets rotated ({n_sets * B * 921600 / 1e6:.0f} MB) and ~8 GB of activations rewritten per step >> 126 MB L2"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": B * 480 * 640 * 3 * world,
                    "d2h_bytes_per_step": B * _lib.RECORD_BYTES * world},
            "gpu_launches": eng.launches_per_step * K,
            "launches_per_step": eng.launches_per_step,
            "flops_per_image": eng.flops_per_image,
            "net_tflops": eng.flops_per_image * value / world / 1e12,
            "poses_in_last_batch": int((st == 1).sum()),
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }
        if scaling_diag:
            out["scaling_diag"] = scaling_diag
        out.update(sub)
        out.update(extra)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def linemod13_kp_models():
    """[13,50,3] key-point models of the 13 LineMod objects from the reference's shipped PLYs
    (1_keypoint_designator/assets/sifts/*.ply x 0.001, committed as tests/golden/kp_models.npz); object 10 ships with 17
    points and is padded to 50 by cycling (model3d.load_kp_model(short="cycle"))."""
    from betapose_b200 import model3d

    g = np.load(os.path.join(ROOT, "tests", "golden", "kp_models.npz"))
    return np.stack([g[f"obj_{o}"] if g[f"obj_{o}"].shape[0] == 50 else model3d.pad_by_cycling(g[f"obj_{o}"], 50) for o in LINEMOD_IDS])


def bench_configs3(args, world, rank, dev, timed, Gather):
    """BASELINE.json configs[3]: all 13 LineMod objects mixed, batch 256 sharded across 8 GPUs = 32 frames per GPU, the
    object of every frame drawn uniformly.  13 detector + key-point network pairs resident per GPU (the
    synthetic object variants of betapose_b200/synth.py), the shipped key-point models, slots run concurrently on side
    streams inside one CUDA graph per slot assignment.  Weak scaling: 32 frames per GPU at every N."""
    import torch

    from betapose_b200 import synth, yolo_cfg
    from betapose_b200.engine import BetaposeEngine

    Bm = 32
    t_build = time.time()
    blocks = yolo_cfg.parse_cfg_text(yolo_cfg.default_cfg_text())
    ys0, ks0 = synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000)
    eng = BetaposeEngine(Bm, [(lambda v=v: synth.variant_yolo_weights(ys0, v, blocks)) for v in range(13)],
                         [(lambda v=v: synth.variant_kpd_state_dict(ks0, v)) for v in range(13)], linemod13_kp_models())
    t_build = time.time() - t_build
    frames = torch.from_numpy(synth.synth_frames(Bm, seed=300 + rank)).to(dev)
    rng = np.random.default_rng(13 + rank)
    assigns = []
    for _ in range(3):  # three different object draws rotated through (one captured graph each)
        slots = np.sort(rng.integers(0, 13, Bm))
        groups, s0 = [], 0
        for s in np.unique(slots):
            c = int((slots == s).sum())
            groups.append((int(s), s0, c))
            s0 += c
        assigns.append(groups)
    gather = Gather(Bm)

    def step(i):
        eng.frames.copy_(frames)
        gather(eng.run_device(Bm, groups=assigns[i % len(assigns)], graph=True))

    K, W = max(6, args.steps // 2), max(3, len(assigns))
    ms, _, _, _ = timed(step, gather, steps=K, warm=W)
    torch.cuda.synchronize()
    st = eng.status.cpu().numpy()
    out = {"workload": "all 13 LineMod objects mixed (object drawn uniformly per frame), 32 frames per GPU (= batch 256 over 8 GPUs), "
                       "13 detector + key-point network pairs resident, shipped key-point models (obj_10 padded by cycling)",
           "value": world * Bm * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "steps": K, "batch_per_gpu": Bm,
           "global_batch": Bm * world, "objects_in_batch": [len(g) for g in assigns], "poses_found_last_batch": int((st == 1).sum()),
           "pnp_failed_last_batch": int((st == -1).sum()), "launches_per_step": eng.launches_per_step, "build_s": round(t_build, 1)}
    del eng
    torch.cuda.empty_cache()
    return out


def bench_configs4(args, world, rank, dev, timed, Gather):
    """BASELINE.json configs[4]: the Occlusion-LineMod path (occlusion_betapose_evaluate.py: DataWriter(cam_K, left_keypoints = 10,
    ...)), batch 128 over 8 GPUs = 16 frames per GPU.  Weak scaling: 16 frames per GPU at every N."""
    import torch

    from betapose_b200 import synth
    from betapose_b200.engine import PipelinedEngine

    Bo = 16
    pipe = PipelinedEngine(args.lanes, Bo, synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50),
                           left_number=10)
    eng = pipe.lanes[0]
    sets = [torch.from_numpy(synth.synth_frames(Bo, seed=400 + rank + 31 * s)).to(dev) for s in range(2)]
    gather = Gather(Bo)

    def step(i):
        pipe.submit_device(i, sets[i & 1], graph=True, after_step=gather)

    K = max(10, args.steps)
    ms, _, _, _ = timed(step, gather, steps=K, warm=4, pipe=pipe)
    torch.cuda.synchronize()
    st = eng.status.cpu().numpy()
    sel = eng.selected.cpu().numpy()
    out = {"workload": "Occlusion-LineMod variant (left_keypoints = 10: PnP on the 10 best-scored key-points), 16 frames per GPU "
                       "(= batch 128 over 8 GPUs)",
           "value": world * Bo * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "steps": K, "batch_per_gpu": Bo,
           "global_batch": Bo * world, "poses_found_last_batch": int((st == 1).sum()), "pnp_failed_last_batch": int((st == -1).sum()),
           "keypoints_selected_per_frame": int(sel[st != 0].sum(1).max()) if (st != 0).any() else 0}
    del eng, pipe
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--lanes", type=int, default=2, help="batches in flight on the GPU (PipelinedEngine lanes)")
    ap.add_argument("--e2e-first", type=int, default=0, help="measure the end-to-end stream before the device-resident loop (see ours_arm)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[3] / configs[4] sub-records (quick kernel iteration)")
    ap.add_argument("--dump-ops", default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 10:
            args.steps = 10
        return reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr",
               "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    ours_arm(args)


if __name__ == "__main__":
    main()
