#!/usr/bin/env python
"""Benchmark of the Betapose per-frame evaluate hot path on B200 (contract: see the task prompt / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, weak scaling)

A step = one batch of B synthetic 640x480 RGB frames per GPU through a1..a12 (resize, YOLOv3-416, decode+arg-max,
crop, FastPose, heat-map decode, pose-NMS, PnP, record packing) and, for N > 1, one NCCL all-gather of the packed
result records.  `value` is timed with the frames already in HBM; `e2e` goes through BetaposeEngine.run with pinned
HOST frames (H2D + D2H inside the timed region).  Rank 0 prints ONE JSON line.
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

METRIC = "linemod_640x480_images_per_sec"
UNIT = "images/s"


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
def cpu_pipeline_factory(threads: int):
    """The oracle port of the whole path on host cores: torch-CPU fp32 networks (oracle/nets.py) + numpy stages
    (oracle/restate.py, oracle/pnp.py).  Returns f(frames_u8) -> number of poses."""
    import torch

    from betapose_b200 import synth, yolo_cfg
    from oracle import nets as onets
    from oracle import pnp as opnp
    from oracle import restate as R

    # the reference's own third-party calls where this box has them (Pillow for the bicubic squash, OpenCV for
    # solvePnPRansac); otherwise the numpy restatements (bit-identical to Pillow; same algorithm as OpenCV but slower)
    try:
        from PIL import Image
    except Exception:
        Image = None
    try:
        import cv2
    except Exception:
        cv2 = None
    run_info = {"resize": "Pillow" if Image else "numpy port", "pnp": "cv2.solvePnPRansac" if cv2 else "numpy port"}
    torch.set_num_threads(threads)
    blocks = yolo_cfg.parse_cfg_text(yolo_cfg.default_cfg_text())
    yparams, _ = onets.split_darknet_weights(blocks, synth.cached_yolo_weights(1000))
    sd = synth.cached_kpd_state_dict(2000)
    kp = synth.synth_kp_model(1, 50)

    def run(frames):
        n_pose = 0
        with torch.no_grad():
            for fr in frames:  # the reference processes one frame at a time (--detbatch 1, one detection per frame)
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
                _, pi, mv, _, _ = R.get_prediction(hm, pt1[None], pt2[None])
                ref = R.pose_nms_single(float(scores[0, 0]), pi[0], mv[0])
                if ref is None:
                    continue
                keep = R.select_keypoints(ref[1], 50)
                if cv2 is not None:
                    ok = cv2.solvePnPRansac(kp[keep], np.ascontiguousarray(ref[0][keep], dtype=np.float32), R.CAM_K,
                                            np.zeros((8, 1), np.float32), reprojectionError=12.0)[0]
                    n_pose += int(bool(ok))
                else:
                    n_pose += int(opnp.solve_pnp(kp[keep], ref[0][keep], R.CAM_K, mode=0, thr=12.0, n_hyp=64, seed=0)["ok"])
        return n_pose

    run.info = run_info
    return run


def time_cpu(frames, threads: int, steps: int, warmup: int):
    run = cpu_pipeline_factory(threads)
    for _ in range(warmup):
        run(frames[:1])
    t0 = time.perf_counter()
    for _ in range(steps):
        run(frames)
    dt = time.perf_counter() - t0
    return steps * len(frames) / dt, dt / steps, run.info


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from betapose_b200 import synth

    cores = os.cpu_count() or 1
    per_step = 2  # frames per step: a bounded sample of the batch-64 workload (the CPU path is ~0.1-0.3 s / frame)
    frames = synth.synth_frames(per_step, seed=100)
    ips, s_per_step, info = time_cpu(frames, cores, args.steps, max(1, min(args.warmup, 2)))
    sample = (f"{per_step} frames/step x {args.steps} steps of the same synthetic 640x480 stream, one frame at a time; "
              f"resize={info['resize']}, pnp={info['pnp']}")
    out = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "obj_01 synthetic 640x480 frames (BASELINE.json configs[2] stream), CPU oracle port: torch-CPU "
                               "fp32 YOLOv3 + FastPose, numpy stages, fp64 RANSAC-EPnP+LM", "frames_per_step": per_step},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
    fx, fy, cx, cy = R.CAM_K[0, 0], R.CAM_K[1, 1], R.CAM_K[0, 2], R.CAM_K[1, 2]
    Rg, tg, uv = [], [], []
    for _ in range(B):
        rv = rng.standard_normal(3)
        rv *= rng.uniform(0, np.pi) / np.linalg.norm(rv)
        th = np.linalg.norm(rv)
        k = rv / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        Rm = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        t = np.array([rng.choice([-0.1, 0.1]), rng.choice([-0.1, 0.1]), rng.uniform(0.6, 1.2)])
        pc = kp3d @ Rm.T + t
        p = np.stack([fx * pc[:, 0] / pc[:, 2] + cx, fy * pc[:, 1] / pc[:, 2] + cy], 1) + rng.normal(0, 1.0, (K, 2))
        out = rng.choice(K, 5, replace=False)
        p[out] += rng.normal(0, 60.0, (5, 2))
        Rg.append(Rm); tg.append(t); uv.append(p.astype(np.float32))
    Rg, tg, uv = np.array(Rg), np.array(tg), np.array(uv)
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
        return stages.score_poses(T(Ra), T(ta), box, T(Rb), T(tb), box, model)["add"].cpu().numpy() * 1000.0

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


def ours_arm(args):
    import torch
    import torch.distributed as dist

    from betapose_b200 import _lib, synth
    from betapose_b200.engine import BetaposeEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, K, W = args.batch, args.steps, args.warmup
    peaks = _peaks()

    eng = BetaposeEngine(B, synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50))
    n_sets = 4  # distinct frame batches rotated through (4 x 59 MB > L2; activations rewritten every step anyway)
    host = [torch.from_numpy(synth.synth_frames(min(B, 16), seed=100 + 17 * rank + s)) for s in range(n_sets)]
    host = [h.repeat((B + h.shape[0] - 1) // h.shape[0], 1, 1, 1)[:B].contiguous().pin_memory() for h in host]
    dev_sets = [h.to(dev) for h in host]
    gathered = torch.empty((world * B, _lib.RECORD_BYTES), dtype=torch.uint8, device=dev) if world > 1 else None

    def step_device(i):
        eng.frames.copy_(dev_sets[i % n_sets])  # device->device: inputs already resident in HBM
        rec = eng.run_device(B, graph=args.graph)
        if world > 1:
            dist.all_gather_into_tensor(gathered, rec)

    def after_step(rec):
        if world > 1:
            dist.all_gather_into_tensor(gathered, rec)

    def timed_e2e():
        """The public streaming API (BetaposeEngine.run_stream) on pinned HOST batches: every step's host->device copy
        (on a side stream, overlapping the previous step's compute) and the device->host read of its result records
        are inside the timed region; the caller holds step i's records before step i+1's are requested."""
        last = None
        stream = eng.run_stream((host[i % n_sets] for i in range(W + K)), graph=bool(args.graph), after_step=after_step)
        for _ in range(W):
            last = next(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            last = next(stream)
        torch.cuda.synchronize()
        dt_ms = (time.perf_counter() - t0) * 1e3
        for _ in stream:
            pass
        ms = torch.tensor([dt_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), last

    def timed(fn):
        for i in range(W):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(K):
            fn(W + i)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ms_dev, t0, t1 = timed(step_device)
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_e2e, last_records = timed_e2e()

    value = world * B * K / (ms_dev * 1e-3)
    e2e = world * B * K / (ms_e2e * 1e-3)

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
        t_nets = e0.elapsed_time(e1) / reps * 1e-3
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
                "nets_ms": t_nets * 1e3, "aux_ms_per_op_events": t_aux * 1e3, "traffic": traffic,
                "traffic_source": "profiles/conv_dram_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per conv launch)" if traffic else None}
        slow = sorted(prof, key=lambda p: -p["ms"])[:8]
        extra["top_ops"] = [{"op": f"{p['net']}[{p['i']}] {p['desc']}", "ms": round(p["ms"], 4),
                             "tflops": round(p["flops"] / (p["ms"] * 1e-3) / 1e12, 1) if p["flops"] else None,
                             "gbs": round(p["bytes"] / (p["ms"] * 1e-3) / 1e9, 1)} for p in slow]
        if args.dump_ops:
            with open(args.dump_ops, "w") as f:
                json.dump(prof, f, indent=1)
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
        # ---- CPU baseline beside it (oracle port on this box's host cores, bounded sample)
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            fr = synth.synth_frames(2, seed=100)
            reps = 6
            ips, _, info = time_cpu(fr, cores, reps, 1)
            cpu = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{len(fr) * reps} frames of the same synthetic stream, one at a time: torch-CPU fp32 nets, "
                             f"resize={info['resize']}, numpy decode/crop/heat-map stages, pnp={info['pnp']}"}

            # ---- BASELINE.json's second clause, "ADD(-S)@0.1d vs ref": the engine's poses of one batch against the CPU
            # oracle's poses (and cv2.solvePnPRansac's, the reference's third-party call, where importable) computed from
            # the same key-points, scored on the GPU by bp_score_poses (ADD over the key-point model, d = its diameter)
            extra["add_vs_ref"] = add_vs_ref(eng, B)

    if rank == 0:
        st = last_records["status"]
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": "obj_01 synthetic 640x480, batch 64 per GPU, 50 keypoints (BASELINE.json configs[2]); "
                                   "1 detection per frame", "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "cuda_graph": bool(args.graph), "pnp": "RANSAC-EPnP 64 hyp + LM (fp64)",
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
        out.update(extra)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
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
