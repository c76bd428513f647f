"""Command-line driver of the evaluate path, flag-compatible with the reference's
`python3 betapose_evaluate.py --nClasses 50 --indir <dir> --outdir <dir> --sp [--profile]` (README.md:75-85) and
`occlusion_betapose_evaluate.py` (`--left_keypoints`, `--obj_id`).  Frames -> BetaposeEngine -> Betapose-results.json.

    python -m betapose_b200.evaluate --indir frames/ --outdir out/ --yolo_weights 01.weights --kpd_weights seq1_model.pkl \\
           --kp_model obj_01.ply
    torchrun --nproc-per-node 8 -m betapose_b200.evaluate ...        # frames sharded over ranks, one all-gather of records
    python -m betapose_b200.evaluate --synthetic 256 --outdir out/   # synthetic frames + synthetic weights (no assets needed)

Scoring against LineMod ground truth (ADD / 2-D reprojection / IoU) is the stage after this path.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

from . import yolo_cfg, compat, dist as bdist, model3d, stages, synth
from .engine import BetaposeEngine
from .opt import parse_args


def main(argv=None):
    o = parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group(o.backend, device_id=torch.device("cuda", local))

    occlusion = o.mode == "occlusion"
    left = o.left_keypoints if occlusion else o.nClasses  # DataWriter(cam_K, 50, ...) vs (cam_K, args.left_keypoints, ...)
    bench_info = model_vertices = kp_sixd = blocks = None
    has_camera_yml = bool(o.sixd_base) and os.path.exists(os.path.join(o.sixd_base, "camera.yml"))
    if o.sixd_base:
        # the reference's evaluation set-up (betapose_evaluate.py:86-98, 203-206): models, key-point model and ground truth
        # of sequence --obj_id from the benchmark tree; --indir defaults to the sequence's rgb/ folder
        from . import sixd

        # the Occlusion annotations live in sequence 02 whatever the object (occlusion_betapose_evaluate.py:204)
        bench_info = sixd.load_sixd(o.sixd_base, seq=2 if occlusion else o.obj_id, nr_frames=0)
        model_vertices, kp_sixd, _ = sixd.load_models(o.sixd_base, o.obj_id, o.nClasses)
        if not o.inputpath and not o.inputlist:
            o.inputpath = os.path.join(o.sixd_base, "test", f"{2 if occlusion else o.obj_id:02d}", "rgb")
    if o.synthetic:
        names = [f"synthetic_{i:06d}.png" for i in range(o.synthetic)]
        yolo_stream, kpd_sd, kp3d = synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, o.nClasses)
    else:
        if o.inputlist:
            names = [ln.strip() for ln in open(o.inputlist) if ln.strip()]
        else:
            names = sorted(f for f in os.listdir(o.inputpath) if f.lower().endswith((".png", ".jpg", ".jpeg", ".ppm", ".pgm", ".npy")))
        names = [os.path.join(o.inputpath, n) for n in names]
        if not names:
            raise SystemExit("no frames found")
        if o.synthetic_weights:
            yolo_stream, kpd_sd = synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000)
        else:
            from . import weights

            blocks = yolo_cfg.parse_cfg_text(open(o.yolo_cfg).read() if o.yolo_cfg else yolo_cfg.default_cfg_text())

            def pack_yolo():
                _, st = weights.read_darknet_weights(o.yolo_weights)
                return weights.pack_darknet(blocks, st, int(o.inp_dim))  # validates the stream against the cfg

            def pack_kpd():
                return weights.pack_fastpose(torch.load(o.kpd_weights, map_location="cpu"), o.nClasses, o.inputResH, o.inputResW)

            if o.packed_cache:  # folded + packed weights, memory-mapped; re-packed when the source changes
                tag = f"{int(o.inp_dim)}_{o.nClasses}"
                yolo_stream, hit_y = weights.load_or_pack(os.path.join(o.packed_cache, os.path.basename(o.yolo_weights) + f".{tag}.bppw"),
                                                          o.yolo_weights, pack_yolo)
                kpd_sd, hit_k = weights.load_or_pack(os.path.join(o.packed_cache, os.path.basename(o.kpd_weights) + f".{tag}.bppw"),
                                                     o.kpd_weights, pack_kpd)
                if rank == 0 and o.profile:
                    print(f"packed-weight cache: detector {'hit' if hit_y else 'packed'}, key-point net {'hit' if hit_k else 'packed'}")
            else:
                _, yolo_stream = weights.read_darknet_weights(o.yolo_weights)
                weights.check_darknet_stream(blocks, yolo_stream)
                kpd_sd = torch.load(o.kpd_weights, map_location="cpu")
                weights.check_fastpose_state_dict(kpd_sd, o.nClasses)
        if o.kp_model:
            kp3d = model3d.load_kp_model(o.kp_model, o.nClasses)
        elif kp_sixd is not None:
            kp3d = kp_sixd
        elif o.synthetic_weights:
            kp3d = synth.synth_kp_model(1, o.nClasses)
        else:
            raise SystemExit("need --kp_model or --sixd_base for the key-point model")

    n_total = len(names)
    lo, hi = bdist.shard_range(n_total, rank, world)
    B = min(o.batch, max(1, hi - lo))
    mode = stages.MODE_RANSAC if o.pnp_mode == "ransac" else stages.MODE_ALLPTS
    # PnP intrinsics: the reference hard-codes the LineMod K for the solver (betapose_evaluate.py:59 -> DataWriter); a
    # benchmark tree with its own camera.yml is solved and scored with that camera instead (documented divergence)
    cam_K = bench_info.cam if (bench_info is not None and has_camera_yml) else model3d.CAM_K
    eng = BetaposeEngine(B, yolo_stream, kpd_sd, kp3d, reso=int(o.inp_dim), inp_h=o.inputResH, inp_w=o.inputResW, n_kp=o.nClasses,
                         left_number=left, conf=o.confidence, pnp_mode=mode, cfg_blocks=blocks, frame_h=o.frame_h, frame_w=o.frame_w,
                         cam_K=cam_K)
    recs = []
    t0 = time.time()

    ingest = None
    if o.synthetic:
        def batches():  # generation of batch i+1 overlaps the GPU work of batch i (BetaposeEngine.run_stream)
            for b0 in range(lo, hi, B):
                yield synth.synth_frames(min(hi, b0 + B) - b0, seed=b0, h=o.frame_h, w=o.frame_w)
        stream = batches()
    else:
        # the native decoder pool fills pinned batch buffers `ingest_depth` batches ahead of the GPU
        from .ingest import FrameIngest

        ingest = FrameIngest(o.ingest_threads, o.frame_h, o.frame_w)
        stream = ingest.batches(names[lo:hi], B, depth=o.ingest_depth, in_flight=max(1, o.lanes))
    # two batches in flight on the GPU (engine.py: PipelinedEngine) unless --lanes 1; the second lane is a second engine
    # with its own activation buffers, created only when there is more than one batch to run
    runner = eng
    if o.lanes > 1 and (hi - lo) > B:
        from .engine import PipelinedEngine

        runner = PipelinedEngine.from_engines([eng] + [
            BetaposeEngine(B, yolo_stream, kpd_sd, kp3d, reso=int(o.inp_dim), inp_h=o.inputResH, inp_w=o.inputResW, n_kp=o.nClasses,
                           left_number=left, conf=o.confidence, pnp_mode=mode, cfg_blocks=blocks, frame_h=o.frame_h, frame_w=o.frame_w,
                           cam_K=cam_K) for _ in range(o.lanes - 1)])
    for rec in runner.run_stream(stream, graph=True, image_index0=lo):
        recs.append(rec)
    if ingest is not None:
        ingest.close()
    torch.cuda.synchronize()
    dt = time.time() - t0
    mine = np.concatenate(recs) if recs else np.zeros(0, stages.RECORD_DTYPE)
    local_bytes = bdist.records_to_bytes(mine).cuda()  # [0, RECORD_BYTES] on a rank whose shard is empty (n_total < world)
    allrec = stages.records_to_numpy(bdist.gather_records(local_bytes, n_total))
    if rank == 0:
        results = [compat.result_from_record(allrec[i], names[int(allrec[i]["image_index"])], o.nClasses) for i in range(n_total)]
        out = compat.write_json(results, o.outputpath)
        print(f"{n_total} frames, {len(out)} poses -> {os.path.join(o.outputpath, 'Betapose-results.json')}")
        if bench_info is not None:
            from . import sixd

            # the reference scores with bench_info.cam (betapose_evaluate.py:248-249), which load_sixd fills from
            # <sixd_base>/camera.yml and leaves at identity otherwise; the identity fallback is meaningless for a
            # reprojection in pixels, so without camera.yml the LineMod intrinsics of betapose_evaluate.py:59 are used
            cam = bench_info.cam if has_camera_yml else model3d.CAM_K
            sixd.evaluate_results(results, bench_info, o.obj_id, model_vertices, cam=cam, occlusion=occlusion, left_keypoints=o.left_keypoints)
        if o.profile:
            print(f"rank 0: {hi - lo} frames in {dt:.3f} s ({(hi - lo) / dt:.1f} frames/s incl. host frame generation / decoding)")
            # per-stage readout (betapose_evaluate.py:132-136,178-186 prints det / pose / post wall-clock means): device time of
            # every stage of one batch, CUDA events, on the last batch's frames
            nb = min(B, hi - lo)
            if nb > 0:
                sm = eng.profile_stages(nb)
                print(f"stage times for a batch of {nb} (ms): " + " | ".join(f"{k} {v:.3f}" for k, v in sm.items()))
                print("det time: {dt:.3f} | pose time: {pt:.3f} | post processing: {pn:.3f}   (ms per batch: resize + detector + decode + crop | "
                      "key-point net | heat-map decode + pose-NMS + PnP + pack)".format(
                          dt=sm["resize"] + sm["detector"] + sm["decode_argmax"] + sm["crop"], pt=sm["keypoint_net"],
                          pn=sm["heatmap_decode"] + sm["pose_pnp"] + sm["pack"]))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
