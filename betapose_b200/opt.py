"""Option surface of the evaluate path: the flags of the reference's argparse singleton that matter for inference
(3_6Dpose_estimator/opt.py:4-150), same names and defaults.  Unlike the reference this module does NOT parse
sys.argv at import; call `parse_args(argv)` (the CLI in betapose_b200/evaluate.py does) or use `opt` = defaults.
Training-only flags of the reference parser are accepted and ignored so existing command lines keep working.
"""
from __future__ import annotations

import argparse


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Betapose evaluate path (B200-native engine)")
    # ---- flags read on the evaluate path (opt.py line numbers of the reference in comments)
    p.add_argument("--left_keypoints", type=int, default=10)            # :13  occlusion variant: key-points kept for PnP
    p.add_argument("--obj_id", type=int, default=5)                     # :19
    p.add_argument("--sp", default=False, action="store_true")          # :25  single process (always true here)
    p.add_argument("--profile", default=False, action="store_true")     # :27
    p.add_argument("--nClasses", type=int, default=50)                  # :39
    p.add_argument("--fast_inference", default=True, type=bool)         # :45
    p.add_argument("--inputResH", type=int, default=320)                # :80
    p.add_argument("--inputResW", type=int, default=256)                # :82
    p.add_argument("--outputResH", type=int, default=80)                # :84
    p.add_argument("--outputResW", type=int, default=64)                # :86
    p.add_argument("--indir", dest="inputpath", default="")             # :114
    p.add_argument("--list", dest="inputlist", default="")              # :116
    p.add_argument("--mode", dest="mode", default="normal")             # :118
    p.add_argument("--outdir", dest="outputpath", default="examples/res/")  # :120
    p.add_argument("--inp_dim", dest="inp_dim", type=str, default="416")    # :122
    p.add_argument("--conf", dest="confidence", type=float, default=0.01)   # :124
    p.add_argument("--nms", dest="nms_thesh", type=float, default=0.6)      # :126 (unused: NMS is hard-coded off, util.py:181)
    p.add_argument("--save_img", default=False, action="store_true")    # :128
    p.add_argument("--vis", default=False, action="store_true")         # :130
    p.add_argument("--format", type=str)                                # :132
    p.add_argument("--detbatch", type=int, default=1)                   # :134
    p.add_argument("--posebatch", type=int, default=80)                 # :136
    # ---- distributed flags the reference declares but never reads (opt.py:103-109); here they are live
    p.add_argument("--dist", dest="dist", type=int, default=1)
    p.add_argument("--backend", dest="backend", type=str, default="nccl")
    p.add_argument("--port", dest="port")
    # ---- engine-specific additions
    p.add_argument("--batch", type=int, default=64, help="frames per engine step")
    p.add_argument("--pnp_mode", type=str, default="ransac", choices=["ransac", "allpoints"],
                   help="ransac = cv2.solvePnPRansac(12 px) semantics (utils/utils.py:32-36); allpoints = EPnP on all points + LM")
    p.add_argument("--yolo_weights", type=str, default="models/yolo/01.weights")
    p.add_argument("--yolo_cfg", type=str, default="")
    p.add_argument("--kpd_weights", type=str, default="")
    p.add_argument("--kp_model", type=str, default="")
    p.add_argument("--synthetic", type=int, default=0, help="evaluate N synthetic frames with synthetic weights")
    p.add_argument("--synthetic_weights", default=False, action="store_true",
                   help="real frames, synthetic (seeded random) network weights: plumbing tests without checkpoints")
    p.add_argument("--sixd_base", type=str, default="",
                   help="SIXD / LineMod benchmark root (the reference hard-codes it, betapose_evaluate.py:91): when given, "
                        "frames, models and ground truth come from it and the ADD / 2-D / IoU summary is printed")
    p.add_argument("--frame_h", type=int, default=480, help="frame height (the LineMod sequences are 640x480)")
    p.add_argument("--frame_w", type=int, default=640)
    p.add_argument("--packed_cache", type=str, default="",
                   help="directory for packed-weight files (BN folded, fp16, kernel order): packed once, memory-mapped afterwards")
    p.add_argument("--ingest_threads", type=int, default=0, help="frame decoder threads (0 = one per hardware thread)")
    p.add_argument("--ingest_depth", type=int, default=2, help="batches decoded ahead of the GPU")
    p.add_argument("--lanes", type=int, default=2, help="batches in flight on the GPU (engine.py: PipelinedEngine); 1 = one at a time")
    return p


def parse_args(argv=None):
    o, _unknown = build_parser().parse_known_args(argv)
    o.num_classes = 80  # opt.py:150 (write_results slices [:, 5:85] of a 6-column tensor => effectively one class)
    return o


opt = parse_args([])
