"""TEST / BASELINE INFRASTRUCTURE ONLY -- timing of the reference's vendored darknet C detector on host cores.

`oracle/_ref/darknet` is compiled by oracle/Makefile from 3_6Dpose_estimator/train_YOLO/src as it lies in the reference
tree (CPU-only, AVX + OpenMP, -Ofast).  This module writes the inputs `darknet detector test` needs into a scratch
directory -- the YOLOv3-single cfg (our generated cfg text, which tests/test_oracle_golden.py checks block for block
against the reference's yolo/cfg/yolov3-single.cfg, plus a [net] block with batch=1, as train_YOLO/cfg/yolo-linemod-single.cfg
has for testing), seeded synthetic weights in the .weights layout, 640x480 synthetic frames -- feeds the image paths on
stdin and parses the "Predicted in ... milli-seconds" lines (detector.c:1155: the network_predict call alone).

A timing baseline only: darknet's numerics differ from the PyTorch detector the evaluate path runs (BN epsilon form,
bilinear resize), so nothing here is compared numerically."""
from __future__ import annotations

import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "_ref", "darknet")

NET_BLOCK = """[net]
batch=1
subdivisions=1
width=416
height=416
channels=3
momentum=0.9
decay=0.0005
learning_rate=0.001
max_batches=500200
policy=steps
steps=3000,5000
scales=.1,.1

"""


def available() -> bool:
    return os.path.isfile(BIN) and os.access(BIN, os.X_OK)


def time_forward(frames_u8: np.ndarray, threads: int, timeout_s: float = 300.0) -> dict:
    """frames_u8 [n,480,640,3] RGB; the first frame is the warm-up.  -> dict(ms list, median_ms, threads)."""
    from PIL import Image

    from betapose_b200 import synth, yolo_cfg

    assert available(), "oracle/_ref/darknet is not built (make -C oracle)"
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "test.cfg"), "w").write(NET_BLOCK + yolo_cfg.default_cfg_text())
        open(os.path.join(td, "obj.names"), "w").write("object\n")
        open(os.path.join(td, "obj.data"), "w").write(f"classes=1\nnames={td}/obj.names\n")
        synth.write_darknet_weights(os.path.join(td, "synth.weights"), synth.cached_yolo_weights(1000))
        paths = []
        for i, fr in enumerate(frames_u8):
            p = os.path.join(td, f"f{i}.png")
            Image.fromarray(fr).save(p)
            paths.append(p)
        env = dict(os.environ, OMP_NUM_THREADS=str(int(threads)))
        r = subprocess.run([BIN, "detector", "test", "obj.data", "test.cfg", "synth.weights", "-thresh", "0.01", "-dont_show"],
                           input="\n".join(paths) + "\n", capture_output=True, text=True, cwd=td, env=env, timeout=timeout_s)
        ms = [float(m) for m in re.findall(r"Predicted in ([0-9.]+) milli-seconds", r.stdout)]
        if len(ms) != len(paths):
            raise RuntimeError(f"darknet: {len(ms)} timings for {len(paths)} frames (rc {r.returncode}): {r.stdout[-400:]} {r.stderr[-400:]}")
    timed = ms[1:] if len(ms) > 1 else ms
    return dict(ms=ms, median_ms=float(np.median(timed)), threads=int(threads), frames_timed=len(timed))
