"""Host-side throughput of the frame ingest next to the reference's decoders on the same files.

    python scripts/bench_ingest.py [--frames 256] [--threads 0]

Writes `--frames` 640x480 PNGs (LineMod-like: smooth background + textured object + sensor noise, libpng level 3 as cv2
writes them) to a temporary directory, then times
  reference : cv2.imread + PIL.Image.open per frame on one thread (what ImageLoader.getitem_yolo does, dataloader.py:150-179)
  pillow    : one PIL decode per frame, one thread
  native x1 / native xN : FrameIngest with 1 / N pool threads, decoding into one batch buffer
and prints one JSON line.  No GPU involved."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_frames(n, d):
    import cv2

    rng = np.random.default_rng(0)
    y, x = np.mgrid[0:480, 0:640].astype(np.float32)
    paths = []
    for i in range(n):
        bg = 110 + 40 * np.sin(x / (60 + i % 7)) + 30 * np.cos(y / (45 + i % 5))
        im = np.stack([bg, bg * 0.9 + 10, bg * 0.8 + 25], -1)
        cx, cy = 200 + (i * 37) % 240, 150 + (i * 53) % 180
        m = ((x - cx) ** 2 + (y - cy) ** 2) < 70 ** 2
        tex = rng.integers(0, 256, (480, 640, 3)).astype(np.float32)
        im[m] = 0.5 * im[m] + 0.5 * tex[m]
        im += rng.normal(0, 2.5, im.shape)
        p = os.path.join(d, f"{i:04d}.png")
        cv2.imwrite(p, np.clip(im, 0, 255).astype(np.uint8))
        paths.append(p)
    return paths


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    import cv2
    from PIL import Image

    from betapose_b200.ingest import FrameIngest

    with tempfile.TemporaryDirectory() as d:
        paths = make_frames(a.frames, d)
        mb = sum(os.path.getsize(p) for p in paths) / 1e6
        for p in paths:  # page cache warm for every arm
            open(p, "rb").read()
        n_ref = min(64, a.frames)
        t = time.perf_counter()
        for p in paths[:n_ref]:
            cv2.imread(p)
            np.asarray(Image.open(p))
        ref = n_ref / (time.perf_counter() - t)
        t = time.perf_counter()
        for p in paths[:n_ref]:
            np.asarray(Image.open(p).convert("RGB"))
        pil = n_ref / (time.perf_counter() - t)
        out = np.empty((a.frames, 480, 640, 3), np.uint8)
        res = {}
        for name, nt in (("native_x1", 1), ("native_xN", a.threads)):
            with FrameIngest(nt) as g:
                g.decode_files(paths[:8], out)
                t = time.perf_counter()
                g.decode_files(paths, out)
                res[name] = a.frames / (time.perf_counter() - t)
                res[name + "_threads"] = g.n_threads
        chk = np.asarray(Image.open(paths[-1]).convert("RGB"))
        assert np.array_equal(out[-1], chk)
        print(json.dumps({"frames": a.frames, "png_mb_per_frame": round(mb / a.frames, 3), "reference_two_decodes_fps": round(ref, 1),
                          "pillow_one_thread_fps": round(pil, 1), **{k: round(v, 1) for k, v in res.items()}, "host_cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
