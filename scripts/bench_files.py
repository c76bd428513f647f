"""Frames on disk -> pose records, end to end: PNG files -> native decoder pool -> pinned ring -> H2D on a side stream ->
the whole a1..a12 step -> records on the host (SURVEY.md 8(f) item 2 in front of the benchmarked path).

    python scripts/bench_files.py [--frames 1024] [--batch 64] [--threads 0] [--depth 2]

Writes 128 distinct 640x480 camera-like PNGs (see scripts/bench_ingest.py) and hard-links them up to `--frames` files,
runs one untimed pass over 128 files (graph capture, page cache), then times one pass over all files with the host clock
(the unit of work starts on disk, so the host clock is the honest one) and prints one JSON line.  Needs a B200."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--depth", type=int, default=2)
    a = ap.parse_args()
    from bench_ingest import make_frames

    from betapose_b200 import synth
    from betapose_b200.engine import BetaposeEngine
    from betapose_b200.ingest import FrameIngest

    eng = BetaposeEngine(a.batch, synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50))
    with tempfile.TemporaryDirectory() as d:
        base = make_frames(128, d)
        paths = list(base)
        while len(paths) < a.frames:
            src = base[len(paths) % 128]
            dst = os.path.join(d, f"{len(paths):05d}.png")
            os.link(src, dst)
            paths.append(dst)
        paths = paths[: a.frames]
        with FrameIngest(a.threads) as ing:
            n = sum(len(r) for r in eng.run_stream(ing.batches(paths[:128], a.batch, depth=a.depth), graph=True))
            torch.cuda.synchronize()
            assert n == min(128, len(paths))
            t0 = time.perf_counter()
            recs = list(eng.run_stream(ing.batches(paths, a.batch, depth=a.depth), graph=True))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            n_threads = ing.n_threads
        recs = np.concatenate(recs)
        assert len(recs) == len(paths) and np.array_equal(recs["image_index"], np.arange(len(paths)))
        # the same 128 images repeat: their records must repeat bit for bit (decode + engine are deterministic)
        first = recs[:128]
        for k in range(128, len(recs) - 127, 128):
            for f in ("status", "box", "R", "t"):
                assert np.array_equal(recs[k: k + 128][f], first[f]), f
        print(json.dumps({"metric": "images_per_sec_from_png_files", "value": round(len(paths) / dt, 1), "frames": len(paths),
                          "batch": a.batch, "decoder_threads": n_threads, "ingest_depth": a.depth, "seconds": round(dt, 3),
                          "poses": int((recs["status"] == 1).sum()), "host_cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
