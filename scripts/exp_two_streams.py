"""Experiment: does splitting a 64-frame step into independent sub-batches on concurrent streams raise throughput?
(tails of small layers -- 160 tiles on 148 SMs, SE FCs, decode, PnP -- would overlap with the other sub-batch's kernels)"""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from betapose_b200 import synth
from betapose_b200.engine import BetaposeEngine

def bench(parts, steps=20, warm=3):
    ys, ks, kp = synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50)
    engs = [BetaposeEngine(b, ys, ks, kp) for b in parts]
    fr = torch.from_numpy(synth.synth_frames(16, seed=1)).cuda()
    for e in engs:
        e.frames.copy_(fr.repeat(4, 1, 1, 1)[:e.B])
    streams = [torch.cuda.Stream() for _ in engs]
    main = torch.cuda.current_stream()
    def step():
        for e, s in zip(engs, streams):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                e.run_device(e.B, graph=True)
        for s in streams:
            main.wait_stream(s)
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"parts {parts}: {ms:.3f} ms/step  {sum(parts) / ms * 1e3:.0f} img/s", flush=True)
    del engs
    torch.cuda.empty_cache()

# free-running variant: each stream replays its own graph back to back with no join between steps
def bench_free(parts, steps=20, warm=3):
    ys, ks, kp = synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50)
    engs = [BetaposeEngine(b, ys, ks, kp) for b in parts]
    fr = torch.from_numpy(synth.synth_frames(16, seed=1)).cuda()
    for e in engs:
        e.frames.copy_(fr.repeat(4, 1, 1, 1)[:e.B])
    streams = [torch.cuda.Stream() for _ in engs]
    for e, s in zip(engs, streams):
        with torch.cuda.stream(s):
            for _ in range(warm):
                e.run_device(e.B, graph=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for e, s in zip(engs, streams):
            with torch.cuda.stream(s):
                e.run_device(e.B, graph=True)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    print(f"free-running parts {parts}: {ms:.3f} ms/step  {sum(parts) / ms * 1e3:.0f} img/s", flush=True)
    del engs
    torch.cuda.empty_cache()

import os
for parts in ([64], [64, 64], [64, 64, 64], [64, 64, 64, 64], [128], [128, 128]):
    bench_free(parts)
