"""Experiment: lanes that each size their conv grids for a FRACTION of the SMs (true concurrency instead of time-slicing)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from betapose_b200 import synth
from betapose_b200.engine import PipelinedEngine

ys, ks, kp = synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50)
B = 64
fr = torch.from_numpy(synth.synth_frames(16, seed=1)).cuda().repeat(4, 1, 1, 1)

def run(pipe, tag, steps=30, warm=8):
    for e in pipe.lanes:
        e._graphs.clear()
    for i in range(warm):
        pipe.submit_device(i, fr)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pipe.fork()
        for i in range(steps):
            pipe.submit_device(i, fr)
        pipe.join(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    print(f"{tag}: {best:.3f} ms/step {B / best * 1e3:.0f} img/s", flush=True)

for lanes, shares in ((2, (0, 2, 1.5)), (3, (0, 3, 2)), (4, (4, 3))):
    pipe = PipelinedEngine(lanes, B, ys, ks, kp)
    for sh in shares:
        for e in pipe.lanes:
            for net in e.yolo + e.kpd:
                net.set_share(int(sh * B))
        run(pipe, f"lanes {lanes}, conv grids sized for 148/{sh if sh else 1} SMs" + (" (whole GPU, PDL)" if not sh else " (no PDL)"))
    del pipe
    torch.cuda.empty_cache()
