"""Experiment: what do the non-network stages still cost with two batches in flight?"""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from betapose_b200 import synth
from betapose_b200.engine import PipelinedEngine

ys, ks, kp = synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50)
B = 64
fr = torch.from_numpy(synth.synth_frames(16, seed=1)).cuda().repeat(4, 1, 1, 1)

def run(pipe, steps=30, warm=6, tag=""):
    for i in range(warm):
        pipe.submit_device(i, fr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pipe.fork()
    for i in range(steps):
        pipe.submit_device(i, fr)
    pipe.join(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{tag}: {ms:.3f} ms/step {B / ms * 1e3:.0f} img/s", flush=True)

for lanes in (1, 2):
    pipe = PipelinedEngine(lanes, B, ys, ks, kp)
    run(pipe, tag=f"lanes {lanes} full step")
    for e in pipe.lanes:
        e.pnp_flags = 2  # BP_PNP_NMS_ONLY: pose-NMS + selection, no PnP
        e._graphs.clear()
    run(pipe, tag=f"lanes {lanes} without PnP")
    # networks only
    gs = []
    for e, s in zip(pipe.lanes, pipe.streams):
        g = torch.cuda.CUDAGraph()
        cs = torch.cuda.Stream()
        with torch.cuda.stream(cs):
            e.yolo[0].forward(B); e.kpd[0].forward(B)
        cs.synchronize()
        with torch.cuda.graph(g, stream=cs):
            e.yolo[0].forward(B); e.kpd[0].forward(B)
        gs.append(g)
    streams = [s if s is not None else torch.cuda.current_stream() for s in pipe.streams]
    for _ in range(4):
        for g, s in zip(gs, streams):
            with torch.cuda.stream(s):
                g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 30
    for i in range(n):
        with torch.cuda.stream(streams[i % lanes]):
            gs[i % lanes].replay()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / n
    print(f"lanes {lanes} networks only: {ms:.3f} ms/step {B / ms * 1e3:.0f} img/s", flush=True)
    del pipe, gs
    torch.cuda.empty_cache()
