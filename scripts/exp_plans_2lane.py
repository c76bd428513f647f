"""Experiment: which tile-plan families win when TWO batches are in flight (tails are filled by the other lane, so
bytes-per-flop matters more than wave quantisation)?"""
import sys, os, re
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from betapose_b200 import synth
from betapose_b200.engine import PipelinedEngine

ys, ks, kp = synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50)
B = 64
fr = torch.from_numpy(synth.synth_frames(16, seed=1)).cuda().repeat(4, 1, 1, 1)
pipe = PipelinedEngine(2, B, ys, ks, kp)

def run(tag, steps=30, warm=6):
    for e in pipe.lanes:
        e._graphs.clear()
    for i in range(warm):
        pipe.submit_device(i, fr)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pipe.fork()
        for i in range(steps):
            pipe.submit_device(i, fr)
        pipe.join(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    print(f"{tag}: {best:.3f} ms/step {B / best * 1e3:.0f} img/s", flush=True)
    return best

def nets():
    for e in pipe.lanes:
        yield e.yolo[0]
        yield e.kpd[0]

def reset():
    for n in nets():
        for i in range(n.num_ops):
            if n.op_desc(i)[0].startswith("conv"):
                n.set_op_config(i, B, 0, 0, 0)

def apply(rule):
    cnt = 0
    for n in nets():
        for i in range(n.num_ops):
            d = n.op_desc(i)[0]
            if not d.startswith("conv"):
                continue
            m = re.match(r"conv (\d)x\d/(\d) (\d+)->(\d+) @(\d+)x(\d+)", d)
            k, s, cin, cout, P, Q = map(int, m.groups())
            cfg = n.op_config(i, B)  # bn, cg, mt, bk, st
            num_kb = (k * k * cin + cfg[3] - 1) // cfg[3]
            new = rule(dict(k=k, s=s, cin=cin, cout=cout, P=P, Q=Q, bn=cfg[0], cg=cfg[1], mt=cfg[2], bk=cfg[3], num_kb=num_kb, M=B * P * Q, desc=d))
            if new is not None and n.set_op_config(i, B, *new):
                cnt += 1
    return cnt

run("planner")
run("planner (again)")
for name, rule in [
    ("pairs wherever bn=256 and num_kb >= 8", lambda o: (256, 2, 1) if o["bn"] == 256 and o["cg"] == 1 and o["bk"] == 64 and o["num_kb"] >= 8 and "f32" not in o["desc"] and "ps2" not in o["desc"] and "up2" not in o["desc"] else None),
    ("pairs wherever bn=256 and num_kb >= 4", lambda o: (256, 2, 1) if o["bn"] == 256 and o["cg"] == 1 and o["bk"] == 64 and o["num_kb"] >= 4 and "f32" not in o["desc"] and "ps2" not in o["desc"] and "up2" not in o["desc"] else None),
    ("pairs also for bn=128 (num_kb >= 4)", lambda o: (o["bn"], 2, 1) if o["bn"] >= 128 and o["cg"] == 1 and o["mt"] == 1 and o["bk"] == 64 and o["num_kb"] >= 4 and "f32" not in o["desc"] and "ps2" not in o["desc"] and "up2" not in o["desc"] else None),
    ("no pairs at all", lambda o: (o["bn"], 1, 1) if o["cg"] == 2 else None),
    ("256-pixel tiles for every bn=128 layer", lambda o: (128, 1, 2) if o["bn"] == 128 and o["cg"] == 1 and o["mt"] == 1 else None),
    ("bn=128 instead of 256 on 1x1 layers with K <= 512", lambda o: (128, 1, 1) if o["bn"] == 256 and o["k"] == 1 and o["cin"] <= 512 and o["cg"] == 1 else None),
]:
    reset()
    c = apply(rule)
    run(f"{name} [{c} ops]")
reset()
run("planner (end)")
