"""BASELINE.json configs[3] shape on one GPU: a batch that mixes several LineMod objects (own detector + key-point net per
object), slots run one after the other vs. concurrently on side streams.  python scripts/bench_mixed.py [n_slots] [batch]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/scripts/", 1)[0])
from betapose_b200 import synth  # noqa: E402
from betapose_b200.engine import BetaposeEngine  # noqa: E402

n_slots = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ys = [synth.cached_yolo_weights(1000 + s) for s in range(n_slots)]
ks = [synth.cached_kpd_state_dict(2000 + s) for s in range(n_slots)]
kp = np.stack([synth.synth_kp_model(1 + s, 50) for s in range(n_slots)])
frames = torch.from_numpy(synth.synth_frames(min(B, 16), seed=5)).repeat((B + 15) // 16, 1, 1, 1)[:B].contiguous().cuda()
slots = np.arange(B) % n_slots
out = {}
for conc in (False, True):
    eng = BetaposeEngine(B, ys, ks, kp, concurrent_slots=conc)
    order = np.argsort(slots, kind="stable")
    groups, s0 = [], 0
    for s in range(n_slots):
        c = int((slots == s).sum())
        groups.append((s, s0, c))
        s0 += c
    eng.frames.copy_(frames[torch.from_numpy(order).cuda()])
    eng.model_idx.copy_(torch.from_numpy(np.sort(slots).astype(np.int32)))
    for _ in range(3):
        eng.run_device(B, groups, graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        rec = eng.run_device(B, groups, graph=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out[conc] = (ms, bytes(rec.cpu().numpy().tobytes()))
    print(f"{n_slots} objects, batch {B}, concurrent_slots={conc}: {ms:.3f} ms/step = {B / ms * 1e3:.0f} images/s", flush=True)
    del eng
    torch.cuda.empty_cache()
print("records identical:", out[False][1] == out[True][1])
