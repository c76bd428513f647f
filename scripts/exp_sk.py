"""Experiment: stream-K per-op timing (selected ops, isolated)"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from betapose_b200 import synth, tune
from betapose_b200.engine import BetaposeEngine
B = 64
eng = BetaposeEngine(B, synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50))
eng.frames.copy_(torch.from_numpy(synth.synth_frames(8, seed=1)).cuda().repeat(8, 1, 1, 1))
eng.run_device(B)
torch.cuda.synchronize()
for netname, ops in (("kpd", [59, 60, 61]), ("yolo", [13, 28, 64, 66])):
    net = eng.kpd[0] if netname == "kpd" else eng.yolo[0]
    for i in ops:
        print(netname, i, net.op_desc(i)[0], f"{tune.time_op(net, B, i, 40) * 1e3:.1f} us", flush=True)
