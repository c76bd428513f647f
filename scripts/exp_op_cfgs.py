"""Experiment: time selected ops of the key-point net at batch 64 under every tile configuration (isolated, predecessor re-run)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BP_NO_TUNE"] = "1"
from betapose_b200 import synth, tune
from betapose_b200.engine import BetaposeEngine

B = 64
eng = BetaposeEngine(B, synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50))
eng.frames.copy_(torch.from_numpy(synth.synth_frames(8, seed=1)).cuda().repeat(8, 1, 1, 1))
eng.run_device(B)
torch.cuda.synchronize()
for netname, ops in (("kpd", [58, 59, 60, 61]), ("yolo", [27, 41, 64])):
    net = eng.kpd[0] if netname == "kpd" else eng.yolo[0]
    for i in ops:
        print(netname, i, net.op_desc(i)[0])
        net.set_op_config(i, B, 0, 0, 0)
        print("   planner", net.op_config(i, B), f"{tune.time_op(net, B, i, 30) * 1e3:.1f} us")
        for cand in tune.CANDIDATES:
            if net.set_op_config(i, B, *cand):
                print("   ", cand, net.op_config(i, B), f"{tune.time_op(net, B, i, 30) * 1e3:.1f} us")
        net.set_op_config(i, B, 0, 0, 0)
