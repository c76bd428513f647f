"""One eager step of the batch-64 path (a1..a12, one lane) bracketed by cudaProfilerStart/Stop, for ncu
(`--profile-from-start off`): every kernel of the step exactly once, after warm-up.
    ncu --profile-from-start off --metrics ... python scripts/ncu_step.py [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from betapose_b200 import synth
from betapose_b200.engine import BetaposeEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = BetaposeEngine(B, synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50))
fr = torch.from_numpy(synth.synth_frames(min(B, 16), seed=100)).cuda()
eng.frames.copy_(fr.repeat((B + fr.shape[0] - 1) // fr.shape[0], 1, 1, 1)[:B])
for _ in range(3):
    eng.run_device(B)
torch.cuda.synchronize()
sel = os.environ.get("NCU_OPS")  # e.g. "kpd:60:63": profile only ops [60, 63) of the key-point net (inputs left by the warm-up)
torch.cuda.profiler.start()
if sel:
    name, a, b = sel.split(":")
    (eng.kpd[0] if name == "kpd" else eng.yolo[0]).forward(B, int(a), int(b))
else:
    eng.run_device(B)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per step:", eng.launches_per_step)
