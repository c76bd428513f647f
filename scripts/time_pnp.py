import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import test_pnp_host as T
from betapose_b200 import synth, stages
kp = synth.synth_kp_model(1, 50)
rng = np.random.default_rng(0)
n = 64
def make(sigma, n_out, garbage=False):
    preds = np.zeros((n, 50, 2), np.float32)
    for i in range(n):
        if garbage:
            preds[i] = rng.uniform([200, 100], [400, 350], (50, 2))
        else:
            preds[i] = T.make_case(rng, kp, sigma, n_out)[2] + np.float32(0.3)
    return torch.from_numpy(preds).cuda()
mv = torch.full((n, 50), 0.5).cuda(); det = torch.ones(n).cuda(); kpt = torch.from_numpy(kp).cuda()
for name, p in (('clean s=0.5', make(0.5, 0)), ('outliers s=1,10', make(1.0, 10)), ('garbage', make(0, 0, True))):
    for mode in (0, 1):
        for _ in range(2): out = stages.pose_pnp(p, mv, det, kpt, mode=mode, n_hyp=64, seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out = stages.pose_pnp(p, mv, det, kpt, mode=mode, n_hyp=64, seed=1)
        e1.record(); torch.cuda.synchronize()
        print(f'{name:18s} mode {mode}: {e0.elapsed_time(e1)/5:.3f} ms  status1={int((out["status"]==1).sum())} inl_mean={float(out["inlier"].float().sum(1).mean()):.1f}')
