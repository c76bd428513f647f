"""Diagnostic: how much LM work do the engine's own (random-network) key-points cause?  Dumps them for offline use too."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from betapose_b200 import synth
from betapose_b200.engine import BetaposeEngine
from oracle import pnp as opnp, restate as R

ys, ks, kp = synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50)
e = BetaposeEngine(64, ys, ks, kp)
fr = synth.synth_frames(16, seed=100)
rec = e.run(np.concatenate([fr] * 4))
torch.cuda.synchronize()
pi, mv, ds = e.preds_img.cpu().numpy(), e.maxval.cpu().numpy(), e.det_score.cpu().numpy()
np.savez_compressed("gpurun_out/engine_keypoints_b64.npz", preds_img=pi, maxval=mv, det_score=ds, status=rec["status"], kp3d=kp,
                    inlier=e.inlier.cpu().numpy())
cnt = {"solve": 0, "refits": 0, "lstsq": 0}
orig_solve, orig_refine = np.linalg.solve, opnp.refine_lm
def solve(*a, **k):
    cnt["solve"] += 1
    return orig_solve(*a, **k)
def refine(*a, **k):
    cnt["refits"] += 1
    return orig_refine(*a, **k)
np.linalg.solve, opnp.refine_lm = solve, refine
for b in range(16):
    ref = R.pose_nms_single(ds[b], pi[b], mv[b])
    if ref is None:
        print(b, "rejected"); continue
    cnt["solve"] = cnt["refits"] = 0
    sol = opnp.solve_pnp(kp, ref[0], R.CAM_K, mode=0, n_hyp=64, seed=0)
    print(b, "ok", sol["ok"], "inliers", int(sol["inliers"].sum()), "refits", cnt["refits"], "solves", cnt["solve"], "gpu status", int(rec["status"][b]),
          "gpu inliers", int(e.inlier[b].sum()), flush=True)
