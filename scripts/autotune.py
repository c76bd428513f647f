"""Measure the tile plans of both networks on this GPU and write betapose_b200/tuned/b200_b<batch>.json.
    python scripts/autotune.py [batch ...]        (default: 64)"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BP_NO_TUNE"] = "1"  # start from the planner's own choices
from betapose_b200 import synth, tune
from betapose_b200.engine import BetaposeEngine

batches = [int(a) for a in sys.argv[1:]] or [64]
ys, ks, kp = synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50)
for B in batches:
    eng = BetaposeEngine(B, ys, ks, kp)
    eng.frames.copy_(torch.from_numpy(synth.synth_frames(min(B, 8), seed=1)).cuda().repeat((B + 7) // 8, 1, 1, 1)[:B])
    eng.run_device(B)
    torch.cuda.synchronize()

    def nets_ms():
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            eng.yolo[0].forward(B); eng.kpd[0].forward(B)
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            eng.yolo[0].forward(B); eng.kpd[0].forward(B)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10

    before = nets_ms()
    t0 = time.time()
    ty = tune.tune_net(eng.yolo[0], B, log=print)
    tk = tune.tune_net(eng.kpd[0], B, log=print)
    after = nets_ms()
    p = tune.save(B, ty, tk, {"gpu": torch.cuda.get_device_name(0), "nets_ms_planner": before, "nets_ms_tuned": after,
                              "tuning_s": round(time.time() - t0, 1)})
    print(f"batch {B}: both nets {before:.3f} -> {after:.3f} ms; {len(ty)} + {len(tk)} overrides -> {p}", flush=True)
    del eng
    torch.cuda.empty_cache()
