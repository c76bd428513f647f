"""Tile plans ranked by SUSTAINED time per launch (tune.time_op_sustained: the op launched back to back for ~0.25 s, i.e. at
the clock the power cap allows for that configuration) instead of burst time.  Writes the table to gpurun_out/ (not into
betapose_b200/tuned/: it is adopted only if both networks together get faster with it, measured A / B / A / B here).
    python scripts/autotune_sustained.py [batch]"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BP_NO_TUNE"] = "1"
from betapose_b200 import synth, tune
from betapose_b200.engine import BetaposeEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ys, ks, kp = synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50)
eng = BetaposeEngine(B, ys, ks, kp)
eng.frames.copy_(torch.from_numpy(synth.synth_frames(min(B, 8), seed=1)).cuda().repeat((B + 7) // 8, 1, 1, 1)[:B])
eng.run_device(B)
torch.cuda.synchronize()


def nets_ms(reps=40):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.yolo[0].forward(B); eng.kpd[0].forward(B)
    s.synchronize()
    with torch.cuda.graph(g, stream=s):
        eng.yolo[0].forward(B); eng.kpd[0].forward(B)
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def apply(tab_y, tab_k):
    for net, tab in ((eng.yolo[0], tab_y), (eng.kpd[0], tab_k)):
        for i in range(net.num_ops):
            if net.op_desc(i)[0].startswith("conv"):
                net.set_op_config(i, B, 0, 0, 0)
        tune.apply(net, B, tab)


before = nets_ms()
t0 = time.time()
timer = lambda n, b, i: tune.time_op_sustained(n, b, i, 0.2)  # noqa: E731
ty = tune.tune_net(eng.yolo[0], B, margin=0.04, log=print, timer=timer)
tk = tune.tune_net(eng.kpd[0], B, margin=0.04, log=print, timer=timer)
print("tuning took %.1f s; %d + %d overrides" % (time.time() - t0, len(ty), len(tk)), flush=True)
res = []
for rep in range(3):
    apply({}, {})
    a = nets_ms()
    apply(ty, tk)
    b = nets_ms()
    res.append((a, b))
    print("planner %.3f ms   sustained-tuned %.3f ms" % (a, b), flush=True)
json.dump({"batch": B, "meta": {"gpu": torch.cuda.get_device_name(0), "ab": res, "first": before}, "yolo": ty, "kpd": tk},
          open(os.path.join("gpurun_out", "r03k_sustained_table_b%d.json" % B), "w"), indent=1, sort_keys=True)
