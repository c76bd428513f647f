"""Diagnostic: FastPose / YOLO at batch 64 vs the same images at batch 2 (different tile plans) and vs the fp32 oracle."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from betapose_b200 import synth, _lib, net as bnet, stages, yolo_cfg
from oracle import nets as onets

sd = synth.cached_kpd_state_dict(2000)
B = 64
fr = torch.from_numpy(synth.synth_frames(B, seed=65)).cuda()
rng = np.random.default_rng(3)
x1, y1 = rng.uniform(0, 300, B), rng.uniform(0, 200, B)
box = torch.from_numpy(np.stack([x1, y1, x1 + rng.uniform(80, 300, B), y1 + rng.uniform(80, 260, B)], 1).astype(np.float32)).cuda()
crop = stages.crop_resize(fr, box, torch.arange(B, dtype=torch.int32, device="cuda"), want_f32=True)
n = bnet.Net(B, 320, 256, _lib.IN_F16)
hm_id = bnet.build_fastpose(n, sd, 50)
n.input(B).copy_(stages.net_input_pixels(crop["net"]))
n.forward(B)
torch.cuda.synchronize()
got64 = n.tensor(hm_id, B).permute(0, 3, 1, 2).contiguous().cpu()
# per-op comparison: run batch 2 on images (0,1), (62,63) through a second net and compare every op output tensor
n2 = bnet.Net(2, 320, 256, _lib.IN_F16)
hm2 = bnet.build_fastpose(n2, sd, 50)
for pair in ((0, 1), (62, 63), (20, 21)):
    n2.input(2).copy_(stages.net_input_pixels(crop["net"])[list(pair)])
    n2.forward(2)
    torch.cuda.synchronize()
    g2 = n2.tensor(hm2, 2).permute(0, 3, 1, 2).contiguous().cpu()
    d = (got64[list(pair)] - g2).abs()
    print("pair", pair, "B64 vs B2 heat-maps: max", d.max().item(), "mean", d.mean().item(), "scale", g2.abs().max().item(), flush=True)
    # walk all tensors
    nt = 0
    worst = []
    t = 1
    while True:
        try:
            a = n.tensor(t, B)[list(pair)].float().cpu()
            b = n2.tensor(t, 2).float().cpu()
        except Exception:
            break
        if a.shape == b.shape:
            dd = (a - b).abs().max().item()
            worst.append((dd, t, tuple(a.shape), b.abs().max().item()))
        t += 1
    worst.sort(reverse=True)
    print("  tensors compared", len(worst), "worst:", worst[:6], flush=True)
    first_bad = [w for w in sorted(worst, key=lambda w: w[1]) if w[0] > 1e-2 * max(w[3], 1e-6)]
    print("  first tensors with rel diff > 1e-2:", first_bad[:5], flush=True)
with torch.no_grad():
    ref = onets.fastpose_forward(sd, crop["f32"][[0, 21, 42, 63]].cpu())
g = got64[[0, 21, 42, 63]]
sc = ref.abs().max().item()
print("B64 vs oracle: max", (g - ref).abs().max().item(), "mean", (g - ref).abs().mean().item(), "scale", sc, "rel max", (g - ref).abs().max().item() / sc)
for i, im in enumerate((0, 21, 42, 63)):
    print("  image", im, "max", (g[i] - ref[i]).abs().max().item(), "mean", (g[i] - ref[i]).abs().mean().item())
print([n.op_desc(i)[0] for i in range(n.num_ops)][:12])
