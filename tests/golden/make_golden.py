"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported in place through oracle/ref_shim.py,
2 in-memory textual patches) and the third-party libraries it delegates to (Pillow, OpenCV) in THIS container.
Run:  python tests/golden/make_golden.py        (needs /root/reference; versions are recorded in each file)

The goldens pin oracle/{restate,nets,pnp}.py (tests/test_oracle_golden.py) everywhere /root/reference is absent,
e.g. on the GPU box.  Inputs are stored next to the outputs so nothing depends on RNG reproducibility, except the
FastPose weights (59.6 M parameters) which are regenerated from a numpy seed by `fastpose_det_weights`.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import restate as R  # noqa: E402

MINI_CFG = """
[convolutional]
batch_normalize=1
filters=8
size=3
stride=1
pad=1
activation=leaky

[convolutional]
batch_normalize=1
filters=16
size=3
stride=2
pad=1
activation=leaky

[convolutional]
batch_normalize=1
filters=8
size=1
stride=1
pad=1
activation=leaky

[convolutional]
batch_normalize=1
filters=16
size=3
stride=1
pad=1
activation=leaky

[shortcut]
from=-3
activation=linear

[convolutional]
batch_normalize=1
filters=32
size=3
stride=2
pad=1
activation=leaky

[convolutional]
filters=18
size=1
stride=1
pad=1
activation=linear

[yolo]
mask = 6,7,8
anchors = 10,13,  16,30,  33,23,  30,61,  62,45,  59,119,  116,90,  156,198,  373,326
classes=1
num=9
jitter=.5
ignore_thresh = .7
truth_thresh = 1
random=1

[route]
layers = -3

[convolutional]
batch_normalize=1
filters=8
size=1
stride=1
pad=1
activation=leaky

[upsample]
stride=2

[route]
layers = -1, 4

[convolutional]
filters=18
size=1
stride=1
pad=1
activation=linear

[yolo]
mask = 3,4,5
anchors = 10,13,  16,30,  33,23,  30,61,  62,45,  59,119,  116,90,  156,198,  373,326
classes=1
num=9
jitter=.5
ignore_thresh = .7
truth_thresh = 1
random=1
"""


def mini_stream(blocks, seed):
    rng = np.random.default_rng(seed)
    chunks, cin, chans = [], 3, []
    for i, b in enumerate(blocks):
        cout = cin
        if b["type"] == "convolutional":
            cout, k = int(b["filters"]), int(b["size"])
            if int(b.get("batch_normalize", 0)):
                chunks += [rng.normal(0, 0.2, cout), rng.uniform(0.6, 1.4, cout), rng.normal(0, 0.2, cout), rng.uniform(0.5, 1.5, cout)]
            else:
                chunks.append(rng.normal(0, 0.5, cout))
            chunks.append(rng.normal(0, np.sqrt(2.0 / (cin * k * k)), cout * cin * k * k))
        elif b["type"] == "route":
            ls = [int(x) for x in b["layers"].split(",")]
            cout = chans[i + ls[0]] if len(ls) == 1 else chans[i + ls[0]] + chans[ls[1]]
        chans.append(cout)
        cin = cout
    return np.concatenate(chunks).astype(np.float32)


def fastpose_det_weights(seed=7):
    """Deterministic (numpy-only) FastPose state_dict: He-scaled convs, BN near identity with damped residual branches."""
    rng = np.random.default_rng(seed)
    sd = {}

    def conv(name, co, ci, k):
        sd[name] = torch.from_numpy(rng.normal(0, np.sqrt(2.0 / (ci * k * k)), (co, ci, k, k)).astype(np.float32))

    def bn(name, c, gain=1.0):
        sd[name + ".weight"] = torch.from_numpy((rng.uniform(0.8, 1.2, c) * gain).astype(np.float32))
        sd[name + ".bias"] = torch.from_numpy(rng.normal(0, 0.05, c).astype(np.float32))
        sd[name + ".running_mean"] = torch.from_numpy(rng.normal(0, 0.05, c).astype(np.float32))
        sd[name + ".running_var"] = torch.from_numpy(rng.uniform(0.8, 1.2, c).astype(np.float32))

    conv("preact.conv1.weight", 64, 3, 7)
    bn("preact.bn1", 64)
    inpl = 64
    for li, (nb, pl) in enumerate(zip((3, 4, 23, 3), (64, 128, 256, 512)), start=1):
        for bi in range(nb):
            p = f"preact.layer{li}.{bi}"
            conv(p + ".conv1.weight", pl, inpl, 1); bn(p + ".bn1", pl)
            conv(p + ".conv2.weight", pl, pl, 3); bn(p + ".bn2", pl)
            conv(p + ".conv3.weight", pl * 4, pl, 1); bn(p + ".bn3", pl * 4, 0.3)
            if bi == 0:
                conv(p + ".downsample.0.weight", pl * 4, inpl, 1); bn(p + ".downsample.1", pl * 4, 0.7)
                c = pl * 4
                for j in (0, 2):
                    sd[f"{p}.se.fc.{j}.weight"] = torch.from_numpy(rng.normal(0, 1 / np.sqrt(c), (c, c)).astype(np.float32))
                    sd[f"{p}.se.fc.{j}.bias"] = torch.from_numpy(rng.normal(0, 0.2, c).astype(np.float32))
            inpl = pl * 4
    conv("duc1.conv.weight", 1024, 512, 3); bn("duc1.bn", 1024)
    conv("duc2.conv.weight", 512, 256, 3); bn("duc2.bn", 512)
    conv("conv_out.weight", 50, 128, 3)
    sd["conv_out.bias"] = torch.from_numpy(rng.normal(0, 0.05, 50).astype(np.float32))
    return sd


def versions():
    import cv2
    import PIL

    return np.array([f"torch {torch.__version__}", f"cv2 {cv2.__version__}", f"Pillow {PIL.__version__}", f"numpy {np.__version__}"])


def make_a12_golden():
    """a12: the reference's result assembly and output format -- `DataWriter.update` (dataloader.py:678-741: getPrediction ->
    pose_nms -> key-point selection -> pnp -> final_result.append) and `write_json` (pPose_nms.py:284-371), both exec'd as
    written.  Frames: two with planted poses, one whose peaks are all below the 0.3 threshold (rejected: result = []), one
    without a detection (boxes None: nothing appended).  Run with left_number = 50 (betapose_evaluate.py:139) and 10
    (occlusion_betapose_evaluate.py:139).  The active solver is cv2.solvePnP ITERATIVE (utils/utils.py:25-29); the poses are
    chosen so that it converges to the least-squares minimum (checked below against EPnP + LM on the same points)."""
    import json
    import math
    import time

    import cv2

    from oracle import pnp as opnp

    ref = ref_shim.load_reference()
    DataWriter = ref_shim.load_datawriter()
    kp = R.load_ply_vertices(os.path.join(ref.sift_dir, "1.ply"))
    rng = np.random.default_rng(1212)
    W, H = 640, 480
    frames = []
    for case in range(4):
        for attempt in range(200):
            rv = rng.standard_normal(3)
            rv *= rng.uniform(0.2, math.pi * 0.9) / np.linalg.norm(rv)
            Rm = cv2.Rodrigues(rv)[0]
            t = np.array([rng.uniform(-0.08, 0.08), rng.uniform(-0.06, 0.06), rng.uniform(0.7, 1.0)])
            pc = kp @ Rm.T + t
            uv = np.stack([R.CAM_K[0, 0] * pc[:, 0] / pc[:, 2] + R.CAM_K[0, 2], R.CAM_K[1, 1] * pc[:, 1] / pc[:, 2] + R.CAM_K[1, 2]], 1)
            box = np.array([uv[:, 0].min() - 4, uv[:, 1].min() - 4, uv[:, 0].max() + 4, uv[:, 1].max() + 4], np.float32)
            pt1, pt2 = R.expand_box(box, W, H)
            # inverse of transformBoxInvert_batch (A.5): heat-map position of an image point
            lenH = max(float(pt2[1] - pt1[1]), float(pt2[0] - pt1[0]) * 320 / 256)
            lenW = lenH * 256 / 320
            padx = max(0.0, (lenW - 1) / 2 - (float(pt2[0]) - 1 - float(pt1[0])) / 2)
            pady = max(0.0, (lenH - 1) / 2 - (float(pt2[1]) - 1 - float(pt1[1])) / 2)
            hx = (uv[:, 0] + 0.3 - float(pt1[0]) + padx) * 80 / lenH - 0.2
            hy = (uv[:, 1] + 0.3 - float(pt1[1]) + pady) * 80 / lenH - 0.2
            ix, iy = np.round(hx).astype(int), np.round(hy).astype(int)
            if ix.min() < 1 or iy.min() < 1 or ix.max() > 62 or iy.max() > 78:
                continue
            hm = np.zeros((1, 50, 80, 64), np.float32)  # sparse maps (the file stays small); peaks and their neighbours below
            peak = rng.uniform(0.35, 0.95, 50).astype(np.float16).astype(np.float32)
            if case == 2:
                peak = (peak * 0.25).astype(np.float16).astype(np.float32)  # every peak < 0.3 -> pose rejected
            for k in range(50):
                hm[0, k, iy[k], ix[k]] = peak[k]
                # quarter-pixel refinement towards the true sub-pixel position
                hm[0, k, iy[k], ix[k] + (1 if hx[k] > ix[k] else -1)] = np.float16(0.2 * peak[k])
                hm[0, k, iy[k] + (1 if hy[k] > iy[k] else -1), ix[k]] = np.float16(0.2 * peak[k])
            ph, pi, mv = ref.getPrediction(torch.from_numpy(hm), torch.from_numpy(pt1[None]), torch.from_numpy(pt2[None]), 320, 256, 80, 64)
            kp2d = pi[0].numpy() - np.float32(0.3)
            if case != 2:
                # the active solver must be in its convergence basin on this frame, for 50 and for the top-10 points
                good = True
                for left in (50, 10):
                    keep = R.select_keypoints(mv[0, :, 0].numpy(), left)
                    Ri, ti = ref_shim.load_pnp()(kp[keep], kp2d[keep], R.CAM_K)
                    sol = opnp.solve_pnp(kp[keep], kp2d[keep], R.CAM_K, mode=opnp.MODE_ALLPTS)
                    good &= bool(sol["ok"]) and np.abs(Ri - sol["R"]).max() < 1e-6 and np.abs(ti.reshape(3) - sol["t"]).max() < 1e-6
                if not good:
                    continue
            frames.append(dict(hm=hm, box=box, score=np.float32(rng.uniform(0.3, 0.99)), pt1=pt1, pt2=pt2, R_true=Rm, t_true=t))
            break
        else:
            raise RuntimeError("no usable pose found")
    names = ["0007.png", "0011.png", "0012.png", "0020.png"]
    out = {"kp3d": kp, "names": np.array(names), "versions": versions()}
    for i, f in enumerate(frames):
        out[f"hm{i}"] = f["hm"].astype(np.float16)
        out[f"box{i}"], out[f"score{i}"], out[f"pt1_{i}"], out[f"pt2_{i}"] = f["box"], f["score"], f["pt1"], f["pt2"]
    for left in (50, 10):
        ref.opt.format = None
        w = DataWriter(R.CAM_K, left, kp).start()
        orig = np.zeros((H, W, 3), np.uint8)
        expect = 0
        for i, f in enumerate(frames):
            if i == 3:  # a frame without a detection first (dataloader.py:691: boxes is None -> nothing is appended)
                w.save(None, None, None, None, None, orig, "0015.png")
            w.save(torch.from_numpy(f["box"][None].copy()), torch.from_numpy(np.array([[f["score"]]], np.float32)), torch.from_numpy(f["hm"].copy()),
                   torch.from_numpy(f["pt1"][None].copy()), torch.from_numpy(f["pt2"][None].copy()), orig, names[i])
            expect += 1
        t0 = time.time()
        while len(w.results()) < expect and time.time() - t0 < 60:
            time.sleep(0.05)
        w.stop()
        res = w.results()
        assert len(res) == expect and [r["imgname"] for r in res] == names
        with tempfile.TemporaryDirectory() as td:
            ref.write_json(res, td)
            text = open(os.path.join(td, "Betapose-results.json")).read()
        js = json.loads(text)
        assert len(js) == 3 and [e["image_id"] for e in js] == ["0007.png", "0011.png", "0020.png"]
        out[f"json_left{left}"] = np.array(text)
        for i, r in enumerate(res):
            out[f"left{left}_n{i}"] = len(r["result"])
            if r["result"]:
                h = r["result"][0]
                out[f"left{left}_kp{i}"] = np.asarray(h["keypoints"], np.float32)
                out[f"left{left}_sc{i}"] = np.asarray(h["kp_score"], np.float32)
                out[f"left{left}_prop{i}"] = np.asarray(h["proposal_score"], np.float32)
                out[f"left{left}_bbox{i}"] = np.asarray(h["bbox"], np.float32)
                out[f"left{left}_R{i}"] = np.asarray(r["cam_R"], np.float64)
                out[f"left{left}_t{i}"] = np.asarray(r["cam_t"], np.float64)
    np.savez_compressed(os.path.join(HERE, "a12_golden.npz"), **out)
    print("a12_golden.npz", os.path.getsize(os.path.join(HERE, "a12_golden.npz")))


def main():
    make_metrics_golden()
    make_a12_golden()
    assert ref_shim.available(), "reference tree not found"
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(2024)
    ver = versions()
    torch.manual_seed(0)
    torch.set_num_threads(1)

    # ---------------------------------------------------------------- a1: Pillow bicubic (dataloader.py:94-99,162)
    from PIL import Image

    small = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    small_out = np.asarray(Image.fromarray(small).resize((29, 31), Image.BICUBIC))  # PIL size = (w, h)
    up = rng.integers(0, 256, (20, 24, 3), dtype=np.uint8)
    up_out = np.asarray(Image.fromarray(up).resize((48, 40), Image.BICUBIC))
    frame_seed = 77
    frame = np.random.default_rng(frame_seed).integers(0, 256, (480, 640, 3), dtype=np.uint8)
    big = np.asarray(Image.fromarray(frame).resize((416, 416), Image.BICUBIC))
    np.savez_compressed(os.path.join(HERE, "resize_golden.npz"), small=small, small_out=small_out, up=up, up_out=up_out,
                        frame_seed=frame_seed, big_sha256=hashlib.sha256(big.tobytes()).hexdigest(), big_rows=big[::52], versions=ver)

    # ---------------------------------------------------------------- a2-a5: reference Darknet on a mini cfg
    with tempfile.TemporaryDirectory() as td:
        cfgp = os.path.join(td, "mini.cfg")
        open(cfgp, "w").write(MINI_CFG)
        blocks = ref.parse_cfg(cfgp)
        stream = mini_stream(blocks, 5)
        wp = os.path.join(td, "mini.weights")
        with open(wp, "wb") as f:
            np.array([0, 1, 0, 0], np.int32).tofile(f)
            stream.tofile(f)
        net = ref.Darknet(cfgp, 64)
        net.load_weights(wp)
        net.eval()
        x = torch.from_numpy(rng.uniform(0, 1, (3, 3, 64, 64)).astype(np.float32))
        x[2] *= 0.3  # a darker image (a constant image would make every cell of a head tie on objectness)
        with torch.no_grad():
            pred = net(x)
            dets = ref.dynamic_write_results(pred.clone(), 0.01, 80, nms=True, nms_conf=0.6)
            none = ref.dynamic_write_results(pred.clone() * 0, 0.6, 80, nms=True, nms_conf=0.6)
        assert isinstance(none, int) and none == 0
        np.savez_compressed(os.path.join(HERE, "darknet_mini_golden.npz"), cfg=np.array(MINI_CFG), stream=stream, x=x.numpy(),
                            pred=pred.numpy(), dets=dets.numpy(), versions=ver)

    # full cfg: parse parity of our generated cfg text is checked in tests/test_yolo_cfg.py against these block dicts
    blocks_full = ref.parse_cfg(ref.cfg_path)
    keys = sorted({k for b in blocks_full for k in b})
    np.savez_compressed(os.path.join(HERE, "yolo_cfg_golden.npz"),
                        blocks=np.array([repr(sorted(b.items())) for b in blocks_full]), n_blocks=len(blocks_full), keys=np.array(keys))

    # ---------------------------------------------------------------- a6: crop_from_dets + cropBox
    fr = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    boxes = np.array([[200.3, 120.7, 330.9, 300.2], [10.2, 5.5, 70.8, 90.1], [500.0, 300.0, 655.0, 500.0], [-20.5, -10.0, 90.0, 200.0],
                      [300.0, 200.0, 301.0, 201.0], [0.0, 0.0, 639.0, 479.0], [100.5, 100.5, 201.6, 400.2]], np.float32)
    inp = ref.im_to_torch(fr)  # RGB frame -> CHW float /255 (the caller does the BGR->RGB cvtColor first)
    n = len(boxes)
    inps, pt1, pt2 = torch.zeros(n, 3, 320, 256), torch.zeros(n, 2), torch.zeros(n, 2)
    inps, pt1, pt2 = ref.crop_from_dets(inp, torch.from_numpy(boxes), inps, pt1, pt2)
    np.savez_compressed(os.path.join(HERE, "crop_golden.npz"), frame=fr, boxes=boxes, pt1=pt1.numpy(), pt2=pt2.numpy(),
                        inps_sub=inps.numpy()[:, :, ::7, ::5], inps_sha256=hashlib.sha256(inps.numpy().tobytes()).hexdigest(),
                        inps_mean=inps.numpy().mean(axis=(2, 3)), versions=ver)

    # ---------------------------------------------------------------- a8: getPrediction
    hm = rng.standard_normal((4, 50, 80, 64)).astype(np.float32)
    hm[0, 0] = -1.0
    hm[0, 1] = 0.0
    hm[1, 2, 0, 0] = 9.0
    hm[1, 3, 79, 63] = 9.0
    hm[2, 4, 40, 30] = 9.0
    hm[2, 4, 41, 31] = 9.0
    hm[3, 5, 10, 10] = 9.0
    hm[3, 5, 10, 11] = hm[3, 5, 10, 9] = 1.0
    gp1 = np.array([[50.5, 40.25], [0, 0], [300.7, 200.1], [400.5, 100.5]], np.float32)
    gp2 = gp1 + np.array([[120.3, 200.9], [5, 5], [150.2, 100.8], [200, 200]], np.float32)
    ph, pi, mv = ref.getPrediction(torch.from_numpy(hm), torch.from_numpy(gp1), torch.from_numpy(gp2), 320, 256, 80, 64)
    np.savez_compressed(os.path.join(HERE, "getpred_golden.npz"), hm=hm.astype(np.float16), pt1=gp1, pt2=gp2, preds_hm=ph.numpy(),
                        preds_img=pi.numpy(), maxval=mv.numpy(), versions=ver)

    # ---------------------------------------------------------------- a9: pose_nms (n = 1 and n = 3)
    cases = []
    for nprop in (1, 1, 3):
        bb = rng.uniform(50, 400, (nprop, 4)).astype(np.float32)
        bb[:, 2:] = bb[:, :2] + rng.uniform(60, 200, (nprop, 2)).astype(np.float32)
        bs = rng.uniform(0.1, 1, (nprop, 1)).astype(np.float32)
        pp = rng.uniform(100, 300, (nprop, 50, 2)).astype(np.float32)
        if nprop == 3:
            pp[1] = pp[0] + rng.normal(0, 1.0, (50, 2)).astype(np.float32)  # near-duplicate pose -> merged
        ps = rng.uniform(0.0, 1.0, (nprop, 50, 1)).astype(np.float32)
        ps[0, 3, 0] = 0.0
        cases.append((bb, bs, pp, ps))
    rej = (cases[0][0].copy(), cases[0][1].copy(), cases[0][2].copy(), (cases[0][3] * 0.2).astype(np.float32))  # max < 0.3 -> rejected
    cases.append(rej)
    out = {}
    for i, (bb, bs, pp, ps) in enumerate(cases):
        res = ref.pose_nms(torch.from_numpy(bb.copy()), torch.from_numpy(bs.copy()), torch.from_numpy(pp.copy()), torch.from_numpy(ps.copy()))
        out[f"c{i}_bb"], out[f"c{i}_bs"], out[f"c{i}_pp"], out[f"c{i}_ps"] = bb, bs, pp, ps
        out[f"c{i}_n"] = len(res)
        for j, r in enumerate(res):
            out[f"c{i}_r{j}_kp"] = np.asarray(r["keypoints"], np.float32)
            out[f"c{i}_r{j}_sc"] = np.asarray(r["kp_score"], np.float32)
            out[f"c{i}_r{j}_prop"] = np.asarray(r["proposal_score"], np.float32)
            out[f"c{i}_r{j}_bbox"] = np.asarray(r["bbox"], np.float32)
    np.savez_compressed(os.path.join(HERE, "pose_nms_golden.npz"), n_cases=len(cases), versions=ver, **out)

    # ---------------------------------------------------------------- a11: OpenCV PnP KATs on the shipped obj_01 key-points
    kp = R.load_ply_vertices(os.path.join(ref.sift_dir, "1.ply"))
    assert kp.shape == (50, 3)
    import math

    P2, Rr, tr, Ri, ti, inl, Rg, tg, sig, nout = [], [], [], [], [], [], [], [], [], []
    for sigma, n_out in ((0.0, 0), (0.1, 0), (0.5, 0), (1.0, 5), (2.0, 10)):
        for _ in range(6):
            rv = rng.standard_normal(3)
            rv *= rng.uniform(0.05, math.pi * 0.95) / np.linalg.norm(rv)
            import cv2

            Rm = cv2.Rodrigues(rv)[0]
            t = np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(0.6, 1.2)])
            pc = kp @ Rm.T + t
            uv = np.stack([R.CAM_K[0, 0] * pc[:, 0] / pc[:, 2] + R.CAM_K[0, 2], R.CAM_K[1, 1] * pc[:, 1] / pc[:, 2] + R.CAM_K[1, 2]], 1)
            uv += rng.normal(0, sigma, uv.shape)
            if n_out:
                o = rng.choice(50, n_out, replace=False)
                uv[o] += rng.normal(0, 60, (n_out, 2))
            uv = uv.astype(np.float32)
            R1, t1, i1 = ref_shim.ref_pnp_ransac(kp, uv, R.CAM_K)
            R2, t2 = ref_shim.ref_pnp(kp, uv, R.CAM_K)
            m = np.zeros(50, bool)
            m[i1.reshape(-1)] = True
            P2.append(uv); Rr.append(R1); tr.append(t1.reshape(3)); Ri.append(R2); ti.append(t2.reshape(3)); inl.append(m)
            Rg.append(Rm); tg.append(t); sig.append(sigma); nout.append(n_out)
    np.savez_compressed(os.path.join(HERE, "pnp_golden.npz"), kp3d=kp, uv=np.array(P2), R_ransac=np.array(Rr), t_ransac=np.array(tr),
                        R_iter=np.array(Ri), t_iter=np.array(ti), inliers=np.array(inl), R_true=np.array(Rg), t_true=np.array(tg),
                        sigma=np.array(sig), n_out=np.array(nout), versions=ver)

    # ---------------------------------------------------------------- key-point models (all 13 shipped PLYs, metres)
    kps = {}
    for f in sorted(os.listdir(ref.sift_dir)):
        if f.endswith(".ply"):
            kps["obj_" + f[:-4]] = R.load_ply_vertices(os.path.join(ref.sift_dir, f)).astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "kp_models.npz"), **kps)

    # ---------------------------------------------------------------- a7: reference FastPose, deterministic weights
    sd = fastpose_det_weights(7)
    model = ref.createModel() if False else ref.FastPose()
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all("num_batches_tracked" in k for k in missing.missing_keys), missing
    model.eval()
    torch.set_num_threads(8)
    xin = torch.from_numpy(rng.uniform(-0.5, 0.5, (1, 3, 320, 256)).astype(np.float32))
    with torch.no_grad():
        out_full = model(xin)
    hm50 = out_full.narrow(1, 0, 50).numpy()
    np.savez_compressed(os.path.join(HERE, "fastpose_golden.npz"), seed=7, x=xin.numpy().astype(np.float32), hm_sub=hm50[:, :, ::4, ::4],
                        hm_argmax=hm50.reshape(1, 50, -1).argmax(2), hm_absmax=np.abs(hm50).max(), out_channels=out_full.shape[1],
                        versions=ver)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


def box_nms_cases():
    """Decoded prediction tensors [B,R,6] with clusters of overlapping boxes around a few objects + low-score clutter."""
    rng = np.random.default_rng(21)
    cases = []
    for B, Rn, n_obj, conf, thr in ((3, 300, 3, 0.01, 0.6), (2, 1200, 6, 0.3, 0.4), (1, 10647, 9, 0.01, 0.6), (2, 64, 0, 0.5, 0.6)):
        pred = np.zeros((B, Rn, 6), np.float32)
        for b in range(B):
            pred[b, :, 0] = rng.uniform(0, 416, Rn)
            pred[b, :, 1] = rng.uniform(0, 416, Rn)
            pred[b, :, 2] = rng.uniform(4, 120, Rn)
            pred[b, :, 3] = rng.uniform(4, 120, Rn)
            pred[b, :, 4] = rng.uniform(0, 0.35, Rn) ** 2          # clutter: mostly below the confidence cut
            pred[b, :, 5] = rng.uniform(0.5, 1.0, Rn)
            for o in range(n_obj):
                c = rng.uniform(60, 356, 2)
                wh = rng.uniform(40, 140, 2)
                idx = rng.choice(Rn, size=min(Rn // 8, 24), replace=False)
                pred[b, idx, 0:2] = c + rng.normal(0, 5, (len(idx), 2))
                pred[b, idx, 2:4] = wh * rng.uniform(0.85, 1.15, (len(idx), 2))
                pred[b, idx, 4] = rng.uniform(0.4, 0.999, len(idx))
        if B == 2 and Rn == 64:
            pred[1, :, 4] = 0.2   # an image without any candidate between images with some
            pred[0, :5, 4] = [0.9, 0.8, 0.7, 0.6, 0.55]
        cases.append((pred, conf, thr))
    return cases


def make_box_nms_golden():
    """write_results of the reference with its disabled IoU-NMS branch re-enabled (ref_shim.load_box_nms) on box_nms_cases()."""
    import torch

    wr = ref_shim.load_box_nms()
    out = {}
    for i, (pred, conf, thr) in enumerate(box_nms_cases()):
        dets = wr(torch.from_numpy(pred.copy()), conf, 80, True, thr)
        dets = np.zeros((0, 8), np.float32) if isinstance(dets, int) else dets.numpy().astype(np.float32)
        out[f"pred{i}"], out[f"conf{i}"], out[f"thr{i}"], out[f"dets{i}"] = pred, np.float32(conf), np.float32(thr), dets
        print("box-nms case", i, pred.shape, "->", dets.shape, np.bincount(dets[:, 0].astype(int), minlength=pred.shape[0]))
    np.savez_compressed(os.path.join(HERE, "box_nms_golden.npz"), n_cases=len(box_nms_cases()), versions=np.array(f"torch {torch.__version__}"), **out)


def make_metrics_golden():
    """utils/metrics.py add_err / projection_error_2d / iou of the UNMODIFIED reference on seeded poses."""
    import cv2

    M = ref_shim.load_metrics()
    rng = np.random.default_rng(12)
    model = rng.uniform(-0.06, 0.06, (777, 3))
    cam = R.CAM_K
    gt, est, boxes_gt, boxes_est, add, proj, iou = [], [], [], [], [], [], []
    for i in range(24):
        rv = rng.standard_normal(3)
        rv *= rng.uniform(0, np.pi) / np.linalg.norm(rv)
        Rg = cv2.Rodrigues(rv)[0]
        tg = np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(0.6, 1.2)])
        dr = rng.standard_normal(3) * [0.0, 0.002, 0.02, 0.2][i % 4]
        Re = cv2.Rodrigues(dr)[0] @ Rg
        te = tg + rng.standard_normal(3) * [0.0, 0.0005, 0.005, 0.05][i % 4]
        G, E = np.eye(4), np.eye(4)
        G[:3, :3], G[:3, 3], E[:3, :3], E[:3, 3] = Rg, tg, Re, te
        bg = np.array([100 + 10 * i, 80, 260 + 10 * i, 300], np.float64)
        be = bg + rng.uniform(-60, 60, 4) * (i % 3)
        if i == 5:
            be = np.array([500, 400, 560, 470], np.float64)  # disjoint
        gt.append(G); est.append(E); boxes_gt.append(bg); boxes_est.append(be)
        add.append(M.add_err(G, E, model))
        proj.append(M.projection_error_2d(G, E, model, cam))
        iou.append(M.iou(list(bg), list(be)))
    np.savez_compressed(os.path.join(HERE, "metrics_golden.npz"), model=model, cam=cam, gt=np.array(gt), est=np.array(est),
                        box_gt=np.array(boxes_gt), box_est=np.array(boxes_est), add=np.array(add), proj=np.array(proj),
                        iou=np.array(iou), versions=np.array(f"numpy {np.__version__}"))
    print("metrics_golden.npz", np.array(add)[:4], np.array(iou)[:6])


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "a12":
    make_a12_golden()
    sys.exit(0)
if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "metrics":
        make_metrics_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "box_nms":
        make_box_nms_golden()
        sys.exit(0)
    main()
