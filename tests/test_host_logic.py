"""CPU: host-side logic and the C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from betapose_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "betapose_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bp_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.bp_version() == 100
    # struct layouts shared across the boundary
    assert C.sizeof(_lib.Record) == 4 + 4 + 16 + 4 + 4 + 600 + 72 + 24
    assert C.sizeof(_lib.ConvSpec) == 12 * 4 + 6 * 8 + 8 + 4 * 8  # ... bn_eps (+pad), packed_w, packed_b, packed_w_elems, packed_b_elems


def test_engine_create_fails_loudly_without_gpu():
    from betapose_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = _lib.lib().bp_engine_create(0, C.byref(h))
    assert rc < 0 and b"no CUDA device" in _lib.lib().bp_last_error()
    with pytest.raises(_lib.BetaposeError):
        _lib.Engine.get(0)
    from betapose_b200 import compat

    with pytest.raises(_lib.BetaposeError):
        compat.getPrediction(torch.zeros(1, 50, 80, 64), torch.zeros(1, 2), torch.ones(1, 2), 320, 256, 80, 64)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "betapose_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("oracle/pnp.py", "").replace("oracle/restate.py", ""), f


def test_darknet_cfg_shapes_and_weight_split():
    from betapose_b200 import net as bnet, yolo_cfg

    blocks = yolo_cfg.parse_cfg_text(yolo_cfg.default_cfg_text())
    kinds = [b["type"] for b in blocks]
    assert (kinds.count("convolutional"), kinds.count("shortcut"), kinds.count("route"), kinds.count("upsample"), kinds.count("yolo")) == (75, 23, 4, 2, 3)
    info = bnet.infer_darknet_shapes(blocks, 416)
    assert [(info[i]["H"], info[i]["C"]) for i in (81, 93, 105)] == [(13, 18), (26, 18), (52, 18)]
    assert info[86]["C"] == 768 and info[98]["C"] == 384
    n_floats = 0
    for i, b in enumerate(blocks):
        if b["type"] == "convolutional":
            cin = 3 if i == 0 else info[i - 1]["C"]
            co, k = int(b["filters"]), int(b["size"])
            n_floats += co * cin * k * k + (4 * co if int(b.get("batch_normalize", 0)) else co)
    assert n_floats == 61_576_342  # 61,523,734 parameters + 52,608 BN running statistics
    stream = np.arange(n_floats, dtype=np.float32)
    params, used = bnet.split_darknet_stream(blocks, stream)
    assert used == n_floats
    assert params[0]["bn_bias"][0] == 0 and params[0]["bn_weight"][0] == 32 and params[0]["weight"].shape == (32, 3, 3, 3)
    with pytest.raises(Exception):
        bnet.split_darknet_stream(blocks, stream[:-5])


def test_opt_surface_matches_reference_defaults():
    from betapose_b200 import opt as O

    o = O.parse_args([])
    assert (o.nClasses, o.inputResH, o.inputResW, o.outputResH, o.outputResW) == (50, 320, 256, 80, 64)
    assert (o.inp_dim, o.confidence, o.nms_thesh, o.detbatch, o.posebatch, o.left_keypoints, o.obj_id) == ("416", 0.01, 0.6, 1, 80, 10, 5)
    assert o.num_classes == 80 and o.outputpath == "examples/res/"
    o = O.parse_args(["--nClasses", "50", "--indir", "a", "--outdir", "b", "--sp", "--profile", "--unknown-training-flag", "3"])
    assert o.inputpath == "a" and o.outputpath == "b" and o.sp and o.profile


def test_write_json_format(tmp_path):
    from betapose_b200 import compat

    rec_dt = np.dtype([("image_index", "<i4"), ("status", "<i4"), ("box", "<f4", 4), ("det_score", "<f4"), ("proposal_score", "<f4"),
                       ("keypoints", "<f4", 150), ("R", "<f8", 9), ("t", "<f8", 3)], align=True)
    rec = np.zeros(3, rec_dt)
    rec["status"] = [1, 0, -1]
    rec["keypoints"][0] = np.arange(150)
    rec["R"][0] = np.arange(9)
    rec["t"][0] = [0.1, 0.2, 0.9]
    rec["proposal_score"][0] = 2.5
    rec["R"][2] = np.arange(9) * 0.5   # status -1: the solver's last estimate (a pose-NMS survivor whose PnP found no consensus)
    rec["t"][2] = [0.0, 0.1, 0.8]
    res = [compat.result_from_record(rec[i], f"/data/rgb/{i:04d}.png") for i in range(3)]
    out = compat.write_json(res, str(tmp_path))
    on_disk = json.load(open(tmp_path / "Betapose-results.json"))
    # a rejected frame emits no entry (pPose_nms.py:296); a survivor whose PnP failed still does (dataloader.py:715-727
    # appends bbox, key-points, cam_R and cam_t for every survivor), so it stays in the IoU / ADD statistics as a miss
    assert on_disk == out and [e["image_id"] for e in out] == ["0000.png", "0002.png"]
    assert out[1]["cam_t"] == [0.0, 0.1, 0.8] and out[1]["cam_R"] == [0.5 * k for k in range(9)] and len(out[1]["keypoints"]) == 150
    e = out[0]
    assert e["image_id"] == "0000.png" and e["cam_R"] == list(map(float, range(9))) and e["cam_t"] == [0.1, 0.2, 0.9]
    assert e["keypoints"][:6] == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0] and len(e["keypoints"]) == 150 and e["score"] == 2.5
    out = compat.write_json(res, str(tmp_path), for_eval=True)
    assert out[0]["image_id"] == 0


def test_shard_ranges_cover_everything():
    from betapose_b200 import dist as D

    for n in (0, 1, 7, 64, 128, 256, 1000):
        for w in (1, 2, 3, 4, 8):
            r = [D.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert D.shard_sizes(256, 8) == [32] * 8 and D.shard_sizes(128, 8) == [16] * 8  # BASELINE.json configs 4 and 5


def test_model3d_ply_roundtrip(tmp_path):
    from betapose_b200 import model3d

    g = np.load(os.path.join(ROOT, "tests", "golden", "kp_models.npz"))
    v = g["obj_1"]
    p = tmp_path / "obj_01.ply"
    with open(p, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nend_header\n" % len(v))
        for row in v * 1000.0:
            f.write("%.9g %.9g %.9g\n" % tuple(row))
    out = model3d.load_kp_model(str(p), 50)
    np.testing.assert_allclose(out, v, rtol=1e-7)
    assert model3d.refine(v, 40).shape == (40, 3)
    from oracle import restate as R

    assert np.array_equal(model3d.refine(v, 44), R.refine_vertices(v, 44))
    q = tmp_path / "obj_10.ply"
    with open(q, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 17\nproperty float x\nproperty float y\nproperty float z\nend_header\n")
        for row in g["obj_10"] * 1000.0:
            f.write("%.9g %.9g %.9g\n" % tuple(row))
    with pytest.raises(ValueError):
        model3d.load_kp_model(str(q), 50)   # the shipped obj-10 model has 17 points
    assert np.array_equal(model3d.CAM_K, R.CAM_K)


def test_sixd_loader_and_score_summary(tmp_path):
    """Host side of the evaluation loop: the SIXD tree loader (utils/sixd.py:60-111 semantics: millimetres -> metres, diameters
    indexed by object id, first ground-truth entry per frame) and the reference's summary arithmetic
    (betapose_evaluate.py:259-266) on hand-made per-frame errors."""
    import yaml

    from betapose_b200 import sixd, stages

    base = tmp_path / "sixd"
    (base / "models").mkdir(parents=True)
    (base / "kpmodels").mkdir()
    (base / "test" / "02").mkdir(parents=True)
    yaml.safe_dump({1: {"diameter": 102.1}, 2: {"diameter": 247.5}}, open(base / "models" / "models_info.yml", "w"))
    yaml.safe_dump({"fx": 500.0, "fy": 501.0, "cx": 320.0, "cy": 240.0}, open(base / "camera.yml", "w"))
    yaml.safe_dump({0: {"cam_K": [1, 0, 2, 0, 3, 4, 0, 0, 1]}, 1: {"cam_K": [1, 0, 2, 0, 3, 4, 0, 0, 1]}}, open(base / "test" / "02" / "info.yml", "w"))
    yaml.safe_dump({0: [{"cam_R_m2c": [1, 0, 0, 0, 1, 0, 0, 0, 1], "cam_t_m2c": [10.0, 20.0, 900.0], "obj_bb": [5, 6, 70, 80], "obj_id": 2}],
                    1: [{"cam_R_m2c": [0, 1, 0, -1, 0, 0, 0, 0, 1], "cam_t_m2c": [0.0, 0.0, 1000.0], "obj_bb": [1, 2, 3, 4], "obj_id": 9},
                        {"cam_R_m2c": [1, 0, 0, 0, 1, 0, 0, 0, 1], "cam_t_m2c": [1.0, 1.0, 1.0], "obj_bb": [0, 0, 1, 1], "obj_id": 2}]},
                   open(base / "test" / "02" / "gt.yml", "w"))
    with open(base / "models" / "obj_02.ply", "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\nend_header\n1 2 3\n4 5 6\n7 8 10\n")
    kp = np.random.default_rng(0).uniform(-50, 50, (50, 3))
    with open(base / "kpmodels" / "obj_02.ply", "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 50\nproperty float x\nproperty float y\nproperty float z\nend_header\n")
        f.write("\n".join("%.5f %.5f %.5f" % tuple(p) for p in kp) + "\n")
    b = sixd.load_sixd(str(base), 2)
    assert b.diameter == [10000.0, 102.1, 247.5] and b.cam[0, 0] == 500.0 and b.cam[1, 2] == 240.0 and len(b.frames) == 2
    oid, pose, bb = b.frames[0].gt[0]
    assert oid == 2 and bb == [5.0, 6.0, 70.0, 80.0] and np.allclose(pose[:3, 3], [0.01, 0.02, 0.9]) and np.array_equal(pose[:3, :3], np.eye(3))
    assert b.frames[1].gt[0][0] == 9 and len(b.frames[1].gt) == 2 and b.frames[1].path.endswith("0001.png")
    assert len(sixd.load_sixd(str(base), None).frames) == 0
    verts, kpm, diam = sixd.load_models(str(base), 2, 50)
    assert diam == 247.5 and np.allclose(verts, np.array([[1, 2, 3], [4, 5, 6], [7, 8, 10]]) * 0.001) and np.allclose(kpm, kp * 0.001, atol=1e-8)
    s = stages.summarize_scores(np.array([0.001, 0.030, 0.005, 0.5]), np.array([1.0, 9.0, 4.9, 100.0]), np.array([0.9, 0.7, 0.51, 0.2]),
                                np.array([1, 1, 1, 0], np.uint8), diameter_mm=100.0)
    assert s["n_scored"] == 3 and s["add_accuracy"] == 2 / 3 and s["proj2d_accuracy"] == 2 / 3 and s["iou_accuracy"] == 0.75
    assert abs(s["mean_add_err_mm"] - 12.0) < 1e-9


def test_weight_format_tooling(tmp_path):
    """darknet .weights round trip (header, stream, per-block split and its inverse, size / truncation checks) and the
    FastPose state_dict schema (the 654 reference keys; missing / mis-shaped / unexpected keys are named)."""
    import torch

    from betapose_b200 import net as bnet, synth, weights, yolo_cfg

    blocks = yolo_cfg.parse_cfg_text(yolo_cfg.default_cfg_text())
    assert weights.darknet_stream_size(blocks) == 61523734 + 2 * sum(int(b["filters"]) for b in blocks if b["type"] == "convolutional" and int(b.get("batch_normalize", 0)))
    stream = synth.cached_yolo_weights(1000)
    weights.check_darknet_stream(blocks, stream)
    p = tmp_path / "x.weights"
    weights.write_darknet_weights(str(p), stream, seen=1234)
    hdr, back = weights.read_darknet_weights(str(p))
    assert hdr.tolist() == [0, 1, 0, 1234] and np.array_equal(back, stream)
    params, used = bnet.split_darknet_stream(blocks, stream)
    assert used == stream.size and np.array_equal(weights.darknet_stream_from_params(blocks, params), stream)
    with pytest.raises(ValueError, match="truncated"):
        weights.check_darknet_stream(blocks, stream[:-5])
    with pytest.raises(ValueError, match="left over"):
        weights.check_darknet_stream(blocks, np.concatenate([stream, np.zeros(3, np.float32)]))
    sd = synth.cached_kpd_state_dict(2000)
    exp = weights.fastpose_expected_shapes(50)
    assert len(exp) + sum(1 for k in sd if k.endswith("num_batches_tracked")) == 654 == len(sd)
    weights.check_fastpose_state_dict(sd)
    bad = dict(sd)
    del bad["preact.layer3.7.bn2.running_var"]
    with pytest.raises(ValueError, match="missing 'preact.layer3.7.bn2.running_var'"):
        weights.check_fastpose_state_dict(bad)
    bad = dict(sd)
    bad["duc1.conv.weight"] = torch.zeros(1024, 256, 3, 3)
    with pytest.raises(ValueError, match="duc1.conv.weight"):
        weights.check_fastpose_state_dict(bad)
    bad = dict(sd)
    bad["module.extra"] = torch.zeros(1)
    with pytest.raises(ValueError, match="unexpected key"):
        weights.check_fastpose_state_dict(bad)


# ---------------------------------------------------------------------------------------------- packed-weight cache (8f-4)
def _fold_reference(w, bias, bn, in_scale=1.0):
    """independent numpy statement of what csrc/net.cu:pack_conv_weights computes, before the re-ordering"""
    cout = w.shape[0]
    scale, shift = np.ones(cout), (np.zeros(cout) if bias is None else bias.astype(np.float64))
    if bn is not None:
        g, be, m, v, eps = bn
        inv = g.astype(np.float64) / np.sqrt(v.astype(np.float64) + np.float64(np.float32(eps)))
        scale, shift = inv, be.astype(np.float64) - m.astype(np.float64) * inv + shift * inv
    wf = (w.astype(np.float64) * scale[:, None, None, None] * in_scale).astype(np.float32).astype(np.float16)
    return wf, shift.astype(np.float32)


def test_pack_conv_weights_layouts_bit_exact():
    """bp_pack_conv_weights (host entry point; the same function bp_net_conv packs with) against a numpy statement of the
    layout: ordinary conv (NHWC im2col K order), pixel-shuffle row permutation, both stem kinds, zero padding."""
    from betapose_b200 import _lib, weights

    rng = np.random.default_rng(0)

    def bn_of(c):
        return (rng.uniform(0.5, 1.5, c).astype(np.float32), rng.normal(0, 0.2, c).astype(np.float32),
                rng.normal(0, 0.2, c).astype(np.float32), rng.uniform(0.5, 1.5, c).astype(np.float32), 1e-5)

    cases = [dict(cout=48, cin=32, k=3, store=_lib.STORE_PLAIN, bn=True, bias=False, in_kind=-1),
             dict(cout=18, cin=64, k=1, store=_lib.STORE_PLAIN, bn=False, bias=True, in_kind=-1),
             dict(cout=64, cin=32, k=3, store=_lib.STORE_PIXSHUF2, bn=True, bias=False, in_kind=-1),
             dict(cout=32, cin=3, k=3, store=_lib.STORE_PLAIN, bn=True, bias=False, in_kind=_lib.IN_RAW255),
             dict(cout=64, cin=3, k=7, store=_lib.STORE_PLAIN, bn=True, bias=False, in_kind=_lib.IN_F16)]
    for c in cases:
        w = rng.normal(0, 0.1, (c["cout"], c["cin"], c["k"], c["k"])).astype(np.float32)
        bias = rng.normal(0, 0.5, c["cout"]).astype(np.float32) if c["bias"] else None
        bn = bn_of(c["cout"]) if c["bn"] else None
        rec = weights.PackRecorder(64, 64, c["in_kind"] if c["in_kind"] >= 0 else _lib.IN_F16)
        src = 0 if c["in_kind"] >= 0 else 5
        rec.conv(src, w, bias=bias, bn=bn, store=c["store"])
        e = rec.entries[0]
        stem = c["in_kind"] >= 0
        wf, shift = _fold_reference(w, bias, bn, 1.0 / 255.0 if c["in_kind"] == _lib.IN_RAW255 else 1.0)
        cout, cin, k = c["cout"], c["cin"], c["k"]
        cv = (32 if k * 8 <= 32 else 64) if stem else 0
        K = k * cv if stem else k * k * cin
        wpitch, cout_pad = (K + 7) // 8 * 8, (cout + 255) // 256 * 256
        want_w = np.zeros((cout_pad, wpitch), np.float16)
        want_b = np.zeros(cout_pad, np.float32)
        rows = np.arange(cout) if c["store"] != _lib.STORE_PIXSHUF2 else (np.arange(cout) % 4) * (cout // 4) + np.arange(cout) // 4
        if stem:
            blk = np.zeros((cout, k, cv // 8, 8), np.float16)
            blk[:, :, :k, :cin] = wf.transpose(0, 2, 3, 1)
            want_w[rows, :K] = blk.reshape(cout, K)
        else:
            want_w[rows, :K] = wf.transpose(0, 2, 3, 1).reshape(cout, K)
        want_b[rows] = shift
        assert e["w"].size == want_w.size and e["b"].size == cout_pad
        assert np.array_equal(e["w"].view(np.uint16), want_w.reshape(-1).view(np.uint16)), c
        assert np.array_equal(e["b"], want_b), c
        # and back: the folded fp32 tensors (fp16-rounded), PyTorch layout
        uw, ub = weights.unpack_conv(e)
        scale_back = 255.0 if c["in_kind"] == _lib.IN_RAW255 else 1.0
        assert np.array_equal(uw, wf.astype(np.float32) * np.float32(scale_back)) and np.array_equal(ub, shift)
    # geometry the kernels do not support is refused here too
    with pytest.raises(_lib.BetaposeError):
        weights.PackRecorder(64, 64, _lib.IN_F16).conv(3, np.zeros((8, 20, 1, 1), np.float32))


def test_packed_weights_file_roundtrip_and_replay(tmp_path):
    """pack a small darknet + a FastPose-shaped state_dict through the real builders, save / mmap-load, and replay: the
    builders driven by shape-only placeholders consume exactly the packed entries, in order; mismatches are named."""
    from betapose_b200 import _lib, net as bnet, weights, yolo_cfg

    blocks = yolo_cfg.parse_cfg_text(yolo_cfg.default_cfg_text())
    rng = np.random.default_rng(1)
    stream = rng.normal(0, 0.05, weights.darknet_stream_size(blocks)).astype(np.float32)
    # BN variances must be positive (the per-block arrays are views into the stream)
    params, _ = bnet.split_darknet_stream(blocks, stream)
    for p in params:
        if p and "bn_var" in p:
            p["bn_var"][:] = np.abs(p["bn_var"]) + 0.5
    pw = weights.pack_darknet(blocks, stream, 416, meta=dict(note="test"))
    assert pw.kind == "darknet" and len(pw) == 75 and pw.meta["reso"] == 416 and pw.entries[0]["in_kind"] == _lib.IN_RAW255
    assert all(e["in_kind"] == -1 for e in pw.entries[1:])
    path = str(tmp_path / "yolo.bppw")
    pw.save(path)
    back = weights.PackedWeights.load(path)
    assert back.kind == "darknet" and back.meta["note"] == "test" and len(back) == 75
    for a, b in zip(pw, back):
        assert a["shape"] == b["shape"] and a["store"] == b["store"] and a["in_kind"] == b["in_kind"]
        assert np.array_equal(a["w"].view(np.uint16), np.asarray(b["w"]).view(np.uint16)) and np.array_equal(a["b"], b["b"])
        assert b["w"].ctypes.data % 64 == 0 and b["b"].ctypes.data % 64 == 0

    class Replay(weights.PackRecorder):  # what Net.conv does with `packed`, without CUDA behind it
        def __init__(self, packed, *a):
            super().__init__(*a)
            self.packed, self.taken = iter(packed), 0

        def conv(self, src, weight, bias=None, bn=None, stride=1, pad=0, act=0, res=-1, res_mode=0, dst=-1, dst_coff=0, store=0,
                 out_f32=False):
            bnet.conv_spec(src, weight, bias, bn, stride, pad, act, res, res_mode, dst, dst_coff, store, out_f32, shapes_only=True)
            bnet.take_packed(self.packed, weight, store, self.in_kind if int(src) == 0 else -1)
            self.taken += 1
            return self._new()

    r = Replay(back, 416, 416, _lib.IN_RAW255)
    bnet.build_darknet(r, blocks, weights.darknet_placeholder_params(blocks))
    bnet.packed_exhausted(r.packed)
    assert r.taken == 75
    # a cache packed for another network is refused with the conv named
    other = [b.copy() for b in blocks]
    other[2] = dict(other[2], filters="48")
    with pytest.raises(_lib.BetaposeError, match="do not match the network"):
        bnet.build_darknet(Replay(back, 416, 416, _lib.IN_RAW255), other, weights.darknet_placeholder_params(other))
    with pytest.raises(_lib.BetaposeError, match="end before"):
        bnet.build_darknet(Replay(back.entries[:10], 416, 416, _lib.IN_RAW255), blocks, weights.darknet_placeholder_params(blocks))
    # FastPose: 115 convolutions (SE Linear layers run as 1x1 convs), head sliced to n_maps
    sd = {k: (np.abs(rng.normal(0, 0.05, sh)) + 0.3).astype(np.float32) for k, sh in weights.fastpose_expected_shapes(60).items()}
    pk = weights.pack_fastpose(sd, 50)
    assert len(pk) == 115 and pk.entries[0]["in_kind"] == _lib.IN_F16 and pk.entries[-1]["shape"] == (50, 128, 3, 3)
    assert sum(e["store"] == _lib.STORE_PIXSHUF2 for e in pk) == 2
    r = Replay(pk, 320, 256, _lib.IN_F16)
    bnet.build_fastpose(r, weights.fastpose_placeholder_state_dict(50), 50)
    bnet.packed_exhausted(r.packed)
    # load_or_pack: second call hits the cache; touching the source invalidates it
    src = str(tmp_path / "a.weights")
    weights.write_darknet_weights(src, stream)
    calls = []

    def pack():
        calls.append(1)
        return weights.pack_darknet(blocks, weights.read_darknet_weights(src)[1])

    c = str(tmp_path / "cache" / "a.bppw")
    p1, hit1 = weights.load_or_pack(c, src, pack)
    p2, hit2 = weights.load_or_pack(c, src, pack)
    assert (hit1, hit2, len(calls)) == (False, True, 1)
    assert all(np.array_equal(a["w"].view(np.uint16), np.asarray(b["w"]).view(np.uint16)) for a, b in zip(p1, p2))
    import os
    os.utime(src, ns=(1, 1))
    _, hit3 = weights.load_or_pack(c, src, pack)
    assert not hit3 and len(calls) == 2
    open(str(tmp_path / "junk.bppw"), "wb").write(b"not a cache")
    with pytest.raises(ValueError):
        weights.PackedWeights.load(str(tmp_path / "junk.bppw"))


def test_scoring_pair_selection_standard_and_occlusion():
    """Which (ground truth, estimate) pairs get scored: betapose_evaluate.py:216-240 looks at a frame's first ground-truth
    entry only; occlusion_betapose_evaluate.py:216-236 at every entry of the wanted object."""
    from betapose_b200 import sixd

    b = sixd.Benchmark()
    eye = np.identity(4)

    def pose(z):
        p = eye.copy()
        p[2, 3] = z
        return p

    gts = [[(5, pose(0.5), [10, 20, 30, 40]), (2, pose(0.6), [1, 2, 3, 4])],      # wanted object first
           [(2, pose(0.7), [1, 2, 3, 4]), (5, pose(0.8), [50, 60, 10, 10])],      # wanted object second
           [(5, pose(0.9), [0, 0, 5, 5]), (5, pose(1.0), [7, 7, 5, 5])],          # two instances
           [(5, pose(1.1), [0, 0, 5, 5])]]                                        # no pose estimated
    for i, g in enumerate(gts):
        fr = sixd.Frame(i, f"{i:04d}.png", np.identity(3))
        fr.gt = list(g)
        b.frames.append(fr)
    res = [{"bbox": np.array([1.0, 2.0, 3.0, 4.0])}]
    final = [{"imgname": f"/x/{i:04d}.png", "result": res if i != 3 else [], "cam_R": np.identity(3), "cam_t": np.full((3, 1), float(i))}
             for i in range(4)]
    Rg, tg, bg, Re, te, be = sixd.collect_scoring_pairs(final, b, 5, occlusion=False)
    assert [t[2] for t in tg] == [0.5, 0.9] and bg[0] == [10, 20, 40, 60] and [t[0] for t in te] == [0.0, 2.0]
    Rg, tg, bg, Re, te, be = sixd.collect_scoring_pairs(final, b, 5, occlusion=True)
    assert [t[2] for t in tg] == [0.5, 0.8, 0.9, 1.0] and bg[1] == [50, 60, 60, 70] and [t[0] for t in te] == [0.0, 1.0, 2.0, 2.0]
    assert len(be) == 4 and all(np.array_equal(x, [1, 2, 3, 4]) for x in be)
    assert sixd.collect_scoring_pairs(final, b, 9, occlusion=True)[0] == []
