"""End-to-end: BetaposeEngine.run on synthetic frames, every stage checked against the oracle with the engine's own
upstream tensors as the oracle's input (stage-wise "teacher forcing": the only way index decisions can be bit-exact
downstream of fp16 convolutions)."""
import numpy as np
import pytest
import torch

from oracle import pnp as opnp
from oracle import restate as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(yolo_stream, kpd_sd, kp_model):
    from betapose_b200.engine import BetaposeEngine

    e = BetaposeEngine(8, yolo_stream, kpd_sd, kp_model, seed=5)
    e._test_weights = (yolo_stream, kpd_sd, kp_model)
    return e


def test_engine_stagewise_parity(engine, frames8, kp_model):
    e = engine
    rec = e.run(frames8)
    torch.cuda.synchronize()
    B = 8
    # a1
    yin = e.yolo[0].input(B).float().cpu().numpy()  # raw 0..255 pixel values, exact in fp16
    for b in range(B):
        assert np.array_equal(yin[b, :, :, :3], R.pil_resize_bicubic(frames8[b], 416, 416))
    # a3-a5 from the engine's fp32 heads
    heads = [e.yolo[0].tensor(h["tensor"], B).permute(0, 3, 1, 2).contiguous().cpu().numpy() for h in e.heads[0]]
    pred = R.yolo_decode(heads)
    dets, rows = R.write_results(pred, 0.01)
    assert dets is not None and len(rows) == B
    assert np.array_equal(e.row.cpu().numpy(), rows.astype(np.int32))
    boxes, scores = R.rescale_boxes(dets, 640, 480)
    np.testing.assert_allclose(e.box.cpu().numpy(), boxes, rtol=2e-6, atol=3e-5)
    # a6 from the engine's boxes
    box_g = e.box.cpu().numpy()
    kin = e.kpd[0].input(B).float().cpu().numpy()
    for b in range(B):
        pt1, pt2 = R.expand_box(box_g[b], 640, 480)
        assert np.array_equal(e.pt1[b].cpu().numpy(), pt1) and np.array_equal(e.pt2[b].cpu().numpy(), pt2)
        ref = R.crop_box(frames8[b], pt1, pt2)
        np.testing.assert_allclose(kin[b, :, :, :3].transpose(2, 0, 1), ref, atol=5e-4)
    # a8 from the engine's heat-maps
    hm = e.kpd[0].tensor(e.hm_id[0], B).permute(0, 3, 1, 2).contiguous().cpu().numpy()
    ph, pi, mv, idx, _ = R.get_prediction(hm, e.pt1.cpu().numpy(), e.pt2.cpu().numpy())
    assert np.array_equal(e.hm_idx.cpu().numpy(), idx.astype(np.int32))
    assert np.array_equal(e.maxval.cpu().numpy(), mv[..., 0])
    assert np.array_equal(e.preds_hm.cpu().numpy(), ph)
    np.testing.assert_allclose(e.preds_img.cpu().numpy(), pi, atol=6.2e-5)
    # a9-a10 from the engine's key-points.  (a11: the key-points of a randomly initialised network fit no pose, so the
    # consensus problem is ill-posed and which local solution wins is not comparable; PnP parity proper is
    # test_planted_pose_tail below and tests/test_stages_gpu.py.)
    pi_g, mv_g, sc_g = e.preds_img.cpu().numpy(), e.maxval.cpu().numpy(), e.det_score.cpu().numpy()
    for b in range(B):
        ref = R.pose_nms_single(sc_g[b], pi_g[b], mv_g[b])
        if ref is None:
            assert rec["status"][b] == 0
            continue
        kps, sc, prop = ref
        kp_rec = rec["keypoints"][b].reshape(50, 3)
        assert np.array_equal(kp_rec[:, :2], kps) and np.array_equal(kp_rec[:, 2], sc)
        np.testing.assert_allclose(rec["proposal_score"][b], prop, rtol=1e-6)
        assert rec["status"][b] in (1, -1)
        assert np.array_equal(rec["box"][b], box_g[b]) and rec["det_score"][b] == sc_g[b]
    assert rec["image_index"].tolist() == list(range(B))


def test_engine_graph_replay_identical(engine, frames8):
    a = engine.run(frames8, graph=False).copy()
    b = engine.run(frames8, graph=True).copy()
    c = engine.run(frames8, graph=True).copy()
    assert a.tobytes() == b.tobytes() == c.tobytes()


def test_engine_partial_batch_and_determinism(engine, frames8):
    full = engine.run(frames8).copy()
    part = engine.run(frames8[:3]).copy()
    assert part.tobytes() == full[:3].tobytes()


def _project(kp, Rm, t):
    pc = kp @ Rm.T + t
    return np.stack([R.CAM_K[0, 0] * pc[:, 0] / pc[:, 2] + R.CAM_K[0, 2], R.CAM_K[1, 1] * pc[:, 1] / pc[:, 2] + R.CAM_K[1, 2]], 1)


def test_planted_pose_tail(kp_model, frames8):
    """a6 geometry -> planted heat-maps (Gaussian peaks where the true pose projects the key-points) -> a8 decode ->
    a9-a11: the GPU chain equals the oracle chain (indices exact, R/t to 1e-6) and recovers the planted pose to the
    accuracy the 80x64 heat-map grid allows."""
    import math

    from betapose_b200 import stages

    rng = np.random.default_rng(4)
    n, K = 6, 50
    poses, boxes = [], []
    for i in range(n):
        rv = rng.standard_normal(3)
        rv *= rng.uniform(0.2, 2.8) / np.linalg.norm(rv)
        th = np.linalg.norm(rv)
        k = rv / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        Rm = np.eye(3) + math.sin(th) * Kx + (1 - math.cos(th)) * Kx @ Kx
        t = np.array([rng.uniform(-0.12, 0.12), rng.uniform(-0.08, 0.08), rng.uniform(0.5, 0.9)])
        uv = _project(kp_model, Rm, t)
        boxes.append([uv[:, 0].min() - 4, uv[:, 1].min() - 4, uv[:, 0].max() + 4, uv[:, 1].max() + 4])
        poses.append((Rm, t, uv))
    boxes = np.array(boxes, np.float32)
    dev = "cuda"
    crop = stages.crop_resize(torch.from_numpy(frames8).to(dev), torch.from_numpy(boxes).to(dev),
                              torch.arange(n, dtype=torch.int32, device=dev))
    pt1, pt2 = crop["pt1"].cpu().numpy(), crop["pt2"].cpu().numpy()
    # forward crop transform (inverse of transformBoxInvert_batch): image px -> heat-map coordinates
    hm = np.zeros((n, K, 80, 64), np.float32)
    yy, xx = np.mgrid[0:80, 0:64].astype(np.float32)
    for i in range(n):
        ul, br = pt1[i], pt2[i]
        lenH = max(br[1] - ul[1], (br[0] - ul[0]) * 320 / 256)
        lenW = lenH * 256 / 320
        cx, cy = (br[0] - 1 - ul[0]) / 2, (br[1] - 1 - ul[1]) / 2
        offx, offy = max(0, (lenW - 1) / 2 - cx), max(0, (lenH - 1) / 2 - cy)
        for k in range(K):
            hx = (poses[i][2][k, 0] - ul[0] + offx) * 80 / lenH
            hy = (poses[i][2][k, 1] - ul[1] + offy) * 80 / lenH
            hm[i, k] = np.exp(-((xx - hx) ** 2 + (yy - hy) ** 2) / (2 * 1.5 ** 2)) * rng.uniform(0.5, 0.95)
    dec = stages.heatmap_decode(torch.from_numpy(hm).to(dev), crop["pt1"], crop["pt2"], layout="nchw")
    det = torch.ones(n, device=dev)
    kp3d = torch.from_numpy(kp_model).to(dev)
    pose = stages.pose_pnp(dec["preds_img"], dec["maxval"].reshape(n, K), det, kp3d, seed=9)
    torch.cuda.synchronize()
    ph, pi, mv, idx, _ = R.get_prediction(hm, pt1, pt2)
    assert np.array_equal(dec["idx"].cpu().numpy(), idx.astype(np.int32))
    pi_g = dec["preds_img"].cpu().numpy()
    np.testing.assert_allclose(pi_g, pi, atol=6.2e-5)
    for i in range(n):
        kps, sc, prop = R.pose_nms_single(1.0, pi_g[i], mv[i])
        sol = opnp.solve_pnp(kp_model, kps, R.CAM_K, mode=0, thr=12.0, n_hyp=64, seed=9)
        assert sol["ok"] and int(pose["status"][i]) == 1
        assert np.array_equal(pose["inlier"][i].cpu().numpy().astype(bool), sol["inliers"])
        Rg, tg = pose["R"][i].cpu().numpy().reshape(3, 3), pose["t"][i].cpu().numpy()
        np.testing.assert_allclose(Rg, sol["R"], atol=1e-6)
        np.testing.assert_allclose(tg, sol["t"], atol=1e-6)
        # planted pose recovered: key-points are quantised to quarter cells of a ~(lenH/80) px grid
        assert sol["inliers"].sum() >= 45
        assert np.abs(Rg - poses[i][0]).max() < 0.08 and np.abs(tg - poses[i][1]).max() < 0.05


def test_evaluate_cli_synthetic(tmp_path):
    """The command-line path with synthetic frames + weights: Betapose-results.json in the reference's format."""
    import json

    from betapose_b200 import evaluate

    assert evaluate.main(["--synthetic", "6", "--batch", "4", "--outdir", str(tmp_path), "--nClasses", "50", "--sp"]) == 0
    res = json.load(open(tmp_path / "Betapose-results.json"))
    assert isinstance(res, list) and len(res) <= 6
    for e in res:
        assert set(e) == {"image_id", "cam_R", "cam_t", "keypoints", "score"}
        assert len(e["cam_R"]) == 9 and len(e["cam_t"]) == 3 and len(e["keypoints"]) == 150
        assert e["image_id"].startswith("synthetic_")


def test_run_stream_matches_run(engine, frames8):
    """The pipelined streaming API (side-stream uploads, graph replay, ragged last batch) returns exactly what the
    plain per-batch call returns, in order, with global image indices."""
    batches = [frames8[0:3], frames8[3:8], frames8[1:2]]
    ref = [engine.run(b, image_index0=i0).copy() for b, i0 in zip(batches, (10, 13, 18))]
    got = list(engine.run_stream(iter(batches), graph=True, image_index0=10))
    assert len(got) == 3
    for g, r in zip(got, ref):
        assert g.dtype == r.dtype and len(g) == len(r)
        for f in g.dtype.names:
            assert np.array_equal(g[f], r[f]), f


def test_run_stream_from_png_files_through_the_native_ingest(engine, frames8, tmp_path):
    """frames on disk -> decoder pool -> pinned ring -> run_stream.  Same records as feeding the arrays,
    over enough batches that every ring buffer is reused while uploads of earlier batches are still in flight."""
    from PIL import Image

    from betapose_b200.ingest import FrameIngest

    order = [(3 * i) % 8 for i in range(19)]
    paths = []
    for j, i in enumerate(order):
        p = str(tmp_path / f"{j:03d}.png")
        Image.fromarray(frames8[i]).save(p, compress_level=1)
        paths.append(p)
    arrays = [frames8[order[b0: b0 + 2]] for b0 in range(0, len(order), 2)]
    ref = np.concatenate(list(engine.run_stream(iter(arrays), graph=True, image_index0=5)))
    for depth in (0, 2):
        with FrameIngest(3) as ing:
            got = np.concatenate(list(engine.run_stream(ing.batches(paths, 2, depth=depth), graph=True, image_index0=5)))
        assert len(got) == len(ref) == 19
        for f in got.dtype.names:
            assert np.array_equal(got[f], ref[f]), (depth, f)
    # two lanes: the ring must keep the last two batches intact while they upload (FrameIngest.batches(in_flight=2))
    from betapose_b200.engine import BetaposeEngine, PipelinedEngine

    second = BetaposeEngine(engine.B, engine._test_weights[0], engine._test_weights[1], engine._test_weights[2], seed=engine.seed)
    pipe = PipelinedEngine.from_engines([engine, second])
    with FrameIngest(3) as ing:
        got = np.concatenate(list(pipe.run_stream(ing.batches(paths, 2, depth=1, in_flight=2), graph=True, image_index0=5)))
    for f in got.dtype.names:
        assert np.array_equal(got[f], ref[f]), ("two lanes", f)
    del pipe, second


def test_engine_from_packed_weight_cache_is_bit_identical(engine, yolo_stream, kpd_sd, kp_model, frames8, tmp_path):
    """both networks packed on the host, saved, memory-mapped back and uploaded as they are
    (bp_conv_spec.packed_w) -- the engine built that way returns the records of the engine built from fp32 weights."""
    from betapose_b200 import weights, yolo_cfg
    from betapose_b200.engine import BetaposeEngine

    blocks = yolo_cfg.parse_cfg_text(yolo_cfg.default_cfg_text())
    weights.pack_darknet(blocks, yolo_stream).save(str(tmp_path / "y.bppw"))
    weights.pack_fastpose(kpd_sd, 50).save(str(tmp_path / "k.bppw"))
    py, pk = weights.PackedWeights.load(str(tmp_path / "y.bppw")), weights.PackedWeights.load(str(tmp_path / "k.bppw"))
    e2 = BetaposeEngine(8, py, pk, kp_model, seed=5)
    ref = engine.run(frames8).copy()
    got = e2.run(frames8)
    for f in got.dtype.names:
        assert np.array_equal(got[f], ref[f]), f
    assert (ref["status"] == 1).any()


def test_multi_instance_path(engine, frames8):
    """SURVEY 8(f) item 3 end to end (betapose_b200/multi.py): box NMS -> crop per detection -> key-point net -> general
    pose-NMS -> PnP per pose.  With max_det = 1 it is the reference's behaviour (arg-max detection, identity merge) and must
    return the poses of BetaposeEngine.run; with more detections every frame yields between 1 and max_det poses."""
    from betapose_b200.multi import MultiInstance

    n = 4
    ref = engine.run(frames8[:n]).copy()
    one = MultiInstance(engine, nms_thr=0.6, max_det=1).run(frames8[:n])
    assert (ref["status"] == 1).any()
    for b in range(n):
        if ref["status"][b] == 0:
            assert one[b] == []
            continue
        assert len(one[b]) == 1
        p = one[b][0]
        kp = ref["keypoints"][b].reshape(50, 3)
        assert p["status"] == ref["status"][b]
        assert np.array_equal(p["bbox"], ref["box"][b]) and np.array_equal(p["bbox_pick"], ref["box"][b])
        assert p["det_score"] == ref["det_score"][b]
        assert np.array_equal(p["keypoints"], kp[:, :2]) and np.array_equal(p["kp_score"][:, 0], kp[:, 2])
        assert abs(p["proposal_score"] - ref["proposal_score"][b]) < 1e-5
        if p["status"] == 1:
            np.testing.assert_allclose(p["cam_R"].reshape(-1), ref["R"][b], atol=1e-9)
            np.testing.assert_allclose(p["cam_t"].reshape(-1), ref["t"][b], atol=1e-9)
    many = MultiInstance(engine, nms_thr=0.45, max_det=5).run(frames8[:n])
    assert sum(len(m) for m in many) >= 1
    for b in range(n):
        assert len(many[b]) <= 5
        for p in many[b]:
            assert np.isfinite(p["keypoints"]).all() and np.isfinite(p["kp_score"]).all() and p["kp_score"].max() >= 0.3
            assert np.array_equal(p["bbox"], ref["box"][b])  # pose_nms reports the frame's first (= best) box for every pose
            if p["status"] == 1:
                R3 = p["cam_R"]
                assert np.isfinite(R3).all() and np.allclose(R3 @ R3.T, np.eye(3), atol=1e-6) and abs(np.linalg.det(R3) - 1) < 1e-6


def test_engine_batch64_permutation_invariance(yolo_stream, kpd_sd, kp_model):
    """BASELINE.json configs[2] size (batch 64).  Frames are processed independently, and every output element of the
    convolutions accumulates in a fixed order whatever tile it lands in, so permuting the batch must permute the
    records bit for bit (size-independent property; the oracle is too slow for 64 frames)."""
    from betapose_b200 import synth
    from betapose_b200.engine import BetaposeEngine

    e = BetaposeEngine(64, yolo_stream, kpd_sd, kp_model, seed=5)
    frames = synth.synth_frames(64, seed=21)
    a = e.run(frames, graph=True).copy()
    perm = np.random.default_rng(0).permutation(64)
    b = e.run(frames[perm], graph=True).copy()
    assert (a["status"] == 1).sum() >= 32  # the synthetic stream is not silently all-rejected
    for f in a.dtype.names:
        if f == "image_index":
            continue
        assert np.array_equal(a[f][perm], b[f]), f
    # idempotence: same input, same bytes
    assert e.run(frames, graph=True).tobytes() == a.tobytes()
    del e
    torch.cuda.empty_cache()


@pytest.mark.parametrize("concurrent", [True, False])
def test_engine_mixed_objects_and_occlusion(yolo_stream, kpd_sd, kp_model, concurrent):
    """configs[3] / configs[4] shape: a batch that mixes two object slots (own detector + key-point net + key-point
    model each, shared activation buffers) and the Occlusion-LineMod setting left_keypoints = 10.  Every frame must get
    exactly what a single-object engine of its slot produces; the 10 selected points are the 10 best-scored ones.  Both
    schedules: slots concurrently on side streams with their own activation buffers, and one after the other on shared
    buffers."""
    from betapose_b200 import synth
    from betapose_b200.engine import BetaposeEngine

    ys2, ks2 = synth.cached_yolo_weights(1001), synth.cached_kpd_state_dict(2001)
    kp2 = synth.synth_kp_model(2, 50)
    frames = synth.synth_frames(6, seed=33)
    slots = np.array([0, 1, 1, 0, 1, 0])
    mixed = BetaposeEngine(6, [yolo_stream, ys2], [kpd_sd, ks2], np.stack([kp_model, kp2]), left_number=10, seed=5,
                           concurrent_slots=concurrent)
    got = mixed.run(frames, obj_slots=slots, graph=concurrent).copy()
    sel = mixed.selected.cpu().numpy()
    score = mixed.kp_score.cpu().numpy()
    order = np.argsort(slots, kind="stable")  # the engine processes the batch grouped by slot
    del mixed
    torch.cuda.empty_cache()
    for s, (yw, kw, kp) in enumerate([(yolo_stream, kpd_sd, kp_model), (ys2, ks2, kp2)]):
        single = BetaposeEngine(6, yw, kw, kp, left_number=10, seed=5)
        idx = np.nonzero(slots == s)[0]
        ref = single.run(frames[idx]).copy()
        for f in ref.dtype.names:
            if f == "image_index":
                continue
            assert np.array_equal(got[f][idx], ref[f]), (s, f)
        del single
        torch.cuda.empty_cache()
    assert got["image_index"].tolist() == list(range(6))
    for j, b in enumerate(order):  # row j of the engine's buffers holds frame order[j]
        if got["status"][b] == 0:
            continue
        assert sel[j].sum() == 10
        top = np.sort(np.argsort(-score[j], kind="stable")[:10])
        assert np.array_equal(np.nonzero(sel[j])[0], top)


def _mini_sixd(base, seq, rng):
    """a miniature SIXD tree: 5 frames, models/, kpmodels/, models_info.yml, test/<seq>/{rgb,info.yml,gt.yml}"""
    import yaml
    from PIL import Image

    from betapose_b200 import synth

    (base / "models").mkdir(parents=True)
    (base / "kpmodels").mkdir()
    (base / "test" / f"{seq:02d}" / "rgb").mkdir(parents=True)
    verts_mm = rng.uniform(-40, 40, (300, 3))
    kp_mm = synth.synth_kp_model(1, 50) * 1000.0

    def ply(path, v):
        with open(path, "w") as f:
            f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nend_header\n" % len(v))
            for p in v:
                f.write("%.6f %.6f %.6f\n" % tuple(p))

    ply(base / "models" / f"obj_{seq:02d}.ply", verts_mm)
    ply(base / "kpmodels" / f"obj_{seq:02d}.ply", kp_mm)
    yaml.safe_dump({i: {"diameter": 100.0 + i} for i in range(1, 8)}, open(base / "models" / "models_info.yml", "w"))
    frames = synth.synth_frames(5, seed=77)
    info, gts = {}, {}
    for i in range(5):
        Image.fromarray(frames[i]).save(base / "test" / f"{seq:02d}" / "rgb" / f"{i:04d}.png")
        info[i] = {"cam_K": [572.4114, 0.0, 325.2611, 0.0, 573.57043, 242.04899, 0.0, 0.0, 1.0], "depth_scale": 1.0}
        Rm = np.linalg.qr(rng.standard_normal((3, 3)))[0]
        gts[i] = [{"cam_R_m2c": [float(v) for v in Rm.reshape(-1)], "cam_t_m2c": [10.0 * i, -20.0, 800.0 + 10 * i],
                   "obj_bb": [100, 80, 200, 240], "obj_id": seq if i != 3 else 2}]
    yaml.safe_dump(info, open(base / "test" / f"{seq:02d}" / "info.yml", "w"))
    yaml.safe_dump(gts, open(base / "test" / f"{seq:02d}" / "gt.yml", "w"))
    return verts_mm * 0.001, gts


def test_sixd_scoring_loop_and_cli(tmp_path, capsys):
    """The evaluation that follows the per-frame path (betapose_evaluate.py:203-266) on a miniature SIXD tree: the scoring
    loop against an independent recomputation with the oracle, then the whole CLI (frames + ground truth from the tree,
    synthetic weights): Betapose-results.json plus the reference's three summary lines."""
    from betapose_b200 import evaluate, sixd

    rng = np.random.default_rng(8)
    base, seq = tmp_path / "sixd", 5
    verts, gts = _mini_sixd(base, seq, rng)
    bench = sixd.load_sixd(str(base), seq)
    mv, kp, diam = sixd.load_models(str(base), seq, 50)
    assert len(bench.frames) == 5 and bench.diameter[seq] == 105.0 and diam == 105.0 and mv.shape == (300, 3) and kp.shape == (50, 3)
    np.testing.assert_allclose(mv, verts, atol=1e-9)
    # hand-made results: frame 0 near-perfect, 1 off by 3 cm, 2 no pose, 3 other object (skipped), 4 good pose but wrong box
    final, exp_add, exp_proj, exp_iou = [], [], [], []
    for i in range(5):
        G = bench.frames[i].gt[0][1]
        E = G.copy()
        E[:3, 3] += [0.0005, 0.03, 0.0, 0.0, 0.001][i]
        box = np.array([100, 80, 300, 320], np.float64) + ([0, 0, 0, 0] if i != 4 else [400, 300, 400, 300])
        res = [] if i == 2 else [{"bbox": box, "keypoints": np.zeros((50, 2)), "kp_score": np.ones((50, 1)), "proposal_score": 1.0}]
        final.append({"imgname": f"{i:04d}.png", "result": res, "cam_R": E[:3, :3], "cam_t": E[:3, 3].reshape(3, 1)})
        if i in (2, 3):
            continue
        io = R.box_iou([100, 80, 300, 320], list(box))
        exp_iou.append(io)
        if io >= 0.5:
            exp_add.append(R.add_err(G, E, mv) * 1000.0)
            exp_proj.append(R.projection_error_2d(G, E, mv, R.CAM_K))
    out = sixd.evaluate_results(final, bench, seq, mv, log=lambda *_: None)
    assert out["n_frames"] == 3 and out["n_scored"] == len(exp_add) == 2
    assert out["add_accuracy"] == float(np.mean(np.array(exp_add) < 10.5)) == 0.5
    assert out["proj2d_accuracy"] == float(np.mean(np.array(exp_proj) < 5.0))
    assert out["iou_accuracy"] == float(np.mean(np.array(exp_iou) > 0.5))
    np.testing.assert_allclose(out["mean_add_err_mm"], np.mean(exp_add), rtol=1e-9)
    # the CLI end to end
    out_dir = tmp_path / "out"
    assert evaluate.main(["--sixd_base", str(base), "--obj_id", str(seq), "--synthetic_weights", "--batch", "4", "--outdir", str(out_dir),
                          "--nClasses", "50", "--sp"]) == 0
    printed = capsys.readouterr().out
    assert (out_dir / "Betapose-results.json").exists()
    for line in ("Mean add accuracy for seq 05 is:", "2d reprojection accuracy for seq 05 is:", "Mean IoU for seq 05 is:"):
        assert line in printed


def _stagewise_oracle_check(e, frames, sample, slot_of=None, fp32_nets=None):
    """Stage-wise parity of an engine run on `frames` for the images in `sample` (indices into the engine's buffers):
    every stage's output against the oracle fed with the engine's own upstream tensors."""
    n = len(frames)
    slot_of = slot_of if slot_of is not None else [0] * n
    box_g = e.box.cpu().numpy()
    pi_g, mv_g, sc_g = e.preds_img.cpu().numpy(), e.maxval.cpu().numpy(), e.det_score.cpu().numpy()
    for b in sample:
        s = slot_of[b]
        # a1 (each slot's detector input buffer holds its own group, rows relative to the group start)
        g0 = min(i for i in range(n) if slot_of[i] == s) if e.concurrent_slots or e.n_slots == 1 else None
        if g0 is not None:
            yin = e.yolo[s].input(e.B)[b - g0].float().cpu().numpy()
            assert np.array_equal(yin[:, :, :3], R.pil_resize_bicubic(frames[b], 416, 416))
            heads = [e.yolo[s].tensor(h["tensor"], e.B)[b - g0:b - g0 + 1].permute(0, 3, 1, 2).contiguous().cpu().numpy() for h in e.heads[s]]
            dets, rows = R.write_results(R.yolo_decode(heads), 0.01)
            assert int(e.row[b]) == int(rows[0])
            boxes, _ = R.rescale_boxes(dets, 640, 480)
            np.testing.assert_allclose(box_g[b], boxes[0], rtol=2e-6, atol=3e-5)
        pt1, pt2 = R.expand_box(box_g[b], 640, 480)
        assert np.array_equal(e.pt1[b].cpu().numpy(), pt1) and np.array_equal(e.pt2[b].cpu().numpy(), pt2)
        if g0 is not None:
            kin = e.kpd[s].input(e.B)[b - g0].float().cpu().numpy()
            np.testing.assert_allclose(kin[:, :, :3].transpose(2, 0, 1), R.crop_box(frames[b], pt1, pt2), atol=5e-4)
            hm = e.kpd[s].tensor(e.hm_id[s], e.B)[b - g0:b - g0 + 1].permute(0, 3, 1, 2).contiguous().cpu().numpy()
            ph, pi, mv, idx, _ = R.get_prediction(hm, pt1[None], pt2[None])
            assert np.array_equal(e.hm_idx[b].cpu().numpy(), idx[0].astype(np.int32))
            assert np.array_equal(mv_g[b], mv[0, :, 0])
            np.testing.assert_allclose(pi_g[b], pi[0], atol=6.2e-5)
            if fp32_nets is not None:  # the fp16 networks themselves against the fp32 oracle nets on the same inputs
                fp32_nets(s, b, yin, kin, hm)
        ref = R.pose_nms_single(sc_g[b], pi_g[b], mv_g[b])
        st = int(e.status[b])
        assert (ref is None) == (st == 0)
        if ref is not None:
            assert np.array_equal(e.keypoints[b].cpu().numpy(), ref[0])
            keep = R.select_keypoints(ref[1], e.left_number)
            assert np.array_equal(np.nonzero(e.selected[b].cpu().numpy())[0], keep)
            sol = opnp.solve_pnp(e.kp3d[s].cpu().numpy()[keep], ref[0][keep], e.cam_K, mode=0, thr=e.reproj_thr, n_hyp=e.n_hyp, seed=e.seed)
            # junk key-points of random networks: the consensus problem is ill-posed, so only well-posed frames (a clear
            # consensus in both) are compared on R, t; status must agree whenever the oracle is confident either way
            if sol["ok"] and st == 1 and np.array_equal(e.inlier[b].cpu().numpy().astype(bool)[keep], sol["inliers"]):
                np.testing.assert_allclose(e.R[b].cpu().numpy().reshape(3, 3), sol["R"], atol=1e-3)  # tolerance of record; 1e-6 on well-posed inputs: test_stages_gpu.py
                np.testing.assert_allclose(e.t[b].cpu().numpy(), sol["t"], atol=1e-3)


def test_engine_batch64_stagewise_oracle_sample(yolo_blocks, yolo_stream, kpd_sd, kp_model):
    """BASELINE.json configs[2] as benchmarked: batch 64, CUDA-graph replay -- stage by stage against the oracle on four
    sampled frames (the CPU port does ~9 images/s, four frames are seconds), including the fp16 networks (CTA-pair and
    256 / 512-pixel-tile plans) against the fp32 oracle networks on the engine's own inputs."""
    from betapose_b200 import synth
    from betapose_b200.engine import BetaposeEngine
    from oracle import nets as onets

    e = BetaposeEngine(64, yolo_stream, kpd_sd, kp_model, seed=5)
    frames = synth.synth_frames(64, seed=77)
    rec = e.run(frames, graph=True)
    rec2 = e.run(frames, graph=True)   # replay
    torch.cuda.synchronize()
    assert rec.tobytes() == rec2.tobytes()
    yparams, _ = onets.split_darknet_weights(yolo_blocks, yolo_stream)

    acc = {"hm": [], "heads": []}

    def fp32_nets(s, b, yin, kin, hm):
        with torch.no_grad():
            ref_heads = onets.darknet_forward(yolo_blocks, yparams, torch.from_numpy(yin[None, :, :, :3]).permute(0, 3, 1, 2) / 255.0)
            ref_hm = onets.fastpose_forward(kpd_sd, torch.from_numpy(kin[None, :, :, :3]).permute(0, 3, 1, 2))
        acc["heads"].append([(e.yolo[s].tensor(h["tensor"], 64)[b:b + 1].permute(0, 3, 1, 2).cpu(), r) for h, r in zip(e.heads[s], ref_heads)])
        acc["hm"].append((torch.from_numpy(hm), ref_hm))

    _stagewise_oracle_check(e, frames, (0, 21, 42, 63), fp32_nets=fp32_nets)
    # fp16 networks vs fp32 oracle networks over the four sampled frames, errors relative to the largest magnitude of the
    # sample as in tests/test_nets_gpu.py (fp16 activations through 100+ layers: ~3e-2 max / ~4e-3 mean of scale measured;
    # the same images at batch 2 give the same bits -- test_fastpose_batch64_plans_vs_oracle -- so this is the number
    # format, not the batch-64 tile plans)
    for hi in range(3):
        g = torch.cat([fr[hi][0] for fr in acc["heads"]])
        r = torch.cat([fr[hi][1] for fr in acc["heads"]])
        scale = r.abs().max().item()
        assert (g - r).abs().max().item() <= 4e-2 * scale and (g - r).abs().mean().item() <= 5e-3 * scale
    g, r = torch.cat([p[0] for p in acc["hm"]]), torch.cat([p[1] for p in acc["hm"]])
    scale = r.abs().max().item()
    d = (g - r).abs()
    assert d.max().item() <= 4e-2 * scale and d.mean().item() <= 5e-3 * scale, (d.max().item(), d.mean().item(), scale)
    assert (rec["status"] == 1).sum() >= 32
    del e
    torch.cuda.empty_cache()


def test_graph_replay_survives_scratch_growth(yolo_stream, kpd_sd, kp_model, frames8):
    """Kernel scratch (heat-map slices, PnP hypothesis rows) is per stream and grow-only: a graph captured at n = 2 must
    still replay correctly after a later, larger batch (n = 8) and a second engine on the same device made the scratch grow
    -- the old blocks may not be freed while graphs point into them."""
    from betapose_b200.engine import BetaposeEngine

    e = BetaposeEngine(8, yolo_stream, kpd_sd, kp_model, seed=5)
    want2 = e.run(frames8[:2]).copy()                # eager reference, n = 2
    want8 = e.run(frames8).copy()
    a = e.run(frames8[:2], graph=True).copy()        # captures n = 2
    b = e.run(frames8, graph=True).copy()            # captures n = 8: scratch grows
    big = BetaposeEngine(16, yolo_stream, kpd_sd, kp_model, seed=5)  # second engine, larger batch, same native engine
    big.run(np.concatenate([frames8, frames8]), graph=True)
    c = e.run(frames8[:2], graph=True).copy()        # replays the n = 2 graph
    d = e.run(frames8, graph=True).copy()
    torch.cuda.synchronize()
    assert a.tobytes() == want2.tobytes() == c.tobytes()
    assert b.tobytes() == want8.tobytes() == d.tobytes()
    del e, big
    torch.cuda.empty_cache()


def test_model_idx_follows_the_grouping(yolo_stream, kpd_sd, kp_model):
    """A plain run after a mixed-object run must solve every frame against slot 0's key-point model again."""
    from betapose_b200 import synth
    from betapose_b200.engine import BetaposeEngine

    kp2 = synth.synth_kp_model(2, 50)
    frames = synth.synth_frames(4, seed=9)
    e = BetaposeEngine(4, [yolo_stream, synth.variant_yolo_weights(yolo_stream, 1)], [kpd_sd, synth.variant_kpd_state_dict(kpd_sd, 1)],
                       np.stack([kp_model, kp2]), seed=5)
    single = BetaposeEngine(4, yolo_stream, kpd_sd, kp_model, seed=5)
    want = single.run(frames).copy()
    e.run(frames, obj_slots=[1, 0, 1, 1])
    got = e.run(frames).copy()
    assert got.tobytes() == want.tobytes()
    del e, single
    torch.cuda.empty_cache()


LINEMOD_IDS = (1, 2, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14, 15)


def linemod13_kp_models():
    """[13,50,3] key-point models from the shipped PLYs (tests/golden/kp_models.npz, metres); object 10 has 17 points and
    is padded by cycling (model3d.load_kp_model(short='cycle'))."""
    import os

    from betapose_b200 import model3d

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kp_models.npz"))
    out = []
    for oid in LINEMOD_IDS:
        v = g[f"obj_{oid}"]
        assert v.shape[0] == (17 if oid == 10 else 50)
        out.append(v if v.shape[0] == 50 else model3d.pad_by_cycling(v, 50))
    return np.stack(out)


def test_configs3_thirteen_objects_batch32(yolo_blocks, yolo_stream, kpd_sd):
    """BASELINE.json configs[3] on one rank: all 13 LineMod objects mixed, 256 / 8 = 32 frames per GPU, object drawn
    uniformly per frame, one detector + key-point network + shipped key-point model per object (object 10:
    17 points padded by cycling).  Three sampled objects (among them object 10) are checked against single-object engines
    bit for bit, and sampled frames stage by stage against the oracle."""
    from betapose_b200 import synth
    from betapose_b200.engine import BetaposeEngine

    kps = linemod13_kp_models()
    ys = [synth.variant_yolo_weights(yolo_stream, v, yolo_blocks) for v in range(13)]
    ks = [synth.variant_kpd_state_dict(kpd_sd, v) for v in range(13)]
    B = 32
    rng = np.random.default_rng(256)
    slots = rng.integers(0, 13, B)
    slots[:3] = [7, 0, 12]  # make sure object 10 (slot 7) and the two ends are present
    frames = synth.synth_frames(B, seed=91)
    e = BetaposeEngine(B, ys, ks, kps, seed=5)
    assert e.concurrent_slots
    got = e.run(frames, obj_slots=slots, graph=True).copy()
    again = e.run(frames, obj_slots=slots, graph=True).copy()
    assert got.tobytes() == again.tobytes() and got["image_index"].tolist() == list(range(B))
    order = np.argsort(slots, kind="stable")            # the engine's buffers hold the batch grouped by slot
    slot_sorted = slots[order].tolist()
    check = [j for j in range(B) if slot_sorted[j] in (7, 0, 12)][:6]
    _stagewise_oracle_check(e, frames[order], check, slot_of=slot_sorted)
    del e
    torch.cuda.empty_cache()
    for s in (7, 0, 12):
        single = BetaposeEngine(B, ys[s], ks[s], kps[s], seed=5)
        idx = np.nonzero(slots == s)[0]
        ref = single.run(frames[idx]).copy()
        for f in ref.dtype.names:
            if f != "image_index":
                assert np.array_equal(got[f][idx], ref[f]), (s, f)
        del single
        torch.cuda.empty_cache()


def test_configs4_occlusion_batch16(yolo_stream, kpd_sd, kp_model):
    """BASELINE.json configs[4] on one rank: the Occlusion-LineMod variant (occlusion_betapose_evaluate.py:139:
    DataWriter(cam_K, args.left_keypoints, ...), default 10), 128 / 8 = 16 frames per GPU; stage by stage against the
    oracle, selection = the 10 best-scored key-points."""
    from betapose_b200 import synth
    from betapose_b200.engine import BetaposeEngine

    e = BetaposeEngine(16, yolo_stream, kpd_sd, kp_model, left_number=10, seed=5)
    frames = synth.synth_frames(16, seed=128)
    rec = e.run(frames, graph=True).copy()
    _stagewise_oracle_check(e, frames, range(16))
    live = rec["status"] != 0
    assert live.sum() >= 8 and (e.selected.cpu().numpy()[live].sum(1) == 10).all()
    del e
    torch.cuda.empty_cache()


def test_pipelined_engine_two_lanes_matches_single(yolo_stream, kpd_sd, kp_model):
    """PipelinedEngine: two batches in flight on two lanes must return, batch by batch and in order, exactly the records
    a single engine returns; a ragged last batch and host / pinned inputs included."""
    from betapose_b200 import synth
    from betapose_b200.engine import BetaposeEngine, PipelinedEngine

    frames = synth.synth_frames(22, seed=55)
    batches = [frames[0:6], frames[6:12], frames[12:18], frames[18:22]]
    single = BetaposeEngine(6, yolo_stream, kpd_sd, kp_model, seed=5)
    want = [single.run(b).copy() for b in batches]
    del single
    torch.cuda.empty_cache()
    pipe = PipelinedEngine(2, 6, yolo_stream, kpd_sd, kp_model, seed=5)
    pulled = []

    def gen():
        for i, b in enumerate(batches):
            pulled.append(i)
            yield torch.from_numpy(b).pin_memory() if i % 2 else b

    got = []
    for k, rec in enumerate(pipe.run_stream(gen(), graph=True, image_index0=100)):
        assert len(pulled) >= min(k + 2, len(batches))  # the iterator runs ahead of the records: two batches in flight
        got.append(rec)
    assert len(got) == len(want)
    first = 100
    for g, w in zip(got, want):
        assert g["image_index"].tolist() == list(range(first, first + len(w)))
        first += len(w)
        for f in w.dtype.names:
            if f != "image_index":
                assert np.array_equal(g[f], w[f]), f
    # device-side submission (bench.py's `value` loop): fork, two steps on two lanes, join
    dev = torch.from_numpy(frames[:6]).cuda()
    pipe.fork()
    r0 = pipe.submit_device(0, dev)
    r1 = pipe.submit_device(1, dev)
    pipe.join()
    torch.cuda.synchronize()
    assert torch.equal(r0, r1) and r0.data_ptr() != r1.data_ptr()
    del pipe
    torch.cuda.empty_cache()


def test_async_tail_many_steps_and_sync_tail_agree(yolo_stream, kpd_sd, kp_model):
    """PipelinedEngine runs pose-NMS + PnP + record packing of step j on a tail stream while step j + 1 is under way, alternating
    between two sets of small per-frame tensors.  Seven DIFFERENT batches through one lane and through two lanes (every set is
    reused at least once, main part of step j + 2 waits for the tail of step j), graph replay and eager: records identical to
    the plain per-batch call and to a PipelinedEngine with the tail on the lane's own stream (async_tail=False)."""
    from betapose_b200 import synth
    from betapose_b200.engine import BetaposeEngine, PipelinedEngine

    frames = synth.synth_frames(28, seed=77)
    batches = [frames[4 * i:4 * i + 4] for i in range(7)]
    e0 = BetaposeEngine(4, yolo_stream, kpd_sd, kp_model, seed=9)
    want = [e0.run(b).copy() for b in batches]
    e1 = BetaposeEngine(4, yolo_stream, kpd_sd, kp_model, seed=9)

    def check(got):
        assert len(got) == len(want)
        for k, (g, w) in enumerate(zip(got, want)):
            assert g["image_index"].tolist() == list(range(4 * k, 4 * k + 4))
            for f in w.dtype.names:
                if f != "image_index":
                    assert np.array_equal(g[f], w[f]), (k, f)

    for engines in ([e0], [e0, e1]):
        for async_tail in (True, False):
            pipe = PipelinedEngine.from_engines(engines, async_tail=async_tail)
            assert pipe.async_tail == async_tail
            for graph in (True, False):
                check(list(pipe.run_stream(iter(batches), graph=graph)))
    # device-side submission, three steps per lane, then join: the last records of each lane are those of batches 4 and 5
    pipe = PipelinedEngine.from_engines([e0, e1])
    devb = [torch.from_numpy(b).cuda() for b in batches]
    pipe.fork()
    recs = [pipe.submit_device(k, devb[k]) for k in range(6)]
    pipe.join()
    torch.cuda.synchronize()
    from betapose_b200 import stages
    for k in (2, 3, 4, 5):  # per lane: set 0 now holds its third step's records (batches 4 / 5), set 1 its second step's (2 / 3)
        g = stages.records_to_numpy(recs[k]).copy()
        for f in want[k].dtype.names:
            if f != "image_index":
                assert np.array_equal(g[f], want[k][f]), (k, f)


def test_run_stream_abandoned_early_then_plain_run(engine, frames8):
    """A consumer that stops pulling records while batches are still in flight (lane and tail streams busy) and then makes a
    plain per-batch call: the generator's clean-up orders the in-flight work before the caller's stream, so the plain call --
    which shares the engine's buffers and small-tensor set 0 -- returns what it always returns."""
    want = engine.run(frames8[2:6]).copy()
    batches = [frames8[0:4], frames8[4:8], frames8[1:5], frames8[3:7], frames8[0:4]]
    gen = engine.run_stream(iter(batches), graph=True)
    first = next(gen)
    assert len(first) == 4
    gen.close()  # two more batches are in flight
    got = engine.run(frames8[2:6])
    for f in want.dtype.names:
        assert np.array_equal(got[f], want[f]), f
