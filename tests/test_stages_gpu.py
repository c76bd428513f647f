"""GPU parity of the stage kernels (through the C ABI) against the CPU oracle (oracle/restate.py, oracle/pnp.py).
Integer / index outputs must be bit-exact; fp32 affine outputs exact or within the stated ulp-level tolerance."""
import numpy as np
import pytest
import torch

from oracle import pnp as opnp
from oracle import restate as R

pytestmark = pytest.mark.gpu


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------------------ a1
@pytest.mark.parametrize("shape,out", [((480, 640), (416, 416)), ((37, 53), (416, 416)), ((600, 800), (208, 320)), ((416, 416), (416, 416))])
def test_resize_bicubic_bit_exact(shape, out):
    from betapose_b200 import stages

    rng = np.random.default_rng(0)
    fr = rng.integers(0, 256, (3,) + shape + (3,), dtype=np.uint8)
    buf, f32 = stages.resize_bicubic(_cuda(fr), out[0], out[1], want_net=True, want_f32=True)
    u8 = stages.net_input_pixels(buf)
    assert float(buf[..., 3:].abs().max()) == 0.0 and float(buf[:, :, :3].abs().max()) == 0.0  # pad channels / columns
    torch.cuda.synchronize()
    for b in range(fr.shape[0]):
        ref = R.pil_resize_bicubic(fr[b], out[0], out[1])
        assert np.array_equal(u8[b].float().cpu().numpy(), ref.astype(np.float32))
        assert np.array_equal(f32[b].cpu().numpy(), (ref.astype(np.float32) / np.float32(255)).transpose(2, 0, 1))


def test_resize_matches_pillow_directly():
    PIL = pytest.importorskip("PIL.Image")
    from betapose_b200 import stages

    fr = np.random.default_rng(1).integers(0, 256, (1, 480, 640, 3), dtype=np.uint8)
    buf, _ = stages.resize_bicubic(_cuda(fr), 416, 416)
    u8 = stages.net_input_pixels(buf)
    ref = np.asarray(PIL.fromarray(fr[0]).resize((416, 416), PIL.BICUBIC))
    assert np.array_equal(u8[0].float().cpu().numpy(), ref.astype(np.float32))


# ------------------------------------------------------------------------------------------------ a3-a5
def _rand_heads(rng, B, scale=1.0, obj_shift=0.0):
    heads = []
    for g in (13, 26, 52):
        h = (rng.standard_normal((B, 18, g, g)) * scale).astype(np.float32)
        h[:, 4::6] += obj_shift
        heads.append(h)
    return heads


def _run_decode(heads, B, conf=0.01, want_decoded=False):
    from betapose_b200 import stages

    nhwc = []
    for h in heads:
        t = torch.zeros((B, h.shape[2], h.shape[3], 20), dtype=torch.float32, device="cuda")  # pitch 20 like the net
        t[..., :18] = _cuda(h).permute(0, 2, 3, 1)
        nhwc.append(t[..., :18])
    anchors = [R.YOLO_ANCHORS[32], R.YOLO_ANCHORS[16], R.YOLO_ANCHORS[8]]
    out = stages.yolo_decode_argmax(nhwc, anchors, B, conf=conf, want_decoded=want_decoded)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_yolo_decode_argmax_row_exact(seed):
    rng = np.random.default_rng(seed)
    B = 5
    heads = _rand_heads(rng, B)
    out = _run_decode(heads, B, want_decoded=True)
    pred = R.yolo_decode(heads)
    dec = out["decoded"].cpu().numpy()
    # the objectness column decides the winner: must be bit-identical to the numpy fp32 evaluation or the row may flip
    np.testing.assert_allclose(dec, pred, rtol=2e-6, atol=1e-5)
    dets, rows = R.write_results(pred, 0.01)
    assert np.array_equal(out["row"].cpu().numpy(), rows.astype(np.int32))
    assert out["valid"].cpu().numpy().all()
    np.testing.assert_allclose(out["det"].cpu().numpy(), dets, rtol=2e-6, atol=2e-5)
    boxes, scores = R.rescale_boxes(dets, 640, 480)
    np.testing.assert_allclose(out["box"].cpu().numpy(), boxes, rtol=2e-6, atol=3e-5)
    np.testing.assert_allclose(out["score"].cpu().numpy(), scores[:, 0], rtol=2e-6)


def test_yolo_decode_no_candidate_and_ties():
    rng = np.random.default_rng(5)
    B = 3
    heads = _rand_heads(rng, B, obj_shift=-12.0)  # sigmoid(obj) << 0.01 everywhere
    heads[1][1, 4, 7, 9] = 3.0   # image 1 gets exactly one candidate: head 16, anchor 0, y=7, x=9
    heads[2][2, 10, 3, 4] = 2.0  # image 2: two equal maxima -> lowest flat row wins
    heads[0][2, 4, 1, 1] = 2.0
    out = _run_decode(heads, B)
    assert out["valid"].cpu().tolist() == [0, 1, 1]
    rows = out["row"].cpu().tolist()
    assert rows[0] == -1
    assert rows[1] == 507 + 0 * 26 * 26 + 7 * 26 + 9
    assert rows[2] == 0 * 169 + 1 * 13 + 1  # head-32 candidate precedes the head-8 one
    pred = R.yolo_decode(heads)
    dets, rr = R.write_results(pred, 0.01)
    assert rr.tolist() == rows[1:]


# ------------------------------------------------------------------------------------------------ a4 with box NMS (8f-3)
def _flat_nms(res):
    det, row, cnt = res["det"].cpu().numpy(), res["row"].cpu().numpy(), res["count"].cpu().numpy()
    return (np.concatenate([det[b, : cnt[b]] for b in range(len(cnt))]).reshape(-1, 8),
            np.concatenate([row[b, : cnt[b]] for b in range(len(cnt))]), cnt)


def test_box_nms_bit_exact_against_reference_branch_goldens():
    """bp_write_results_nms against what the reference's own (shipped, disabled) IoU-NMS branch produces
    (tests/golden/box_nms_golden.npz), including the full 10647-row case (16384-key shared-memory sort, 1024 threads)."""
    import os

    from betapose_b200 import stages

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "box_nms_golden.npz"))
    for i in range(int(g["n_cases"])):
        pred, want = g[f"pred{i}"], g[f"dets{i}"]
        res = stages.write_results_nms(_cuda(pred), float(g[f"conf{i}"]), float(g[f"thr{i}"]), max_det=pred.shape[1])
        d, r, cnt = _flat_nms(res)
        assert np.array_equal(d, want), i
        assert np.array_equal(res["total"].cpu().numpy(), cnt)
        _, rows, counts = R.write_results_nms(pred, float(g[f"conf{i}"]), float(g[f"thr{i}"]))
        assert np.array_equal(r, rows) and np.array_equal(cnt, counts)


def test_box_nms_on_decoded_heads_cap_and_seam():
    """decode (bp_yolo_decode_argmax, decoded tensor) -> bp_write_results_nms on real head shapes; max_det caps the
    output but not the count; the dynamic_write_results seam with its > 100 detections retry."""
    from betapose_b200 import compat, stages

    rng = np.random.default_rng(11)
    B = 3
    heads = _rand_heads(rng, B, obj_shift=-4.0)
    out = _run_decode(heads, B, want_decoded=True)
    pred_dev = out["decoded"]
    pred = pred_dev.cpu().numpy()
    res = stages.write_results_nms(pred_dev, 0.05, 0.45, max_det=4096)
    d, r, cnt = _flat_nms(res)
    d2, r2, c2 = R.write_results_nms(pred, 0.05, 0.45)
    assert cnt.min() > 20
    assert np.array_equal(cnt, c2) and np.array_equal(r, r2) and np.array_equal(d, d2)
    capped = stages.write_results_nms(pred_dev, 0.05, 0.45, max_det=7)
    assert np.array_equal(capped["count"].cpu().numpy(), np.minimum(c2, 7)) and np.array_equal(capped["total"].cpu().numpy(), c2)
    assert np.array_equal(capped["det"].cpu().numpy()[1, :7], res["det"].cpu().numpy()[1, :7])
    # seam: more than 100 detections in the batch -> one retry with nms_conf - 0.05 (yolo/util.py:111-113)
    dets = compat.dynamic_write_results(pred_dev, 0.05, 80, True, 0.45, box_nms=True)
    want = d2 if len(d2) <= 100 else R.write_results_nms(pred, 0.05, 0.45 - 0.05)[0]  # the retry subtracts in double, as the reference does
    assert len(d2) > 100 and np.array_equal(dets.cpu().numpy(), want)
    few = compat.dynamic_write_results(pred_dev, 0.9999, 80, True, 0.45, box_nms=True)
    w2 = R.write_results_nms(pred, 0.9999, 0.45)[0]
    assert (isinstance(few, int) and few == 0 and len(w2) == 0) or np.array_equal(few.cpu().numpy(), w2)
    assert compat.dynamic_write_results(pred_dev, 2.0, 80, True, 0.45, box_nms=True) == 0
    # default behaviour of the seam is untouched: one arg-max row per image
    one = compat.dynamic_write_results(pred_dev, 0.05, 80, True, 0.45)
    assert one.shape == (B, 8)


# ------------------------------------------------------------------------------------------------ a6
BOXES = np.array([
    [200.3, 120.7, 330.9, 300.2],    # ordinary, w > 100
    [10.2, 5.5, 70.8, 90.1],         # small (30 % expansion), clamps at 0
    [500.0, 300.0, 655.0, 500.0],    # beyond the right/bottom edge
    [-20.5, -10.0, 90.0, 200.0],     # negative corner
    [300.0, 200.0, 301.0, 201.0],    # degenerate -> min size 5
    [0.0, 0.0, 639.0, 479.0],        # whole frame
    [100.5, 100.5, 201.6, 400.2],    # wb = 101-ish (the int-vs-true division case of SURVEY A.4)
], np.float32)


def test_crop_resize_matches_oracle(frames8):
    from betapose_b200 import stages

    n = len(BOXES)
    idx = np.arange(n, dtype=np.int32) % frames8.shape[0]
    out = stages.crop_resize(_cuda(frames8), _cuda(BOXES), _cuda(idx), want_f16=True, want_f32=True)
    torch.cuda.synchronize()
    for i in range(n):
        pt1, pt2 = R.expand_box(BOXES[i], 640, 480)
        assert np.array_equal(out["pt1"][i].cpu().numpy(), pt1) and np.array_equal(out["pt2"][i].cpu().numpy(), pt2)
        ref = R.crop_box(frames8[idx[i]], pt1, pt2)
        got = out["f32"][i].cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=0, atol=2.4e-7)  # <= 2 fp32 ulp at |x| <= 1
        got16 = stages.net_input_pixels(out["net"])[i].float().cpu().numpy().transpose(2, 0, 1)
        np.testing.assert_allclose(got16, ref, rtol=0, atol=5e-4)    # fp16 rounding of values in [-0.5, 0.6]
        assert float(out["net"][i, :, :, 3:].abs().max()) == 0.0


def test_crop_resize_invalid_rows_are_zero(frames8):
    from betapose_b200 import stages

    valid = np.array([1, 0, 1], np.uint8)
    out = stages.crop_resize(_cuda(frames8), _cuda(BOXES[:3]), _cuda(np.zeros(3, np.int32)), valid=_cuda(valid))
    torch.cuda.synchronize()
    assert float(out["net"][1].abs().max()) == 0.0 and float(out["net"][0].abs().max()) > 0.0


# ------------------------------------------------------------------------------------------------ a8
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("K", [50, 17])
def test_heatmap_decode_exact(layout, K):
    from betapose_b200 import stages

    rng = np.random.default_rng(7)
    n = 6
    hm = rng.standard_normal((n, K, 80, 64)).astype(np.float32)
    hm[0, 0] = -1.0                      # all non-positive -> (0,0) + 0.2
    hm[0, 1] = 0.0
    hm[1, 2, 0, 0] = 9.0                 # border peak: no quarter-pixel refinement
    hm[1, 3, 79, 63] = 9.0
    hm[2, 4, 40, 30] = 9.0; hm[2, 4, 41, 31] = 9.0   # tie -> lowest flat index
    hm[3, 5, 10, 10] = 9.0; hm[3, 5, 10, 11] = hm[3, 5, 10, 9] = 1.0  # zero horizontal gradient -> sign 0
    pt1 = np.array([[50.5, 40.25], [0, 0], [300.7, 200.1], [10, 10], [100, 50], [400.5, 100.5]], np.float32)
    pt2 = pt1 + np.array([[120.3, 200.9], [5, 5], [150.2, 100.8], [300, 100], [80.4, 300.6], [200, 200]], np.float32)
    if layout == "nchw":
        t = _cuda(hm)
    else:
        buf = torch.zeros((n, 80, 64, (K + 3) // 4 * 4), dtype=torch.float32, device="cuda")
        buf[..., :K] = _cuda(hm).permute(0, 2, 3, 1)
        t = buf[..., :K]
    out = stages.heatmap_decode(t, _cuda(pt1), _cuda(pt2), layout=layout)
    torch.cuda.synchronize()
    ph, pi, mv, idx, sign = R.get_prediction(hm, pt1, pt2)
    assert np.array_equal(out["idx"].cpu().numpy(), idx.astype(np.int32))
    assert np.array_equal(out["maxval"].cpu().numpy(), mv)
    assert np.array_equal(out["preds_hm"].cpu().numpy(), ph)
    np.testing.assert_allclose(out["preds_img"].cpu().numpy(), pi, rtol=0, atol=6.2e-5)  # 1 fp32 ulp at 640
    # exactness where numpy and the kernel round identically (no fused multiply-add on either side)
    assert np.mean(out["preds_img"].cpu().numpy() == pi) > 0.99


# ------------------------------------------------------------------------------------------------ a9-a11
def _pnp_case(rng, kp, sigma, n_out=0):
    import math

    rv = rng.standard_normal(3)
    rv = rv / np.linalg.norm(rv) * rng.uniform(0, math.pi * 0.95)
    th = np.linalg.norm(rv)
    k = rv / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    Rm = np.eye(3) + math.sin(th) * Kx + (1 - math.cos(th)) * Kx @ Kx
    t = np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(0.6, 1.2)])
    pc = kp @ Rm.T + t
    uv = np.stack([R.CAM_K[0, 0] * pc[:, 0] / pc[:, 2] + R.CAM_K[0, 2], R.CAM_K[1, 1] * pc[:, 1] / pc[:, 2] + R.CAM_K[1, 2]], 1)
    uv = uv + rng.normal(0, sigma, uv.shape)
    if n_out:
        o = rng.choice(len(kp), n_out, replace=False)
        uv[o] += rng.normal(0, 60, (n_out, 2))
    return Rm, t, uv.astype(np.float32)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("sigma,n_out", [(0.0, 0), (0.5, 0), (1.0, 5), (2.0, 10)])
def test_pose_pnp_matches_oracle(kp_model, mode, sigma, n_out):
    from betapose_b200 import stages

    if mode == 1 and n_out:
        pytest.skip("all-points mode has no outlier rejection (mirrors cv2.solvePnP)")
    rng = np.random.default_rng(100 + int(sigma * 10) + n_out)
    n, K = 24, 50
    preds = np.zeros((n, K, 2), np.float32)
    maxval = rng.uniform(0.35, 0.95, (n, K)).astype(np.float32)
    det_score = rng.uniform(0.2, 1.0, n).astype(np.float32)
    for i in range(n):
        _, _, uv = _pnp_case(rng, kp_model, sigma, n_out)
        preds[i] = uv + np.float32(0.3)  # the kernel subtracts 0.3 (pose_nms)
    maxval[3] = 0.1                       # rejected by pose-NMS (max score < 0.3)
    maxval[4, 7] = 0.0                    # zero score -> 1e-5
    out = stages.pose_pnp(_cuda(preds), _cuda(maxval), _cuda(det_score), _cuda(kp_model), mode=mode, n_hyp=64, seed=11)
    torch.cuda.synchronize()
    status = out["status"].cpu().numpy()
    for i in range(n):
        ref = R.pose_nms_single(det_score[i], preds[i], maxval[i])
        if ref is None:
            assert status[i] == 0 and i == 3
            continue
        kps, sc, prop = ref
        assert status[i] == 1
        assert np.array_equal(out["keypoints"][i].cpu().numpy(), kps)
        assert np.array_equal(out["kp_score"][i].cpu().numpy(), sc)
        np.testing.assert_allclose(float(out["proposal"][i]), prop, rtol=1e-6)
        sol = opnp.solve_pnp(kp_model, kps, R.CAM_K, mode=mode, thr=12.0, n_hyp=64, seed=11)
        assert sol["ok"]
        assert np.array_equal(out["inlier"][i].cpu().numpy().astype(bool), sol["inliers"])
        # tolerance of record (BASELINE.json north_star): 1e-3 on R and t; the fp64 kernel is far inside it
        np.testing.assert_allclose(out["R"][i].cpu().numpy().reshape(3, 3), sol["R"], atol=1e-6)
        np.testing.assert_allclose(out["t"][i].cpu().numpy(), sol["t"], atol=1e-6)


def test_pose_pnp_selection_left_number(kp_model):
    from betapose_b200 import stages

    rng = np.random.default_rng(9)
    n, K = 8, 50
    preds = np.zeros((n, K, 2), np.float32)
    maxval = rng.uniform(0.35, 0.95, (n, K)).astype(np.float32)
    maxval[:, 10] = maxval[:, 20]  # ties: the first arg-min goes first
    for i in range(n):
        preds[i] = _pnp_case(rng, kp_model, 0.3)[2] + np.float32(0.3)
    det = np.ones(n, np.float32)
    out = stages.pose_pnp(_cuda(preds), _cuda(maxval), _cuda(det), _cuda(kp_model), left_number=10, mode=0, seed=2)
    torch.cuda.synchronize()
    for i in range(n):
        keep = R.select_keypoints(maxval[i], 10)
        sel = np.nonzero(out["selected"][i].cpu().numpy())[0]
        assert np.array_equal(sel, keep)
        kps = preds[i] - np.float32(0.3)
        sol = opnp.solve_pnp(kp_model[keep], kps[keep], R.CAM_K, mode=0, n_hyp=64, seed=2)
        np.testing.assert_allclose(out["R"][i].cpu().numpy().reshape(3, 3), sol["R"], atol=1e-6)
        np.testing.assert_allclose(out["t"][i].cpu().numpy(), sol["t"], atol=1e-6)


def test_pack_records_roundtrip(kp_model):
    from betapose_b200 import stages

    rng = np.random.default_rng(3)
    n, K = 5, 50
    preds = np.stack([_pnp_case(rng, kp_model, 0.2)[2] for _ in range(n)]) + np.float32(0.3)
    maxval = rng.uniform(0.4, 0.9, (n, K)).astype(np.float32)
    det = rng.uniform(0.2, 1, n).astype(np.float32)
    box = rng.uniform(0, 400, (n, 4)).astype(np.float32)
    pose = stages.pose_pnp(_cuda(preds), _cuda(maxval), _cuda(det), _cuda(kp_model))
    rec = stages.records_to_numpy(stages.pack_records(1000, _cuda(box), _cuda(det), pose))
    assert rec["image_index"].tolist() == list(range(1000, 1000 + n))
    assert np.array_equal(rec["box"], box) and np.array_equal(rec["det_score"], det)
    assert np.array_equal(rec["R"], pose["R"].cpu().numpy()) and np.array_equal(rec["t"], pose["t"].cpu().numpy())
    kp = rec["keypoints"].reshape(n, 50, 3)
    assert np.array_equal(kp[:, :, :2], pose["keypoints"].cpu().numpy()) and np.array_equal(kp[:, :, 2], pose["kp_score"].cpu().numpy())


# ------------------------------------------------------------------------------------------------ scoring (8(f) item 1)
def test_score_poses_matches_reference_goldens():
    """bp_score_poses against the reference's own metrics.py outputs (golden) and the oracle: fp64, 1e-12 relative."""
    import os

    from betapose_b200 import compat, stages

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics_golden.npz"))
    n = len(g["gt"])
    out = stages.score_poses(_cuda(g["est"][:, :3, :3].copy()), _cuda(g["est"][:, :3, 3].copy()), _cuda(g["box_est"].astype(np.float32)),
                             _cuda(g["gt"][:, :3, :3].copy()), _cuda(g["gt"][:, :3, 3].copy()), _cuda(g["box_gt"].astype(np.float32)),
                             _cuda(g["model"]), cam_K=g["cam"])
    torch.cuda.synchronize()
    np.testing.assert_allclose(out["add"].cpu().numpy(), g["add"], rtol=1e-12, atol=1e-16)
    np.testing.assert_allclose(out["proj"].cpu().numpy(), g["proj"], rtol=1e-11, atol=1e-12)
    iou32 = np.array([R.box_iou(g["box_gt"][i].astype(np.float32).astype(np.float64), g["box_est"][i].astype(np.float32).astype(np.float64))
                      for i in range(n)])
    np.testing.assert_allclose(out["iou"].cpu().numpy(), iou32, rtol=2e-7, atol=0)
    assert np.array_equal(out["scored"].cpu().numpy().astype(bool), out["iou"].cpu().numpy() >= 0.5)
    # the drop-in seams
    assert abs(compat.add_err(g["gt"][3], g["est"][3], g["model"]) - g["add"][3]) <= 1e-12 * g["add"][3]
    assert abs(compat.projection_error_2d(g["gt"][3], g["est"][3], g["model"], g["cam"]) - g["proj"][3]) <= 1e-11 * g["proj"][3]
    assert abs(compat.iou(g["box_gt"][1], g["box_est"][1]) - g["iou"][1]) <= 1e-6
    s = stages.summarize_scores(out["add"], out["proj"], out["iou"], out["scored"], diameter_mm=102.0)
    sc = g["iou"] >= 0.5
    assert s["n_scored"] == int(sc.sum()) and abs(s["add_accuracy"] - float((g["add"][sc] * 1000 < 10.2).mean())) < 1e-12


def test_score_poses_full_size_properties(kp_model):
    """BASELINE-size batch (256 poses, a 20k-vertex model): identical poses score exactly 0; a pure translation d scores
    ADD = |d| for every model (size-independent properties)."""
    from betapose_b200 import stages

    rng = np.random.default_rng(5)
    n, V = 256, 20000
    model = rng.uniform(-0.08, 0.08, (V, 3))
    Rg = np.stack([np.linalg.qr(rng.standard_normal((3, 3)))[0] for _ in range(n)])
    Rg[np.linalg.det(Rg) < 0, :, 0] *= -1
    tg = np.stack([rng.uniform(-0.1, 0.1, n), rng.uniform(-0.1, 0.1, n), rng.uniform(0.6, 1.2, n)], 1)
    box = np.tile(np.array([[100, 100, 300, 300]], np.float32), (n, 1))
    same = stages.score_poses(_cuda(Rg), _cuda(tg), _cuda(box), _cuda(Rg), _cuda(tg), _cuda(box), _cuda(model))
    assert float(same["add"].abs().max()) == 0.0 and float(same["proj"].abs().max()) == 0.0 and float(same["iou"].min()) == 1.0
    d = rng.standard_normal((n, 3)) * 0.01
    sh = stages.score_poses(_cuda(Rg), _cuda(tg + d), _cuda(box), _cuda(Rg), _cuda(tg), _cuda(box), _cuda(model))
    np.testing.assert_allclose(sh["add"].cpu().numpy(), np.linalg.norm(d, axis=1), rtol=1e-9)


# ------------------------------------------------------------------------------------------------ general pose-NMS (8(f) item 3)
def _nms_clusters(rng, n_clusters, per_cluster, K=50):
    """proposals scattered around `n_clusters` distinct poses: members of a cluster are within a few pixels (they are
    suppressed and merged), clusters are far apart."""
    bb, bs, pp, ps = [], [], [], []
    for c in range(n_clusters):
        centre = rng.uniform(80, 400, (K, 2)).astype(np.float32)
        for _ in range(per_cluster):
            p = centre + rng.normal(0, 1.5, (K, 2)).astype(np.float32)
            x1, y1, x2, y2 = p[:, 0].min() - 5, p[:, 1].min() - 5, p[:, 0].max() + 5, p[:, 1].max() + 5
            bb.append([x1, y1, x2, y2]); bs.append(rng.uniform(0.3, 0.99))
            pp.append(p); ps.append(rng.uniform(0.05, 0.95, (K, 1)).astype(np.float32))
    return np.float32(bb), np.float32(bs).reshape(-1, 1), np.float32(pp), np.float32(ps)


def _check_nms(out, first, ref, K=50):
    cnt = int(out["count"])
    assert cnt == len(ref)
    for j, r in enumerate(ref):
        np.testing.assert_allclose(out["keypoints"][first + j].cpu().numpy(), r["keypoints"], rtol=0, atol=2e-4)
        np.testing.assert_allclose(out["kp_score"][first + j].cpu().numpy(), r["kp_score"][:, 0], rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(float(out["proposal"][first + j]), float(r["proposal_score"][0]), rtol=2e-5)


@pytest.mark.parametrize("n_clusters,per_cluster", [(1, 1), (1, 4), (3, 3), (6, 2), (16, 4)])
def test_pose_nms_general_matches_oracle(n_clusters, per_cluster):
    from betapose_b200 import stages

    rng = np.random.default_rng(100 * n_clusters + per_cluster)
    bb, bs, pp, ps = _nms_clusters(rng, n_clusters, per_cluster)
    ref = R.pose_nms(bb, bs, pp, ps)
    out = stages.pose_nms(_cuda(bb), _cuda(bs), _cuda(pp), _cuda(ps[..., 0]))
    torch.cuda.synchronize()
    assert len(ref) == n_clusters  # every cluster collapses to one merged pose
    _check_nms({k: (v[0] if k == "count" else v) for k, v in out.items()}, 0, ref)


def test_pose_nms_reference_goldens_and_batched():
    """The reference's own pose_nms outputs (tests/golden/pose_nms_golden.npz: n = 1, n = 3 with a merge, rejected), all
    four cases in ONE launch (several images concatenated), and the drop-in seam for n = 3."""
    import os

    from betapose_b200 import compat, stages

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pose_nms_golden.npz"))
    nc = int(g["n_cases"])
    bb = np.concatenate([g[f"c{i}_bb"] for i in range(nc)]).astype(np.float32)
    bs = np.concatenate([g[f"c{i}_bs"] for i in range(nc)]).astype(np.float32)
    pp = np.concatenate([g[f"c{i}_pp"] for i in range(nc)]).astype(np.float32)
    ps = np.concatenate([g[f"c{i}_ps"] for i in range(nc)]).astype(np.float32)
    counts = [len(g[f"c{i}_bb"]) for i in range(nc)]
    out = stages.pose_nms(_cuda(bb), _cuda(bs), _cuda(pp), _cuda(ps[..., 0]), counts=counts)
    torch.cuda.synchronize()
    first = out["first"].cpu().numpy()
    for i in range(nc):
        n_res = int(g[f"c{i}_n"])
        assert int(out["count"][i]) == n_res
        for j in range(n_res):
            np.testing.assert_allclose(out["keypoints"][first[i] + j].cpu().numpy(), g[f"c{i}_r{j}_kp"], rtol=0, atol=2e-4)
            np.testing.assert_allclose(out["kp_score"][first[i] + j].cpu().numpy(), g[f"c{i}_r{j}_sc"][:, 0], rtol=2e-5, atol=1e-6)
            np.testing.assert_allclose(float(out["proposal"][first[i] + j]), float(g[f"c{i}_r{j}_prop"][0]), rtol=2e-5)
    res = compat.pose_nms(torch.from_numpy(g["c2_bb"]), torch.from_numpy(g["c2_bs"]), torch.from_numpy(g["c2_pp"]), torch.from_numpy(g["c2_ps"].copy()))
    assert len(res) == int(g["c2_n"]) and np.array_equal(res[0]["bbox"].numpy(), g["c2_r0_bbox"])
    np.testing.assert_allclose(res[1]["keypoints"].numpy(), g["c2_r1_kp"], atol=2e-4)
    assert res[0]["kp_score"].shape == (50, 1) and res[0]["proposal_score"].shape == (1,)
