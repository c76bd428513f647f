"""GPU: the drop-in seam functions (betapose_b200/compat.py) called the way the reference's stage objects call
their own (CPU tensors in, same return types / sentinels), checked against the oracle."""
import numpy as np
import pytest
import torch

from oracle import nets as onets
from oracle import pnp as opnp
from oracle import restate as R

pytestmark = pytest.mark.gpu


def test_darknet_seam_and_write_results(yolo_blocks, yolo_stream, frames8, tmp_path):
    from betapose_b200 import compat, synth

    wpath = str(tmp_path / "synth.weights")
    synth.write_darknet_weights(wpath, yolo_stream, seen=123)
    det_model = compat.Darknet(None, reso=416)
    det_model.load_weights(wpath)
    det_model.net_info["height"] = "416"        # dataloader.py:296
    det_model.cuda().eval()
    assert det_model.seen == 123
    x = torch.from_numpy(np.stack([R.yolo_input_from_frame(f) for f in frames8[:2]]))
    pred = det_model(x, CUDA=True)
    assert pred.shape == (2, 10647, 6) and pred.device == x.device
    params, _ = onets.split_darknet_weights(yolo_blocks, yolo_stream)
    with torch.no_grad():
        ref = R.yolo_decode([h.numpy() for h in onets.darknet_forward(yolo_blocks, params, x)])
    # objectness / class columns are sigmoids in [0,1]; boxes are pixels: compare in their own scales
    d = np.abs(pred.numpy()[..., 4:] - ref[..., 4:])   # fp16 network (fp16 input here) vs fp32 oracle, after the sigmoid
    assert d.max() < 5e-2 and d.mean() < 3e-3
    dets = compat.dynamic_write_results(pred, 0.01, 80, nms=True, nms_conf=0.6)
    ref_dets, rows = R.write_results(pred.numpy(), 0.01)   # a4 on the seam's own prediction tensor: exact
    assert isinstance(dets, torch.Tensor) and dets.shape == (2, 8)
    np.testing.assert_array_equal(dets.numpy(), ref_dets)
    assert compat.dynamic_write_results(pred * 0, 0.6, 80) == 0  # sentinel


def test_crop_seam_mutates_and_fills(frames8):
    from betapose_b200 import compat

    fr = frames8[0]
    img = torch.from_numpy(fr.astype(np.float32).transpose(2, 0, 1) / np.float32(255))
    orig = img.clone()
    boxes = torch.tensor([[200.3, 120.7, 330.9, 300.2], [10.2, 5.5, 70.8, 90.1]])
    inps, pt1, pt2 = torch.zeros(2, 3, 320, 256), torch.zeros(2, 2), torch.zeros(2, 2)
    a, b, c = compat.crop_from_dets(img, boxes, inps, pt1, pt2)
    assert a is inps and b is pt1 and c is pt2
    for i in range(2):
        p1, p2 = R.expand_box(boxes[i].numpy(), 640, 480)
        assert np.array_equal(pt1[i].numpy(), p1) and np.array_equal(pt2[i].numpy(), p2)
        np.testing.assert_allclose(inps[i].numpy(), R.crop_box(fr, p1, p2), atol=2.4e-7)
    np.testing.assert_allclose(img[0].numpy(), orig[0].numpy() - np.float32(0.406), atol=1e-7)  # in-place mean subtraction
    np.testing.assert_allclose(img[2].numpy(), orig[2].numpy() - np.float32(0.480), atol=1e-7)


def test_kpd_seam_and_get_prediction(kpd_sd, frames8):
    from betapose_b200 import compat

    x = torch.from_numpy(np.stack([R.crop_box(frames8[i], *R.expand_box(np.float32([150, 80, 420, 400]), 640, 480)) for i in range(2)]))
    pose_model = compat.InferenNet_fast(5, 1, None, state_dict=kpd_sd).cuda().eval()
    hm = pose_model(x)
    assert hm.shape == (2, 50, 80, 64) and hm.device == x.device
    with torch.no_grad():
        ref = onets.fastpose_forward(kpd_sd, x)
    assert (hm - ref).abs().max().item() <= 3e-2 * ref.abs().max().item()
    pt1 = torch.tensor([[96.0, 16.0], [96.0, 16.0]])
    pt2 = torch.tensor([[474.0, 464.0], [474.0, 464.0]])
    ph, pi, mv = compat.getPrediction(hm, pt1, pt2, 320, 256, 80, 64)
    rh, ri, rm, _, _ = R.get_prediction(hm.numpy(), pt1.numpy(), pt2.numpy())
    assert np.array_equal(ph.numpy(), rh) and np.array_equal(mv.numpy(), rm)
    np.testing.assert_allclose(pi.numpy(), ri, atol=6.2e-5)


def test_pose_nms_and_pnp_seams(kp_model):
    from betapose_b200 import compat

    rng = np.random.default_rng(0)
    Rm = np.array([[0.36, 0.48, -0.8], [-0.8, 0.6, 0.0], [0.48, 0.64, 0.6]])
    t = np.array([0.03, -0.02, 0.8])
    pc = kp_model @ Rm.T + t
    uv = np.stack([R.CAM_K[0, 0] * pc[:, 0] / pc[:, 2] + R.CAM_K[0, 2], R.CAM_K[1, 1] * pc[:, 1] / pc[:, 2] + R.CAM_K[1, 2]], 1).astype(np.float32)
    uv += rng.normal(0, 0.3, uv.shape).astype(np.float32)
    scores = rng.uniform(0.4, 0.9, (1, 50, 1)).astype(np.float32)
    scores[0, 7, 0] = 0.0
    ps = torch.from_numpy(scores.copy())
    res = compat.pose_nms(torch.tensor([[10.0, 20.0, 200.0, 220.0]]), torch.tensor([[0.8]]), torch.from_numpy(uv[None] + np.float32(0.3)), ps)
    ref = R.pose_nms(np.float32([[10, 20, 200, 220]]), np.float32([[0.8]]), uv[None] + np.float32(0.3), scores)
    assert len(res) == len(ref) == 1 and float(ps[0, 7, 0]) == pytest.approx(1e-5)   # mutated in place like the reference
    assert np.array_equal(res[0]["keypoints"].numpy(), ref[0]["keypoints"])
    np.testing.assert_allclose(res[0]["kp_score"].numpy(), ref[0]["kp_score"], rtol=1e-6)
    np.testing.assert_allclose(res[0]["proposal_score"].numpy(), ref[0]["proposal_score"], rtol=1e-6)
    assert compat.pose_nms(torch.tensor([[10.0, 20.0, 200.0, 220.0]]), torch.tensor([[0.8]]), torch.from_numpy(uv[None]), torch.full((1, 50, 1), 0.2)) == []
    # pnp seam: numpy in, (R [3,3] f64, t [3,1] f64) out
    Rg, tg = compat.pnp(kp_model, uv, R.CAM_K)
    assert Rg.shape == (3, 3) and tg.shape == (3, 1) and Rg.dtype == np.float64
    sol = opnp.solve_pnp(kp_model, uv, R.CAM_K, mode=0, n_hyp=64, seed=0)
    np.testing.assert_allclose(Rg, sol["R"], atol=1e-6)
    np.testing.assert_allclose(tg[:, 0], sol["t"], atol=1e-6)
    assert np.abs(Rg - Rm).max() < 2e-2
    with pytest.raises(AssertionError):
        compat.pnp(kp_model[:10], uv, R.CAM_K)


@pytest.mark.parametrize("left", [50, 10])
def test_a12_datawriter_assembly_and_json_vs_reference_golden(tmp_path, left):
    """a12 through the product: the seams DataWriter.update calls (getPrediction -> pose_nms -> selection -> pnp) and
    write_json, and the engine's own tail (bp_pose_pnp -> bp_pack_records -> result_from_record -> write_json), on the
    heat-maps of tests/golden/a12_golden.npz; expected = the JSON the reference's own DataWriter + write_json produced."""
    import json
    import os

    from betapose_b200 import compat, stages
    from test_oracle_golden import check_a12_json, load

    g = load("a12_golden.npz")
    kp3d = g["kp3d"]
    names = [str(n) for n in g["names"]]
    # ---- (i) seam by seam, CPU tensors in, like DataWriter.update
    final = []
    for i, name in enumerate(names):
        hm = torch.from_numpy(g[f"hm{i}"].astype(np.float32))
        boxes, scores = torch.from_numpy(g[f"box{i}"][None].copy()), torch.tensor([[float(g[f"score{i}"])]])
        ph, pi, ps = compat.getPrediction(hm, torch.from_numpy(g[f"pt1_{i}"][None].copy()), torch.from_numpy(g[f"pt2_{i}"][None].copy()), 320, 256, 80, 64)
        result = {"imgname": name, "result": compat.pose_nms(boxes, scores, pi, ps)}
        if result["result"]:
            kp_score = np.array(result["result"][0]["kp_score"][:, 0])
            kp_2d = np.array(result["result"][0]["keypoints"])
            kp_3d = np.array(kp3d)
            while len(kp_2d) > left:               # dataloader.py:720-724 verbatim semantics
                d = np.argmin(kp_score, axis=0)
                kp_score, kp_2d, kp_3d = np.delete(kp_score, d), np.delete(kp_2d, d, axis=0), np.delete(kp_3d, d, axis=0)
            Rm, t = compat.pnp(kp_3d, kp_2d, R.CAM_K)
            result.update({"cam_R": Rm, "cam_t": t})
        else:
            result.update({"cam_R": [], "cam_t": []})
        final.append(result)
    compat.write_json(final, str(tmp_path / "seams"))
    check_a12_json(json.load(open(tmp_path / "seams" / "Betapose-results.json")), g, left, pose_tol=1e-5)
    # ---- (ii) the engine's tail in one batch: heat-map decode -> bp_pose_pnp (pose-NMS n=1, selection, PnP) -> records
    n = len(names)
    dev = torch.device("cuda")
    hm = torch.from_numpy(np.concatenate([g[f"hm{i}"].astype(np.float32) for i in range(n)])).to(dev)
    pt1 = torch.from_numpy(np.stack([g[f"pt1_{i}"] for i in range(n)])).to(dev)
    pt2 = torch.from_numpy(np.stack([g[f"pt2_{i}"] for i in range(n)])).to(dev)
    box = torch.from_numpy(np.stack([g[f"box{i}"] for i in range(n)])).to(dev)
    det = torch.tensor([float(g[f"score{i}"]) for i in range(n)], device=dev)
    dec = stages.heatmap_decode(hm, pt1, pt2, layout="nchw")
    pose = stages.pose_pnp(dec["preds_img"], dec["maxval"].reshape(n, 50).contiguous(), det, torch.from_numpy(kp3d).to(dev), left_number=left)
    rec = stages.records_to_numpy(stages.pack_records(0, box, det, pose))
    assert rec["status"].tolist() == [1, 1, 0, 1]
    res = [compat.result_from_record(rec[i], os.path.join("/data/rgb", names[i])) for i in range(n)]
    compat.write_json(res, str(tmp_path / "engine"))
    check_a12_json(json.load(open(tmp_path / "engine" / "Betapose-results.json")), g, left, pose_tol=1e-5)
