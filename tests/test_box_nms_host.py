"""CPU: the product's block-level IoU-NMS (betapose_b200/csrc/box_nms.cuh -- the source the CUDA kernel instantiates --
compiled for the host by `make pnp_host` with a one-thread block) against the oracle restatement and the goldens from the
reference's own branch.  Bit-exact: detections, order, fp32 values, rows, counts."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import restate as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "betapose_b200", "csrc", "build", "libbp_box_nms_host.so")
G = os.path.join(ROOT, "tests", "golden", "box_nms_golden.npz")


@pytest.fixture(scope="module")
def host():
    if not os.path.isfile(SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "betapose_b200", "csrc"), "pnp_host"])
    L = C.CDLL(SO)
    L.bp_box_nms_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int] + [C.c_void_p] * 4
    L.bp_box_iou_host.argtypes = [C.c_void_p, C.c_void_p]
    L.bp_box_iou_host.restype = C.c_float
    return L


def run(L, pred, conf, thr, max_det):
    pred = np.ascontiguousarray(pred, np.float32)
    B, Rn, A = pred.shape
    det = np.zeros((B, max_det, 8), np.float32)
    row = np.full((B, max_det), -1, np.int32)
    cnt = np.zeros(B, np.int32)
    tot = np.zeros(B, np.int32)
    assert L.bp_box_nms_host(pred.ctypes.data, B, Rn, A, conf, thr, max_det, det.ctypes.data, row.ctypes.data, cnt.ctypes.data, tot.ctypes.data) == 0
    return det, row, cnt, tot


def flat(det, row, cnt):
    return (np.concatenate([det[b, : cnt[b]] for b in range(len(cnt))]).reshape(-1, 8),
            np.concatenate([row[b, : cnt[b]] for b in range(len(cnt))]))


def test_host_build_matches_reference_goldens(host):
    g = np.load(G)
    for i in range(int(g["n_cases"])):
        pred, want = g[f"pred{i}"], g[f"dets{i}"]
        conf, thr = float(g[f"conf{i}"]), float(g[f"thr{i}"])
        det, row, cnt, tot = run(host, pred, conf, thr, max_det=pred.shape[1])
        d, r = flat(det, row, cnt)
        assert np.array_equal(d, want), i
        assert np.array_equal(cnt, tot)
        _, rows, counts = R.write_results_nms(pred, conf, thr)
        assert np.array_equal(r, rows) and np.array_equal(cnt, counts)


def test_host_build_cap_ties_multiclass_and_degenerate_boxes(host):
    rng = np.random.default_rng(3)
    g = np.load(G)
    pred = g["pred0"]
    # max_det caps what is written, not what is counted
    det, row, cnt, tot = run(host, pred, 0.01, 0.6, max_det=10)
    full, frow, fcnt, _ = run(host, pred, 0.01, 0.6, max_det=300)
    assert np.array_equal(cnt, np.minimum(fcnt, 10)) and np.array_equal(tot, fcnt)
    for b in range(len(pred)):
        assert np.array_equal(det[b, : cnt[b]], full[b, :10]) and np.array_equal(row[b, : cnt[b]], frow[b, :10])
    # ties in objectness: lower row first (the documented tie-break), identical boxes suppress each other
    p = np.zeros((1, 40, 6), np.float32)
    p[0, :, :4] = [100, 100, 50, 50]
    p[0, :, 4] = 0.5
    p[0, :, 5] = 0.9
    p[0, 20:, 0] = 300  # a second group, far away
    det, row, cnt, _ = run(host, p, 0.1, 0.6, 40)
    assert cnt[0] == 2 and list(row[0, :2]) == [0, 20]
    d2, r2, c2 = R.write_results_nms(p, 0.1, 0.6)
    assert np.array_equal(det[0, :2], d2) and list(r2) == [0, 20]
    # multi-class rows: only rows whose class arg-max is 0 are candidates (yolo/util.py:166-167)
    q = rng.uniform(0, 1, (2, 200, 9)).astype(np.float32)
    q[..., :2] *= 400
    q[..., 2:4] = q[..., 2:4] * 60 + 5
    det, row, cnt, _ = run(host, q, 0.2, 0.5, 200)
    keep = q[..., 5:].argmax(-1) == 0
    q1 = np.where(keep[..., None], q, 0)[..., :6].astype(np.float32)
    d2, r2, c2 = R.write_results_nms(q1, 0.2, 0.5)
    d, r = flat(det, row, cnt)
    assert np.array_equal(d, d2) and np.array_equal(r, r2) and np.array_equal(cnt, c2)
    # zero-size and negative-size boxes, NaN objectness (never a candidate), negative confidence threshold
    z = rng.uniform(0, 1, (1, 120, 6)).astype(np.float32)
    z[0, :, :2] *= 50
    z[0, :40, 2:4] = 0
    z[0, 40:80, 2:4] *= -30
    z[0, 80:, 2:4] *= 30
    z[0, 5, 4] = np.nan
    z[0, 7, 4] = 0.0
    det, row, cnt, _ = run(host, z, -1.0, 0.3, 120)
    d2, r2, c2 = R.write_results_nms(z, -1.0, 0.3)
    d, r = flat(det, row, cnt)
    assert np.array_equal(r, r2) and np.array_equal(d, d2, equal_nan=True) and 5 not in r and 7 in r
    # nothing above the threshold
    det, row, cnt, tot = run(host, z, 2.0, 0.3, 8)
    assert cnt[0] == 0 and tot[0] == 0


def test_host_iou_is_the_reference_formula(host):
    rng = np.random.default_rng(4)
    for _ in range(2000):
        a = np.sort(rng.uniform(0, 400, 4).astype(np.float32).reshape(2, 2), 0).T.reshape(-1)[[0, 2, 1, 3]].copy()
        b = (a + rng.normal(0, 30, 4)).astype(np.float32)
        got = host.bp_box_iou_host(a.ctypes.data, b.ctypes.data)
        want = R.bbox_iou_plus1(a, b[None])[0]
        assert got == want or (np.isnan(got) and np.isnan(want))


def test_host_build_random_sweep_against_oracle(host):
    """Many small random scenes (heavy overlap, coarse coordinates so that IoUs land exactly on the threshold, duplicated
    objectness values): detections, order and counts equal the oracle's every time."""
    rng = np.random.default_rng(6)
    for trial in range(120):
        Rn = int(rng.integers(1, 90))
        p = np.zeros((2, Rn, 6), np.float32)
        p[..., 0:2] = rng.integers(0, 12, (2, Rn, 2)) * 8.0
        p[..., 2:4] = rng.integers(1, 6, (2, Rn, 2)) * 8.0 - 1.0      # (w+1)(h+1) areas are multiples of 64: exact IoU fractions
        p[..., 4] = rng.integers(1, 20, (2, Rn)) / 20.0 if trial % 2 else rng.uniform(0, 1, (2, Rn))
        p[..., 5] = rng.uniform(0, 1, (2, Rn))
        conf = float(rng.choice([0.0, 0.1, 0.5]))
        thr = float(rng.choice([0.25, 0.5, 0.6, 1.0 / 3.0]))
        det, row, cnt, tot = run(host, p, conf, thr, Rn)
        d2, r2, c2 = R.write_results_nms(p, conf, thr)
        d, r = flat(det, row, cnt)
        assert np.array_equal(cnt, c2) and np.array_equal(tot, c2), trial
        assert np.array_equal(r, r2) and np.array_equal(d, d2), trial
