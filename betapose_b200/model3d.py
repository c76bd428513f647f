"""Key-point model loading (a13): ASCII PLY -> [N,3] float64 metres and the greedy closest-pair refinement down to
nClasses points.  Replaces Model3D.load / Model3D.refine (3_6Dpose_estimator/utils/model.py:29-46,79-85) and the
key-point part of load_sixd_models (betapose_evaluate.py:53-84).  Load-time host code."""
from __future__ import annotations

import numpy as np

CAM_K = np.array([[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]], np.float64)
"""LineMod intrinsics the reference hard-codes (betapose_evaluate.py:59)."""


def load_ply(path: str, scale: float = 0.001) -> np.ndarray:
    with open(path, "rb") as f:
        head = f.read(4096)
    if b"format ascii" not in head:
        raise ValueError(f"{path}: only ASCII PLY key-point models are supported")
    with open(path) as f:
        lines = f.read().split("\n")
    nv, i = 0, 0
    while lines[i].strip() != "end_header":
        t = lines[i].split()
        if len(t) == 3 and t[0] == "element" and t[1] == "vertex":
            nv = int(t[2])
        i += 1
    v = np.array([[float(x) for x in ln.split()[:3]] for ln in lines[i + 1:i + 1 + nv]], np.float64)
    return v * scale


def refine(vertices: np.ndarray, n_keep: int) -> np.ndarray:
    """Repeatedly delete one point of the closest pair until n_keep remain (no-op for the shipped 50-point models)."""
    v = np.array(vertices, np.float64, copy=True)
    while v.shape[0] > n_keep:
        d = np.sqrt(((v[:, None] - v[None]) ** 2).sum(axis=2))
        np.fill_diagonal(d, np.inf)
        i, j = np.unravel_index(np.argmin(d), d.shape)
        if not d[i, j] < 100.0:
            i = 0
        v = np.delete(v, i, axis=0)
    return v


def load_kp_model(path: str, n_classes: int = 50, short: str = "error") -> np.ndarray:
    """Key-point model of one object: PLY x 0.001 -> refine(n_classes) (betapose_evaluate.py:75-83).

    A model with FEWER than n_classes points cannot be used by the reference at all: its pnp() asserts equal counts of
    3-D and 2-D points (utils/utils.py:23) and the key-point network always predicts n_classes maps.  That is the case
    of the one shipped PLY with 17 vertices, 1_keypoint_designator/assets/sifts/10.ply (LineMod object 10, the
    `Semmetry_obj10` entry of KPD/src/main_fast_inference.py:29-32).
      short = "error" (default): refuse at load time instead of failing per frame, like the reference would;
      short = "cycle": pad to n_classes by cycling, kp3d[k] = ply[k % n_points] -- heat-map k is then read as a detector
                       of model point k % n_points.  This is the documented choice for running all 13 objects from the
                       shipped fixtures (BASELINE.json configs[3]; SURVEY.md 8(d) "pad to 50 by cycling or exclude")."""
    v = refine(load_ply(path), n_classes)
    if v.shape[0] != n_classes:
        if short == "cycle" and 0 < v.shape[0] < n_classes:
            return pad_by_cycling(v, n_classes)
        # the reference's pnp() asserts equal counts (utils/utils.py:23); fail at load time instead of per frame
        raise ValueError(f"{path}: {v.shape[0]} key-points, the key-point network predicts {n_classes}")
    return v


def pad_by_cycling(v: np.ndarray, n_classes: int) -> np.ndarray:
    v = np.asarray(v, np.float64)
    return v[np.arange(n_classes) % v.shape[0]].copy()
