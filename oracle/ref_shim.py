"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference (sjtuytc/betapose).

Loads the reference's own Python modules from ``$BETAPOSE_REF`` or ``/root/reference`` so that
``oracle/restate.py`` can be pinned against them and golden vectors can be generated
(``tests/golden/make_golden.py``).  Nothing is copied: source text is read, patched *in memory*
(2 textual patches) and exec'd.  The reference tree does not exist on the GPU box, so nothing that runs
there (``-m gpu`` tests, ``smoke()``, ``bench.py``) may import this module; callers must check
``available()`` first.

Patches:
  * ``KPD/src/utils/img.py:268,313``  ``x.cuda(async=True)`` is a SyntaxError on py>=3.7 -> ``non_blocking=True``
  * ``yolo/darknet.py:152-154,160,327``, ``yolo/bbox.py:69``  hard ``.cuda()`` calls -> stripped for CPU runs
Stubs: IPython.embed, matplotlib(.pyplot), visdom, torchsample.transforms.{SpecialCrop,Pad} (un-vendored
third-party; restated from ``KPD/src/utils/img.py:198-213`` which proves centred padding).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

_E = "3_6Dpose_estimator"


def ref_root() -> str | None:
    for cand in (os.environ.get("BETAPOSE_REF"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, _E, "opt.py")):
            return cand
    return None


def available() -> bool:
    return ref_root() is not None


_loaded: dict[str, types.ModuleType] = {}


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


class _SpecialCrop:
    """torchsample.transforms.SpecialCrop(size, crop_type=1): top-left crop (silently clipped)."""

    def __init__(self, size, crop_type=0):
        assert crop_type == 1
        self.size = size

    def __call__(self, x):
        h, w = int(self.size[0]), int(self.size[1])
        return x[:, :h, :w]


class _Pad:
    """torchsample.transforms.Pad(size): centred zero pad of a CHW tensor, ceil before / floor after."""

    def __init__(self, size):
        self.size = size

    def __call__(self, x):
        import torch

        a = x.numpy()
        shape_diffs = [int(np.ceil((int(s) - a_s))) for s, a_s in zip(self.size, a.shape)]
        shape_diffs = np.maximum(shape_diffs, 0)
        pad_sizes = [(int(np.ceil(s / 2.0)), int(np.floor(s / 2.0))) for s in shape_diffs]
        return torch.from_numpy(np.pad(a, pad_sizes, mode="constant"))


def _install_stubs() -> None:
    _stub("IPython", embed=lambda *a, **k: None)
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    _stub("visdom")
    ts = _stub("torchsample")
    tst = _stub("torchsample.transforms", SpecialCrop=_SpecialCrop, Pad=_Pad)
    ts.transforms = tst


def _load(dotted: str, relpath: str, cpu: bool = True, aliases=()) -> types.ModuleType:
    if dotted in _loaded:
        return _loaded[dotted]
    root = ref_root()
    assert root, "reference tree not found"
    path = os.path.join(root, _E, relpath)
    src = open(path).read()
    src = src.replace("async=True", "non_blocking=True")
    if cpu:
        src = src.replace(".cuda()", "")
    mod = types.ModuleType(dotted)
    mod.__file__ = path
    # make parent packages resolvable for `from a.b import c`
    parts = dotted.split(".")
    for i in range(1, len(parts)):
        pk = ".".join(parts[:i])
        if pk not in sys.modules:
            p = types.ModuleType(pk)
            p.__path__ = []
            sys.modules[pk] = p
    sys.modules[dotted] = mod
    for a in aliases:
        sys.modules[a] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    if len(parts) > 1:
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)
    _loaded[dotted] = mod
    return mod


def load_reference() -> types.SimpleNamespace:
    """Returns a namespace with the reference's live-path symbols (CPU)."""
    if "ns" in _loaded:
        return _loaded["ns"]  # type: ignore[return-value]
    _install_stubs()
    argv = sys.argv
    sys.argv = ["x", "--sp"]
    try:
        opt_mod = _load("opt", "opt.py")
    finally:
        sys.argv = argv
    bbox = _load("yolo.bbox", "yolo/bbox.py", aliases=("bbox",))
    util = _load("yolo.util", "yolo/util.py", aliases=("util",))
    darknet = _load("yolo.darknet", "yolo/darknet.py")
    se_mod = _load("KPD.src.models.layers.SE_module", "KPD/src/models/layers/SE_module.py")
    se_res = _load("KPD.src.models.layers.SE_Resnet", "KPD/src/models/layers/SE_Resnet.py")
    duc = _load("KPD.src.models.layers.DUC", "KPD/src/models/layers/DUC.py")
    fastpose = _load("KPD.src.models.FastPose", "KPD/src/models/FastPose.py")
    img = _load("KPD.src.utils.img", "KPD/src/utils/img.py", aliases=("utils.img",))
    ev = _load("KPD.src.utils.eval", "KPD/src/utils/eval.py")
    nms = _load("pPose_nms", "pPose_nms.py")

    # crop_from_dets: exec only the tail of dataloader.py (importing all of it drags in renderer/vispy)
    import torch

    root = ref_root()
    dl_src = open(os.path.join(root, _E, "dataloader.py")).read()
    tail = dl_src[dl_src.index("def crop_from_dets"):]
    dl_ns = {"torch": torch, "cropBox": img.cropBox, "opt": opt_mod.opt}
    exec(compile(tail, "dataloader.py<crop_from_dets>", "exec"), dl_ns)

    ns = types.SimpleNamespace(
        opt=opt_mod.opt,
        Darknet=darknet.Darknet,
        parse_cfg=darknet.parse_cfg,
        dynamic_write_results=util.dynamic_write_results,
        write_results=util.write_results,
        FastPose=fastpose.FastPose,
        createModel=fastpose.createModel,
        cropBox=img.cropBox,
        im_to_torch=img.im_to_torch,
        transformBoxInvert_batch=img.transformBoxInvert_batch,
        getPrediction=ev.getPrediction,
        pose_nms=nms.pose_nms,
        write_json=nms.write_json,
        crop_from_dets=dl_ns["crop_from_dets"],
        cfg_path=os.path.join(root, _E, "yolo/cfg/yolov3-single.cfg"),
        sift_dir=os.path.join(root, "1_keypoint_designator/assets/sifts"),
    )
    _loaded["ns"] = ns  # type: ignore[assignment]
    return ns


def load_box_nms():
    """`write_results` of yolo/util.py with its shipped IoU-NMS branch switched back on: the two statements that disable
    it (`nms = False`, :181) and that afterwards keep only the arg-max row (:210-211) are neutralised IN MEMORY; every other
    line, `bbox_iou` included, runs as written.  Used only to pin oracle/restate.py:write_results_nms."""
    load_reference()
    root = ref_root()
    src = open(os.path.join(root, _E, "yolo/util.py")).read().replace(".cuda()", "")
    a = "            nms = False\n"
    b = "            best_idx = np.argmax(out[:,5])\n            out = out[best_idx].view(1,-1)\n"
    assert src.count(a) == 1 and src.count(b) == 1, "reference yolo/util.py changed"
    src = src.replace(a, "").replace(b, "")
    mod = types.ModuleType("yolo.util_box_nms")
    exec(compile(src, "yolo/util.py<box-nms on>", "exec"), mod.__dict__)
    return mod.write_results


def load_metrics() -> types.ModuleType:
    """The reference's scoring functions (3_6Dpose_estimator/utils/metrics.py) unmodified; pyquaternion (only used by
    rot_error, which the evaluate loop never calls) is stubbed."""
    _install_stubs()
    if "pyquaternion" not in sys.modules:
        _stub("pyquaternion", Quaternion=object)
    return _load("utils.metrics", "utils/metrics.py")


def load_pnp():
    """The reference's own ``pnp`` (3_6Dpose_estimator/utils/utils.py:17-41): the module cannot be imported (it pulls
    renderer / vispy / scipy.misc at import), so the text of that one function is exec'd as written."""
    import cv2

    root = ref_root()
    src = open(os.path.join(root, _E, "utils/utils.py")).read()
    a = src.index("def pnp(")
    b = src.index("    return R, t", a) + len("    return R, t")
    ns = {"np": np, "cv2": cv2}
    exec(compile(src[a:b] + "\n", "utils/utils.py<pnp>", "exec"), ns)
    return ns["pnp"]


def load_datawriter():
    """``class DataWriter`` of 3_6Dpose_estimator/dataloader.py:649-763 exec'd as written (the module as a whole drags in
    renderer / vispy), with the names it uses bound to the reference's own functions: getPrediction, pose_nms, pnp, opt.
    Used to generate the a12 golden (result assembly + write_json), tests/golden/make_golden.py."""
    import json
    import time
    from queue import Queue
    from threading import Thread

    import cv2
    import torch

    ref = load_reference()
    root = ref_root()
    src = open(os.path.join(root, _E, "dataloader.py")).read()
    a = src.index("class DataWriter:")
    b = src.index("class Mscoco", a)
    ns = {"np": np, "cv2": cv2, "os": os, "time": time, "json": json, "torch": torch, "Queue": Queue, "Thread": Thread,
          "opt": ref.opt, "getPrediction": ref.getPrediction, "pose_nms": ref.pose_nms, "pnp": load_pnp(), "write_json": ref.write_json}
    exec(compile(src[a:b], "dataloader.py<DataWriter>", "exec"), ns)
    return ns["DataWriter"]


def ref_pnp(points_3D, points_2D, cameraMatrix):
    """The reference's ``pnp`` body (3_6Dpose_estimator/utils/utils.py:17-41) cannot be imported (the module
    pulls renderer/vispy at import); its two live statements are the cv2 calls below."""
    import cv2

    dist = np.zeros((8, 1), dtype="float32")
    _, rvec, t = cv2.solvePnP(points_3D, np.ascontiguousarray(points_2D[:, :2]).reshape((-1, 1, 2)), cameraMatrix, dist)
    R, _ = cv2.Rodrigues(rvec)
    return R, t


def ref_pnp_ransac(points_3D, points_2D, cameraMatrix):
    """The commented variant (utils/utils.py:32-36) -- the PnP oracle of record."""
    import cv2

    dist = np.zeros((8, 1), dtype="float32")
    ok, rvec, t, inl = cv2.solvePnPRansac(
        points_3D, np.ascontiguousarray(points_2D[:, :2], dtype=np.float32), cameraMatrix, dist, reprojectionError=12.0
    )
    R, _ = cv2.Rodrigues(rvec)
    return R, t, inl
