"""Weight-format tooling (host code): the two on-disk formats of the path and their checks.

* darknet `.weights` (detector): 16-byte header {major, minor, revision, seen} (int32) + one fp32 stream, per conv block
  in cfg order: [bn_bias, bn_weight, bn_mean, bn_var | bias] then the kernel [Cout, Cin, k, k]
  (3_6Dpose_estimator/yolo/darknet.py:365-432, the layout darknet's save_weights writes).
* PyTorch state_dict (key-point net): the 654 tensors of FastPose / SE-ResNet-101 + 2 DUC + head
  (KPD/src/models/FastPose.py:13-35); checkpoints written by torch 0.4 lack the
  `num_batches_tracked` entries.
Both are validated BEFORE anything is uploaded, with messages that name the block / key, instead of failing somewhere
inside a reshape."""
from __future__ import annotations

import numpy as np

from . import net as _net

FASTPOSE_PLANES = (64, 128, 256, 512)


def read_darknet_weights(path: str):
    """-> (header int32[4], stream fp32[...])."""
    with open(path, "rb") as f:
        header = np.fromfile(f, dtype=np.int32, count=4)
        stream = np.fromfile(f, dtype=np.float32)
    if header.size != 4:
        raise ValueError(f"{path}: shorter than the 16-byte darknet header")
    return header, stream


def write_darknet_weights(path: str, stream: np.ndarray, seen: int = 0, version=(0, 1, 0)) -> None:
    with open(path, "wb") as f:
        np.array([version[0], version[1], version[2], seen], np.int32).tofile(f)
        np.ascontiguousarray(stream, np.float32).tofile(f)


def darknet_stream_size(blocks) -> int:
    """number of fp32 values the cfg's conv blocks consume"""
    info = _net.infer_darknet_shapes(blocks)
    n = 0
    for i, b in enumerate(blocks):
        if b["type"] != "convolutional":
            continue
        cin = 3 if i == 0 else info[i - 1]["C"]
        cout, k = int(b["filters"]), int(b["size"])
        n += (4 * cout if int(b.get("batch_normalize", 0)) else cout) + cout * cin * k * k
    return n


def check_darknet_stream(blocks, stream: np.ndarray) -> None:
    need = darknet_stream_size(blocks)
    if stream.size < need:
        raise ValueError(f"darknet weights: {stream.size} values, the cfg needs {need} (truncated file or wrong cfg)")
    if stream.size > need:
        raise ValueError(f"darknet weights: {stream.size - need} values left over after the last conv block (wrong cfg?)")
    if not np.isfinite(stream).all():
        raise ValueError("darknet weights: non-finite values")


def darknet_stream_from_params(blocks, params) -> np.ndarray:
    """inverse of net.split_darknet_stream: per-block parameter dicts -> the fp32 stream (for writing a .weights file)."""
    out = []
    for b, p in zip(blocks, params):
        if b["type"] != "convolutional":
            continue
        if int(b.get("batch_normalize", 0)):
            out += [p["bn_bias"], p["bn_weight"], p["bn_mean"], p["bn_var"]]
        else:
            out.append(p["bias"])
        out.append(np.asarray(p["weight"], np.float32).reshape(-1))
    return np.concatenate([np.asarray(a, np.float32).reshape(-1) for a in out])


def fastpose_expected_shapes(n_out: int = 50) -> dict:
    """key -> shape of every learnable / running tensor of FastPose (without num_batches_tracked)."""
    sh = {"preact.conv1.weight": (64, 3, 7, 7)}

    def bn(pre, c):
        for k in ("weight", "bias", "running_mean", "running_var"):
            sh[f"{pre}.{k}"] = (c,)

    bn("preact.bn1", 64)
    inplanes = 64
    for li, (nb, planes) in enumerate(zip(_net.FASTPOSE_LAYERS, FASTPOSE_PLANES), start=1):
        for bi in range(nb):
            pre = f"preact.layer{li}.{bi}"
            sh[f"{pre}.conv1.weight"] = (planes, inplanes, 1, 1)
            bn(f"{pre}.bn1", planes)
            sh[f"{pre}.conv2.weight"] = (planes, planes, 3, 3)
            bn(f"{pre}.bn2", planes)
            sh[f"{pre}.conv3.weight"] = (planes * 4, planes, 1, 1)
            bn(f"{pre}.bn3", planes * 4)
            if bi == 0:
                sh[f"{pre}.downsample.0.weight"] = (planes * 4, inplanes, 1, 1)
                bn(f"{pre}.downsample.1", planes * 4)
                c = planes * 4
                sh[f"{pre}.se.fc.0.weight"], sh[f"{pre}.se.fc.0.bias"] = (c, c), (c,)
                sh[f"{pre}.se.fc.2.weight"], sh[f"{pre}.se.fc.2.bias"] = (c, c), (c,)
            inplanes = planes * 4
    sh["duc1.conv.weight"] = (1024, 512, 3, 3)
    bn("duc1.bn", 1024)
    sh["duc2.conv.weight"] = (512, 256, 3, 3)
    bn("duc2.bn", 512)
    sh["conv_out.weight"], sh["conv_out.bias"] = (n_out, 128, 3, 3), (n_out,)
    return sh


def check_fastpose_state_dict(sd: dict, n_out: int = 50) -> None:
    """raises with the first missing / mis-shaped key; `module.` prefixes (DataParallel checkpoints) are not accepted"""
    exp = fastpose_expected_shapes(n_out)
    for k, shape in exp.items():
        if k not in sd:
            raise ValueError(f"FastPose state_dict: missing '{k}'")
        got = tuple(sd[k].shape)
        if k.startswith("conv_out."):
            if got[0] < n_out or got[1:] != shape[1:]:
                raise ValueError(f"FastPose state_dict: '{k}' has shape {got}, need at least {shape}")
        elif got != shape:
            raise ValueError(f"FastPose state_dict: '{k}' has shape {got}, expected {shape}")
    extra = [k for k in sd if k not in exp and not k.endswith("num_batches_tracked")]
    if extra:
        raise ValueError(f"FastPose state_dict: unexpected key '{extra[0]}' ({len(extra)} in total)")
