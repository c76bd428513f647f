"""Seeded synthetic assets (no trained weights, images or ground truth are shipped with the reference, and there
is no network here): detector weights in darknet ``.weights`` layout, keypoint-network weights as a state_dict
with the reference's 654 keys, 640x480 RGB frames, PnP test geometry.

Plain i.i.d. random weights make both networks almost input-independent after ~100 layers, so
BatchNorm running statistics are *calibrated*: each conv's running_mean/var are set to the statistics of its own
output on a small seeded calibration batch (LSUV-style), and the last BN of every residual branch gets a small
gain so activations neither explode nor die.  The result behaves like a trained network numerically (unit-scale
activations, input-dependent heads, fp16-safe) while staying reproducible from a seed.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F

from . import yolo_cfg

BN_EPS = 1e-5


def _rng(seed):
    return np.random.default_rng(seed)


def synth_frames(n: int, seed: int = 0, h: int = 480, w: int = 640) -> np.ndarray:
    """uint8 [n,h,w,3] RGB: smooth random blobs + texture + noise (so resize / crop see real structure)."""
    r = _rng(seed)
    out = np.empty((n, h, w, 3), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for i in range(n):
        img = np.zeros((h, w, 3), np.float32)
        for _ in range(6):
            cx, cy = r.uniform(0, w), r.uniform(0, h)
            s = r.uniform(30, 160)
            col = r.uniform(0, 255, 3).astype(np.float32)
            g = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
            img += g[..., None] * col
        fx, fy = r.uniform(0.02, 0.3, 2)
        img += 30 * np.sin(xx * fx + yy * fy)[..., None]
        img += r.normal(0, 12, (h, w, 3)).astype(np.float32)
        out[i] = np.clip(img, 0, 255).astype(np.uint8)
    return out


def _calib_frames(n: int, seed: int, out_h: int, out_w: int, crop: bool = False) -> torch.Tensor:
    """Calibration inputs drawn from the same distribution as `synth_frames` (resized; for the keypoint net a
    centre crop with the reference's channel means removed)."""
    fr = torch.from_numpy(synth_frames(n, seed)).permute(0, 3, 1, 2).float() / 255.0
    if crop:
        fr = fr[:, :, 90:390, 200:440] - torch.tensor([0.406, 0.457, 0.480])[None, :, None, None]
    return F.interpolate(fr, size=(out_h, out_w), mode="bilinear", align_corners=True)


# ------------------------------------------------------------------------------------------------ detector
def _calib_bn(y: torch.Tensor, r, gain_scale: float = 1.0):
    c = y.shape[1]
    mean = y.mean(dim=(0, 2, 3))
    var = y.var(dim=(0, 2, 3), unbiased=False).clamp_min(1e-6)
    gamma = torch.from_numpy(r.uniform(0.7, 1.3, c).astype(np.float32)) * gain_scale
    beta = torch.from_numpy(r.normal(0, 0.15, c).astype(np.float32))
    out = (y - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + BN_EPS)
    out = out * gamma[None, :, None, None] + beta[None, :, None, None]
    return out, gamma, beta, mean, var


def synth_yolo_weights(seed: int = 1000, blocks: list[dict] | None = None, reso: int = 416, calib_n: int = 12):
    """-> fp32 stream (darknet order, without the 16-byte header)."""
    blocks = blocks if blocks is not None else yolo_cfg.parse_cfg_text(yolo_cfg.default_cfg_text())
    r = _rng(seed)
    x = _calib_frames(calib_n, seed + 7, reso, reso)
    chunks: list[np.ndarray] = []
    outputs: dict[int, torch.Tensor] = {}
    with torch.no_grad():
        for i, b in enumerate(blocks):
            t = b["type"]
            if t == "convolutional":
                cin, cout, k = x.shape[1], int(b["filters"]), int(b["size"])
                pad = (k - 1) // 2 if int(b["pad"]) else 0
                fan_in = cin * k * k
                w = torch.from_numpy(r.normal(0, 1.0 / np.sqrt(fan_in), (cout, cin, k, k)).astype(np.float32))
                y = F.conv2d(x, w, None, int(b["stride"]), pad)
                if int(b.get("batch_normalize", 0)):
                    nxt = blocks[i + 1]["type"] if i + 1 < len(blocks) else ""
                    y, gamma, beta, mean, var = _calib_bn(y, r, 0.35 if nxt == "shortcut" else 1.0)
                    chunks += [beta.numpy(), gamma.numpy(), mean.numpy(), var.numpy()]
                else:
                    # detection head: unit-scale logits, objectness spread wide enough to have a clear winner
                    # (box logits kept small so exp(tw), exp(th) stay sane, as in a trained detector)
                    std = y.std(dim=(0, 2, 3)).clamp_min(1e-6)
                    nattr = cout // 3
                    target = torch.tensor([0.15 if (c % nattr) < 4 else 0.35 for c in range(cout)])
                    w = w * (target / std)[:, None, None, None]
                    y = y * (target / std)[None, :, None, None]
                    bias = torch.from_numpy(r.normal(0, 0.1, cout).astype(np.float32))
                    y = y + bias[None, :, None, None]
                    chunks.append(bias.numpy())
                chunks.append(w.numpy().reshape(-1))
                if b["activation"] == "leaky":
                    y = F.leaky_relu(y, 0.1)
                x = y
            elif t == "upsample":
                x = F.interpolate(x, scale_factor=int(b["stride"]), mode="nearest")
            elif t == "shortcut":
                x = outputs[i - 1] + outputs[i + int(b["from"])]
            elif t == "route":
                ls = [int(v) for v in b["layers"].split(",")] if isinstance(b["layers"], str) else [int(v) for v in b["layers"]]
                x = outputs[i + ls[0]] if len(ls) == 1 else torch.cat((outputs[i + ls[0]], outputs[ls[1]]), 1)
            elif t == "yolo":
                x = outputs[i - 1]
            outputs[i] = x
    return np.concatenate([c.astype(np.float32).reshape(-1) for c in chunks])


def write_darknet_weights(path: str, stream: np.ndarray, seen: int = 0) -> None:
    """16-byte header {major=0, minor=1, revision=0, seen} + fp32 stream (what darknet.py:377-380 reads)."""
    with open(path, "wb") as f:
        np.array([0, 1, 0, seen], np.int32).tofile(f)
        stream.astype(np.float32).tofile(f)


# ------------------------------------------------------------------------------------------------ keypoint net
FASTPOSE_LAYERS = (3, 4, 23, 3)
FASTPOSE_PLANES = (64, 128, 256, 512)


def synth_kpd_state_dict(seed: int = 2000, n_classes: int = 50, calib_n: int = 12) -> dict:
    """state_dict with the reference FastPose keys, torch fp32 tensors."""
    r = _rng(seed)
    sd: dict[str, torch.Tensor] = {}
    x = _calib_frames(calib_n, seed + 7, 320, 256, crop=True)

    def conv_w(name, cout, cin, k):
        w = torch.from_numpy(r.normal(0, 1.0 / np.sqrt(cin * k * k), (cout, cin, k, k)).astype(np.float32))
        sd[name] = w
        return w

    def bn(name, y, gain=1.0):
        out, gamma, beta, mean, var = _calib_bn(y, r, gain)
        sd[name + ".weight"], sd[name + ".bias"] = gamma, beta
        sd[name + ".running_mean"], sd[name + ".running_var"] = mean, var
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        return out

    with torch.no_grad():
        x = F.relu(bn("preact.bn1", F.conv2d(x, conv_w("preact.conv1.weight", 64, 3, 7), None, 2, 3)))
        x = F.max_pool2d(x, 3, 2, 1)
        inplanes = 64
        for li, (nb, planes) in enumerate(zip(FASTPOSE_LAYERS, FASTPOSE_PLANES), start=1):
            for bi in range(nb):
                pre = f"preact.layer{li}.{bi}"
                stride = 2 if (bi == 0 and li > 1) else 1
                o = F.relu(bn(pre + ".bn1", F.conv2d(x, conv_w(pre + ".conv1.weight", planes, inplanes, 1))))
                o = F.relu(bn(pre + ".bn2", F.conv2d(o, conv_w(pre + ".conv2.weight", planes, planes, 3), None, stride, 1)))
                o = bn(pre + ".bn3", F.conv2d(o, conv_w(pre + ".conv3.weight", planes * 4, planes, 1)), 0.35)
                res = x
                if bi == 0:
                    c = planes * 4
                    for j in (0, 2):
                        sd[f"{pre}.se.fc.{j}.weight"] = torch.from_numpy(
                            r.normal(0, 1.0 / np.sqrt(c), (c, c)).astype(np.float32))
                        sd[f"{pre}.se.fc.{j}.bias"] = torch.from_numpy(r.normal(0, 0.2, c).astype(np.float32))
                    y = o.mean(dim=(2, 3))
                    y = F.relu(F.linear(y, sd[pre + ".se.fc.0.weight"], sd[pre + ".se.fc.0.bias"]))
                    y = torch.sigmoid(F.linear(y, sd[pre + ".se.fc.2.weight"], sd[pre + ".se.fc.2.bias"]))
                    o = o * y[:, :, None, None]
                    res = bn(pre + ".downsample.1",
                             F.conv2d(x, conv_w(pre + ".downsample.0.weight", planes * 4, inplanes, 1), None, stride))
                x = F.relu(o + res)
                inplanes = planes * 4
        x = F.pixel_shuffle(x, 2)
        for name, cin, cout in (("duc1", 512, 1024), ("duc2", 256, 512)):
            x = F.pixel_shuffle(F.relu(bn(name + ".bn", F.conv2d(x, conv_w(name + ".conv.weight", cout, cin, 3), None, 1, 1))), 2)
        w = conv_w("conv_out.weight", n_classes, 128, 3)
        y = F.conv2d(x, w, None, 1, 1)
        std = y.std().clamp_min(1e-6)
        sd["conv_out.weight"] = w / (std * 2.0)  # heat-map logits ~ N(0, 0.5): peaks comfortably above 0.3
        sd["conv_out.bias"] = torch.from_numpy(r.normal(0, 0.05, n_classes).astype(np.float32))
    return sd


# ------------------------------------------------------------------------------------------------ caching
def cache_dir() -> str:
    d = os.environ.get("BETAPOSE_B200_CACHE", os.path.join(os.path.expanduser("~"), ".cache", "betapose_b200"))
    os.makedirs(d, exist_ok=True)
    return d


def _publish(tmp: str, final: str) -> None:
    """atomic: several ranks of one node may build the same cache entry at the same time (torchrun on a fresh box)"""
    os.replace(tmp, final)


def cached_yolo_weights(seed: int = 1000) -> np.ndarray:
    p = os.path.join(cache_dir(), f"yolo_synth_{seed}.npy")
    if os.path.exists(p):
        try:
            return np.load(p)
        except Exception:
            pass  # unreadable entry: rebuild
    s = synth_yolo_weights(seed)
    tmp = f"{p}.{os.getpid()}.tmp.npy"
    np.save(tmp, s)
    _publish(tmp, p)
    return s


def cached_kpd_state_dict(seed: int = 2000) -> dict:
    p = os.path.join(cache_dir(), f"kpd_synth_{seed}.pt")
    if os.path.exists(p):
        try:
            return torch.load(p)
        except Exception:
            pass
    sd = synth_kpd_state_dict(seed)
    tmp = f"{p}.{os.getpid()}.tmp"
    torch.save(sd, tmp)
    _publish(tmp, p)
    return sd


# ------------------------------------------------------------------------------------------------ per-object variants
# BASELINE.json configs[3] mixes the 13 LineMod objects, each with its own detector and key-point network.  Calibrating 13
# random weight sets costs ~15 s of CPU each; instead object v > 0 gets object 0's networks seen through a fixed input
# symmetry: every convolution kernel mirrored left-right and / or up-down and the colour channels of the stem permuted,
#   f_v(x) ~ mirror(f_0(mirror(permute_rgb(x)))).
# The weights are different bytes and the networks different functions of the frame, while every layer keeps the
# activation statistics object 0 was calibrated to (the synthetic frames have no preferred direction or colour), so the
# variants are as fp16-safe and input-dependent as the original.  (Mirroring is exact only up to the one-pixel alignment
# shift of stride-2 layers on even sizes; that does not matter for statistics.)
_RGB_PERMS = ((0, 1, 2), (1, 2, 0), (2, 0, 1), (0, 2, 1), (2, 1, 0), (1, 0, 2))


def _variant_ops(v: int):
    v = int(v)
    return _RGB_PERMS[v % 6], bool((v // 6) & 1) != bool(v & 1), bool((v // 12) & 1) != bool((v >> 1) & 1)  # perm, flip_lr, flip_ud


def _variant_kernel(w: np.ndarray, flip_lr: bool, flip_ud: bool) -> np.ndarray:
    if flip_lr:
        w = w[:, :, :, ::-1]
    if flip_ud:
        w = w[:, :, ::-1, :]
    return w


def variant_yolo_weights(stream: np.ndarray, v: int, blocks: list[dict] | None = None) -> np.ndarray:
    """darknet fp32 stream of object-variant v (v = 0: the stream itself)."""
    if v == 0:
        return stream
    from . import net as _net, weights as _weights

    blocks = blocks if blocks is not None else yolo_cfg.parse_cfg_text(yolo_cfg.default_cfg_text())
    perm, lr, ud = _variant_ops(v)
    params, used = _net.split_darknet_stream(blocks, np.asarray(stream, np.float32))
    assert used == stream.size
    first = True
    out = []
    for p in params:
        if p is None:
            out.append(None)
            continue
        q = dict(p)
        w = _variant_kernel(p["weight"], lr, ud)
        if first:
            w = w[:, list(perm)]
            first = False
        q["weight"] = np.ascontiguousarray(w)
        out.append(q)
    return _weights.darknet_stream_from_params(blocks, out)


def variant_kpd_state_dict(sd: dict, v: int) -> dict:
    """FastPose state_dict of object-variant v (v = 0: the dict itself)."""
    if v == 0:
        return sd
    perm, lr, ud = _variant_ops(v)
    out = {}
    for k, t in sd.items():
        if t.dim() == 4:
            w = t
            if lr:
                w = w.flip(3)
            if ud:
                w = w.flip(2)
            if k == "preact.conv1.weight":
                w = w[:, list(perm)]
            out[k] = w.contiguous()
        else:
            out[k] = t
    return out


# ------------------------------------------------------------------------------------------------ keypoint models
def synth_kp_model(seed: int = 1, n: int = 50, radius: float = 0.045) -> np.ndarray:
    """[n,3] float64 metres: points on a bumpy ellipsoid of LineMod-object size (the 13 designated-keypoint PLYs
    live in the reference tree and are only available where it is mounted)."""
    r = _rng(seed)
    v = r.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v *= np.array([1.0, 0.85, 1.2]) * radius * r.uniform(0.6, 1.0, (n, 1))
    return v.astype(np.float64)
