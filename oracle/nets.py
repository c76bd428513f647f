"""TEST INFRASTRUCTURE ONLY -- torch fp32 (CPU) functional restatement of the two networks on the path.

  * darknet_forward   <- 3_6Dpose_estimator/yolo/darknet.py:223-363 (build_model + forward) and :365-432
                         (load_weights: 16-byte header, then per conv [bn_bias, bn_weight, bn_mean, bn_var] or
                         [conv_bias], then weights [Cout,Cin,kh,kw])
  * fastpose_forward  <- KPD/src/models/FastPose.py:13-35, layers/SE_Resnet.py:6-99, SE_module.py:4-19, DUC.py:5-23
                         and main_fast_inference.py:42-46 (narrow to the first 50 maps)

These are floating-point kernels, so the oracle is a plain torch fp32 reference (conv2d / batch_norm eval
formula); pinned against the reference's own nn.Modules (same weights) by tests/test_oracle_pinned.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


# ---------------------------------------------------------------------------------------------- YOLOv3
def split_darknet_weights(blocks: list[dict], stream: np.ndarray, in_channels: int = 3):
    """stream: fp32 array *after* the 16-byte header -> per-block dict of tensors (None for non-conv blocks)."""
    out, ptr = [], 0
    chans, cin = [], in_channels
    for i, b in enumerate(blocks):
        cout = cin
        if b["type"] == "convolutional":
            cout, k = int(b["filters"]), int(b["size"])
            d = {}
            if int(b.get("batch_normalize", 0)):
                for name in ("bn_bias", "bn_weight", "bn_mean", "bn_var"):
                    d[name] = torch.from_numpy(stream[ptr:ptr + cout].copy())
                    ptr += cout
            else:
                d["bias"] = torch.from_numpy(stream[ptr:ptr + cout].copy())
                ptr += cout
            n = cout * cin * k * k
            d["weight"] = torch.from_numpy(stream[ptr:ptr + n].copy()).view(cout, cin, k, k)
            ptr += n
            out.append(d)
        else:
            if b["type"] == "route":
                ls = [int(x) for x in (b["layers"] if isinstance(b["layers"], list) else b["layers"].split(","))]
                if len(ls) == 1:
                    cout = chans[i + ls[0]]
                else:
                    cout = chans[i + ls[0]] + chans[ls[1]]
            out.append(None)
        chans.append(cout)
        cin = cout
    return out, ptr


def darknet_forward(blocks: list[dict], params: list, x: torch.Tensor):
    """x [B,3,reso,reso] fp32 -> list of raw head tensors [B,18,g,g] in network order (stride 32, 16, 8)."""
    outputs, heads = {}, []
    for i, b in enumerate(blocks):
        t = b["type"]
        if t == "convolutional":
            p = params[i]
            k = int(b["size"])
            pad = (k - 1) // 2 if int(b["pad"]) else 0
            x = F.conv2d(x, p["weight"], p.get("bias"), stride=int(b["stride"]), padding=pad)
            if "bn_weight" in p:
                x = (x - p["bn_mean"][None, :, None, None]) / torch.sqrt(p["bn_var"][None, :, None, None] + BN_EPS)
                x = x * p["bn_weight"][None, :, None, None] + p["bn_bias"][None, :, None, None]
            if b["activation"] == "leaky":
                x = F.leaky_relu(x, 0.1)
        elif t == "upsample":
            x = F.interpolate(x, scale_factor=int(b["stride"]), mode="nearest")
        elif t == "shortcut":
            x = outputs[i - 1] + outputs[i + int(b["from"])]
        elif t == "route":
            ls = [int(v) for v in (b["layers"] if isinstance(b["layers"], list) else b["layers"].split(","))]
            if len(ls) == 1:
                x = outputs[i + ls[0]]
            else:
                x = torch.cat((outputs[i + ls[0]], outputs[ls[1]]), 1)
        elif t == "yolo":
            heads.append(x)
            x = outputs[i - 1]
        outputs[i] = x
    return heads


# ---------------------------------------------------------------------------------------------- FastPose
def _bn(x, sd, pre):
    return F.batch_norm(x, sd[pre + ".running_mean"], sd[pre + ".running_var"], sd[pre + ".weight"], sd[pre + ".bias"],
                        False, 0.0, BN_EPS)


def _bottleneck(x, sd, pre, stride, first):
    out = F.relu(_bn(F.conv2d(x, sd[pre + ".conv1.weight"]), sd, pre + ".bn1"))
    out = F.relu(_bn(F.conv2d(out, sd[pre + ".conv2.weight"], stride=stride, padding=1), sd, pre + ".bn2"))
    out = _bn(F.conv2d(out, sd[pre + ".conv3.weight"]), sd, pre + ".bn3")
    res = x
    if first:
        y = out.mean(dim=(2, 3))
        y = F.relu(F.linear(y, sd[pre + ".se.fc.0.weight"], sd[pre + ".se.fc.0.bias"]))
        y = torch.sigmoid(F.linear(y, sd[pre + ".se.fc.2.weight"], sd[pre + ".se.fc.2.bias"]))
        out = out * y[:, :, None, None]
        res = _bn(F.conv2d(x, sd[pre + ".downsample.0.weight"], stride=stride), sd, pre + ".downsample.1")
    return F.relu(out + res)


FASTPOSE_LAYERS = (3, 4, 23, 3)


def fastpose_forward(sd: dict, x: torch.Tensor, n_maps: int = 50, return_stages: bool = False):
    """x [N,3,320,256] fp32 -> heat-maps [N,n_maps,80,64] fp32."""
    stages = {}
    x = F.relu(_bn(F.conv2d(x, sd["preact.conv1.weight"], stride=2, padding=3), sd, "preact.bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    stages["stem"] = x
    for li, nb in enumerate(FASTPOSE_LAYERS, start=1):
        for bi in range(nb):
            stride = 2 if (bi == 0 and li > 1) else 1
            x = _bottleneck(x, sd, f"preact.layer{li}.{bi}", stride, bi == 0)
        stages[f"layer{li}"] = x
    x = F.pixel_shuffle(x, 2)
    for d in ("duc1", "duc2"):
        x = F.pixel_shuffle(F.relu(_bn(F.conv2d(x, sd[d + ".conv.weight"], padding=1), sd, d + ".bn")), 2)
        stages[d] = x
    x = F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
    x = x.narrow(1, 0, n_maps)
    return (x, stages) if return_stages else x
