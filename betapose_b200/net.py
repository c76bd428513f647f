"""Python side of bp_net: graph construction for the two networks on the path.

`build_darknet` walks a parsed darknet cfg the way the reference's `Darknet.build_model` / `forward` do
(3_6Dpose_estimator/yolo/darknet.py:223-363) but emits *fused* ops: conv+BN+leaky(+shortcut) is one launch,
nearest-x2 upsample and channel concat ([route] with two layers) are folded into the producers' store
addresses.  `build_fastpose` does the same for FastPose (KPD/src/models/FastPose.py:13-35 and layers/*.py):
conv+BN+ReLU(+residual) in one launch, PixelShuffle folded into the DUC convs, the SE fully-connected pair run
through the same tensor-core kernel as 1x1 convs over a [N,1,1,C] tensor.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, RES_AFTER_ACT, RES_BEFORE_ACT, RES_NONE, STORE_PIXSHUF2,
                   STORE_PLAIN, STORE_UPSAMPLE2)


def _f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _shape4(weight):
    sh = tuple(weight.shape)
    return sh + (1, 1) if len(sh) == 2 else sh


def conv_spec(src, weight, bias, bn, stride, pad, act, res, res_mode, dst, dst_coff, store, out_f32, shapes_only=False):
    """-> (ConvSpec, arrays to keep alive while it is used).  shapes_only: do not touch parameter data (placeholders)."""
    sh = _shape4(weight)
    s = _lib.ConvSpec()
    s.src, s.cout, s.ksize, s.stride, s.pad = int(src), int(sh[0]), int(sh[2]), int(stride), int(pad)
    s.act, s.res, s.res_mode, s.dst, s.dst_coff = int(act), int(res), int(res_mode), int(dst), int(dst_coff)
    s.store_mode, s.out_f32 = int(store), int(bool(out_f32))
    keep = []
    if shapes_only:
        return s, keep
    w = _f32(weight).reshape(sh)
    keep.append(w)
    s.weight = w.ctypes.data_as(C.c_void_p)
    if bias is not None:
        b = _f32(bias)
        keep.append(b)
        s.bias = b.ctypes.data_as(C.c_void_p)
    if bn is not None:
        g, be, m, v = (_f32(x) for x in bn[:4])
        keep += [g, be, m, v]
        s.bn_gamma, s.bn_beta = g.ctypes.data_as(C.c_void_p), be.ctypes.data_as(C.c_void_p)
        s.bn_mean, s.bn_var = m.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p)
        s.bn_eps = float(bn[4]) if len(bn) > 4 else 1e-5
    return s, keep


def take_packed(packed, weight, store, in_kind):
    """next entry of a PackedWeights iterator, checked against what the builder is about to emit"""
    ent = next(packed, None)
    want = tuple(int(v) for v in _shape4(weight))
    if ent is None:
        raise _lib.BetaposeError(f"packed weights end before the network does (next conv: {want})")
    if tuple(ent["shape"]) != want or ent["store"] != int(store) or ent["in_kind"] != int(in_kind):
        raise _lib.BetaposeError(f"packed weights do not match the network: conv {ent['index']} was packed for {tuple(ent['shape'])} "
                                 f"(store {ent['store']}, input kind {ent['in_kind']}), the builder asks for {want} "
                                 f"(store {int(store)}, input kind {int(in_kind)})")
    return ent["w"], ent["b"]


def packed_exhausted(packed) -> None:
    if next(packed, None) is not None:
        raise _lib.BetaposeError("packed weights hold more convolutions than the network (wrong cfg / n_maps?)")


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias engine-owned device memory without a copy."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class Net:
    def __init__(self, max_batch: int, in_h: int, in_w: int, in_kind: int, share: "Net | None" = None, device=None):
        self.engine = _lib.Engine.get(device)
        self.max_batch, self.in_h, self.in_w, self.in_kind = int(max_batch), int(in_h), int(in_w), int(in_kind)
        h = C.c_void_p()
        _lib.check(_lib.lib().bp_net_create(self.engine.handle, self.max_batch, self.in_h, self.in_w, self.in_kind,
                                            share.handle if share is not None else None, C.byref(h)), "bp_net_create")
        self.handle = h
        self._keep = share  # the sharing net must outlive us
        self.device = torch.device("cuda", self.engine.device)

    def set_op_config(self, op: int, batch: int, block_n: int = 0, cg: int = 0, mt: int = 0) -> bool:
        """Force the tile plan of convolution `op` at `batch` (bp_net_set_op_config); False if the layer cannot run that way."""
        rc = _lib.lib().bp_net_set_op_config(self.handle, int(op), int(batch), int(block_n), int(cg), int(mt))
        if rc == _lib.ERR_UNSUPPORTED:
            return False
        _lib.check(rc, "bp_net_set_op_config")
        return True

    def op_config(self, op: int, batch: int) -> tuple:
        """(BLOCK_N, cg, mt, BLOCK_K, stages) of the plan `op` runs at `batch`; zeros for aux ops."""
        cfg = (C.c_int * 5)()
        _lib.check(_lib.lib().bp_net_op_config(self.handle, int(op), int(batch), cfg), "bp_net_op_config")
        return tuple(cfg)

    def set_share(self, share_batch: int) -> None:
        """This net runs concurrently with other nets whose batches sum to share_batch images (bp_net_set_share)."""
        _lib.check(_lib.lib().bp_net_set_share(self.handle, int(share_batch)), "bp_net_set_share")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().bp_net_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- builders -------------------------------------------------------------------------------------
    def conv(self, src, weight, bias=None, bn=None, stride=1, pad=0, act=ACT_NONE, res=-1, res_mode=RES_NONE, dst=-1,
             dst_coff=0, store=STORE_PLAIN, out_f32=False) -> int:
        """`self.packed` (an iterator over a weights.PackedWeights' entries, set by the caller before building) makes
        every conv take its folded + packed weights from there; `weight` / `bias` / `bn` then only provide shapes."""
        packed = getattr(self, "packed", None)
        s, keep = conv_spec(src, weight, bias, bn, stride, pad, act, res, res_mode, dst, dst_coff, store, out_f32,
                            shapes_only=packed is not None)
        if packed is not None:
            w, b = take_packed(packed, weight, store, self.in_kind if int(src) == 0 else -1)
            keep += [w, b]
            s.packed_w, s.packed_b = w.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)
            s.packed_w_elems, s.packed_b_elems = w.size, b.size
        return _lib.check(_lib.lib().bp_net_conv(self.handle, C.byref(s)), "bp_net_conv")

    def alloc_tensor(self, h, w, c) -> int:
        return _lib.check(_lib.lib().bp_net_alloc_tensor(self.handle, int(h), int(w), int(c)), "bp_net_alloc_tensor")

    def view(self, t, coff, c) -> int:
        return _lib.check(_lib.lib().bp_net_view(self.handle, int(t), int(coff), int(c)), "bp_net_view")

    def maxpool3x3s2(self, src) -> int:
        return _lib.check(_lib.lib().bp_net_maxpool3x3s2(self.handle, int(src)), "bp_net_maxpool3x3s2")

    def global_avgpool(self, src) -> int:
        return _lib.check(_lib.lib().bp_net_global_avgpool(self.handle, int(src)), "bp_net_global_avgpool")

    def scale_add_relu(self, y, gates, skip) -> int:
        return _lib.check(_lib.lib().bp_net_scale_add_relu(self.handle, int(y), int(gates), int(skip)), "bp_net_scale_add_relu")

    def pixel_shuffle2(self, src) -> int:
        return _lib.check(_lib.lib().bp_net_pixel_shuffle2(self.handle, int(src)), "bp_net_pixel_shuffle2")

    def upsample2(self, src, dst=-1, dst_coff=0) -> int:
        return _lib.check(_lib.lib().bp_net_upsample2(self.handle, int(src), int(dst), int(dst_coff)), "bp_net_upsample2")

    def copy_channels(self, src, dst, dst_coff) -> int:
        return _lib.check(_lib.lib().bp_net_copy_channels(self.handle, int(src), int(dst), int(dst_coff)), "bp_net_copy_channels")

    def add(self, a, b) -> int:
        return _lib.check(_lib.lib().bp_net_add(self.handle, int(a), int(b)), "bp_net_add")

    # ---- introspection --------------------------------------------------------------------------------
    def tensor_info(self, t):
        dims = (C.c_int * 8)()
        p = C.c_void_p()
        _lib.check(_lib.lib().bp_net_tensor_info(self.handle, int(t), dims, C.byref(p)), "bp_net_tensor_info")
        return dict(H=dims[0], W=dims[1], C=dims[2], pitch=dims[3], f32=bool(dims[4]), coff=dims[5], row_px=dims[6],
                    col0=dims[7], ptr=p.value)

    def tensor(self, t, batch=None) -> torch.Tensor:
        """torch view [batch, H, W, C] (strided over the pitch) aliasing the engine's buffer."""
        i = self.tensor_info(t)
        n = self.max_batch if batch is None else int(batch)
        if t == 0:  # network input: fp16 [n, H, W + pad columns, 8]; the view covers the data pixels, channels R,G,B
            full = torch.as_tensor(_DevArray(i["ptr"], (n, i["H"], i["row_px"], 8), "<f2"), device=self.device)
            return full[:, :, i["col0"]:i["col0"] + i["W"], :3]
        typ = "<f4" if i["f32"] else "<f2"
        full = torch.as_tensor(_DevArray(i["ptr"], (n * i["H"] * i["W"] * i["pitch"],), typ), device=self.device)
        return full.as_strided((n, i["H"], i["W"], i["C"]), (i["H"] * i["W"] * i["pitch"], i["W"] * i["pitch"], i["pitch"], 1))

    def input(self, batch=None) -> torch.Tensor:
        return self.tensor(0, batch)

    @property
    def num_ops(self) -> int:
        return _lib.lib().bp_net_num_ops(self.handle)

    @property
    def flops_per_image(self) -> float:
        return float(_lib.lib().bp_net_flops_per_image(self.handle))

    def op_desc(self, i):
        buf = C.create_string_buffer(256)
        fl, by = C.c_double(), C.c_double()
        _lib.check(_lib.lib().bp_net_op_desc(self.handle, int(i), buf, 256, C.byref(fl), C.byref(by)), "bp_net_op_desc")
        return buf.value.decode(), fl.value, by.value

    def forward(self, batch: int, first: int | None = None, last: int | None = None) -> None:
        """Enqueue all (or ops [first,last)) on the current torch stream; input must already be in `self.input()`."""
        if first is None and last is None:
            _lib.check(_lib.lib().bp_net_forward(self.handle, int(batch), _lib.stream_ptr()), "bp_net_forward")
        else:
            _lib.check(_lib.lib().bp_net_forward_range(self.handle, int(batch), int(first or 0),
                                                       int(self.num_ops if last is None else last), _lib.stream_ptr()),
                       "bp_net_forward_range")


# ======================================================================================================
# YOLOv3 from a darknet cfg
# ======================================================================================================
def _route_layers(b, i):
    ls = b["layers"] if isinstance(b["layers"], list) else b["layers"].split(",")
    return [(i + int(v)) if int(v) < 0 else int(v) for v in ls]


def infer_darknet_shapes(blocks, reso=416, in_c=3):
    info = []
    for i, b in enumerate(blocks):
        t = b["type"]
        p = info[i - 1] if i > 0 else dict(C=in_c, H=reso, W=reso)
        if t == "convolutional":
            k, s = int(b["size"]), int(b["stride"])
            pad = (k - 1) // 2 if int(b["pad"]) else 0
            d = dict(C=int(b["filters"]), H=(p["H"] + 2 * pad - k) // s + 1, W=(p["W"] + 2 * pad - k) // s + 1, srcs=[i - 1], pad=pad)
        elif t == "upsample":
            s = int(b["stride"])
            d = dict(C=p["C"], H=p["H"] * s, W=p["W"] * s, srcs=[i - 1])
        elif t == "shortcut":
            d = dict(C=p["C"], H=p["H"], W=p["W"], srcs=[i - 1, i + int(b["from"])])
        elif t == "route":
            ls = _route_layers(b, i)
            d = dict(C=sum(info[l]["C"] for l in ls), H=info[ls[0]]["H"], W=info[ls[0]]["W"], srcs=ls)
        elif t == "yolo":
            d = dict(C=p["C"], H=p["H"], W=p["W"], srcs=[i - 1])
        else:
            raise _lib.BetaposeError(f"unsupported darknet block type '{t}' (block {i})")
        d["type"] = t
        info.append(d)
    return info


def split_darknet_stream(blocks, stream: np.ndarray, in_c=3):
    """fp32 stream after the 16-byte header -> per-block parameter dicts (darknet.py:365-432 order)."""
    info = infer_darknet_shapes(blocks)
    out, ptr = [], 0
    for i, b in enumerate(blocks):
        if b["type"] != "convolutional":
            out.append(None)
            continue
        cin = in_c if i == 0 else info[i - 1]["C"]
        cout, k = int(b["filters"]), int(b["size"])
        d = {}
        if int(b.get("batch_normalize", 0)):
            for name in ("bn_bias", "bn_weight", "bn_mean", "bn_var"):
                d[name] = stream[ptr:ptr + cout]
                ptr += cout
        else:
            d["bias"] = stream[ptr:ptr + cout]
            ptr += cout
        n = cout * cin * k * k
        if ptr + n > stream.size:
            raise _lib.BetaposeError(f"weights file too short at conv block {i}: need {ptr + n} floats, have {stream.size}")
        d["weight"] = stream[ptr:ptr + n].reshape(cout, cin, k, k)
        ptr += n
        out.append(d)
    return out, ptr


def build_darknet(net: Net, blocks, params):
    """Emit fused ops. Returns list of heads: dict(tensor=id, grid=g, anchors=[(w,h)x3], classes=int)."""
    info = infer_darknet_shapes(blocks, net.in_h)
    n = len(blocks)
    consumers = [[] for _ in range(n)]
    for i, d in enumerate(info):
        for s in d["srcs"]:
            if s >= 0:
                consumers[s].append(i)

    # ---- plan concat fusion for two-layer routes: route r = cat(upsample(conv u-1), block b)
    concat_of = {}   # block index -> (route index, channel offset)   (the block's value lands in the concat buffer)
    up_fused = set()
    for r, b in enumerate(blocks):
        if b["type"] != "route":
            continue
        ls = _route_layers(b, r)
        if len(ls) != 2:
            continue
        a, bb = ls
        ok = (blocks[a]["type"] == "upsample" and int(blocks[a]["stride"]) == 2 and consumers[a] == [r]
              and blocks[a - 1]["type"] == "convolutional" and consumers[a - 1] == [a] and info[a - 1]["C"] % 8 == 0)
        tb = blocks[bb]["type"]
        ok = ok and bb < a - 1 and info[bb]["C"] % 8 == 0 and (
            tb == "convolutional" or (tb == "shortcut" and blocks[bb - 1]["type"] == "convolutional" and consumers[bb - 1] == [bb]))
        if ok:
            concat_of[a] = (r, 0)
            concat_of[bb] = (r, info[a]["C"])
            up_fused.add(a)
    concat_tensor = {}

    def concat_id(r):
        if r not in concat_tensor:
            concat_tensor[r] = net.alloc_tensor(info[r]["H"], info[r]["W"], info[r]["C"])
        return concat_tensor[r]

    out_id: list = [None] * n
    heads = []
    skip = set()
    for i, b in enumerate(blocks):
        if i in skip:
            continue
        t = b["type"]
        if t == "convolutional":
            p = params[i]
            src = 0 if i == 0 else out_id[i - 1]
            act = ACT_LEAKY if b["activation"] == "leaky" else ACT_NONE
            bn = (p["bn_weight"], p["bn_bias"], p["bn_mean"], p["bn_var"], 1e-5) if "bn_weight" in p else None
            kw = dict(bias=p.get("bias"), bn=bn, stride=int(b["stride"]), pad=info[i]["pad"], act=act)
            nxt = blocks[i + 1]["type"] if i + 1 < n else ""
            if nxt == "shortcut" and consumers[i] == [i + 1] and blocks[i + 1].get("activation", "linear") == "linear":
                j = i + 1
                res = out_id[j + int(blocks[j]["from"])]
                if j in concat_of:
                    r, coff = concat_of[j]
                    out_id[j] = net.conv(src, p["weight"], res=res, res_mode=RES_AFTER_ACT, dst=concat_id(r), dst_coff=coff, **kw)
                else:
                    out_id[j] = net.conv(src, p["weight"], res=res, res_mode=RES_AFTER_ACT, **kw)
                skip.add(j)
            elif nxt == "upsample" and (i + 1) in up_fused:
                r, coff = concat_of[i + 1]
                out_id[i + 1] = net.conv(src, p["weight"], dst=concat_id(r), dst_coff=coff, store=STORE_UPSAMPLE2, **kw)
                skip.add(i + 1)
            elif nxt == "yolo":
                out_id[i] = net.conv(src, p["weight"], out_f32=True, **kw)
            elif i in concat_of:
                r, coff = concat_of[i]
                out_id[i] = net.conv(src, p["weight"], dst=concat_id(r), dst_coff=coff, **kw)
            else:
                out_id[i] = net.conv(src, p["weight"], **kw)
        elif t == "shortcut":
            out_id[i] = net.add(out_id[i - 1], out_id[i + int(b["from"])])
        elif t == "upsample":
            out_id[i] = net.upsample2(out_id[i - 1])
        elif t == "route":
            ls = _route_layers(b, i)
            if len(ls) == 1:
                out_id[i] = out_id[ls[0]]
            elif i in concat_tensor:
                out_id[i] = concat_tensor[i]
            else:
                cat = net.alloc_tensor(info[i]["H"], info[i]["W"], info[i]["C"])
                off = 0
                for l in ls:
                    net.copy_channels(out_id[l], cat, off)
                    off += info[l]["C"]
                out_id[i] = cat
        elif t == "yolo":
            mask = [int(x) for x in b["mask"].split(",")]
            a = [int(x) for x in b["anchors"].split(",")]
            anchors = [(a[2 * m], a[2 * m + 1]) for m in mask]
            heads.append(dict(tensor=out_id[i - 1], grid=info[i]["H"], anchors=anchors, classes=int(b["classes"])))
            out_id[i] = out_id[i - 2] if i >= 2 else None
    return heads


# ======================================================================================================
# FastPose (SE-ResNet-101 + DUC x2 + head)
# ======================================================================================================
FASTPOSE_LAYERS = (3, 4, 23, 3)


def build_fastpose(net: Net, sd: dict, n_maps: int = 50) -> int:
    """Returns the tensor id of the fp32 heat-maps [N,80,64,n_maps]."""

    def bn(pre):
        return (sd[pre + ".weight"], sd[pre + ".bias"], sd[pre + ".running_mean"], sd[pre + ".running_var"], 1e-5)

    x = net.conv(0, sd["preact.conv1.weight"], bn=bn("preact.bn1"), stride=2, pad=3, act=ACT_RELU)
    x = net.maxpool3x3s2(x)
    for li, nb in enumerate(FASTPOSE_LAYERS, start=1):
        for bi in range(nb):
            pre = f"preact.layer{li}.{bi}"
            stride = 2 if (bi == 0 and li > 1) else 1
            o = net.conv(x, sd[pre + ".conv1.weight"], bn=bn(pre + ".bn1"), act=ACT_RELU)
            o = net.conv(o, sd[pre + ".conv2.weight"], bn=bn(pre + ".bn2"), stride=stride, pad=1, act=ACT_RELU)
            if bi == 0:
                y = net.conv(o, sd[pre + ".conv3.weight"], bn=bn(pre + ".bn3"), act=ACT_NONE)
                skip = net.conv(x, sd[pre + ".downsample.0.weight"], bn=bn(pre + ".downsample.1"), stride=stride, act=ACT_NONE)
                g = net.global_avgpool(y)
                g = net.conv(g, sd[pre + ".se.fc.0.weight"], bias=sd[pre + ".se.fc.0.bias"], act=ACT_RELU)
                g = net.conv(g, sd[pre + ".se.fc.2.weight"], bias=sd[pre + ".se.fc.2.bias"], act=ACT_SIGMOID)
                x = net.scale_add_relu(y, g, skip)
            else:
                x = net.conv(o, sd[pre + ".conv3.weight"], bn=bn(pre + ".bn3"), act=ACT_RELU, res=x, res_mode=RES_BEFORE_ACT)
    x = net.pixel_shuffle2(x)
    x = net.conv(x, sd["duc1.conv.weight"], bn=bn("duc1.bn"), pad=1, act=ACT_RELU, store=STORE_PIXSHUF2)
    x = net.conv(x, sd["duc2.conv.weight"], bn=bn("duc2.bn"), pad=1, act=ACT_RELU, store=STORE_PIXSHUF2)
    w = sd["conv_out.weight"][:n_maps]
    b = sd["conv_out.bias"][:n_maps]
    return net.conv(x, w, bias=b, pad=1, act=ACT_NONE, out_f32=True)
