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


# ======================================================================================================
# Packed-weight cache (device half): the form bp_net_conv uploads, computed once on the host
# ======================================================================================================
# Loading a detector + key-point net from their training formats costs a pass over 121 M fp32 parameters (BN folding in
# fp64, fp16 conversion, re-ordering to the kernels' K order) before anything reaches the GPU.  A PackedWeights file
# holds the result: per convolution the fp16 rows [cout_pad][K] and the fp32 bias, in builder order, memory-mapped on
# load and handed to bp_net_conv as they are (bp_conv_spec.packed_w / packed_b).  Packing runs the SAME builder code as
# the engine against a recorder that has no CUDA behind it, so it works (and is tested) without a GPU.
import ctypes as _C
import json as _json
import os as _os

from . import _lib

_MAGIC = b"BPPW\x01\0\0\0"
PACKED_LAYOUT = 1  # bump when csrc/net.cu:pack_conv_weights changes the layout


class PackRecorder:
    """Stands in for net.Net while build_darknet / build_fastpose run: every conv is folded + packed on the host
    (bp_pack_conv_weights), every other builder call only hands out tensor ids."""

    def __init__(self, in_h: int, in_w: int, in_kind: int):
        self.in_h, self.in_w, self.in_kind = int(in_h), int(in_w), int(in_kind)
        self.entries: list[dict] = []
        self._next = 1  # tensor 0 is the network input

    def _new(self) -> int:
        self._next += 1
        return self._next - 1

    def conv(self, src, weight, bias=None, bn=None, stride=1, pad=0, act=0, res=-1, res_mode=0, dst=-1, dst_coff=0, store=0,
             out_f32=False) -> int:
        s, keep = _net.conv_spec(src, weight, bias, bn, stride, pad, act, res, res_mode, dst, dst_coff, store, out_f32)
        sh = tuple(int(v) for v in _net._shape4(weight))
        in_kind = self.in_kind if int(src) == 0 else -1
        nw, nb = _C.c_size_t(), _C.c_size_t()
        L = _lib.lib()
        _lib.check(L.bp_pack_conv_weights(_C.byref(s), sh[1], in_kind, None, None, _C.byref(nw), _C.byref(nb)), "bp_pack_conv_weights")
        w = np.empty(nw.value, np.float16)
        b = np.empty(nb.value, np.float32)
        _lib.check(L.bp_pack_conv_weights(_C.byref(s), sh[1], in_kind, w.ctypes.data_as(_C.c_void_p), b.ctypes.data_as(_C.c_void_p),
                                          _C.byref(nw), _C.byref(nb)), "bp_pack_conv_weights")
        self.entries.append(dict(index=len(self.entries), shape=sh, store=int(store), in_kind=in_kind, w=w, b=b))
        return self._new()

    def alloc_tensor(self, h, w, c) -> int:
        return self._new()

    def view(self, t, coff, c) -> int:
        return self._new()

    def maxpool3x3s2(self, src) -> int:
        return self._new()

    global_avgpool = pixel_shuffle2 = maxpool3x3s2

    def scale_add_relu(self, y, gates, skip) -> int:
        return self._new()

    def upsample2(self, src, dst=-1, dst_coff=0) -> int:
        return self._new()

    def copy_channels(self, src, dst, dst_coff) -> int:
        return int(dst)

    def add(self, a, b) -> int:
        return self._new()


class PackedWeights:
    """The packed convolutions of one network, in builder order.  kind: "darknet" | "fastpose"."""

    def __init__(self, kind: str, entries, meta: dict | None = None):
        self.kind, self.entries, self.meta = kind, list(entries), dict(meta or {})
        self.meta.setdefault("layout", PACKED_LAYOUT)
        self.meta.setdefault("lib_version", int(_lib.lib().bp_version()))

    def __iter__(self):
        return iter(self.entries)

    def __len__(self):
        return len(self.entries)

    @property
    def nbytes(self) -> int:
        return sum(e["w"].nbytes + e["b"].nbytes for e in self.entries)

    def save(self, path: str) -> None:
        """one file: magic, u64 header length, JSON header, 64-byte aligned blobs; written to a temporary name and renamed"""
        table, off = [], 0
        for e in self.entries:
            rec = dict(shape=list(e["shape"]), store=e["store"], in_kind=e["in_kind"])
            for k in ("w", "b"):
                off = (off + 63) // 64 * 64
                rec[k + "_off"], rec[k + "_elems"] = off, int(e[k].size)
                off += e[k].nbytes
            table.append(rec)
        head = _json.dumps(dict(kind=self.kind, meta=self.meta, entries=table)).encode()
        base = (len(_MAGIC) + 8 + len(head) + 63) // 64 * 64
        tmp = f"{path}.tmp{_os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(_MAGIC + np.uint64(len(head)).tobytes() + head)
            for e, rec in zip(self.entries, table):
                for k in ("w", "b"):
                    f.seek(base + rec[k + "_off"])
                    f.write(np.ascontiguousarray(e[k]).tobytes())
            f.truncate(base + off)
        _os.replace(tmp, path)

    @classmethod
    def load(cls, path: str, mmap: bool = True) -> "PackedWeights":
        with open(path, "rb") as f:
            if f.read(len(_MAGIC)) != _MAGIC:
                raise ValueError(f"{path}: not a packed-weight file")
            n = int(np.frombuffer(f.read(8), np.uint64)[0])
            head = _json.loads(f.read(n).decode())
        if head["meta"].get("layout") != PACKED_LAYOUT:
            raise ValueError(f"{path}: packed layout {head['meta'].get('layout')}, this build reads layout {PACKED_LAYOUT}")
        base = (len(_MAGIC) + 8 + n + 63) // 64 * 64
        raw = np.memmap(path, np.uint8, "r") if mmap else np.fromfile(path, np.uint8)
        entries = []
        for i, rec in enumerate(head["entries"]):
            w = raw[base + rec["w_off"]: base + rec["w_off"] + 2 * rec["w_elems"]].view(np.float16)
            b = raw[base + rec["b_off"]: base + rec["b_off"] + 4 * rec["b_elems"]].view(np.float32)
            entries.append(dict(index=i, shape=tuple(rec["shape"]), store=rec["store"], in_kind=rec["in_kind"], w=w, b=b))
        return cls(head["kind"], entries, head["meta"])


def _placeholder(shape):
    return np.broadcast_to(np.float32(0), tuple(shape))  # shape without storage: packed builds never read parameter data


def darknet_placeholder_params(blocks, in_c: int = 3):
    """what net.split_darknet_stream returns, with shapes only"""
    info = _net.infer_darknet_shapes(blocks)
    out = []
    for i, b in enumerate(blocks):
        if b["type"] != "convolutional":
            out.append(None)
            continue
        cin = in_c if i == 0 else info[i - 1]["C"]
        cout, k = int(b["filters"]), int(b["size"])
        d = {n: _placeholder((cout,)) for n in (("bn_bias", "bn_weight", "bn_mean", "bn_var") if int(b.get("batch_normalize", 0)) else ("bias",))}
        d["weight"] = _placeholder((cout, cin, k, k))
        out.append(d)
    return out


def fastpose_placeholder_state_dict(n_out: int = 50) -> dict:
    return {k: _placeholder(sh) for k, sh in fastpose_expected_shapes(n_out).items()}


def pack_darknet(blocks, stream: np.ndarray, reso: int = 416, meta: dict | None = None) -> PackedWeights:
    check_darknet_stream(blocks, np.asarray(stream, np.float32))
    params, _ = _net.split_darknet_stream(blocks, np.asarray(stream, np.float32))
    rec = PackRecorder(reso, reso, _lib.IN_RAW255)
    _net.build_darknet(rec, blocks, params)
    return PackedWeights("darknet", rec.entries, dict(meta or {}, reso=int(reso)))


def pack_fastpose(sd: dict, n_maps: int = 50, inp_h: int = 320, inp_w: int = 256, meta: dict | None = None) -> PackedWeights:
    check_fastpose_state_dict(sd, n_maps)
    rec = PackRecorder(inp_h, inp_w, _lib.IN_F16)
    _net.build_fastpose(rec, sd, n_maps)
    return PackedWeights("fastpose", rec.entries, dict(meta or {}, n_maps=int(n_maps)))


def unpack_conv(entry: dict):
    """Inverse of the packing for one convolution: -> (weight fp32 [cout, cin, k, k] with BN folded in, bias fp32 [cout]).
    (The hand-back direction of SURVEY 8(f) item 4: a packed network can be written out again as BN-free fp32 tensors.)"""
    cout, cin, k, _ = entry["shape"]
    stem = entry["in_kind"] >= 0
    cv = (32 if k * 8 <= 32 else 64) if stem else 0
    K = k * cv if stem else k * k * cin
    wpitch = (K + 7) // 8 * 8
    rows = np.asarray(entry["w"], np.float16).reshape(-1, wpitch)[:, :K].astype(np.float32)
    bias = np.asarray(entry["b"], np.float32)
    if entry["store"] == _lib.STORE_PIXSHUF2:  # kernel row sub*(cout/4) + c holds channel 4c + sub
        o = np.arange(cout)
        perm = (o % 4) * (cout // 4) + o // 4
        rows, bias = rows[perm], bias[perm]
    else:
        rows, bias = rows[:cout], bias[:cout]
    if stem:
        w = rows.reshape(cout, k, cv // 8, 8)[:, :, :k, :cin].transpose(0, 3, 1, 2)
        if entry["in_kind"] == _lib.IN_RAW255:
            w = w * np.float32(255.0)
    else:
        w = rows.reshape(cout, k, k, cin).transpose(0, 3, 1, 2)
    return np.ascontiguousarray(w, np.float32), bias.copy()


def file_fingerprint(path: str) -> dict:
    st = _os.stat(path)
    return dict(name=_os.path.basename(path), size=int(st.st_size), mtime_ns=int(st.st_mtime_ns))


def load_or_pack(cache_path: str, source_path: str, pack_fn):
    """PackedWeights from `cache_path` when it was packed from this very `source_path` (name, size, mtime) by this layout,
    else pack_fn() -> PackedWeights, saved for next time."""
    fp = file_fingerprint(source_path)
    if _os.path.isfile(cache_path):
        try:
            pw = PackedWeights.load(cache_path)
            if pw.meta.get("source") == fp:
                return pw, True
        except ValueError:
            pass
    pw = pack_fn()
    pw.meta["source"] = fp
    _os.makedirs(_os.path.dirname(_os.path.abspath(cache_path)), exist_ok=True)
    pw.save(cache_path)
    return pw, False
