"""ctypes binding of libbetapose_b200.so (C ABI: include/betapose_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C betapose_b200/csrc`.  There is no fallback:
if the shared object is missing, or a compute entry point is called without a CUDA device, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BP_LIB_PATH") or os.path.join(_HERE, "libbetapose_b200.so")  # BP_LIB_PATH: an experimental build of the same library

# every symbol include/betapose_b200.h declares (tests check the .so exports all of them)
EXPORTS = (
    "bp_last_error", "bp_version", "bp_engine_create", "bp_engine_destroy", "bp_net_create", "bp_net_destroy",
    "bp_net_input_ptr", "bp_net_conv", "bp_net_alloc_tensor", "bp_net_view", "bp_net_maxpool3x3s2",
    "bp_net_global_avgpool", "bp_net_scale_add_relu", "bp_net_pixel_shuffle2", "bp_net_upsample2",
    "bp_net_copy_channels", "bp_net_add", "bp_net_tensor_info", "bp_net_num_launches", "bp_net_flops_per_image",
    "bp_net_forward", "bp_net_forward_range", "bp_net_num_ops", "bp_net_op_desc", "bp_resize_bicubic",
    "bp_yolo_decode_argmax", "bp_write_results", "bp_crop_resize", "bp_heatmap_decode", "bp_pose_pnp", "bp_pack_records",
    "bp_score_poses", "bp_pose_nms", "bp_ingest_create", "bp_ingest_destroy", "bp_ingest_num_threads", "bp_png_info",
    "bp_png_decode", "bp_ingest_submit", "bp_ingest_wait", "bp_zlib_inflate",
    "bp_write_results_nms", "bp_pack_conv_weights", "bp_frame_decode", "bp_net_set_share", "bp_net_set_op_config", "bp_net_op_config",
)
ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_IO = -1, -2, -3, -4
ORDER_RGB, ORDER_BGR = 0, 1  # frame ingest channel orders

ACT_NONE, ACT_LEAKY, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3
RES_NONE, RES_AFTER_ACT, RES_BEFORE_ACT = 0, 1, 2
STORE_PLAIN, STORE_UPSAMPLE2, STORE_PIXSHUF2 = 0, 1, 2
IN_RAW255, IN_F16 = 0, 1  # network-input kinds (include/betapose_b200.h)
IN_PAD_LEFT, IN_PAD_COLS = 3, 8  # network-input buffers are fp16 [N, H, W + IN_PAD_COLS, 8], data from column IN_PAD_LEFT


class ConvSpec(C.Structure):
    _fields_ = [
        ("src", C.c_int), ("cout", C.c_int), ("ksize", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("act", C.c_int), ("res", C.c_int), ("res_mode", C.c_int), ("dst", C.c_int), ("dst_coff", C.c_int),
        ("store_mode", C.c_int), ("out_f32", C.c_int),
        ("weight", C.c_void_p), ("bias", C.c_void_p), ("bn_gamma", C.c_void_p), ("bn_beta", C.c_void_p),
        ("bn_mean", C.c_void_p), ("bn_var", C.c_void_p), ("bn_eps", C.c_float),
        ("packed_w", C.c_void_p), ("packed_b", C.c_void_p), ("packed_w_elems", C.c_size_t), ("packed_b_elems", C.c_size_t),
    ]


class Record(C.Structure):
    _fields_ = [
        ("image_index", C.c_int32), ("status", C.c_int32), ("box", C.c_float * 4), ("det_score", C.c_float),
        ("proposal_score", C.c_float), ("keypoints", C.c_float * 150), ("R", C.c_double * 9), ("t", C.c_double * 3),
    ]


RECORD_BYTES = C.sizeof(Record)


class BetaposeError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise BetaposeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C betapose_b200/csrc`. betapose_b200 has no CPU / eager fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i, f, d = C.c_void_p, C.c_int, C.c_float, C.c_double
    L.bp_last_error.restype = C.c_char_p
    L.bp_version.restype = i
    L.bp_engine_create.argtypes = [i, C.POINTER(vp)]
    L.bp_engine_destroy.argtypes = [vp]
    L.bp_engine_destroy.restype = None
    L.bp_net_create.argtypes = [vp, i, i, i, i, vp, C.POINTER(vp)]
    L.bp_net_destroy.argtypes = [vp]
    L.bp_net_destroy.restype = None
    L.bp_net_input_ptr.argtypes = [vp]
    L.bp_net_input_ptr.restype = vp
    L.bp_net_conv.argtypes = [vp, C.POINTER(ConvSpec)]
    L.bp_pack_conv_weights.argtypes = [C.POINTER(ConvSpec), i, i, vp, vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.bp_net_alloc_tensor.argtypes = [vp, i, i, i]
    L.bp_net_view.argtypes = [vp, i, i, i]
    L.bp_net_maxpool3x3s2.argtypes = [vp, i]
    L.bp_net_global_avgpool.argtypes = [vp, i]
    L.bp_net_scale_add_relu.argtypes = [vp, i, i, i]
    L.bp_net_pixel_shuffle2.argtypes = [vp, i]
    L.bp_net_upsample2.argtypes = [vp, i, i, i]
    L.bp_net_copy_channels.argtypes = [vp, i, i, i]
    L.bp_net_add.argtypes = [vp, i, i]
    L.bp_net_tensor_info.argtypes = [vp, i, C.POINTER(i), C.POINTER(vp)]
    L.bp_net_num_launches.argtypes = [vp]
    L.bp_net_flops_per_image.argtypes = [vp]
    L.bp_net_flops_per_image.restype = d
    L.bp_net_forward.argtypes = [vp, i, vp]
    L.bp_net_forward_range.argtypes = [vp, i, i, i, vp]
    L.bp_net_set_share.argtypes = [vp, i]
    L.bp_net_set_op_config.argtypes = [vp, i, i, i, i, i]
    L.bp_net_op_config.argtypes = [vp, i, i, C.POINTER(C.c_int)]
    L.bp_net_num_ops.argtypes = [vp]
    L.bp_net_op_desc.argtypes = [vp, i, C.c_char_p, i, C.POINTER(d), C.POINTER(d)]
    L.bp_resize_bicubic.argtypes = [vp, vp, i, i, i, i, i, vp, vp, vp]
    L.bp_yolo_decode_argmax.argtypes = [vp, C.POINTER(vp), C.POINTER(i), C.POINTER(i), i, C.POINTER(f), i, i, i, f, i, i,
                                        vp, vp, vp, vp, vp, vp, vp]
    L.bp_crop_resize.argtypes = [vp, vp, i, i, vp, vp, vp, i, i, i, vp, vp, vp, vp, vp]
    L.bp_heatmap_decode.argtypes = [vp, vp, C.c_long, C.c_long, C.c_long, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp]
    L.bp_write_results.argtypes = [vp, vp, i, i, i, f, vp, vp, vp, vp]
    L.bp_write_results_nms.argtypes = [vp, vp, i, i, i, f, f, i, vp, vp, vp, vp, vp]
    L.bp_pose_pnp.argtypes = [vp, vp, vp, vp, vp, i, i, vp, vp, vp, i, i, i, f, i, C.c_uint32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.bp_pack_records.argtypes = [vp, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.bp_pose_nms.argtypes = [vp, i, vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.bp_score_poses.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp]
    L.bp_ingest_create.argtypes = [i, C.POINTER(vp)]
    L.bp_ingest_destroy.argtypes = [vp]
    L.bp_ingest_destroy.restype = None
    L.bp_ingest_num_threads.argtypes = [vp]
    L.bp_png_info.argtypes = [vp, C.c_size_t, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    L.bp_png_decode.argtypes = [vp, C.c_size_t, i, i, i, vp, C.c_size_t]
    L.bp_frame_decode.argtypes = [vp, C.c_size_t, i, i, i, vp, C.c_size_t]
    L.bp_ingest_submit.argtypes = [vp, C.POINTER(C.c_char_p), i, i, i, i, vp, C.c_size_t, vp]
    L.bp_ingest_submit.restype = C.c_int64
    L.bp_ingest_wait.argtypes = [vp, C.c_int64]
    L.bp_zlib_inflate.argtypes = [vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    _lib = L
    return L


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        msg = lib().bp_last_error().decode("utf-8", "replace")
        raise BetaposeError(f"{what}: {msg} (code {rc})")
    return rc


def require_cuda() -> None:
    import torch

    if not torch.cuda.is_available():
        raise BetaposeError("betapose_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


class Engine:
    """One per process/device. Owns the native engine handle."""

    _instances: dict[int, "Engine"] = {}

    def __init__(self, device: int = 0):
        require_cuda()
        self.device = int(device)
        h = C.c_void_p()
        check(lib().bp_engine_create(self.device, C.byref(h)), "bp_engine_create")
        self.handle = h

    @classmethod
    def get(cls, device: int | None = None) -> "Engine":
        import torch

        require_cuda()
        dev = torch.cuda.current_device() if device is None else int(device)
        if dev not in cls._instances:
            cls._instances[dev] = Engine(dev)
        return cls._instances[dev]


def ptr(t) -> C.c_void_p:
    """device/host address of a torch tensor (or None)."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr() -> C.c_void_p:
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
