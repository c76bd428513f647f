"""Frame ingest: image files -> uint8 [n, H, W, 3] RGB frames in pinned host memory, decoded by
the native thread pool of libbetapose_b200.so (csrc/ingest.cu) ahead of the GPU.

The reference decodes every frame twice (cv2.imread for `orig_img`, PIL.Image.open for the detector input) on the one
Python thread of `ImageLoader.getitem_yolo` (dataloader.py:150-179; yolo/preprocess.py:34-46).  Here one decode per frame
lands directly in the buffer `BetaposeEngine.run_stream` uploads from; `FrameIngest.batches` keeps `depth` batches in
flight so decoding, the host->device copy and the GPU step of three different batches overlap.

Files the native decoder does not handle (JPEG, Adam7-interlaced PNG) are decoded with Pillow on the calling thread --
host-side image decoding is outside the CUDA path either way; corrupt files raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class _Ticket:
    __slots__ = ("id", "paths", "c_paths", "status", "out", "n")


class FrameIngest:
    def __init__(self, n_threads: int = 0, frame_h: int = 480, frame_w: int = 640, order: str = "rgb"):
        self.H, self.W = int(frame_h), int(frame_w)
        self.order = {"rgb": _lib.ORDER_RGB, "bgr": _lib.ORDER_BGR}[order]
        h = C.c_void_p()
        _lib.check(_lib.lib().bp_ingest_create(int(n_threads), C.byref(h)), "bp_ingest_create")
        self.handle = h
        self.n_threads = _lib.lib().bp_ingest_num_threads(h)

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().bp_ingest_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: module globals may already be gone
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ single in-memory stream (calling thread)
    @staticmethod
    def png_info(data: bytes) -> dict:
        vals = [C.c_int() for _ in range(4)]
        buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
        _lib.check(_lib.lib().bp_png_info(buf, len(data), *[C.byref(v) for v in vals]), "bp_png_info")
        return dict(zip(("H", "W", "channels", "depth"), (v.value for v in vals)))

    def decode_bytes(self, data: bytes, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.H, self.W, 3), np.uint8)
        assert out.dtype == np.uint8 and out.shape == (self.H, self.W, 3) and out.strides[1:] == (3, 1)
        buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
        _lib.check(_lib.lib().bp_frame_decode(buf, len(data), self.H, self.W, self.order, C.c_void_p(out.ctypes.data), out.strides[0]),
                   "bp_frame_decode")
        return out

    # ------------------------------------------------------------------ files, through the pool
    def _frame_buffer(self, n: int) -> torch.Tensor:
        t = torch.empty((n, self.H, self.W, 3), dtype=torch.uint8)
        return t.pin_memory() if torch.cuda.is_available() else t

    def submit(self, paths, out: torch.Tensor | np.ndarray) -> _Ticket:
        """Queue `paths` for decoding into out[:len(paths)] (uint8 [>=n, H, W, 3], C-contiguous).  Returns at once."""
        n = len(paths)
        assert n > 0 and tuple(out.shape[1:]) == (self.H, self.W, 3) and out.shape[0] >= n
        if isinstance(out, torch.Tensor):
            assert out.dtype == torch.uint8 and out.is_contiguous() and not out.is_cuda
            addr = out.data_ptr()
        else:
            assert out.dtype == np.uint8 and out.flags.c_contiguous
            addr = out.ctypes.data
        t = _Ticket()
        t.paths, t.n, t.out = list(paths), n, out  # keep the buffers alive until wait()
        t.c_paths = (C.c_char_p * n)(*[str(p).encode() for p in paths])
        t.status = np.zeros(n, np.int32)
        t.id = _lib.lib().bp_ingest_submit(self.handle, t.c_paths, n, self.H, self.W, self.order, C.c_void_p(addr), 0,
                                           C.c_void_p(t.status.ctypes.data))
        _lib.check(int(t.id), "bp_ingest_submit")
        return t

    def wait(self, t: _Ticket):
        """Block until the ticket's files are decoded; returns out[:n]."""
        rc = _lib.lib().bp_ingest_wait(self.handle, t.id)
        if rc != 0:
            msg = _lib.lib().bp_last_error().decode("utf-8", "replace")
            hard = np.flatnonzero((t.status != 0) & (t.status != _lib.ERR_UNSUPPORTED))
            if hard.size:
                raise _lib.BetaposeError(f"frame ingest: {hard.size} file(s) failed, first: {msg} (code {rc})")
            for i in np.flatnonzero(t.status == _lib.ERR_UNSUPPORTED):
                self._decode_with_pillow(t.paths[i], t.out[i])
        return t.out[: t.n]

    def _decode_with_pillow(self, path, dst):
        from PIL import Image

        if str(path).lower().endswith(".npy"):
            im = np.load(path)
            if im.dtype != np.uint8:
                raise _lib.BetaposeError(f"{path}: frames must be uint8, got {im.dtype}")
            im = np.ascontiguousarray(im if im.ndim == 3 else np.repeat(im[..., None], 3, -1))
        else:
            im = np.asarray(Image.open(path).convert("RGB"))
        if im.shape != (self.H, self.W, 3):
            raise _lib.BetaposeError(f"{path}: expected a {self.W}x{self.H} frame, got {im.shape[1]}x{im.shape[0]}")
        if self.order == _lib.ORDER_BGR:
            im = im[:, :, ::-1]
        if isinstance(dst, torch.Tensor):
            dst.copy_(torch.from_numpy(np.ascontiguousarray(im)))
        else:
            dst[...] = im

    def decode_files(self, paths, out=None):
        if out is None:
            out = np.empty((len(paths), self.H, self.W, 3), np.uint8)
        return self.wait(self.submit(paths, out))

    def batches(self, paths, batch: int, depth: int = 2, in_flight: int = 1):
        """Yield uint8 [n <= batch, H, W, 3] host tensors (pinned when CUDA is present) for consecutive slices of `paths`,
        with up to `depth` later batches decoding in the pool meanwhile.  A yielded tensor stays valid until `in_flight + 1`
        more batches have been requested -- what run_stream needs with `in_flight` lanes (engine.py: PipelinedEngine; 1 for a
        plain BetaposeEngine): when it asks for batch k it may still be uploading batches k - in_flight .. k - 1 and has
        synchronised on everything older."""
        paths = list(paths)
        starts = list(range(0, len(paths), batch))
        in_flight = max(1, int(in_flight))
        n_buf = min(len(starts), depth + in_flight + 2)
        # the pinned ring is kept between calls (pinning ~60 MB buffers costs tens of ms each): one stream at a time
        if getattr(self, "_streaming", False):
            raise _lib.BetaposeError("FrameIngest.batches: a previous stream of this FrameIngest is still being consumed")
        ring = getattr(self, "_ring", [])
        if len(ring) < n_buf or (ring and ring[0].shape[0] != batch):
            ring = [self._frame_buffer(batch) for _ in range(n_buf)]
            self._ring = ring
        ring = ring[:max(n_buf, 1)]
        tickets = {}

        def submit(j):
            tickets[j] = self.submit(paths[starts[j]: starts[j] + batch], ring[j % len(ring)])

        self._streaming = True
        try:
            for j in range(min(depth + 1, len(starts))):
                submit(j)
            for j in range(len(starts)):
                fr = self.wait(tickets.pop(j))
                nxt = j + depth + 1
                if nxt < len(starts):
                    submit(nxt)  # reuses the buffer of batch nxt - n_buf = j - in_flight - 1: the caller is done with it (see above)
                yield fr
        finally:
            for t in tickets.values():  # abandoned early: let queued decodes finish before their buffers can be reused
                _lib.lib().bp_ingest_wait(self.handle, t.id)
            self._streaming = False


def convert_sequence(paths, out_dir: str, fmt: str = "ppm", frame_h: int = 480, frame_w: int = 640, n_threads: int = 0, chunk: int = 64):
    """Decode a sequence once and write it as binary PPM (or .npy) frames, which the pool then delivers at memory speed
    (no entropy decoding): the way to keep a B200 fed from disk when the host has fewer than ~20 cores per GPU.
    Returns the new paths, same order."""
    import os

    assert fmt in ("ppm", "npy")
    os.makedirs(out_dir, exist_ok=True)
    paths, out_paths = list(paths), []
    head = f"P6\n{frame_w} {frame_h}\n255\n".encode()
    with FrameIngest(n_threads, frame_h, frame_w) as ing:
        buf = np.empty((chunk, frame_h, frame_w, 3), np.uint8)
        for c0 in range(0, len(paths), chunk):
            part = paths[c0: c0 + chunk]
            frames = ing.decode_files(part, buf)
            for p, fr in zip(part, frames):
                q = os.path.join(out_dir, os.path.splitext(os.path.basename(p))[0] + "." + fmt)
                if fmt == "ppm":
                    with open(q, "wb") as f:
                        f.write(head)
                        f.write(fr.tobytes())
                else:
                    np.save(q, fr)
                out_paths.append(q)
    return out_paths


def _main(argv=None):
    """python -m betapose_b200.ingest INDIR OUTDIR [--fmt ppm|npy] [--frame_h 480 --frame_w 640] [--threads 0]
    Decodes every frame of INDIR once (native pool; Pillow for what it does not read) into OUTDIR as PPM / .npy."""
    import argparse
    import os
    import time

    ap = argparse.ArgumentParser(prog="python -m betapose_b200.ingest")
    ap.add_argument("indir")
    ap.add_argument("outdir")
    ap.add_argument("--fmt", choices=("ppm", "npy"), default="ppm")
    ap.add_argument("--frame_h", type=int, default=480)
    ap.add_argument("--frame_w", type=int, default=640)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args(argv)
    names = sorted(f for f in os.listdir(a.indir) if f.lower().endswith((".png", ".jpg", ".jpeg", ".ppm", ".pgm", ".npy")))
    if not names:
        raise SystemExit(f"no frames in {a.indir}")
    t0 = time.perf_counter()
    out = convert_sequence([os.path.join(a.indir, n) for n in names], a.outdir, a.fmt, a.frame_h, a.frame_w, a.threads)
    dt = time.perf_counter() - t0
    print(f"{len(out)} frames -> {a.outdir} ({a.fmt}) in {dt:.2f} s ({len(out) / dt:.0f} frames/s)")
    return 0


if __name__ == "__main__":
    import sys

    sys.exit(_main())
