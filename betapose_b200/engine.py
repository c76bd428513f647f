"""BetaposeEngine: the whole per-frame evaluate path as one enqueue-only pipeline.

    frames uint8 [B,480,640,3] RGB
      -> bp_resize_bicubic      (a1  dataloader.py:94-99,162)          writes the detector input in place
      -> bp_net_forward (YOLO)  (a2  yolo/darknet.py:319-363)          75 fused tcgen05 conv launches
      -> bp_yolo_decode_argmax  (a3-a5 darknet.py:129-169, util.py:118-223, dataloader.py:350-364)
      -> bp_crop_resize         (a6  dataloader.py:794-835, img.py:242-262) writes the keypoint-net input in place
      -> bp_net_forward (KPD)   (a7  KPD/src/models/FastPose.py:28-35)
      -> bp_heatmap_decode      (a8  KPD/src/utils/eval.py:113-147)
      -> bp_pose_pnp            (a9-a11 pPose_nms.py:24-122, dataloader.py:715-726, utils/utils.py:17-41)
      -> bp_pack_records        (a12 dataloader.py:708-730)
    -> records uint8 [B, sizeof(bp_record)]

Nothing synchronises with the host; every buffer is allocated once in __init__ so a step can be captured in a
CUDA graph (`capture=True`).  One detector + one keypoint network per object id ("slot"); slots share activation
buffers (bp_net_create(share_buffers_with=...)) and a mixed batch is processed slot by slot.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib, net as _net, stages, weights as _weights, yolo_cfg


class BetaposeEngine:
    def __init__(self, max_batch: int, yolo_streams, kpd_state_dicts, kp3d, cfg_blocks=None, frame_h: int = 480,
                 frame_w: int = 640, reso: int = 416, inp_h: int = 320, inp_w: int = 256, n_kp: int = 50,
                 left_number: int = 50, conf: float = 0.01, pnp_mode: int = stages.MODE_RANSAC, reproj_thr: float = 12.0,
                 n_hyp: int = 64, seed: int = 0, cam_K=stages.CAM_K, device=None, concurrent_slots: bool | None = None):
        """yolo_streams: fp32 darknet weight stream (or list, one per object slot); kpd_state_dicts: FastPose
        state_dict (or list); either may be a weights.PackedWeights (packed-weight cache) instead, or a zero-argument
        callable returning one (evaluated when the slot is built); kp3d: float64 [K,3] (or [n_slots,K,3]) key-point model
        in metres."""
        _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.B = int(max_batch)
        self.frame_h, self.frame_w, self.reso, self.inp_h, self.inp_w = frame_h, frame_w, reso, inp_h, inp_w
        self.K, self.left_number, self.conf = int(n_kp), int(left_number), float(conf)
        self.pnp_mode, self.reproj_thr, self.n_hyp, self.seed = int(pnp_mode), float(reproj_thr), int(n_hyp), int(seed)
        self.cam_K = np.asarray(cam_K, np.float64)
        if not isinstance(yolo_streams, (list, tuple)):
            yolo_streams = [yolo_streams]
        if not isinstance(kpd_state_dicts, (list, tuple)):
            kpd_state_dicts = [kpd_state_dicts]
        assert len(yolo_streams) == len(kpd_state_dicts)
        self.n_slots = len(yolo_streams)
        # Mixed-object batches (BASELINE configs[3]): every slot (object) sees only a few frames, so its convolutions fill a
        # fraction of the machine.  With concurrent_slots the slots get their own activation buffers and run on their
        # own CUDA streams (forked from / joined into the caller's stream, also inside graph capture), so their small
        # grids overlap.  Costs (max_batch x 133 MB) of activations per extra slot; default: on for up to 16 slots.
        self.concurrent_slots = (1 < self.n_slots <= 16) if concurrent_slots is None else bool(concurrent_slots and self.n_slots > 1)
        blocks = cfg_blocks if cfg_blocks is not None else yolo_cfg.parse_cfg_text(yolo_cfg.default_cfg_text())
        self.blocks = blocks

        self.yolo: list[_net.Net] = []
        self.kpd: list[_net.Net] = []
        self.heads: list[list[dict]] = []
        self.hm_id: list[int] = []
        with torch.cuda.device(self.device):
            yolo_streams, kpd_state_dicts = list(yolo_streams), list(kpd_state_dicts)
            for s in range(self.n_slots):
                # a slot's weights may be given as a zero-argument callable: produced when the slot is built and released
                # right after the upload (13 objects x ~0.5 GB of fp32 parameters need not sit in host memory at once)
                if callable(yolo_streams[s]):
                    yolo_streams[s] = yolo_streams[s]()
                if callable(kpd_state_dicts[s]):
                    kpd_state_dicts[s] = kpd_state_dicts[s]()
                shared = None if self.concurrent_slots else (self.yolo[0] if s else None)
                y = _net.Net(self.B, reso, reso, _lib.IN_RAW255, share=shared, device=self.device.index)
                if isinstance(yolo_streams[s], _weights.PackedWeights):  # packed-weight cache: shapes from the cfg, data as packed
                    assert yolo_streams[s].kind == "darknet" and yolo_streams[s].meta.get("reso") == reso
                    params, y.packed = _weights.darknet_placeholder_params(blocks), iter(yolo_streams[s])
                else:
                    params, used = _net.split_darknet_stream(blocks, np.asarray(yolo_streams[s], np.float32))
                self.heads.append(_net.build_darknet(y, blocks, params))
                if getattr(y, "packed", None) is not None:
                    _net.packed_exhausted(y.packed)
                self.yolo.append(y)
                k = _net.Net(self.B, inp_h, inp_w, _lib.IN_F16, share=None if self.concurrent_slots else (self.kpd[0] if s else None),
                             device=self.device.index)
                sd = kpd_state_dicts[s]
                if isinstance(sd, _weights.PackedWeights):
                    assert sd.kind == "fastpose" and sd.meta.get("n_maps") == self.K
                    sd, k.packed = _weights.fastpose_placeholder_state_dict(self.K), iter(sd)
                self.hm_id.append(_net.build_fastpose(k, sd, self.K))
                if getattr(k, "packed", None) is not None:
                    _net.packed_exhausted(k.packed)
                self.kpd.append(k)
                yolo_streams[s] = kpd_state_dicts[s] = None  # release the host copy of this slot's parameters
            # measured tile plans for this batch size, where a table exists (betapose_b200/tune.py)
            from . import tune as _tune

            tab = None if self.concurrent_slots else _tune.load(self.B)
            self.tuned_ops = 0
            if tab:
                for net, key in [(n, "yolo") for n in self.yolo] + [(n, "kpd") for n in self.kpd]:
                    self.tuned_ops += _tune.apply(net, self.B, tab.get(key, {}))
            if self.concurrent_slots:
                # the slots of a mixed batch split the machine in proportion to their frame counts (bp_net_set_share)
                for net in self.yolo + self.kpd:
                    net.set_share(self.B)
            kp3d = np.asarray(kp3d, np.float64)
            if kp3d.ndim == 2:
                kp3d = np.broadcast_to(kp3d[None], (self.n_slots,) + kp3d.shape)
            assert kp3d.shape == (self.n_slots, self.K, 3), kp3d.shape
            self.kp3d = torch.from_numpy(np.array(kp3d, np.float64, copy=True)).to(self.device)
            dev, B, K = self.device, self.B, self.K
            f32, i32, u8 = torch.float32, torch.int32, torch.uint8
            # The small per-frame tensors exist TWICE ("parity" 0 / 1): with the PnP tail of step j running on a side stream
            # while the main part of step j + 1 is already under way (PipelinedEngine, async tail), consecutive steps of an
            # engine must not share them.  `self.det`, `self.records`, ... are rebound to one set by _select(parity); every
            # single-stream entry point uses set 0.
            def small():
                return dict(
                    det=torch.zeros((B, 8), dtype=f32, device=dev), box=torch.zeros((B, 4), dtype=f32, device=dev),
                    det_score=torch.zeros((B,), dtype=f32, device=dev), row=torch.zeros((B,), dtype=i32, device=dev),
                    valid=torch.zeros((B,), dtype=u8, device=dev), pt1=torch.zeros((B, 2), dtype=f32, device=dev),
                    pt2=torch.zeros((B, 2), dtype=f32, device=dev), preds_hm=torch.zeros((B, K, 2), dtype=f32, device=dev),
                    preds_img=torch.zeros((B, K, 2), dtype=f32, device=dev), maxval=torch.zeros((B, K), dtype=f32, device=dev),
                    hm_idx=torch.zeros((B, K), dtype=i32, device=dev), keypoints=torch.zeros((B, K, 2), dtype=f32, device=dev),
                    kp_score=torch.zeros((B, K), dtype=f32, device=dev), proposal=torch.zeros((B,), dtype=f32, device=dev),
                    selected=torch.zeros((B, K), dtype=u8, device=dev), R=torch.zeros((B, 9), dtype=torch.float64, device=dev),
                    t=torch.zeros((B, 3), dtype=torch.float64, device=dev), inlier=torch.zeros((B, K), dtype=u8, device=dev),
                    status=torch.zeros((B,), dtype=i32, device=dev),
                    records=torch.zeros((B, _lib.RECORD_BYTES), dtype=u8, device=dev),
                    model_idx=torch.zeros((B,), dtype=i32, device=dev))

            self._sets = [small(), small()]
            self._select(0)
            self.img_idx = torch.arange(B, dtype=i32, device=dev)
            self.frames = torch.zeros((B, frame_h, frame_w, 3), dtype=u8, device=dev)
        self._cam = (C.c_double * 4)(*stages.pinhole4(self.cam_K))
        self._head_args = []
        for s in range(self.n_slots):
            hs = self.heads[s]
            infos = [self.yolo[s].tensor_info(h["tensor"]) for h in hs]
            nh = len(hs)
            self._head_args.append(dict(
                ptrs=(C.c_void_p * nh)(*[i["ptr"] for i in infos]),
                grids=(C.c_int * nh)(*[h["grid"] for h in hs]),
                pitches=(C.c_int * nh)(*[i["pitch"] for i in infos]),
                anchors=(C.c_float * (nh * 6))(*[float(v) for h in hs for wh in h["anchors"] for v in wh]),
                n=nh, n_attr=5 + hs[0]["classes"]))
        self._graphs: dict = {}
        self.launches_per_step = 0

    # ------------------------------------------------------------------------------------------------
    def _select(self, parity: int) -> None:
        """bind self.det, self.box, ..., self.records to small-tensor set `parity`"""
        self.__dict__.update(self._sets[parity & 1])
        self._parity = parity & 1

    @property
    def flops_per_image(self) -> float:
        return self.yolo[0].flops_per_image + self.kpd[0].flops_per_image

    def count_launches(self, n_slots_in_batch: int = 1) -> int:
        """kernel launches one step enqueues: per slot the (fused two-pass) resize + detector ops + decode + crop +
        keypoint-net ops + heat-map decode; then pnp (hypotheses + refine) + pack once."""
        per_slot = 1 + self.yolo[0].num_ops + 1 + 1 + self.kpd[0].num_ops + 1
        return per_slot * n_slots_in_batch + 3

    def _enqueue_slot(self, slot: int, b0: int, n: int, st) -> None:
        """Run images [b0, b0+n) of the current batch through slot `slot`'s networks (frames already on device)."""
        L, e = _lib.lib(), self.yolo[slot].engine.handle
        yolo, kpd = self.yolo[slot], self.kpd[slot]
        off = lambda t: C.c_void_p(t.data_ptr() + b0 * t.stride(0) * t.element_size())  # noqa: E731
        # a1: resize straight into the detector's input buffer
        _lib.check(L.bp_resize_bicubic(e, off(self.frames), n, self.frame_h, self.frame_w, self.reso, self.reso,
                                       C.c_void_p(yolo.tensor_info(0)["ptr"]), None, st), "bp_resize_bicubic")
        _lib.check(L.bp_net_forward(yolo.handle, n, st), "bp_net_forward(yolo)")
        ha = self._head_args[slot]
        _lib.check(L.bp_yolo_decode_argmax(e, ha["ptrs"], ha["grids"], ha["pitches"], ha["n"], ha["anchors"], ha["n_attr"], n,
                                           self.reso, self.conf, self.frame_w, self.frame_h, off(self.det), off(self.box),
                                           off(self.det_score), off(self.row), off(self.valid), None, st), "bp_yolo_decode_argmax")
        # a6: crop straight into the keypoint net's input buffer (img_idx is relative to the frame slice)
        _lib.check(L.bp_crop_resize(e, off(self.frames), self.frame_h, self.frame_w, off(self.box), _lib.ptr(self.img_idx),
                                    off(self.valid), n, self.inp_h, self.inp_w, C.c_void_p(kpd.tensor_info(0)["ptr"]), None,
                                    off(self.pt1), off(self.pt2), st), "bp_crop_resize")
        _lib.check(L.bp_net_forward(kpd.handle, n, st), "bp_net_forward(kpd)")

    def _enqueue_decode(self, slot: int, b0: int, n: int, st) -> None:
        """a8 for the images of one slot (kept on the caller's stream: the engine-level scratch of the sliced arg-max is
        shared between calls)."""
        L, e = _lib.lib(), self.yolo[slot].engine.handle
        kpd = self.kpd[slot]
        off = lambda t: C.c_void_p(t.data_ptr() + b0 * t.stride(0) * t.element_size())  # noqa: E731
        hi = kpd.tensor_info(self.hm_id[slot])
        _lib.check(L.bp_heatmap_decode(e, C.c_void_p(hi["ptr"]), hi["H"] * hi["W"] * hi["pitch"], 1, hi["pitch"], n, self.K,
                                       hi["H"], hi["W"], self.inp_h, self.inp_w, off(self.pt1), off(self.pt2),
                                       off(self.preds_hm), off(self.preds_img), off(self.maxval), off(self.hm_idx), st),
                   "bp_heatmap_decode")

    def _enqueue_tail(self, n: int, image_index0: int, st) -> None:
        self._enqueue_pnp(n, st)
        self._enqueue_pack(n, image_index0, st)

    def _enqueue_pnp(self, n: int, st) -> None:
        L, e = _lib.lib(), self.yolo[0].engine.handle
        _lib.check(L.bp_pose_pnp(e, _lib.ptr(self.preds_img), _lib.ptr(self.maxval), _lib.ptr(self.det_score),
                                 _lib.ptr(self.valid), n, self.K, _lib.ptr(self.kp3d), _lib.ptr(self.model_idx),
                                 C.cast(self._cam, C.c_void_p), self.left_number, self.pnp_mode, getattr(self, "pnp_flags", 0), self.reproj_thr, self.n_hyp,
                                 self.seed & 0xFFFFFFFF, _lib.ptr(self.keypoints), _lib.ptr(self.kp_score),
                                 _lib.ptr(self.proposal), _lib.ptr(self.selected), _lib.ptr(self.R), _lib.ptr(self.t),
                                 _lib.ptr(self.inlier), _lib.ptr(self.status), st), "bp_pose_pnp")

    def _enqueue_pack(self, n: int, image_index0: int, st) -> None:
        L, e = _lib.lib(), self.yolo[0].engine.handle
        _lib.check(L.bp_pack_records(e, n, self.K, int(image_index0), _lib.ptr(self.box), _lib.ptr(self.det_score),
                                     _lib.ptr(self.keypoints), _lib.ptr(self.kp_score), _lib.ptr(self.proposal),
                                     _lib.ptr(self.R), _lib.ptr(self.t), _lib.ptr(self.status), _lib.ptr(self.records), st),
                   "bp_pack_records")

    def _enqueue(self, n: int, groups, image_index0: int, parts: str = "all") -> None:
        """parts: "all" = the whole step; "main" = a1-a8 (resize ... heat-map decode); "tail" = a9-a12 (pose-NMS + PnP + record
        packing), which only reads the small tensors the main part left (PipelinedEngine runs it on a side stream)."""
        st = _lib.stream_ptr()
        if parts == "tail":
            self._enqueue_tail(n, image_index0, st)
            return
        if self.n_slots > 1:
            # the key-point model each frame is solved against follows from the grouping of THIS call (a stale
            # model_idx from an earlier mixed-object batch would pair slot s's networks with another object's model)
            for slot, b0, cnt in groups:
                self.model_idx[b0:b0 + cnt].fill_(int(slot))
        if self.concurrent_slots and len(groups) > 1:
            main = torch.cuda.current_stream()
            if not hasattr(self, "_slot_streams"):
                self._slot_streams = [torch.cuda.Stream() for _ in range(self.n_slots)]
            for slot, b0, cnt in groups:
                side = self._slot_streams[slot]
                side.wait_stream(main)  # fork (also valid under CUDA-graph capture)
                with torch.cuda.stream(side):
                    self._enqueue_slot(slot, b0, cnt, _lib.stream_ptr())
            for slot, _, _ in groups:
                main.wait_stream(self._slot_streams[slot])  # join
            for slot, b0, cnt in groups:
                self._enqueue_decode(slot, b0, cnt, st)
        else:
            for slot, b0, cnt in groups:  # slots share activation buffers: decode a slot's heat-maps before the next slot runs
                self._enqueue_slot(slot, b0, cnt, st)
                self._enqueue_decode(slot, b0, cnt, st)
        if parts == "all":
            self._enqueue_tail(n, image_index0, st)

    # ------------------------------------------------------------------------------------------------
    def run_device(self, n: int | None = None, groups=None, image_index0: int = 0, graph: bool = False, parts: str = "all",
                   parity: int = 0) -> torch.Tensor:
        """Process the first n frames already resident in `self.frames` (grouped by slot: list of (slot, first, count),
        frames of one slot contiguous; `self.model_idx` is rewritten from the grouping).  Returns the records tensor view [n, RECORD_BYTES]
        (device); nothing has synchronised.  `parts` / `parity`: run only the main part or only the tail of the step, on small-tensor
        set `parity` (see _enqueue, _select); the caller orders "tail" after "main" of the same parity (PipelinedEngine does)."""
        n = self.B if n is None else int(n)
        groups = [(0, 0, n)] if groups is None else list(groups)
        assert parts in ("all", "main", "tail")
        with torch.cuda.device(self.device):
            self._select(parity)
            try:
                if not graph:
                    self._enqueue(n, groups, image_index0, parts)
                else:
                    key = (n, tuple(groups), image_index0) if (parts == "all" and not parity) else (n, tuple(groups), image_index0, parts, parity & 1)
                    g = self._graphs.get(key)
                    if g is None:
                        # warm-up ON the capture stream (the library keeps its kernel scratch per stream: lazy allocations,
                        # function attributes and descriptor builds for this batch size all happen here, outside capture),
                        # then capture on that same stream.  One capture stream per engine: graphs of different engines
                        # never share scratch.
                        if not hasattr(self, "_capture_stream"):
                            self._capture_stream = torch.cuda.Stream()
                        cap = self._capture_stream
                        cap.wait_stream(torch.cuda.current_stream())
                        with torch.cuda.stream(cap):
                            self._enqueue(n, groups, image_index0, parts)
                        cap.synchronize()
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=cap):
                            self._enqueue(n, groups, image_index0, parts)
                        self._graphs[key] = g
                    g.replay()
                rec = self.records[:n]
            finally:
                self._select(0)
        self.launches_per_step = self.count_launches(len(groups))
        return rec

    def profile_stages(self, n: int | None = None, reps: int = 5) -> dict:
        """Device time of every stage of one step on the frames resident in `self.frames` (slot 0), CUDA events between the
        stages, median of `reps` eager passes: the reference's `--profile` readout (betapose_evaluate.py:132-136,178-186:
        detection / pose / post-processing wall-clock lists without any synchronisation) measured where the work runs.
        -> {'resize', 'detector', 'decode_argmax', 'crop', 'keypoint_net', 'heatmap_decode', 'pose_pnp', 'pack', 'total'} in ms."""
        n = self.B if n is None else int(n)
        L, e, st = _lib.lib(), self.yolo[0].engine.handle, _lib.stream_ptr()
        yolo, kpd, ha = self.yolo[0], self.kpd[0], self._head_args[0]
        names = ["resize", "detector", "decode_argmax", "crop", "keypoint_net", "heatmap_decode", "pose_pnp", "pack"]
        with torch.cuda.device(self.device):
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)] for _ in range(reps + 1)]
            for r in range(reps + 1):  # pass 0 = warm-up
                k = iter(ev[r])
                next(k).record()
                _lib.check(L.bp_resize_bicubic(e, _lib.ptr(self.frames), n, self.frame_h, self.frame_w, self.reso, self.reso,
                                               C.c_void_p(yolo.tensor_info(0)["ptr"]), None, st), "bp_resize_bicubic")
                next(k).record()
                _lib.check(L.bp_net_forward(yolo.handle, n, st), "bp_net_forward(yolo)")
                next(k).record()
                _lib.check(L.bp_yolo_decode_argmax(e, ha["ptrs"], ha["grids"], ha["pitches"], ha["n"], ha["anchors"], ha["n_attr"], n,
                                                   self.reso, self.conf, self.frame_w, self.frame_h, _lib.ptr(self.det), _lib.ptr(self.box),
                                                   _lib.ptr(self.det_score), _lib.ptr(self.row), _lib.ptr(self.valid), None, st), "bp_yolo_decode_argmax")
                next(k).record()
                _lib.check(L.bp_crop_resize(e, _lib.ptr(self.frames), self.frame_h, self.frame_w, _lib.ptr(self.box), _lib.ptr(self.img_idx),
                                            _lib.ptr(self.valid), n, self.inp_h, self.inp_w, C.c_void_p(kpd.tensor_info(0)["ptr"]), None,
                                            _lib.ptr(self.pt1), _lib.ptr(self.pt2), st), "bp_crop_resize")
                next(k).record()
                _lib.check(L.bp_net_forward(kpd.handle, n, st), "bp_net_forward(kpd)")
                next(k).record()
                self._enqueue_decode(0, 0, n, st)
                next(k).record()
                self._enqueue_pnp(n, st)
                next(k).record()
                self._enqueue_pack(n, 0, st)
                next(k).record()
            torch.cuda.synchronize()
        out = {nm: float(np.median([ev[r][i].elapsed_time(ev[r][i + 1]) for r in range(1, reps + 1)])) for i, nm in enumerate(names)}
        out["total"] = float(np.median([ev[r][0].elapsed_time(ev[r][-1]) for r in range(1, reps + 1)]))
        return out

    def run_stream(self, batches, graph: bool = True, image_index0: int = 0, after_step=None):
        """Pipelined evaluation of a stream of frame batches on this engine alone (one batch on the GPU at a time; the
        next batch's host->device copy overlaps it).  See PipelinedEngine.run_stream, which this delegates to with a
        single lane; PipelinedEngine(2, ...) keeps two batches in flight on the GPU."""
        if not hasattr(self, "_single_lane"):
            self._single_lane = PipelinedEngine.from_engines([self])
        yield from self._single_lane.run_stream(batches, graph=graph, image_index0=image_index0, after_step=after_step)

    def run(self, frames_u8, obj_slots=None, image_index0: int = 0, graph: bool = False) -> np.ndarray:
        """Public end-to-end call: frames uint8 [n,H,W,3] RGB (host numpy / pinned torch / cuda tensor), optional
        per-frame object slot ids.  Returns the host structured array of bp_record (one per frame, input order)."""
        fr = torch.as_tensor(frames_u8)
        n = int(fr.shape[0])
        assert n <= self.B and tuple(fr.shape[1:]) == (self.frame_h, self.frame_w, 3) and fr.dtype == torch.uint8
        order = None
        groups = [(0, 0, n)]
        if obj_slots is not None and self.n_slots > 1:
            slots = np.asarray(obj_slots, np.int64)
            order = np.argsort(slots, kind="stable")
            groups, s0 = [], 0
            sorted_slots = slots[order]
            for s in np.unique(sorted_slots):
                cnt = int((sorted_slots == s).sum())
                groups.append((int(s), s0, cnt))
                s0 += cnt
            idx = torch.from_numpy(order)
            fr = fr[idx.to(fr.device)] if fr.is_cuda else fr[idx]
        self.frames[:n].copy_(fr, non_blocking=True)
        rec = self.run_device(n, groups, image_index0, graph=graph)
        out = stages.records_to_numpy(rec)
        if order is not None:
            inv = np.empty_like(order)
            inv[order] = np.arange(n)
            out = out[inv]
            out["image_index"] = image_index0 + np.arange(n)
        return out


class PipelinedEngine:
    """L independent BetaposeEngine "lanes" (own activation buffers, own captured graph, own CUDA stream) that take the
    batches of a stream in turn, so L batches are in flight on the GPU at once.

    Why: one batch alone cannot keep 148 SMs busy all the time -- the second wave of a 160-tile layer, the 1x1-spatial SE
    layers, the decode / PnP kernels (fp64 latency chains on a few warps) and every kernel boundary leave SMs idle.  With
    a second batch in flight the hardware block scheduler fills those holes with the other lane's CTAs (measured on B200,
    batch 64: 6 598 -> 7 143 images/s with two lanes; a third lane adds nothing; splitting ONE batch into halves loses,
    `scripts/exp_two_streams.py`).  The reference is a pipeline of stages too (dataloader.py: ImageLoader ->
    DetectionLoader -> DetectionProcessor -> DataWriter threads with queues between them); this is its device-side
    equivalent.  Costs one extra set of activation buffers (~133 MB per frame of batch) and one extra weight copy."""

    def __init__(self, lanes: int, max_batch: int, *args, async_tail: bool = True, **kw):
        assert lanes >= 1
        self._init([BetaposeEngine(max_batch, *args, **kw) for _ in range(int(lanes))], async_tail)

    @classmethod
    def from_engines(cls, engines, async_tail: bool = True):
        self = cls.__new__(cls)
        self._init(list(engines), async_tail)
        return self

    def _init(self, engines, async_tail: bool = True):
        self.lanes = engines
        e0 = engines[0]
        self.device, self.B = e0.device, e0.B
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream() for _ in engines] if len(engines) > 1 else [None]
            # Asynchronous tail: pose-NMS + PnP + record packing of a lane's step j (two fp64 latency chains on a few warps,
            # ~0.8 ms at batch 64, 1 KB of data per frame) run on the lane's high-priority TAIL stream while the lane's main
            # stream already works on step j + 1 (resize, detector, ...): the tail is off the critical path of the lane.
            # Consecutive steps of a lane alternate between the engine's two small-tensor sets; the main part of step j + 2
            # waits for the tail of step j.
            self.async_tail = bool(async_tail) and os.environ.get("BP_ASYNC_TAIL", "1") != "0"  # env: experiment switch
            self.tail_streams = [torch.cuda.Stream(priority=-1) for _ in engines] if self.async_tail else [None] * len(engines)
            self._ev_main = [[torch.cuda.Event() for _ in range(2)] for _ in engines]
            self._ev_tail = [[torch.cuda.Event() for _ in range(2)] for _ in engines]
            self._lane_steps = [0] * len(engines)
            self._primed = [set() for _ in engines]
            self._copy_stream = torch.cuda.Stream()
            self._stage = [torch.empty_like(e.frames) for e in engines]
            # two host record buffers per lane: batch k's records may still be in the caller's hands when batch k + L lands
            self._rec_host = [[torch.empty((e.B, _lib.RECORD_BYTES), dtype=torch.uint8).pin_memory() for _ in range(2)] for e in engines]

    @property
    def launches_per_step(self) -> int:
        return self.lanes[0].launches_per_step

    @property
    def flops_per_image(self) -> float:
        return self.lanes[0].flops_per_image

    def lane_stream(self, k: int):
        """(engine, stream) that batch number k runs on; stream is None for a single lane (= the caller's stream)."""
        i = k % len(self.lanes)
        return self.lanes[i], self.streams[i]

    def _step(self, i: int, lane_stream, n: int, graph: bool, after_step=None, rec_host=None, ev_done=None):
        """Enqueue one step of lane i whose frames are already in the lane's `frames` buffer (ordered on `lane_stream`):
        main part on the lane's stream, tail (+ after_step + the records' device->host copy) on the tail stream when the
        tail is asynchronous, else everything on the lane's stream.  Returns the device records view."""
        eng = self.lanes[i]
        if not self.async_tail:
            rec = eng.run_device(n, graph=graph)
            tail = lane_stream
        else:
            par = self._lane_steps[i] & 1
            self._lane_steps[i] += 1
            tail = self.tail_streams[i]
            if graph and n not in self._primed[i]:
                # first step of this lane at this batch size: capture the graphs of BOTH small-tensor sets now (the other
                # set's main + tail run once on the same frames), so that no later step pays for a capture
                self._primed[i].add(n)
                # (both sets are written here: the tails of the lane's earlier steps -- other batch sizes -- must be done with them)
                lane_stream.wait_event(self._ev_tail[i][0])
                lane_stream.wait_event(self._ev_tail[i][1])
                with torch.cuda.stream(lane_stream):
                    eng.run_device(n, graph=True, parts="main", parity=par ^ 1)
                    eng.run_device(n, graph=True, parts="tail", parity=par ^ 1)
                    eng.run_device(n, graph=True, parts="main", parity=par)
                    eng.run_device(n, graph=True, parts="tail", parity=par)
            lane_stream.wait_event(self._ev_tail[i][par])  # tail of this lane's step j - 2 has released small-tensor set `par`
            eng.run_device(n, graph=graph, parts="main", parity=par)
            self._ev_main[i][par].record(lane_stream)
            tail.wait_event(self._ev_main[i][par])
            with torch.cuda.stream(tail):
                rec = eng.run_device(n, graph=graph, parts="tail", parity=par)
        with torch.cuda.stream(tail):
            if after_step is not None:
                after_step(rec)
            if rec_host is not None:
                rec_host[:n].copy_(rec, non_blocking=True)
            if self.async_tail:
                self._ev_tail[i][par].record(tail)
            if ev_done is not None:
                ev_done.record(tail)
        return rec

    def submit_device(self, k: int, frames_dev: torch.Tensor | None = None, n: int | None = None, graph: bool = True,
                      after_step=None) -> torch.Tensor:
        """Enqueue batch k (frames already in HBM, or already in the lane's `frames` buffer when frames_dev is None) on its
        lane's stream(s); returns the lane's device records view (valid once the lane's tail stream has run: `join()`).  The
        caller orders the lane streams against its own stream (fork / join) when it needs to."""
        eng, st = self.lane_stream(k)
        n = eng.B if n is None else int(n)
        with torch.cuda.device(self.device):
            lane_stream = st if st is not None else torch.cuda.current_stream()
            with torch.cuda.stream(lane_stream):
                if frames_dev is not None:
                    eng.frames[:n].copy_(frames_dev[:n], non_blocking=True)
                rec = self._step(k % len(self.lanes), lane_stream, n, graph, after_step)
        return rec

    def fork(self) -> None:
        """lane streams wait for everything enqueued so far on the caller's stream"""
        cur = torch.cuda.current_stream()
        for s in list(self.streams) + list(self.tail_streams):
            if s is not None:
                s.wait_stream(cur)

    def join(self) -> None:
        """the caller's stream waits for everything enqueued so far on the lanes (main and tail streams)"""
        cur = torch.cuda.current_stream()
        for s in list(self.streams) + list(self.tail_streams):
            if s is not None:
                cur.wait_stream(s)

    def run_stream(self, batches, graph: bool = True, image_index0: int = 0, after_step=None):
        """Pipelined evaluation of a stream of frame batches.  `batches` yields uint8 [n <= max_batch, H, W, 3] RGB host
        arrays / tensors (pinned memory for a truly asynchronous copy).  Batch k runs on lane k % L: its host->device copy
        goes over a side stream into the lane's staging buffer while earlier batches compute, the lane's stream then copies
        it into place, replays the lane's step graph and reads the records back.  With L lanes the iterator is pulled L
        batches ahead of the records being yielded.  Yields one host record array per batch, in input order.
        `after_step(records_device)` is enqueued on the lane's stream after each batch (e.g. the all-gather)."""
        from collections import deque

        L = len(self.lanes)
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream()
            streams = [s if s is not None else cur for s in self.streams]
            ev_h2d = [torch.cuda.Event() for _ in range(L)]
            ev_free = [torch.cuda.Event() for _ in range(L)]   # staging buffer consumed by the lane's stream
            ev_done = [[torch.cuda.Event() for _ in range(2)] for _ in range(L)]
            for i in range(L):
                if streams[i] is not cur:
                    streams[i].wait_stream(cur)
                if self.tail_streams[i] is not None:
                    self.tail_streams[i].wait_stream(cur)
                ev_free[i].record(streams[i])
            pending = deque()
            it = iter(batches)
            k, idx0, exhausted = 0, int(image_index0), False

            def launch(fr, k, idx0):
                fr = torch.as_tensor(fr)
                n = int(fr.shape[0])
                i = k % L
                eng = self.lanes[i]
                assert n <= eng.B and tuple(fr.shape[1:]) == (eng.frame_h, eng.frame_w, 3) and fr.dtype == torch.uint8
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(ev_free[i])
                    self._stage[i][:n].copy_(fr, non_blocking=True)
                    ev_h2d[i].record(self._copy_stream)
                h = (k // L) & 1
                with torch.cuda.stream(streams[i]):
                    streams[i].wait_event(ev_h2d[i])
                    eng.frames[:n].copy_(self._stage[i][:n], non_blocking=True)  # device->device, ~0.03 ms for 64 frames
                    ev_free[i].record(streams[i])
                    # one captured graph (pair) per batch size; image indices are fixed up on the host below
                    self._step(i, streams[i], n, graph, after_step, self._rec_host[i][h], ev_done[i][h])
                pending.append((i, h, n, idx0))

            try:
                while True:
                    while not exhausted and len(pending) < L:
                        try:
                            nxt = next(it)           # host work (frame decoding, ...) overlaps the GPU
                        except StopIteration:
                            exhausted = True
                            break
                        launch(nxt, k, idx0)
                        idx0 += int(torch.as_tensor(nxt).shape[0])
                        k += 1
                    if not pending:
                        break
                    i, h, n, first = pending.popleft()
                    # keep the pipe full while the oldest batch finishes: the next batch of this lane can be fetched and uploaded now
                    if not exhausted and len(pending) < L:
                        try:
                            nxt = next(it)
                            launch(nxt, k, idx0)
                            idx0 += int(torch.as_tensor(nxt).shape[0])
                            k += 1
                        except StopIteration:
                            exhausted = True
                    ev_done[i][h].synchronize()
                    out = stages.records_to_numpy(self._rec_host[i][h][:n]).copy()
                    out["image_index"] = first + np.arange(n)
                    yield out
            finally:
                # also when the consumer abandons the stream early (generator closed): whatever is still in flight on the lane
                # and tail streams is ordered before the caller's next work on its own stream (e.g. a plain run() afterwards,
                # which uses small-tensor set 0 on the caller's stream)
                for s in streams + [t for t in self.tail_streams if t is not None]:
                    if s is not cur:
                        cur.wait_stream(s)


class _nullctx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False
