"""Measured tile plans for the convolution kernel (replace the cost-model guess by a table).

conv_plan.cuh chooses (BLOCK_N, CTA pairs, pixels per tile) with a small cost model.  `tune_net` times every legal
configuration of every convolution of a built network at one batch size on the GPU it runs on (the layer right before is
re-run each time so the input sits in L2 the way it does inside the real chain), keeps the winners that beat the planner's
choice by a margin, and `save` / `apply` persist them as JSON keyed by the layer's shape signature.  A table lives in
betapose_b200/tuned/<gpu>_b<batch>.json and is applied by BetaposeEngine when the batch size matches (BP_NO_TUNE=1 turns
that off).  No reference counterpart: the reference leaves algorithm choice to cuDNN's heuristics."""
from __future__ import annotations

import json
import os
import re

import numpy as np
import torch

CANDIDATES = [(bn, cg, mt) for bn in (256, 128, 64, 32) for (cg, mt) in ((1, 1), (2, 1), (1, 2), (1, 4))]
_PLAN_RE = re.compile(r" bn\d+ bk\d+ st\d+( cg2| mt\d)?( g4)?( sk)?")
TUNED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tuned")


def shape_key(desc: str) -> str:
    """op description without the plan marks: 'conv 3x3/1 128->256 @52x52 +res'"""
    return _PLAN_RE.sub("", desc)


def time_op(net, batch: int, i: int, reps: int = 12) -> float:
    """median device time (ms) of op i, its predecessor re-run before every sample"""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        if i > 0:
            net.forward(batch, i - 1, i)
        a.record()
        net.forward(batch, i, i + 1)
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev][2:]))


def time_op_sustained(net, batch: int, i: int, dur_s: float = 0.25, per_graph: int = 40) -> float:
    """device time (ms) per launch of op i launched back to back for about `dur_s` seconds (a CUDA graph of `per_graph`
    launches replayed).  On a power-capped GPU this is what a launch costs inside a long-running stream: the clock settles
    where the layer's power draw meets the cap, so the figure ranks configurations by ENERGY per launch, not by cycles at
    the boost clock -- a burst measurement (time_op) favours the configuration with the fewest cycles even when it moves
    more bytes per flop and therefore runs at a lower sustained clock."""
    if i > 0:
        net.forward(batch, i - 1, i)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        net.forward(batch, i, i + 1)
    st.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(per_graph):
            net.forward(batch, i, i + 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    t1 = a.elapsed_time(b) / per_graph  # ms per launch, cold clocks
    n = max(2, int(dur_s * 1e3 / max(t1, 1e-3) / per_graph))
    for _ in range(max(1, n // 3)):  # let the clock settle
        g.replay()
    a.record()
    for _ in range(n):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (n * per_graph)


def tune_net(net, batch: int, margin: float = 0.03, reps: int = 12, log=None, timer=None) -> dict:
    """-> {shape_key: {"cfg": [bn, cg, mt], "ms": best, "default": [bn, cg, mt], "default_ms": ...}} for the convolutions
    where a measured configuration beats the planner's by more than `margin`.  Leaves the winners applied to `net`."""
    net.forward(batch)
    torch.cuda.synchronize()
    if timer is not None:  # e.g. time_op_sustained
        time_op_ = lambda net_, batch_, i_, reps_: timer(net_, batch_, i_)  # noqa: E731
    else:
        time_op_ = time_op
    table: dict = {}
    seen: dict = {}
    for i in range(net.num_ops):
        desc = net.op_desc(i)[0]
        if not desc.startswith("conv"):
            continue
        key = shape_key(desc)
        if key in seen:  # same layer shape again (the nets repeat their blocks): reuse the verdict
            if seen[key] is not None:
                net.set_op_config(i, batch, *seen[key])
            continue
        net.set_op_config(i, batch, 0, 0, 0)
        d_cfg = net.op_config(i, batch)[:3]
        d_ms = time_op_(net, batch, i, reps)
        best, best_ms = None, d_ms
        for cand in CANDIDATES:
            if tuple(cand) == tuple(d_cfg) or not net.set_op_config(i, batch, *cand):
                continue
            try:
                ms = time_op_(net, batch, i, reps)
            except Exception:
                continue
            if ms < best_ms:
                best, best_ms = cand, ms
        if best is not None and best_ms < (1.0 - margin) * d_ms:
            # confirm against the default once more (the first sample of a layer sometimes carries a clock ramp)
            net.set_op_config(i, batch, 0, 0, 0)
            d2 = time_op_(net, batch, i, reps)
            net.set_op_config(i, batch, *best)
            b2 = time_op_(net, batch, i, reps)
            if b2 < (1.0 - margin) * d2:
                table[key] = {"cfg": list(best), "ms": round(b2, 5), "default": list(d_cfg), "default_ms": round(d2, 5)}
                seen[key] = best
                if log:
                    log(f"{key}: {d_cfg} {d2:.4f} ms -> {best} {b2:.4f} ms")
                continue
        net.set_op_config(i, batch, 0, 0, 0)
        seen[key] = None
    return table


def apply(net, batch: int, table: dict) -> int:
    """force the table's configurations on the matching convolutions of `net` at `batch`; returns how many were applied"""
    n = 0
    for i in range(net.num_ops):
        desc = net.op_desc(i)[0]
        ent = table.get(shape_key(desc)) if desc.startswith("conv") else None
        if ent and net.set_op_config(i, batch, *ent["cfg"]):
            n += 1
    return n


def table_path(batch: int, gpu: str = "b200") -> str:
    return os.path.join(TUNED_DIR, f"{gpu}_b{int(batch)}.json")


def load(batch: int):
    """{'yolo': table, 'kpd': table} for this batch size, or None"""
    if os.environ.get("BP_NO_TUNE"):
        return None
    p = table_path(batch)
    if not os.path.isfile(p):
        return None
    try:
        return json.load(open(p))
    except Exception:
        return None


def save(batch: int, yolo: dict, kpd: dict, meta: dict | None = None) -> str:
    os.makedirs(TUNED_DIR, exist_ok=True)
    p = table_path(batch)
    with open(p, "w") as f:
        json.dump({"batch": int(batch), "meta": meta or {}, "yolo": yolo, "kpd": kpd}, f, indent=1, sort_keys=True)
    return p
