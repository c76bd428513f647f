"""Darknet .cfg handling for the detector.

`parse_cfg` accepts the same file format the reference reads (3_6Dpose_estimator/yolo/darknet.py:45-74): one
dict per ``[section]``, string values, comments and blank lines skipped.  `default_cfg_text` *generates* the
single-class YOLOv3 topology the reference ships as ``yolo/cfg/yolov3-single.cfg`` from its structural
description (107 blocks = 75 conv, 23 shortcut, 4 route, 2 upsample, 3 yolo) so tests and the
benchmark do not depend on the reference tree; tests/test_yolo_cfg.py checks both parse to identical blocks
where the reference is available.
"""
from __future__ import annotations

ANCHORS = "10,13,  16,30,  33,23,  30,61,  62,45,  59,119,  116,90,  156,198,  373,326"


def parse_cfg_text(text: str) -> list[dict]:
    blocks: list[dict] = []
    cur: dict = {}
    for raw in text.split("\n"):
        line = raw.strip()
        if not line or line[0] == "#":
            continue
        if line[0] == "[":
            if cur:
                blocks.append(cur)
            cur = {"type": line[1:-1].rstrip()}
        else:
            key, value = line.split("=", 1)
            cur[key.rstrip()] = value.lstrip()
    if cur:
        blocks.append(cur)
    return blocks


def parse_cfg(cfgfile: str) -> list[dict]:
    with open(cfgfile, "r") as f:
        return parse_cfg_text(f.read())


def _conv(filters: int, size: int, stride: int = 1, bn: bool = True, act: str = "leaky") -> str:
    s = "[convolutional]\n"
    if bn:
        s += "batch_normalize=1\n"
    s += f"filters={filters}\nsize={size}\nstride={stride}\npad=1\nactivation={act}\n\n"
    return s


def _residual(ch: int) -> str:
    return _conv(ch // 2, 1) + _conv(ch, 3) + "[shortcut]\nfrom=-3\nactivation=linear\n\n"


def _yolo(mask: str, classes: int) -> str:
    return (f"[yolo]\nmask = {mask}\nanchors = {ANCHORS}\nclasses={classes}\nnum=9\njitter=.5\n"
            "ignore_thresh = .7\ntruth_thresh = 1\nrandom=1\n\n")


def default_cfg_text(classes: int = 1) -> str:
    """YOLOv3 (Darknet-53 backbone, 3 heads) with ``classes`` classes and no ``[net]`` section."""
    head_filters = 3 * (5 + classes)
    t = _conv(32, 3)
    for ch, reps in ((64, 1), (128, 2), (256, 8), (512, 8), (1024, 4)):
        t += _conv(ch, 3, stride=2)
        for _ in range(reps):
            t += _residual(ch)
    # head at stride 32
    for _ in range(3):
        t += _conv(512, 1) + _conv(1024, 3)
    t += _conv(head_filters, 1, bn=False, act="linear") + _yolo("6,7,8", classes)
    # head at stride 16
    t += "[route]\nlayers = -4\n\n" + _conv(256, 1) + "[upsample]\nstride=2\n\n[route]\nlayers = -1, 61\n\n"
    for _ in range(3):
        t += _conv(256, 1) + _conv(512, 3)
    t += _conv(head_filters, 1, bn=False, act="linear") + _yolo("3,4,5", classes)
    # head at stride 8
    t += "[route]\nlayers = -4\n\n" + _conv(128, 1) + "[upsample]\nstride=2\n\n[route]\nlayers = -1, 36\n\n"
    for _ in range(3):
        t += _conv(128, 1) + _conv(256, 3)
    t += _conv(head_filters, 1, bn=False, act="linear") + _yolo("0,1,2", classes)
    return t
