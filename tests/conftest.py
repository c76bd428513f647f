"""Shared pytest plumbing: the `gpu` marker, repo-root imports, seeded synthetic assets."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run on the B200 box with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def yolo_blocks():
    from betapose_b200 import yolo_cfg

    return yolo_cfg.parse_cfg_text(yolo_cfg.default_cfg_text())


@pytest.fixture(scope="session")
def yolo_stream():
    from betapose_b200 import synth

    return synth.cached_yolo_weights(1000)


@pytest.fixture(scope="session")
def kpd_sd():
    from betapose_b200 import synth

    return synth.cached_kpd_state_dict(2000)


@pytest.fixture(scope="session")
def kp_model():
    from betapose_b200 import synth

    return synth.synth_kp_model(1, 50)


@pytest.fixture(scope="session")
def frames8():
    from betapose_b200 import synth

    return synth.synth_frames(8, seed=3)
