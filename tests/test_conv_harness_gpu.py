"""GPU: the stand-alone conv harness (tests/harness/conv_harness.cu, built in-tree by __graft_entry__.build()): every
tile configuration of conv_umma_kernel -- single CTA, CTA pairs (cta_group::2), 256-pixel tiles, the packed stems,
residual / fused-store epilogues, ragged and odd tile counts -- against a naive fp32 GPU convolution."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "betapose_b200", "csrc", "build", "conv_harness")


def test_conv_harness_quick():
    if not os.path.isfile(EXE):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "betapose_b200", "csrc"), "harness"])
    out = subprocess.run([EXE, "quick"], capture_output=True, text=True, timeout=600)
    tail = "\n".join(out.stdout.splitlines()[-25:])
    assert out.returncode == 0, tail
    assert "TOTAL FAILURES: 0" in out.stdout and "probe failures: 0" in out.stdout, tail
    assert out.stdout.count(" ok") >= 50, tail  # the whole matrix ran
