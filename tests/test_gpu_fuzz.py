"""Randomised parity: tools/fuzz_parity.py (fused and dense paths vs the CPU oracle over random shapes, class counts,
thresholds, scene densities and a few absurd logits), a fixed seed so that failures reproduce."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_fuzz_parity_fixed_seed():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "60", "3"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_fuzz_loss_and_build_target_fixed_seed():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_loss.py"), "25", "5"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
