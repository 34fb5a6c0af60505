"""Every selectable kernel form must give the oracle's answer bit for bit: the default split filter, the fused TMA pipeline
and the fused register-staged filter."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("env", [{}, {"YL_FILTER": "fused"}, {"YL_FILTER": "fused", "YL_NO_TMA": "1"}],
                         ids=["default", "fused-tma", "fused-ldg"])
def test_kernel_forms_match_oracle(env):
    e = dict(os.environ)
    e.update(env)
    p = subprocess.run([sys.executable, os.path.join(HERE, "alt_path_check.py")], env=e, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "MISMATCH" not in p.stdout
