"""GPU: the rollout kernel against the oracle on fresh games beyond the fixtures (tests/parity_at_scale.py at suite size)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_rollouts_match_oracle_on_fresh_games():
    n = os.environ.get("SB_SCALE_N", "6000")
    tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), "parity_at_scale.py")
    out = subprocess.run([sys.executable, tool, n], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if "mismatches" in l]
    assert len(lines) == 4 and all(" 0 mismatches" in l for l in lines), out.stdout[-2000:]
