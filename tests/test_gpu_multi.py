"""GPU (needs >= 2 devices, skipped otherwise): the N-rank evaluation over NCCL gives the SAME integer counts as one rank.
BASELINE config 3 (population 256 x 256 games vs a heuristic baseline), world sizes 1, 2 and -- when the box has them -- 4 and 8.
The same tool is run through `gpurun --gpus N` and its output kept under profiles/ (profiles/r2_nccl_parity.txt)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, out, P, G):
    tool = os.path.join(ROOT, "tools", "nccl_parity.py")
    if world == 1:
        cmd = [sys.executable, tool, out, str(P), str(G)]
    else:
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(port), tool, out, str(P), str(G)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return np.load(out)


def test_counts_identical_for_every_world_size(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    ref = _run(1, str(tmp_path / "w1.npy"), 256, 256)
    assert ref.shape == (256, 3) and ref.sum() == 256 * 256
    for world in [w for w in (2, 4, 8) if w <= n]:
        got = _run(world, str(tmp_path / ("w%d.npy" % world)), 256, 256)
        assert np.array_equal(ref, got), world
