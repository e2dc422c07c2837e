"""-m gpu, needs >= 2 GPUs (skipped otherwise): sharding the global batch of 64 over 2 ranks with ONE gradient all-reduce
per optimiser step + SyncBN reproduces the single-GPU step (same costs, same updated parameters)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_data_parallel_matches_single_gpu(tmp_path):
    single, dp = str(tmp_path / "single.npz"), str(tmp_path / "dp2.npz")
    script = os.path.join(ROOT, "tools", "dp_check.py")
    env = dict(os.environ)
    r = subprocess.run([sys.executable, script, "--out", single], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), script, "--out", dp], capture_output=True, text=True, timeout=300,
                       env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    a, b = np.load(single), np.load(dp)
    print("costs single", a["costs"], "dp2", b["costs"])
    assert np.allclose(a["costs"], b["costs"], rtol=2e-3)
    for k in a.files:
        if k == "costs":
            continue
        diff = np.abs(a[k] - b[k])
        # Adam's first steps move every weight by ~lr*sign(g): an entry whose gradient sits at the rounding-noise level may
        # step the other way on the two runs (different split-K / reduction orders).  Measured: 0.2-0.8 % of entries differ
        # by 2*lr = 4e-4, everything else agrees to < 1e-4.
        assert diff.max() <= 1.5e-3 and (diff > 1e-4).mean() < 0.03, (k, diff.max(), (diff > 1e-4).mean())
