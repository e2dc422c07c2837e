"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel path (graphical-gan_b200/gg/dist.py) —
batch sharding, the flat gradient bucket with ONE summing all-reduce per optimiser step and 1/P scaling, and the SyncBN
statistic exchange (sum / sum-of-squares partials) — without a GPU."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, os.path.join(%r, "graphical-gan_b200"))
    import numpy as np, torch
    from gg import dist as ggdist
    rank, world = ggdist.init_from_env(backend="gloo")
    assert world == 2 and ggdist.world_size() == 2 and ggdist.rank() == rank
    rs = np.random.RandomState(0)
    X = rs.randn(64, 10)                       # the global batch (same on both ranks)
    W = [rs.randn(10, 3), rs.randn(3)]         # replicated parameters
    lo, hi = ggdist.shard_bounds(64)
    assert (lo, hi) == (rank * 32, rank * 32 + 32)
    xs = ggdist.shard(X)
    # local gradients of the LOCAL mean loss 0.5*mean((xW+b)^2)
    def grads(x):
        y = x @ W[0] + W[1]
        return [x.T @ y / x.shape[0] / 1.0, y.mean(0)]
    g_local = grads(xs)
    offs, total = ggdist.bucket_offsets([g.size for g in g_local])
    assert offs == [0, 30] and total == 33
    flat = torch.zeros(total, dtype=torch.float64)
    for o, g in zip(offs, g_local):
        flat[o:o + g.size] = torch.from_numpy(g.reshape(-1))
    ggdist.all_reduce_sum(flat)                # the ONE exchange of the step
    flat *= 1.0 / world                        # grad_scale of gg_adam_multi
    g_full = grads(X)
    for o, g in zip(offs, g_full):
        assert np.allclose(flat[o:o + g.size].numpy(), g.reshape(-1), rtol=1e-12), "bucketed DP gradient != full-batch gradient"
    # SyncBN: all-reduced [sum x, sum x^2] give the full-batch mean / biased variance
    part = torch.from_numpy(np.stack([xs.sum(0), (xs ** 2).sum(0)]))
    ggdist.all_reduce_sum(part)
    mean = part[0] / 64
    var = part[1] / 64 - mean ** 2
    assert np.allclose(mean.numpy(), X.mean(0)) and np.allclose(var.numpy(), X.var(0))
    try:
        ggdist.shard_bounds(63)
        raise SystemExit("expected ValueError for an uneven batch")
    except ValueError:
        pass
    torch.distributed.barrier()
    print("rank" + str(rank) + "-ok", flush=True)
''') % ROOT


def test_two_rank_gloo_bucket_allreduce_and_syncbn(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "rank0-ok" in out.stdout and "rank1-ok" in out.stdout, out.stdout


def test_single_process_defaults():
    sys.path.insert(0, os.path.join(ROOT, "graphical-gan_b200"))
    from gg import dist as ggdist
    assert ggdist.world_size() == 1 and ggdist.rank() == 0
    assert ggdist.shard_bounds(64) == (0, 64)
    a = np.arange(12).reshape(6, 2)
    assert np.array_equal(ggdist.shard(a, 1, 3), a[2:4])
