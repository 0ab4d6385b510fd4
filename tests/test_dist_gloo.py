"""N > 1 host path on CPU: world_size 2, gloo backend (the GPU path uses the same code with NCCL)."""
import os
import subprocess
import sys

import pytest

from madronalib_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_voices():
    for V in (1, 7, 64, 65536):
        for world in (1, 2, 3, 8):
            r = [parallel.shard_range(V, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == V
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
    with pytest.raises(ValueError):
        parallel.shard_range(8, 2, 2)


def test_two_rank_gloo_mix_bus():
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DIST_OK 2" in r.stdout
