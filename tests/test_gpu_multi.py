"""Multi-GPU mix bus on real hardware (needs >= 2 GPUs on the box: run under `gpurun --gpus 2`; skipped
otherwise).  The host-side sharding and the NCCL/gloo fallback are covered on CPU by test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_memory_mix_bus_all_reduce(gpu, world):
    if gpu.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29611 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_OK %d" % world in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
