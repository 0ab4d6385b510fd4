"""Worker for tests/test_dist_gloo.py: run under torch.distributed.run with the gloo backend.

Each rank owns a contiguous shard of the voices (parallel.shard_range), computes its shard
(the CPU oracle stands in for the CUDA kernel -- this test is about the host-side sharding and
the collective), all-reduces the mix bus through parallel.MixBusReducer, and rank 0 checks the
result against the unsharded computation.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from madronalib_b200 import parallel, workloads as wl  # noqa: E402
from oracle.bindings import PortOracle  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    for full, T in ((wl.config_a(256), 3), (wl.config_6(96), 24)):  # the headline chain; Aaltoverb (delay memory, feedback)
        run_sharded(full, T, rank, world)
    if rank == 0:
        print("DIST_OK", world)
    dist.barrier()
    dist.destroy_process_group()


def run_sharded(full, T, rank, world):
    V = full.n_voices
    inp_full = full.inputs(T)
    v0, v1 = parallel.shard_range(V, rank, world)
    shard = full.shard(rank, world)
    assert shard.n_voices == v1 - v0
    P = PortOracle()
    out, mix, st = P.run(shard.spec, shard.n_voices, T, np.ascontiguousarray(inp_full[:, :, v0:v1]),
                         shard.state, shard.coef, want_mix=True, mix_mode=1)
    red = parallel.MixBusReducer(dist)
    assert red.active
    bufs = [torch.from_numpy(mix.copy()), torch.from_numpy(mix.copy())]
    for i in range(3):  # double-buffered submit/wait protocol, as bench.py drives it
        red.wait(i)  # the all-reduce that last used this buffer must be complete
        bufs[i & 1].copy_(torch.from_numpy(mix))
        red.submit(i, bufs[i & 1])
    red.drain()
    gathered = [torch.zeros_like(torch.from_numpy(out)) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(out))
    if rank == 0:
        want_out, want_mix, _ = P.run(full.spec, V, T, inp_full, full.state, full.coef,
                                      want_mix=True, mix_mode=1, n_shards=world)
        got_out = np.concatenate([g.numpy() for g in gathered], axis=2)
        assert np.array_equal(got_out.view(np.uint32), want_out.view(np.uint32)), "sharded out"
        for b in bufs[:2]:
            assert np.array_equal(b.numpy().view(np.uint32), want_mix.view(np.uint32)), "mix bus"
        _, ref_order, _ = P.run(full.spec, V, T, inp_full, full.state, full.coef, want_mix=True,
                                mix_mode=0)
        tol = V * np.finfo(np.float32).eps * np.abs(want_out).sum(axis=2).max()
        assert np.abs(bufs[0].numpy() - ref_order).max() <= tol
    dist.barrier()


if __name__ == "__main__":
    main()
