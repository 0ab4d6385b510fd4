"""Worker for tests/test_gpu_multi.py: run under torch.distributed.run, one rank per GPU (NCCL).

Every rank owns a contiguous voice shard of ONE bank (parallel.shard_range), runs it on its GPU with the
peer-memory mix bus attached (parallel.PeerMixBus: the all-reduce happens inside mix_reduce_kernel over
NVLink), and checks: its own rows against the CPU checker; the all-reduced mix bus against the checker's
sharded tree (shards summed in rank order) -- bit-exact, identical on every rank; and against a NCCL
all-reduce of the local mix buses (same bits at world = 2, within rounding of a different order above).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from madronalib_b200 import api, parallel, workloads as wl  # noqa: E402
from oracle.bindings import PortOracle  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    api.init(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    P = PortOracle()
    for full, T, n_calls, use_async in ((wl.config_a(2048 + 77), 5, 3, False), (wl.config_a(40000), 8, 2, False),
                                        (wl.config_4(96), 6, 2, False), (wl.config_a(2048 + 77), 5, 5, True),
                                        (wl.config_a(40000), 8, 4, True)):
        V = full.n_voices
        v0, v1 = parallel.shard_range(V, rank, world)
        shard = full.shard(rank, world)
        inp_full = full.inputs(T * n_calls)
        want_out, want_mix, _ = P.run(full.spec, V, T * n_calls, inp_full, full.state, full.coef, want_mix=True,
                                      mix_mode=1, n_shards=world, nthreads=8)
        g = api.VoiceGraph(shard.spec, shard.n_voices)
        g.set_coefs(shard.coef)
        g.set_state(shard.state)
        n_out = full.spec.n_out
        bus = parallel.PeerMixBus(dist, api, g, T * n_out * 64, async_completion=use_async)
        if use_async and V > 30000:
            g.set_mix_async(True)  # the local reduction on the graph's own stream as well
        d_in = torch.from_numpy(np.ascontiguousarray(inp_full[:, :, v0:v1])).to(dev)
        d_out = torch.empty((T * n_calls, n_out, v1 - v0, 64), dtype=torch.float32, device=dev)
        d_mix = torch.zeros((T * n_calls, n_out, 64), dtype=torch.float32, device=dev)
        sh = torch.cuda.current_stream().cuda_stream
        for c in range(n_calls):  # back-to-back calls: both parities of the exchange buffer, no host sync between
            g.process_device(d_in[c * T:(c + 1) * T], d_out[c * T:(c + 1) * T], d_mix[c * T:(c + 1) * T], T, sh)
        g.mix_wait(sh)  # async completion: the last call's sum runs on the bus's stream (no-op otherwise)
        torch.cuda.synchronize()
        got_out, got_mix = d_out.cpu().numpy(), d_mix.cpu().numpy()
        assert np.array_equal(got_out.view(np.uint32), want_out[:, :, v0:v1].view(np.uint32)), "shard rows"
        assert np.array_equal(got_mix.view(np.uint32), want_mix.view(np.uint32)), \
            "peer mix bus != checker (rank %d, %s)" % (rank, full.name)
        # the NCCL route gives the same sum
        bus.close()
        g2 = api.VoiceGraph(shard.spec, shard.n_voices)
        g2.set_coefs(shard.coef)
        g2.set_state(shard.state)
        d_mix2 = torch.zeros_like(d_mix)
        g2.process_device(d_in, d_out, d_mix2, T * n_calls, sh)
        dist.all_reduce(d_mix2)
        torch.cuda.synchronize()
        if world == 2:
            assert torch.equal(d_mix2, d_mix), "NCCL all-reduce differs from the peer bus at world 2"
        else:
            tol = V * np.finfo(np.float32).eps * np.abs(want_out).sum(axis=2).max()
            assert float((d_mix2 - d_mix).abs().max()) <= tol
        g.close()
        g2.close()
        dist.barrier()
    if rank == 0:
        print("MULTI_OK", world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
