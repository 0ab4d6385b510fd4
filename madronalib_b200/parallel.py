"""Multi-GPU plumbing: one process per GPU (torch.distributed), voices sharded, one collective.

Voices are independent units (reference: Bank rows never exchange data,
source/DSP/MLDSPFunctional.h:328-337), so the data path needs NO collective; the only
exchange is the sum of the per-rank mix buses (reference pattern: Synth::processVector
accumulating voices into the outputs, source/app/MLSynth.h:36-60), a [T][n_out][64] f32
all-reduce -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def shard_range(n_voices: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous voice range [v0, v1) owned by `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return n_voices * rank // world, n_voices * (rank + 1) // world


class MixBusReducer:
    """Double-buffered asynchronous all-reduce of the mix bus.

    ``submit(i, tensor)`` issues the all-reduce of step i's partial mix bus on the process
    group's stream and returns immediately, so it overlaps the next step's kernel;
    ``wait(i)`` (called before buffer i&1 is reused, or at the end) completes it.
    """

    def __init__(self, dist_module=None):
        self.dist = dist_module
        self.pending: List[Optional[object]] = [None, None]

    @property
    def active(self) -> bool:
        return self.dist is not None and self.dist.is_initialized() and self.dist.get_world_size() > 1

    def wait(self, i: int) -> None:
        w = self.pending[i & 1]
        if w is not None:
            w.wait()
            self.pending[i & 1] = None

    def submit(self, i: int, tensor) -> None:
        if not self.active:
            return
        self.wait(i)
        self.pending[i & 1] = self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM, async_op=True)

    def drain(self) -> None:
        self.wait(0)
        self.wait(1)


class PeerMixBus:
    """The mix-bus all-reduce fused into the kernel that finishes the local sum (api.MixBus): every rank
    creates an exchange buffer, the CUDA IPC handles travel through ``dist.all_gather_object`` once, and from
    then on a process call with a `mix` output needs NO collective call -- the kernel writes its 64-sample
    rows into every peer's buffer over NVLink and sums the world's rows in rank order.  ``MixBusReducer``
    (NCCL / gloo) remains the checked fallback and the CPU-test path."""

    def __init__(self, dist_module, api_module, graph, max_floats: int, async_completion: bool = False):
        self.dist, self.graph = dist_module, graph
        rank, world = dist_module.get_rank(), dist_module.get_world_size()
        self.bus = api_module.MixBus(rank, world, max_floats)
        if async_completion:  # wait-for-peers + sum on the bus's own stream: graph.mix_wait(stream) completes it
            self.bus.set_async(True)
        handles = [None] * world
        dist_module.all_gather_object(handles, self.bus.handle())
        self.bus.connect(handles)
        dist_module.barrier()  # nobody writes into a peer before every peer has mapped and zeroed its buffer
        graph.attach_mixbus(self.bus)

    def close(self) -> None:
        self.graph.attach_mixbus(None)
        self.dist.barrier()  # no peer may still be writing into this buffer
        self.bus.close()
