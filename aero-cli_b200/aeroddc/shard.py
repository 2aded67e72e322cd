"""VFO sharding across the GPUs of one node: VFO v runs on rank v mod N, the raw
IQ block is broadcast from rank 0 (NCCL over NVLink on GPUs), every rank returns its own payloads.
There is no reduction and no other collective on the data path.

Pure plumbing over torch.distributed; the per-rank compute object is injected (`aeroddc.Bank` on a
GPU; the CPU tests inject a stand-in) so the partition/broadcast/collect logic runs under gloo too.
"""
import numpy as np


def shard_vfos(n_vfos, world, rank):
    """Indices of the VFOs rank `rank` owns: v mod world == rank."""
    return [v for v in range(n_vfos) if v % world == rank]


def owner_of(v, world):
    return v % world


class ShardedBank:
    """One rank's view of a VFO bank sharded over `world` ranks.

    make_bank(local_vfo_descs) -> object with process(block ndarray) and output(i) -> (bytes, rate).
    """

    def __init__(self, vfo_descs, make_bank, dist=None, world=1, rank=0):
        self.world, self.rank, self.dist = world, rank, dist
        self.n = len(vfo_descs)
        self.mine = shard_vfos(self.n, world, rank)
        self.bank = make_bank([vfo_descs[v] for v in self.mine])

    def broadcast_block(self, block_tensor):
        """Rank 0's tensor content is sent to every rank (in place)."""
        if self.world > 1:
            self.dist.broadcast(block_tensor, src=0)
        return block_tensor

    def process(self, block_tensor):
        """Broadcast + local compute. Returns {global vfo index: (payload bytes, rate)} of this rank."""
        self.broadcast_block(block_tensor)
        self.bank.process(block_tensor.numpy() if hasattr(block_tensor, "numpy") else np.asarray(block_tensor))
        return {v: self.bank.output(i) for i, v in enumerate(self.mine)}

    def gather_outputs(self, local):
        """Collect every rank's payloads on rank 0 (host-side; the publisher of a multi-GPU box)."""
        if self.world == 1:
            return dict(local)
        parts = [None] * self.world
        self.dist.all_gather_object(parts, local)
        merged = {}
        for p in parts:
            merged.update(p)
        return merged
