"""CPU suite, part 3: the N > 1 path (VFO sharding v mod N, raw-block broadcast from rank 0, payload
collection) under torch.distributed/gloo with world_size 2. The per-rank compute object is an
oracle-backed stand-in (this is a test), so the result must equal the single-process chain."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aeroddc.shard import ShardedBank, owner_of, shard_vfos
from oracle_bind import Oracle, synth_anchor

FS, BLK = 288000, 57600
DESCS = [dict(mixer=-34567.0, D=1, L=6, gain=0.5, bw=0), dict(mixer=20000.0, D=0, L=6, gain=0.25, bw=3000), dict(mixer=51234.0, D=4, L=0, gain=0.4, bw=0),
         dict(mixer=-99999.0, D=2, L=0, gain=0.3, bw=0), dict(mixer=7.0, D=3, L=5, gain=0.2, bw=0)]


class OracleBank:
    def __init__(self, descs):
        self.o = [Oracle(FS, BLK, d["D"], d["L"], d["mixer"], d["gain"], d["bw"]) for d in descs]
        self.last = []

    def process(self, block):
        self.last = [o.process(block) for o in self.o]

    def output(self, i):
        return self.last[i], self.o[i].out_rate


def test_partition_is_a_disjoint_cover():
    for n in (1, 5, 128, 1024, 1027):
        for world in (1, 2, 4, 8):
            parts = [shard_vfos(n, world, r) for r in range(world)]
            flat = sorted(v for p in parts for v in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            assert all(owner_of(v, world) == r for r, p in enumerate(parts) for v in p)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sb = ShardedBank(DESCS, OracleBank, dist=dist, world=world, rank=rank)
    results = []
    for b in range(3):
        # only rank 0 owns the data; the others start from garbage and must receive the broadcast
        x = torch.from_numpy(synth_anchor(b * BLK, BLK)) if rank == 0 else torch.full((2 * BLK,), 123.0)
        local = sb.process(x)
        assert sorted(local) == shard_vfos(len(DESCS), world, rank)
        merged = sb.gather_outputs(local)
        results.append({v: (p.hex()[:64], len(p), r) for v, (p, r) in merged.items()})
    if rank == 0:
        q.put(results)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_matches_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = ShardedBank(DESCS, OracleBank)
    for b in range(3):
        want = single.process(torch.from_numpy(synth_anchor(b * BLK, BLK)))
        assert sorted(got[b]) == list(range(len(DESCS)))
        for v, (p, r) in want.items():
            assert got[b][v] == (p.hex()[:64], len(p), r)
