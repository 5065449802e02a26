"""The N>1 host logic (shard ownership, root all-gather, proof gather + assembly) on CPU with the
gloo backend, world_size 2 and 3, driven by a stand-in machine (the real prover needs a GPU; the
GPU test test_multi_shard_proof_is_bit_exact checks the same two-phase API against lsk_prove)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from lurk_b200.machine import PROOF_MAGIC
from lurk_b200.parallel import prove_distributed, shards_of_rank


class FakeShard:
    def __init__(self, index):
        self.index = index


class FakeMachine:
    """commit: root = f(shard); prove: words depend on the shard and on ALL roots (like the transcript)."""

    def commit_shard(self, shard):
        return (np.arange(8, dtype=np.uint32) * 7 + shard.index * 1000 + 1).astype(np.uint32)

    def prove_shard(self, shard, roots):
        assert roots.shape[1] == 8
        n = 5 + 3 * shard.index          # ragged lengths
        mix = int(roots.astype(np.uint64).sum() % 2013265921)
        return (np.arange(n, dtype=np.uint64) * 31 + shard.index + mix).astype(np.uint32)


def expected(num_shards):
    m = FakeMachine()
    shards = [FakeShard(i) for i in range(num_shards)]
    roots = np.stack([m.commit_shard(s) for s in shards])
    parts = [m.prove_shard(s, roots) for s in shards]
    return np.concatenate([np.asarray([PROOF_MAGIC, num_shards], dtype=np.uint32)] + parts)


def _worker(rank, world, port, num_shards, q, gpus_per_group=1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        proof = prove_distributed(FakeMachine(), [FakeShard(i) for i in range(num_shards)],
                                  gpus_per_group=gpus_per_group)
        if rank == 0:
            q.put(proof.words)
        else:
            assert proof is None
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


# (world, shards, ranks per device group): the last two rows are the column-sharded layouts -- one group of 2
# proving every shard together, and 2 groups of 2 taking alternate shards (each rank of a group makes the same calls)
@pytest.mark.parametrize("world,num_shards,group", [(2, 2, 1), (2, 5, 1), (3, 4, 1), (2, 1, 1), (2, 3, 2), (4, 5, 2)])
def test_prove_distributed_matches_sequential(world, num_shards, group):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_shards, q, group)) for r in range(world)]
    for p in procs:
        p.start()
    words = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(words, expected(num_shards))


def test_shard_ownership_is_a_partition():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 4, 9):
            owned = sorted(s for r in range(world) for s in shards_of_rank(n, r, world))
            assert owned == list(range(n))


def test_device_group_partition_is_a_contiguous_cover():
    # lsk_comm_partition: the share of columns / rows / Merkle nodes each rank of a column-sharded prove computes
    from lurk_b200 import _native as N
    for n in (0, 1, 4, 18, 655, 1 << 20, (1 << 22) + 5):
        for size in (1, 2, 4, 8):
            parts = [N.comm_partition(n, r, size) for r in range(size)]
            assert parts[0].start == 0 and parts[-1].stop == n
            for a, b in zip(parts, parts[1:]):
                assert a.stop == b.start
            if n % size == 0:
                assert len({len(p) for p in parts}) == 1      # equal shares: what the in-place all-gathers need
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(N.LurkStarkError):
        N.comm_partition(8, 2, 2)
