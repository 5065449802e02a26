"""Shard-per-GPU proving: one process per GPU, `torch.distributed` for the plumbing.

The reference has no distributed path; its unit of independent work is the
shard (`Shard::shard`, src/lair/execute.rs:186-216).  Rank r proves shards r, r+G, ...: phase 1
commits the main traces, the 8-word roots are all-gathered (the only exchange on the data path:
every rank must feed the same transcript), phase 2 proves each shard against that transcript, and
the shard proofs are gathered to rank 0 and concatenated -- bit-identical to the single-GPU proof.
The collective backend is whatever the process group was created with: "nccl" on GPUs (tensors on
the rank's device), "gloo" in the CPU tests (which drive this module with a stand-in machine).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .machine import MachineProof, assemble_proof


def _device():
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def shards_of_rank(num_shards: int, rank: int, world: int) -> List[int]:
    return list(range(rank, num_shards, world))


def all_gather_roots(local: dict, num_shards: int) -> np.ndarray:
    """local: {shard index: 8-word root} for this rank's shards -> (num_shards, 8) array on every rank."""
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (num_shards + world - 1) // world
    buf = torch.zeros((per, 8), dtype=torch.int64)
    for k, s in enumerate(shards_of_rank(num_shards, rank, world)):
        buf[k] = torch.from_numpy(np.asarray(local[s], dtype=np.int64))
    buf = buf.to(_device())
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    roots = np.zeros((num_shards, 8), dtype=np.uint32)
    for r, o in enumerate(outs):
        o = o.cpu().numpy()
        for k, s in enumerate(shards_of_rank(num_shards, r, world)):
            roots[s] = o[k].astype(np.uint32)
    return roots


def gather_shard_proofs(local: dict, num_shards: int, dst: int = 0) -> Optional[List[np.ndarray]]:
    """local: {shard index: proof words}; returns the ordered list on rank `dst`, None elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = _device()
    mine = shards_of_rank(num_shards, rank, world)
    sizes = torch.zeros(num_shards, dtype=torch.int64)
    for s in mine:
        sizes[s] = len(local[s])
    sizes = sizes.to(dev)
    dist.all_reduce(sizes)                      # every shard is owned by exactly one rank
    sizes = sizes.cpu().numpy()
    per = (num_shards + world - 1) // world
    cap = int(sizes.max()) if num_shards else 0
    buf = torch.zeros((per, cap), dtype=torch.int64)
    for k, s in enumerate(mine):
        buf[k, :len(local[s])] = torch.from_numpy(np.asarray(local[s], dtype=np.int64))
    buf = buf.to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, outs, dst=dst)
    if rank != dst:
        return None
    parts: List[Optional[np.ndarray]] = [None] * num_shards
    for r, o in enumerate(outs):
        o = o.cpu().numpy()
        for k, s in enumerate(shards_of_rank(num_shards, r, world)):
            parts[s] = o[k, :int(sizes[s])].astype(np.uint32)
    return parts


def prove_distributed(machine, shards: Sequence, dst: int = 0) -> Optional[MachineProof]:
    """machine.prove for a record cut into `shards` (Shard::shard), spread over the process group.
    `machine` needs commit_shard(shard) -> root and prove_shard(shard, roots) -> words
    (lurk_b200.machine.StarkMachine).  Returns the MachineProof on rank `dst`."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n = len(shards)
    mine = shards_of_rank(n, rank, world)
    roots = all_gather_roots({s: machine.commit_shard(shards[s]) for s in mine}, n)
    parts = gather_shard_proofs({s: machine.prove_shard(shards[s], roots) for s in mine}, n, dst)
    return assemble_proof(parts) if parts is not None else None
