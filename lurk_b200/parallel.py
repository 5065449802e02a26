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


def all_gather_roots(local: dict, num_shards: int, shards_of=None) -> np.ndarray:
    """local: {shard index: 8-word root} for this rank's shards -> (num_shards, 8) array on every rank.
    shards_of(r): the shards rank r contributes (default: r, r + G, ...; device groups: the group leader's)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if shards_of is None:
        def shards_of(r):
            return shards_of_rank(num_shards, r, world)
    per = max(1, max(len(shards_of(r)) for r in range(world)))
    buf = torch.zeros((per, 8), dtype=torch.int64)
    for k, s in enumerate(shards_of(rank)):
        buf[k] = torch.from_numpy(np.asarray(local[s], dtype=np.int64))
    buf = buf.to(_device())
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    roots = np.zeros((num_shards, 8), dtype=np.uint32)
    for r, o in enumerate(outs):
        o = o.cpu().numpy()
        for k, s in enumerate(shards_of(r)):
            roots[s] = o[k].astype(np.uint32)
    return roots


def gather_shard_proofs(local: dict, num_shards: int, dst: int = 0, shards_of=None) -> Optional[List[np.ndarray]]:
    """local: {shard index: proof words}; returns the ordered list on rank `dst`, None elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = _device()
    if shards_of is None:
        def shards_of(r):
            return shards_of_rank(num_shards, r, world)
    mine = shards_of(rank)
    # one exchange: every rank sends `per` slots of [length, words...] (32-bit patterns; gloo has no uint32, so the
    # words travel as int32), padded to the longest proof -- the lengths ride in the same message
    per = max(1, max(len(shards_of(r)) for r in range(world)))
    mylen = torch.tensor([max([len(local[s]) for s in mine], default=0)], dtype=torch.int64, device=dev)
    dist.all_reduce(mylen, op=dist.ReduceOp.MAX)
    cap = int(mylen.item())
    host = np.zeros((per, cap + 1), dtype=np.int32)
    for k, s in enumerate(mine):
        w = np.ascontiguousarray(local[s], dtype=np.uint32)
        host[k, 0] = len(w)
        host[k, 1:1 + len(w)] = w.view(np.int32)
    buf = torch.from_numpy(host).to(dev)
    if rank == dst:
        out = torch.empty((world, per, cap + 1), dtype=torch.int32, device=dev)
        dist.gather(buf, list(out.unbind(0)), dst=dst)
    else:
        dist.gather(buf, None, dst=dst)
        return None
    got = out.cpu().numpy()
    parts: List[Optional[np.ndarray]] = [None] * num_shards
    for r in range(world):
        for k, s in enumerate(shards_of(r)):
            n = int(got[r, k, 0])
            parts[s] = got[r, k, 1:1 + n].view(np.uint32).copy()
    return parts


def prove_distributed(machine, shards: Sequence, dst: int = 0, gpus_per_group: int = 1) -> Optional[MachineProof]:
    """machine.prove for a record cut into `shards` (Shard::shard), spread over the process group.
    `machine` needs commit_shard(shard) -> root and prove_shard(shard, roots) -> words
    (lurk_b200.machine.StarkMachine).  Returns the MachineProof on rank `dst`.
    gpus_per_group > 1: consecutive ranks form device groups (join_device_group) that prove their shards
    together -- every rank of group j makes the same calls for shards j, j + groups, ..., the leader reports."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n = len(shards)
    assert world % gpus_per_group == 0
    shards_of = group_shards(n, world // gpus_per_group, gpus_per_group)
    mine = list(range(rank // gpus_per_group, n, world // gpus_per_group))
    roots = all_gather_roots({s: machine.commit_shard(shards[s]) for s in mine}, n, shards_of)
    parts = gather_shard_proofs({s: machine.prove_shard(shards[s], roots) for s in mine}, n, dst, shards_of)
    return assemble_proof(parts) if parts is not None else None


# ------------------------------------------------------------------------------------------------
# Column-sharded prove: the ranks of the process group prove every shard TOGETHER
# ------------------------------------------------------------------------------------------------
def join_device_group(ctx, gpus_per_group: Optional[int] = None):
    """Make the ranks of the torch.distributed process group device groups of the library (consecutive ranks,
    `gpus_per_group` each; default: one group of all ranks): afterwards `machine.prove(record)` /
    `commit_shard` / `prove_shard` called by every rank of a group with the same arguments splits the
    throughput-bound stages of each shard over the group's GPUs (LDE by column block, Merkle / quotient / reduced
    openings by row range) and returns the same proof on every rank of the group -- word for word the single-GPU
    proof.  torch.distributed only carries the 128-byte NCCL ids; the collectives on the data path are issued by
    the library on its own stream.  Returns (group index, number of groups, rank within the group)."""
    from . import _native as N
    world, rank = dist.get_world_size(), dist.get_rank()
    g = gpus_per_group or world
    assert world % g == 0 and g & (g - 1) == 0, "group size must be a power of two dividing the number of ranks"
    ids = [None] * world
    dist.all_gather_object(ids, N.comm_unique_id() if rank % g == 0 else None)
    ctx.comm_init(rank % g, g, ids[rank - rank % g])
    return rank // g, world // g, rank % g


def group_shards(num_shards: int, ngroups: int, gpus_per_group: int):
    """shards_of(rank) for device groups: group j proves shards j, j + ngroups, ...; its leader reports them."""
    def shards_of(r):
        return list(range(r // gpus_per_group, num_shards, ngroups)) if r % gpus_per_group == 0 else []
    return shards_of


# ------------------------------------------------------------------------------------------------
# Column-sharded commit of one matrix (the second multi-GPU mode of the north star)
# ------------------------------------------------------------------------------------------------
def column_partition(width: int, world: int, rank: int) -> range:
    """Columns of rank `rank`: equal slices (the all-to-all below needs them equal; pad the matrix with
    zero columns on the caller's side if world does not divide width -- zero columns change the root)."""
    assert width % world == 0, "column_sharded_commit needs world | width"
    per = width // world
    return range(rank * per, (rank + 1) * per)


def column_sharded_commit(ctx, cols_dev: "torch.Tensor", log_n: int, shift_kind: int = 0, timers: Optional[dict] = None):
    """Commit the coset LDE of an (N x w) matrix whose COLUMNS are spread over the ranks.

    cols_dev: this rank's (w / G, N) int32 CUDA tensor in the library's internal layout (column-major, rows
    bit-reversed, Montgomery form; `import_columns`).  NTT/LDE is column-local; Merkle leaves need whole rows,
    so one all-to-all turns the column partition into a row partition (each rank keeps rows
    [d * 2N/G, (d+1) * 2N/G) of ALL columns), each rank hashes its rows and builds its subtree, the G subtree
    roots are all-gathered and the top log2(G) levels are joined on every rank.  Returns the root (8
    canonical words), identical on every rank and identical to the single-GPU commitment.
    """
    import ctypes
    from . import _native as N
    world, rank = dist.get_world_size(), dist.get_rank()
    assert world & (world - 1) == 0, "the number of GPUs must be a power of two"
    wl, n = cols_dev.shape
    assert n == 1 << log_n and cols_dev.dtype == torch.int32 and cols_dev.is_contiguous()
    dev = cols_dev.device
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        lde = torch.empty((wl, 2 * n), dtype=torch.int32, device=dev)
        N.check(N.lib().lsk_lde_columns_dev(ctx.h, ctypes.c_void_p(cols_dev.data_ptr()), ctypes.c_size_t(n),
                                            ctypes.c_void_p(lde.data_ptr()), ctypes.c_size_t(2 * n), log_n, wl,
                                            shift_kind))
        ev[1].record(stream)
        rows = 2 * n // world
        # [wl][G][rows] -> [G][wl][rows]: block d goes to rank d; received blocks are ordered by source rank,
        # i.e. by global column index, so the result is the column-major (w x rows) slab of this rank's rows
        send = lde.view(wl, world, rows).permute(1, 0, 2).contiguous()
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        ev[2].record(stream)
        root = np.zeros(8, dtype=np.uint32)
        N.check(N.lib().lsk_merkle_subtree_root_dev(ctx.h, ctypes.c_void_p(recv.data_ptr()), ctypes.c_size_t(rows),
                                                    (rows - 1).bit_length(), wl * world, N.vp(root)))
        ev[3].record(stream)
        mine = torch.from_numpy(root.astype(np.int64)).to(dev)
        outs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(outs, mine)
        roots = np.stack([o.cpu().numpy().astype(np.uint32) for o in outs])
        ev[4].record(stream)
    out = np.zeros(8, dtype=np.uint32)
    N.check(N.lib().lsk_merkle_join_roots(N.vp(N.u32(roots)), world, N.vp(out)))
    if timers is not None:
        torch.cuda.synchronize()
        for name, a, b in (("lde", 0, 1), ("all_to_all", 1, 2), ("leaves+subtree", 2, 3), ("all_gather_roots", 3, 4)):
            timers[name] = timers.get(name, 0.0) + ev[a].elapsed_time(ev[b])
    return out


def import_columns(ctx, rows_host: np.ndarray, device) -> "torch.Tensor":
    """Canonical row-major (N x w) host matrix -> (w, N) int32 CUDA tensor in the internal layout."""
    import ctypes
    from . import _native as N
    rows_host = N.u32(rows_host)
    h, w = rows_host.shape
    out = torch.empty((w, h), dtype=torch.int32, device=device)
    N.check(N.lib().lsk_import_columns_dev(ctx.h, N.vp(rows_host), ctypes.c_size_t(h), ctypes.c_size_t(w),
                                           ctypes.c_void_p(out.data_ptr())))
    return out
