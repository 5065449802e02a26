#!/usr/bin/env python
"""Column-sharded commit of one wide trace over G GPUs: parity + throughput.

    torchrun --nproc-per-node G tools/column_sharded_commit.py [--log-rows 20] [--width 656] [--steps 5]

Parity: at 2^12 rows the root equals the single-GPU commitment of the same matrix (lsk_coset_lde +
lsk_merkle_root on rank 0, host API).  Throughput: LDE + all-to-all + leaf hashing + subtree of an
(N x width) matrix (default: the Lurk hash4 chip's width rounded up to a multiple of 8 GPUs), timed with
CUDA events, max over ranks; prints one JSON line with the per-stage split and the NVLink bytes moved.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = 2013265921


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-rows", type=int, default=20)
    ap.add_argument("--width", type=int, default=656)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from lurk_b200 import _native as N
    from lurk_b200.parallel import column_partition, column_sharded_commit, import_columns

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = N.Context(local)

    # ---- parity at a size the host API handles quickly
    lg, w = 12, 8 * world
    rng = np.random.default_rng(7)
    full = rng.integers(0, P, size=(1 << lg, w), dtype=np.uint32)
    cols = column_partition(w, world, rank)
    mine = import_columns(ctx, full[:, cols.start:cols.stop], dev)
    root = column_sharded_commit(ctx, mine, lg)
    if rank == 0:
        lde = np.zeros((2 << lg, w), dtype=np.uint32)
        N.check(N.lib().lsk_coset_lde(ctx.h, N.vp(full), ctypes.c_size_t(1 << lg), ctypes.c_size_t(w), 0, N.vp(lde)))
        want = np.zeros(8, dtype=np.uint32)
        ptrs = (ctypes.c_void_p * 1)(lde.ctypes.data)
        hs = (ctypes.c_size_t * 1)(lde.shape[0])
        ws = (ctypes.c_size_t * 1)(w)
        N.check(N.lib().lsk_merkle_root(ctx.h, ptrs, hs, ws, ctypes.c_size_t(1), N.vp(want)))
        assert np.array_equal(root, want), (root, want)
    roots = [None] * world
    dist.all_gather_object(roots, root.tolist())
    assert all(r == roots[0] for r in roots), "ranks disagree on the root"

    # ---- throughput
    L, W = args.log_rows, args.width
    assert W % world == 0
    wl = W // world
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    mine = torch.randint(0, P, (wl, 1 << L), dtype=torch.int32, device=dev, generator=g)
    for _ in range(2):
        column_sharded_commit(ctx, mine, L)
    timers = {}
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        column_sharded_commit(ctx, mine, L, timers=timers)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / args.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        cells = (1 << L) * W
        sent = 8 * cells // world * (world - 1) // world      # bytes each rank sends in the all-to-all
        print(json.dumps({"metric": "column_sharded_commit_cells_per_sec", "n_gpus": world, "log_rows": L, "width": W,
                          "ms_per_commit": float(ms), "value": cells / (float(ms) / 1e3), "unit": "trace cells/s",
                          "all_to_all_bytes_sent_per_gpu": sent,
                          "stage_ms": {k: v / args.steps for k, v in timers.items()},
                          "parity": "root == single-GPU lsk_merkle_root at 2^12 x %d" % w}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
