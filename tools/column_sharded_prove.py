#!/usr/bin/env python
"""Column-sharded PROVE of one shard over G GPUs (BASELINE configs[3]): parity + timing.

    torchrun --nproc-per-node G tools/column_sharded_prove.py [--log-rows 20] [--workload fib|cons] [--steps 3]

Every rank first proves the record alone (no device group), then the ranks join one device group
(lurk_b200.parallel.join_device_group) and prove it together.  Parity: the group's proof equals the single-GPU
proof word for word on EVERY rank, for each parity workload (a fib shard, a multi-shard fib record, the wide
hash4 chip) and for the timed workload; rank 0 also runs machine.verify on it.  Timing: ms per prove alone vs
together (CUDA events around the step, max over ranks), the NCCL share and the bytes each rank received.
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-rows", type=int, default=20)
    ap.add_argument("--workload", default="fib", choices=["fib", "cons"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--skip-parity-set", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from bench import Workload
    from lurk_b200 import _native as N
    from lurk_b200.execute import QueryRecord, Shard, ShardingConfig, execute_by_name
    from lurk_b200.machine import new_machine
    from lurk_b200.parallel import join_device_group

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = N.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def setup(workload, n_arg, shard_size=None):
        wl = Workload(workload)
        top, entry = wl.top, wl.entry
        m = new_machine(top, entry, ctx=ctx)
        pk, vk = m.setup()
        q = QueryRecord.new(top)
        execute_by_name(top, entry, [n_arg], q)
        q.finalize()
        return m, pk, vk, Shard(q, 0, ShardingConfig(shard_size or (1 << 22)))

    def timed(fn, steps):
        ms_all = []
        for _ in range(steps):
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            torch.cuda.synchronize()
            ms = torch.tensor([a.elapsed_time(b)], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            ms_all.append(float(ms))
        return sum(ms_all) / len(ms_all)

    # ---- the workloads: (name, workload, argument, shard size)
    cases = []
    if not args.skip_parity_set:
        cases += [("fib 2^16 rows, one shard", "fib", (1 << 16) - 1, None),
                  ("fib 2^17+5 rows in shards of 2^16", "fib", (1 << 17) + 4, 1 << 16),
                  ("cons 2^15 rows of the hash4 chip", "cons", (1 << 15) - 1, None)]
    cases.append(("timed: %s 2^%d rows, one shard" % (args.workload, args.log_rows), args.workload,
                  (1 << args.log_rows) - 1, max(1 << 22, 1 << args.log_rows)))
    alone = {}
    built = {}
    for name, wl, n_arg, ss in cases:
        m, pk, vk, q = setup(wl, n_arg, ss)
        built[name] = (m, pk, vk, q)
        alone[name] = np.array(m.prove(pk, q).words, copy=True)
    tname = cases[-1][0]
    m, pk, vk, q = built[tname]
    q.queries.pin()                    # the record is the step's input: pinned host memory, uploaded by every prove
    for _ in range(2):
        m.prove(pk, q)
    ms_alone = timed(lambda: m.prove(pk, q), args.steps)

    join_device_group(ctx)
    parity = {}
    for name, wl, n_arg, ss in cases:
        m, pk, vk, q = built[name]
        proof = m.prove(pk, q)
        same = bool(np.array_equal(np.asarray(proof.words), alone[name]))
        flags = [None] * world
        dist.all_gather_object(flags, same)
        parity[name] = all(flags)
        if rank == 0:
            m.verify(vk, proof)
    m, pk, vk, q = built[tname]
    for _ in range(2):
        m.prove(pk, q)
    ms_group = timed(lambda: m.prove(pk, q), args.steps)
    ctx.profile_enable(True)
    m.prove(pk, q)
    prof = ctx.profile_read()
    info = ctx.comm_info()
    ctx.profile_enable(False)
    if rank == 0:
        rows = 1 << args.log_rows
        kern = {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        print(json.dumps({"metric": "column_sharded_prove_rows_per_sec", "n_gpus": world, "workload": tname,
                          "ms_per_prove_one_gpu": ms_alone, "ms_per_prove_group": ms_group,
                          "speedup": ms_alone / ms_group, "value": rows / (ms_group / 1e3), "unit": "rows/s",
                          "scaling": "strong", "nccl_ms": prof.get("nccl", {}).get("ms"),
                          "nccl_bytes_received_per_gpu": info["bytes_received"],
                          "rank0_kernel_ms": kern, "parity_vs_single_gpu_proof": parity}), flush=True)
    ok = all(parity.values())
    dist.barrier()
    ctx.comm_destroy()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit("column-sharded proof differs from the single-GPU proof: %r" % parity)


if __name__ == "__main__":
    main()
