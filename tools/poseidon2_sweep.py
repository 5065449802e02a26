#!/usr/bin/env python
"""Poseidon2 standalone throughput sweep (BASELINE.json configs[4]).

    python tools/poseidon2_sweep.py [--min-log 16] [--max-log 26] [--widths 24,32,40,16] [--out FILE]
    torchrun --nproc-per-node N tools/poseidon2_sweep.py ...      # N GPUs: independent slices, no collective

For every width W (24 / 32 / 40 = hash3 / hash4 / hash5 of content-addressed data,
src/core/zstore.rs:305-333; 16 = the commitment permutation) and n = 2^min .. 2^max permutations:
inputs in[i][j] = splitmix64(seed = 1, i * W + j) mod p generated on the device, hashed by
lsk_poseidon2_hash_dev (inputs and digests resident in HBM), timed with CUDA events on the library's
stream (3 warm-up + 5 timed launches), digests of a sample checked against the CPU oracle.
Reports permutations/s, algorithmic GB/s (4 (W + 8) bytes per permutation) against the measured HBM
peak, and Montgomery-multiplies/s (the pipe that actually bounds the kernel).
One JSON line per (width, n); rank 0 prints the aggregate over all GPUs.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = 2013265921
RP = {16: 13, 24: 21, 32: 30, 40: 38}


def splitmix_inputs(torch, n, w, dev, first_row=0):
    """in[i][j] = splitmix64(1 + (i * w + j)) mod p, computed on the device in int64 arithmetic."""
    idx = torch.arange(first_row * w, (first_row + n) * w, dtype=torch.int64, device=dev)
    z = idx * -7046029254386353131 + (1 + -7046029254386353131)     # seed + (k + 1) * 0x9E3779B97F4A7C15 (wrapping)

    def shr(x, k):   # logical shift right on int64
        return (x >> k) & ((1 << (64 - k)) - 1)
    z = (z ^ shr(z, 30)) * -4658895280553007687                     # 0xBF58476D1CE4E5B9
    z = (z ^ shr(z, 27)) * -7723592293110705685                     # 0x94D049BB133111EB
    z = z ^ shr(z, 31)
    return (shr(z, 1) % P).to(torch.int32).reshape(n, w)             # 63-bit value mod p


def modmuls_per_perm(w):
    return 8 * w * 4 + RP[w] * (4 + w)     # S-boxes (x^7 = 4 products) + internal diagonal


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log", type=int, default=16)
    ap.add_argument("--max-log", type=int, default=26)
    ap.add_argument("--widths", default="32,24,40,16")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from lurk_b200 import _native as N
    from lurk_b200.poseidon import PoseidonChipset
    from oracle import pyoracle
    pyoracle.build()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = N.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    lines = []
    for w in [int(x) for x in args.widths.split(",")]:
        chip = PoseidonChipset(w, ctx)
        for lg in range(args.min_log, args.max_log + 1):
            n_total = 1 << lg
            n = n_total // world                      # this rank's slice
            if n == 0:
                continue
            # generated in slabs of 2^20 rows: the int64 temporaries of a 2^28-row slab would not fit beside the inputs
            x = torch.empty((n, w), dtype=torch.int32, device=dev)
            for r0 in range(0, n, 1 << 20):
                r1 = min(n, r0 + (1 << 20))
                x[r0:r1] = splitmix_inputs(torch, r1 - r0, w, dev, first_row=rank * n + r0)
            out = torch.empty((n, 8), dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            with torch.cuda.stream(stream):
                for _ in range(3):
                    chip.hash_dev(x.data_ptr(), n, out.data_ptr())
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                a.record(stream)
                for _ in range(5):
                    chip.hash_dev(x.data_ptr(), n, out.data_ptr())
                b.record(stream)
            torch.cuda.synchronize()
            ms = torch.tensor([a.elapsed_time(b) / 5], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            ms = float(ms)
            # parity on a sample (first 256 rows of this rank's slice)
            k = min(n, 256)
            got = out[:k].cpu().numpy().view(np.uint32)
            want = pyoracle.poseidon2_hash(w, x[:k].cpu().numpy().view(np.uint32))
            assert np.array_equal(got, want), f"W={w} n=2^{lg}: GPU digests differ from the oracle"
            if rank == 0:
                pps = n_total / (ms / 1e3)
                gbs = pps * 4 * (w + 8) / 1e9
                line = {"metric": "poseidon2_permutations_per_sec", "width": w, "log_n": lg, "n_gpus": world,
                        "ms": ms, "value": pps, "unit": "perm/s", "algorithmic_GBps": gbs,
                        "hbm_frac_of_measured_peak": gbs / (peak * world), "modmul_per_s": pps * modmuls_per_perm(w),
                        "checked_against_oracle": k}
                lines.append(line)
                print(json.dumps(line), flush=True)
            del x, out
    if rank == 0:
        # CPU oracle beside it (one width-32 batch sized for a few seconds)
        t = time.perf_counter()
        xs = np.random.default_rng(1).integers(0, P, size=(1 << 18, 32), dtype=np.uint32)
        pyoracle.poseidon2_hash(32, xs)
        dt = time.perf_counter() - t
        line = {"metric": "poseidon2_permutations_per_sec", "impl": "cpu_oracle", "width": 32, "log_n": 18,
                "value": (1 << 18) / dt, "unit": "perm/s", "cores": os.cpu_count()}
        lines.append(line)
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                for ln in lines:
                    f.write(json.dumps(ln) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
