#!/usr/bin/env python
"""ZStore batch interning: where the GPU batch hasher pays and where it does not.

    python tools/zstore_batch_bench.py [--out FILE]

(1) hash_many on n independent hash4 preimages, host hasher vs GPU launch (host buffers in, digests out), n = 2^4..2^18:
    the crossover batch size.  (2) interning data of two shapes with 2^14 conses each -- a flat list (a chain: every level
    one node wide) and 2^10 sublists of 16 (levels 2^10 wide) -- one by one, deferred on the host, deferred with the GPU.
One JSON line per measurement."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lurk_b200 import _native as N
from lurk_b200.zstore import P, ZStore


def best(f, reps=3):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t)
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    ctx = N.default_context()
    lines = []
    rng = np.random.default_rng(7)
    for lg in range(4, 19, 2):
        n = 1 << lg
        pre = rng.integers(0, P, size=(n, 32), dtype=np.uint32)
        out_h, out_g = np.empty((n, 8), np.uint32), np.empty((n, 8), np.uint32)
        import ctypes
        th = best(lambda: N.check(N.lib().lsk_poseidon2_hash_host(32, N.vp(pre), ctypes.c_size_t(n), N.vp(out_h))))
        tg = best(lambda: N.check(N.lib().lsk_poseidon2_hash(ctx.h, 32, N.vp(pre), ctypes.c_size_t(n), N.vp(out_g))))
        assert np.array_equal(out_h, out_g)
        lines.append({"what": "hash4 batch, host buffers in / digests out", "n": n, "host_us": th * 1e6, "gpu_us": tg * 1e6,
                      "host_us_per_hash": th * 1e6 / n, "gpu_us_per_hash": tg * 1e6 / n})
        print(json.dumps(lines[-1]), flush=True)

    def direct(z, o):
        return z.u64(o) if isinstance(o, int) else z.list([direct(z, x) for x in o])
    shapes = {"flat list of 2^14 (a chain: levels 1 wide)": list(range(1 << 14)),
              "2^10 sublists of 16 (levels 2^10 wide)": [[i * 16 + j for j in range(16)] for i in range(1 << 10)]}
    for name, data in shapes.items():
        za, zb, zc = ZStore(), ZStore(), ZStore()
        for z in (za, zb, zc):
            z.nil()
        t1 = best(lambda: direct(ZStore(), data), 1)
        t2 = best(lambda: ZStore().resolve([ZStore().defer_data(data)]), 1)
        zc_root = None

        def gpu():
            z = ZStore()
            return z.resolve([z.defer_data(data)], ctx=ctx, gpu_min_batch=256), z.gpu_batches
        t = time.perf_counter()
        (root,), batches = gpu()
        t3 = time.perf_counter() - t
        assert root == direct(za, data)
        lines.append({"what": "intern " + name, "one_by_one_ms": t1 * 1e3, "deferred_host_ms": t2 * 1e3, "deferred_gpu_ms": t3 * 1e3,
                      "gpu_batches": batches, "note": "Python bookkeeping included in all three"})
        print(json.dumps(lines[-1]), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
