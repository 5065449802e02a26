#!/usr/bin/env python
"""The reference's benches (benches/fib.rs, lcs.rs, sum.rs) through this repo's path: the program is read from its source
text, executed by `lurk_main` of the packaged Lurk toplevel (lurk_b200/lurk_eval.py), and the record is proved on the GPU
and verified.  One JSON line per program.

    python tools/lurk_prove.py --program fib --arg 20 [--steps 3]
    python tools/lurk_prove.py --program lcs --arg 10        (leading characters of the reference's two strings)
    python tools/lurk_prove.py --program sum --arg 2000
    python tools/lurk_prove.py --exec-only ...                 rows per chip, no GPU

The executor recurses as deep as the Lurk program does (100000 frames for the default fib / sum arguments): run under
`ulimit -s unlimited`, as the measured lines in profiles/r1_lurk_bench_programs.jsonl were.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--program", default="fib", choices=["fib", "lcs", "sum"])
    ap.add_argument("--arg", type=int, default=20)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--exec-only", action="store_true")
    args = ap.parse_args()
    from lurk_b200 import _native as N
    from lurk_b200.bench_programs import setup
    from lurk_b200.execute import Shard, execute_by_name
    from lurk_b200.lurk_eval import build_lurk_toplevel
    top = build_lurk_toplevel()
    inp, q, z = setup(args.program, args.arg, top)
    t0 = time.perf_counter()
    out = [int(v) for v in execute_by_name(top, "lurk_main", inp, q)]
    t_exec = time.perf_counter() - t0
    rows = {f.name: int(N.lib().lsk_record_num_queries(q.h, f.index)) for f in top.funcs}
    tall = max(rows.values())
    pad = 1 << max(tall - 1, 0).bit_length()
    line = {"program": args.program, "arg": args.arg, "result_tag": out[0], "result_digest": out[8:],
            "rows_tallest_chip": tall, "tallest_chip": max(rows, key=rows.get), "padded_rows": pad,
            "rows_all_chips": sum(rows.values()), "executor_s": round(t_exec, 3)}
    if args.exec_only:
        print(json.dumps(line), flush=True)
        return
    import torch
    from lurk_b200.machine import new_machine
    m = new_machine(top, "lurk_main")
    pk, vk = m.setup()
    q.finalize()
    q.pin()
    shard = Shard.new(q)
    proof = m.prove(pk, shard)
    m.verify(vk, proof)
    m.prove(pk, shard)
    times = []
    for _ in range(args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.prove(pk, shard)
        torch.cuda.synchronize()
        times.append(1e3 * (time.perf_counter() - t0))
    ms = sum(times) / len(times)
    line.update({"metric": "lurk_%s_prove_trace_rows_per_sec" % args.program, "ms_per_prove_e2e": round(ms, 3),
                 "value": pad / (ms / 1e3), "unit": "rows/s", "proof_words": int(len(proof.words)), "verified": True,
                 "compiled_constraint_programs": int(N.lib().lsk_machine_jit_chips(m.h))})
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
