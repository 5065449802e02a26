#!/usr/bin/env python
"""benches/fib.rs for real: (letrec ((fib (lambda (n) (if (< n 2) n (+ (fib (- n 1)) (fib (- n 2))))))) (fib N)) evaluated by
the restated Lurk evaluator (lurk_b200/lurk_eval.py::lurk_core_toplevel, all 39 funcs of the reference's toplevel) and proved on the
GPU.  N = 95324 gives 2^20 rows of the `eval` chip (BASELINE configs[1]: "fib at 2^20 trace rows").

    python tools/lurk_fib_prove.py [--n 95324] [--steps 3]          one JSON line
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def fib_expr(z, k):
    B, L, U = z.builtin, z.list, z.u64
    sym = lambda s: z.symbol(["lurk-user", s])
    n, fib = sym("n"), sym("fib")
    return L([B("letrec"), L([L([fib, L([B("lambda"), L([n]),
              L([B("if"), L([B("<"), n, U(2)]), n,
                 L([B("+"), L([fib, L([B("-"), n, U(1)])]), L([fib, L([B("-"), n, U(2)])])])])])])]),
              L([fib, U(k)])])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=95324)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--precompile-only", action="store_true", help="fill the cubin cache for this toplevel (no GPU) and exit")
    args = ap.parse_args()
    import goldens
    from lurk_b200 import _native as N
    from lurk_b200 import jit_warm
    from lurk_b200.zstore import ZStore
    top = goldens.toplevel_for("lurk_core")
    if args.precompile_only:
        os.environ["LSK_JIT"] = "1"
        print(jit_warm.precompile(top, "lurk_eval_demo"))
        return
    import torch
    from lurk_b200.execute import QueryRecord, Shard, execute_by_name
    from lurk_b200.machine import new_machine
    z = ZStore()
    e = fib_expr(z, args.n)
    m = new_machine(top, "lurk_eval_demo")
    pk, vk = m.setup()
    q = QueryRecord.new(top)
    z.inject(q, top)
    t0 = time.perf_counter()
    out = execute_by_name(top, "lurk_eval_demo", [e.tag] + [0] * 7 + list(e.digest), q)
    t_exec = time.perf_counter() - t0
    a, b = 0, 1
    for _ in range(args.n):
        a, b = b, (a + b) % (1 << 64)
    assert int(out[0]) == 0 and sum(int(x) << (8 * i) for i, x in enumerate(out[1:9])) == a, "wrong fib value"
    rows = {f.name: int(N.lib().lsk_record_num_queries(q.h, f.index)) for f in top.funcs}
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
    ctx = N.default_context()
    ctx.profile_enable(True)
    m.prove(pk, shard)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    tall = max(rows.values())
    pad = 1 << (tall - 1).bit_length()
    print(json.dumps({
        "metric": "lurk_fib_prove_trace_rows_per_sec", "workload": "Lurk (fib %d) through the restated evaluator, one shard" % args.n,
        "fib_mod_2_64": a, "rows_tallest_chip": tall, "padded_rows": pad, "rows_all_chips": sum(rows.values()),
        "executor_s": round(t_exec, 2), "ms_per_prove_e2e": sum(times) / len(times), "value": pad / (sum(times) / len(times) / 1e3),
        "unit": "rows/s", "proof_words": int(len(proof.words)), "verified": True,
        "compiled_constraint_programs": int(N.lib().lsk_machine_jit_chips(m.h)),
        "kernel_ms": {k: round(v["ms"], 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}), flush=True)


if __name__ == "__main__":
    main()
