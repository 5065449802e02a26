#!/usr/bin/env python
"""One or two proves of a workload, for running under ncu / compute-sanitizer:
    python tools/prove_once.py [fib|cons|u64|lurk-fib|lurk-lcs|lurk-sum] [log_rows] [repeats]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lurk_b200.execute import QueryRecord, Shard, ShardingConfig, execute_by_name
from lurk_b200.machine import new_machine

workload = sys.argv[1] if len(sys.argv) > 1 else "fib"
log_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 20
repeats = int(sys.argv[3]) if len(sys.argv) > 3 else 2
if workload == "u64":
    from lurk_b200.programs import u64_toplevel
    top, entry, arg = u64_toplevel(), "factu", [20] + [0] * 7
    q = QueryRecord.new(top)
    execute_by_name(top, entry, arg, q)
else:
    from bench import Workload
    wl = Workload(workload)
    top, entry = wl.top, wl.entry
    q, _ = wl.record(wl.arg_for(log_rows))
m = new_machine(top, entry)
pk, vk = m.setup()
for _ in range(repeats):
    proof = m.prove(pk, Shard(q, 0, ShardingConfig(1 << max(log_rows, 10))))
m.verify(vk, proof)
print("ok", workload, log_rows, len(proof.words))
