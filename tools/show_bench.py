#!/usr/bin/env python
"""Pretty-print the last JSON line of a bench.py output file."""
import json, signal, sys
signal.signal(signal.SIGPIPE, signal.SIG_DFL)   # `| head` must not raise
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g rows/s  ms/step %.2f | e2e %.4g rows/s (%.2f ms) | launches %d | clocks %s" % (
    d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d.get('clocks')))
if 'cpu_baseline' in d: print("cpu", d['cpu_baseline'])
tot = 0
for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step']):
    print("%-16s %7.3f ms %4d launches %8.1f GB/s" % (k, v['ms_per_step'], v['launches_per_step'], v['algorithmic_GBps']))
    tot += v['ms_per_step']
print("sum kernels %.2f ms" % tot)
