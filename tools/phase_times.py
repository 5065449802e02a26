import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np
from lurk_b200.execute import QueryRecord, Shard, execute_by_name
from lurk_b200.lair import demo_toplevel
from lurk_b200.machine import new_machine
import lair_programs as lp
which = sys.argv[1]; L = int(sys.argv[2])
if which == 'fib':
    top = demo_toplevel(); entry = 'fib'
else:
    top = lp.hash_toplevel(); entry = which
m = new_machine(top, entry)
pk, vk = m.setup()
q = QueryRecord.new(top)
execute_by_name(top, entry, [(1 << L) - 1], q)
sh = Shard.new(q)
for i in range(3):
    t = time.perf_counter(); p = m.prove(pk, sh); dt = time.perf_counter() - t
    print(f"prove {i}: {dt*1e3:.2f} ms, {len(p.words)} words", file=sys.stderr)
os.environ['LSK_PHASES'] = '1'
# getenv is read per prove call, so setting it here (same process) is enough
import ctypes
ctypes.CDLL(None).setenv(b"LSK_PHASES", b"1", 1)
t = time.perf_counter(); p = m.prove(pk, sh); dt = time.perf_counter() - t
print(f"prove with phase syncs: {dt*1e3:.2f} ms", file=sys.stderr)
ctypes.CDLL(None).unsetenv(b"LSK_PHASES")
ctx = m.ctx
ctx.profile_enable(True)
p = m.prove(pk, sh)
prof = ctx.profile_read()
tot = sum(v['ms'] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
    print(f"{k:18s} {v['ms']:8.3f} ms {v['launches']:5d} launches {v['bytes']/max(v['ms'],1e-9)/1e6:9.1f} GB/s", file=sys.stderr)
print("kernel total", tot, file=sys.stderr)
