import sys
sys.path.insert(0,'/root/repo')
from lurk_b200.execute import QueryRecord, Shard, ShardingConfig, execute_by_name
from lurk_b200.lair import demo_toplevel
from lurk_b200.machine import new_machine
top = demo_toplevel(); m = new_machine(top, "fib"); pk,_ = m.setup()
q = QueryRecord.new(top); execute_by_name(top, "fib", [(1<<20)-1], q)
for _ in range(2): m.prove(pk, Shard(q,0,ShardingConfig(1<<20)))
