"""GPU parity: witness-trace generation through the C ABI vs the reference's golden matrices and
vs the oracle on larger / multi-shard inputs."""
import numpy as np
import pytest

import goldens
from lurk_b200.execute import (BytesChip, MemChip, QueryRecord, Shard, ShardingConfig, execute_by_name)
from lurk_b200.lair import FuncChip, demo_toplevel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", goldens.CASES, ids=[c[0] + ":" + c[1] for c in goldens.CASES])
def test_golden_func_trace(case):
    tname, fname, arg_list, golden = case
    top = goldens.toplevel_for(tname)
    q = QueryRecord.new(top)
    for args in arg_list:
        execute_by_name(top, fname, args, q)
    trace = FuncChip.from_name(fname, top).generate_trace(Shard.new(q))
    assert trace.shape == golden.shape
    assert np.array_equal(trace, golden)


def test_golden_mem_trace():
    top = goldens.toplevel_for("memory")
    q = QueryRecord.new(top)
    execute_by_name(top, "test", [], q)
    assert np.array_equal(MemChip(3).generate_trace(Shard.new(q)), goldens.MEMORY_MEM3)
    t = MemChip(8).generate_trace(Shard.new(q))
    assert t.shape == (4, 12) and not t.any()


def _oracle_rec(oracle, top, fname, args):
    ot = oracle.OracleToplevel(top.serialize())
    rec = oracle.OracleRecord(ot)
    f = top.func_by_name(fname)
    rec.execute(f.index, list(args), f.output_size)
    return rec


def test_fib_2pow16_matches_oracle(oracle):
    top = demo_toplevel()
    n = (1 << 16) - 1
    q = QueryRecord.new(top)
    execute_by_name(top, "fib", [n], q)
    trace = FuncChip.from_name("fib", top).generate_trace(Shard.new(q))
    want = _oracle_rec(oracle, top, "fib", [n]).func_trace(1)
    assert trace.shape == (1 << 16, 18)
    assert np.array_equal(trace, want)


def test_sharded_traces_match_oracle(oracle):
    # reference tests multi-shard with tiny shard sizes (src/air/debug.rs:119-158); ragged last shard
    top = goldens.toplevel_for("ackermann")
    q = QueryRecord.new(top)
    execute_by_name(top, "ackermann", [2, 30], q)
    rec = _oracle_rec(oracle, top, "ackermann", [2, 30])
    cfg = ShardingConfig(64)
    shards = Shard.new(q).shard(cfg)
    assert len(shards) == rec.num_shards(64) and len(shards) > 2
    chip = FuncChip.from_name("ackermann", top)
    for s in shards:
        assert np.array_equal(chip.generate_trace(s), rec.func_trace(0, shard=s.index, shard_size=64))


def test_bytes_traces_match_oracle(oracle):
    top = goldens.toplevel_for("range")
    q = QueryRecord.new(top)
    execute_by_name(top, "range_test", [100, 12, 64], q)
    rec = _oracle_rec(oracle, top, "range_test", [100, 12, 64])
    chip = BytesChip()
    assert np.array_equal(chip.generate_trace(Shard.new(q)), rec.bytes_trace(0))
    assert not chip.generate_trace(Shard(q, 1)).any()
    assert np.array_equal(chip.generate_preprocessed_trace(), oracle.bytes_preprocessed())


def test_hash_chip_trace_matches_oracle(oracle):
    # a func wrapping the hash4 extern chip: width 1 + 32 + 8 + (2 + 611) + 1 = 655 (eval_direct.rs:2053-2055)
    from lurk_b200.lair import ChipInfo, Toplevel, extern_call, func, ret
    chips = [ChipInfo("hasher4", 32, 8, 611, 0, kind=2)]
    f = func("hash4", [("preimg", 32)], 8, [extern_call([("img", 8)], "hasher4", ["preimg"]), ret("img")])
    top = Toplevel([f], chips)
    chip = FuncChip.from_name("hash4", top)
    assert chip.width() == 655
    q = QueryRecord.new(top)
    rng = np.random.default_rng(3)
    ot = oracle.OracleToplevel(top.serialize())
    rec = oracle.OracleRecord(ot)
    for _ in range(5):
        pre = rng.integers(0, 2013265921, size=32, dtype=np.uint32)
        a = execute_by_name(top, "hash4", pre, q)
        b = rec.execute(0, pre, 8)
        assert np.array_equal(a, b)
    assert np.array_equal(chip.generate_trace(Shard.new(q)), rec.func_trace(0))


def test_partial_func_traces_match_oracle(oracle):
    # src/lair/trace.rs:110-121,235-254: depth bytes, require triples of the range checks, per-call
    # callee depth + LessThanWitness + byte-lookup require; Bytes chip gains less-than records
    top = goldens.toplevel_for("partial")
    q = QueryRecord.new(top)
    execute_by_name(top, "pmain", [600, 9], q)
    rec = _oracle_rec(oracle, top, "pmain", [600, 9])
    for name in ("pmain", "pfib", "twice", "pdouble", "phalf"):
        chip = FuncChip.from_name(name, top)
        assert np.array_equal(chip.generate_trace(Shard.new(q)), rec.func_trace(chip.func.index)), name
    assert np.array_equal(BytesChip().generate_trace(Shard.new(q)), rec.bytes_trace(0))


def test_u64_chip_traces_match_oracle(oracle):
    # src/core/u64.rs:130-168 (populate_witness of Sum64 / Diff64 / Compare64) inside func rows
    top = goldens.toplevel_for("u64")
    for fname, args in [("fibu", [300]), ("maxu", [9, 0, 0, 0, 0, 0, 0, 1] + [9, 0, 0, 0, 0, 0, 1, 0]),
                        ("sub", [0] * 8 + [1] + [0] * 7), ("less_than", [5] * 8 + [5] * 8),
                        ("factu", [40] + [0] * 7), ("gcdu", [0, 0, 0, 1, 2, 3, 4, 5] + [0, 7, 0, 0, 9, 0, 0, 0]),
                        ("divrem", [255] * 8 + [1, 1] + [0] * 6), ("is_zero", [0] * 8), ("mul", [255] * 16),
                        ("big_num_less_than", [1, 2, 3, 2013265920, 5, 6, 7, 8] + [1, 2, 3, 0x77FFFFFF, 5, 6, 7, 8]),
                        ("big_num_less_than", [7] * 16)]:
        q = QueryRecord.new(top)
        execute_by_name(top, fname, args, q)
        rec = _oracle_rec(oracle, top, fname, args)
        chip = FuncChip.from_name(fname, top)
        assert np.array_equal(chip.generate_trace(Shard.new(q)), rec.func_trace(chip.func.index)), fname
        assert np.array_equal(BytesChip().generate_trace(Shard.new(q)), rec.bytes_trace(0)), fname


def test_fib_2pow20_trace_matches_oracle(oracle):
    # BASELINE configs[1] at full size: 2^20 rows x 18 columns, GPU == oracle cell for cell
    top = demo_toplevel()
    n = (1 << 20) - 1
    q = QueryRecord.new(top)
    execute_by_name(top, "fib", [n], q)
    trace = FuncChip.from_name("fib", top).generate_trace(Shard.new(q))
    want = _oracle_rec(oracle, top, "fib", [n]).func_trace(1)
    assert trace.shape == (1 << 20, 18)
    assert np.array_equal(trace, want)


def _gpu_record_satisfies_air(oracle, top, q, entry, ss):
    """debug_chip_constraints_and_queries_with_sharding (src/air/debug.rs:119-158) on the GPU's traces, judged by the
    oracle's independent restatement of every chip's Air::eval (oracle/air_check.h)."""
    ot = oracle.OracleToplevel(top.serialize())
    tq = oracle.TraceQueries(ot)
    pv = list(q.expect_public_values())
    cfg = ShardingConfig(ss)
    prep = BytesChip().generate_preprocessed_trace()
    for s in Shard(q, 0, cfg).shard(cfg):
        if s.index == 0:
            tq.check(oracle.AIR_ENTRY, entry.index, np.asarray([pv], dtype=np.uint32), public_values=pv)
        for f in top.funcs:
            if len(s.get_func_range(f.index)):
                tq.check(oracle.AIR_FUNC, f.index, FuncChip(f, top).generate_trace(s))
        if s.index == 0:
            for i, ln in enumerate([2, 3, 4, 5, 6, 8]):
                tq.check(oracle.AIR_MEM, i, MemChip(ln).generate_trace(s))
        tq.check(oracle.AIR_BYTES, 0, BytesChip().generate_trace(s), prep=prep)
    tq.verify()
    return tq


def test_gpu_traces_of_a_lurk_program_satisfy_the_independent_air(oracle):
    # benches/fib.rs through lurk_main, every chip of the 48-chip machine, 2 shards
    from lurk_b200.bench_programs import setup
    from lurk_b200.lurk_eval import build_lurk_toplevel
    top = build_lurk_toplevel()
    inp, q, _ = setup("fib", 40, top)
    execute_by_name(top, "lurk_main", inp, q)
    tq = _gpu_record_satisfies_air(oracle, top, q, top.func_by_name("lurk_main"), 256)
    assert tq.counts()["sends"] > 1000


@pytest.mark.parametrize("tname,fname,args", [("demo", "fib", [300]), ("partial", "pmain", [300, 21]), ("hash", "cons_chain", [9]),
                                              ("u64", "divrem", [255] * 8 + [1, 1] + [0] * 6), ("u64", "gcdu", [0, 0, 0, 1, 2, 3, 4, 5] + [0, 7, 0, 0, 9, 0, 0, 0])])
def test_gpu_traces_satisfy_the_independent_air(oracle, tname, fname, args):
    top = goldens.toplevel_for(tname)
    q = QueryRecord.new(top)
    execute_by_name(top, fname, args, q)
    _gpu_record_satisfies_air(oracle, top, q, top.func_by_name(fname), 64)
