"""The oracle's independent AIR / lookup checker (oracle/air_check.h): a row-by-row restatement of the reference's
`Air::eval` implementations and of its own semantic checker, src/air/debug.rs:91-107,161-206.

It never reads the product's captured constraint programs.  Pins: every golden trace of the reference's tests
satisfies it (the reference runs `debug_constraints_collecting_queries` on the same traces: src/lair/air.rs:599-623,
:676-712 ..., src/lair/memory.rs:131-177), the three memoset keys of `lair_range_test` (src/lair/air.rs:967-977),
and a perturbed cell is rejected.  It is then used two ways:
  * whole records -- every chip of the chip vector, every shard, multiset balance and count chains
    (`debug_chip_constraints_and_queries_with_sharding`, src/air/debug.rs:119-158, shard size 4 as in
    src/core/tests/mod.rs:57-63);
  * against the product: the captured 'MACH' chip vector (C++ `csrc/host_air.cpp` and its Python specification
    `lurk_b200/air.py`) must compute, on random rows, exactly the constraint values and messages of this
    restatement -- SURVEY.md 8a rows A6 / A7 no longer rest on the capture being right by construction."""
import numpy as np
import pytest

import goldens
from lurk_b200.lair import DEPTH_W, FuncChip, demo_toplevel

P = 2013265921
MEM_LENS = [2, 3, 4, 5, 6, 8]


def _record(oracle, top, fname, arg_list):
    ot = oracle.OracleToplevel(top.serialize())
    rec = oracle.OracleRecord(ot)
    f = top.func_by_name(fname)
    for args in arg_list:
        rec.execute(f.index, list(args), f.output_size)
    return ot, rec, f


def check_record(oracle, top, ot, rec, entry, shard_size=None):
    """debug_chip_constraints_and_queries_with_sharding (src/air/debug.rs:119-158) over build_lair_chip_vector
    (src/lair/lair_chip.rs:196-212): Entrypoint, every func, six memory tables, the byte table."""
    ss = shard_size or (1 << 22)
    tq = oracle.TraceQueries(ot)
    pv = rec.public_values()
    prep = oracle.bytes_preprocessed()
    for shard in range(max(1, rec.num_shards(ss))):
        if shard == 0:
            tq.check(oracle.AIR_ENTRY, entry.index, np.asarray([pv], dtype=np.uint32), public_values=pv)
        for f in top.funcs:
            if shard * ss < rec.num_queries(f.index):
                tq.check(oracle.AIR_FUNC, f.index, rec.func_trace(f.index, shard, ss))
        if shard == 0:
            for i in range(6):
                tq.check(oracle.AIR_MEM, i, rec.mem_trace(i))
        tq.check(oracle.AIR_BYTES, 0, rec.bytes_trace(shard), prep=prep)
    tq.verify()
    return tq


@pytest.mark.parametrize("case", goldens.CASES, ids=[c[0] + ":" + c[1] for c in goldens.CASES])
def test_golden_traces_satisfy_the_air(oracle, case):
    tname, fname, arg_list, golden = case
    top = goldens.toplevel_for(tname)
    ot = oracle.OracleToplevel(top.serialize())
    tq = oracle.TraceQueries(ot)
    tq.check(oracle.AIR_FUNC, top.func_by_name(fname).index, golden)      # the literal matrix of the reference's test
    assert tq.counts()["memoset_keys"] > 0


def test_golden_memory_trace_satisfies_the_air(oracle):
    # src/lair/memory.rs:146-176: the func trace and the Mem[3] trace; the reference checks the rows only
    top = goldens.toplevel_for("memory")
    ot = oracle.OracleToplevel(top.serialize())
    tq = oracle.TraceQueries(ot)
    tq.check(oracle.AIR_FUNC, top.func_by_name("test").index, goldens.MEMORY_FUNC)
    tq.check(oracle.AIR_MEM, 1, goldens.MEMORY_MEM3)
    # without the Entrypoint chip the func's provide (last_count = 1, execute.rs:444-447) has no matching require:
    with pytest.raises(oracle.AirCheckError):
        tq.verify()
    pv = [2, 2]
    tq.check(oracle.AIR_ENTRY, 0, np.asarray([pv], dtype=np.uint32), public_values=pv)
    tq.verify()


def test_lair_range_test_memoset_keys(oracle):
    # src/lair/air.rs:942-977
    top = goldens.toplevel_for("range")
    ot = oracle.OracleToplevel(top.serialize())
    tq = oracle.TraceQueries(ot)
    tq.check(oracle.AIR_FUNC, top.func_by_name("range_test").index, goldens.RANGE)
    assert tq.has_memoset_key([0, 0, 100, 12, 64])      # function return provide
    assert tq.has_memoset_key([3, 1, 100, 12])          # byte range checks for 100, 12, 64
    assert tq.has_memoset_key([3, 1, 64, 0])
    assert not tq.has_memoset_key([3, 1, 12, 64])
    assert tq.counts()["memoset_keys"] == 3


_SINGLE = [c for c in goldens.CASES if len(c[2]) == 1]    # one `execute` = one Entrypoint: the whole record balances


@pytest.mark.parametrize("case", _SINGLE, ids=[c[0] + ":" + c[1] for c in _SINGLE])
def test_perturbed_cells_are_rejected(oracle, case):
    # nonce, input, output and selector cells of a real row are all constrained (by an assertion or, as lookup operands,
    # by the balance check): changing any one of them must fail the row check or `verify` of the whole record.  (Auxiliary
    # cells are shared between branches, src/lair/func_chip.rs:146-175: the ones the taken branch does not use are free.)
    tname, fname, arg_list, golden = case
    top = goldens.toplevel_for(tname)
    ot, rec, f = _record(oracle, top, fname, arg_list)
    check_record(oracle, top, ot, rec, f)                   # the honest record passes
    rng = np.random.default_rng(5)
    real_rows = rec.num_queries(f.index)
    rejected = 0
    fixed = set(range(1 + f.input_size + f.output_size)) | set(range(golden.shape[1] - FuncChip(f, top).layout_sizes.sel, golden.shape[1]))
    cells = [(r, c) for r in range(real_rows) for c in sorted(fixed) if not (c == 0 and r == golden.shape[0] - 1)]
    for r, c in [cells[i] for i in rng.permutation(len(cells))[:24]]:
        bad = golden.copy()
        bad[r, c] = (int(bad[r, c]) + 1) % P
        tq = oracle.TraceQueries(ot)
        try:
            pv = rec.public_values()
            tq.check(oracle.AIR_ENTRY, f.index, np.asarray([pv], dtype=np.uint32), public_values=pv)
            tq.check(oracle.AIR_FUNC, f.index, bad)
            for g in top.funcs:
                if g.index != f.index and rec.num_queries(g.index):
                    tq.check(oracle.AIR_FUNC, g.index, rec.func_trace(g.index))
            for i in range(6):
                tq.check(oracle.AIR_MEM, i, rec.mem_trace(i))
            tq.check(oracle.AIR_BYTES, 0, rec.bytes_trace(0), prep=oracle.bytes_preprocessed())
            tq.verify()
        except oracle.AirCheckError:
            rejected += 1
    assert rejected == min(24, len(cells)), (rejected, len(cells))


@pytest.mark.parametrize("tname,fname,args,ss", [
    ("demo", "fib", [(40,)], None), ("demo", "fib", [(40,)], 4), ("demo", "factorial", [(9,)], 4),
    ("demo", "even", [(11,)], 4), ("ackermann", "ackermann", [(2, 9)], 4), ("memory", "test", [()], None),
    ("partial", "pfib", [(12,)], 4), ("partial", "pmain", [(300, 21)], 64),
])
def test_whole_records_balance(oracle, tname, fname, args, ss):
    top = goldens.toplevel_for(tname)
    ot, rec, f = _record(oracle, top, fname, args)
    tq = check_record(oracle, top, ot, rec, f, ss)
    n = tq.counts()
    assert n["sends"] == n["receives"] > 0


def _u64_cases():
    from test_oracle_stark import u64_bytes
    return [("add", u64_bytes(2**64 - 1) + u64_bytes(2)), ("sub", u64_bytes(0) + u64_bytes(1)),
            ("less_than", u64_bytes(0x0102030405060708) + u64_bytes(0x0102030505060708)), ("less_than", u64_bytes(77) + u64_bytes(77)),
            ("is_zero", u64_bytes(0)), ("is_zero", u64_bytes(2**40)),
            ("mul", u64_bytes(0xFFFFFFFFFFFFFFFF) + u64_bytes(0xFFFFFFFFFFFFFFFF)),
            ("divrem", u64_bytes(10**18 + 7) + u64_bytes(12345)), ("divrem", u64_bytes(5) + u64_bytes(2**63)),
            ("big_num_less_than", [1, 2, 3, 4, 5, 6, 7, 2013265920] + [1, 2, 3, 4, 5, 6, 7, 2013265919]),
            ("big_num_less_than", [3] * 8 + [3] * 8), ("gcdu", u64_bytes(2**40 * 3**5 * 7) + u64_bytes(2**20 * 3**9 * 11)),
            ("fibu", [60])]


@pytest.mark.parametrize("fname,args", _u64_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_u64_and_bignum_chipsets_satisfy_their_gadget_airs(oracle, fname, args):
    # src/core/u64.rs, src/core/big_num.rs over src/gadgets/unsigned/*: witness + byte lookups + constraints
    top = goldens.toplevel_for("u64")
    ot, rec, f = _record(oracle, top, fname, [args])
    check_record(oracle, top, ot, rec, f, 4)


@pytest.mark.parametrize("fname,args", [("commit_cons", "kat2"), ("cons_chain", [5]), ("env_chain", [3])])
def test_hasher_chips_satisfy_the_wide_poseidon2_air(oracle, fname, args):
    # src/poseidon/wide/air.rs:11-124 on the witness of src/poseidon/wide/trace.rs:12-82
    from test_oracle_stark import kat2_args
    top = goldens.toplevel_for("hash")
    ot, rec, f = _record(oracle, top, fname, [kat2_args() if args == "kat2" else args])
    check_record(oracle, top, ot, rec, f)


def test_a_lurk_evaluation_satisfies_the_air_of_all_chips(oracle):
    # the reference runs this check on every one of its evaluation tests (src/core/tests/mod.rs:57-63, shard size 4)
    from lurk_b200.zstore import ZStore
    from test_lurk_eval import _cases, _oracle_record
    top = goldens.toplevel_for("lurk_core")
    z = ZStore()
    expr, want = _cases(z)["a fixed point evaluates its body in the extended environment"]
    ot, rec, f, out = _oracle_record(oracle, top, z, expr)
    assert (out[0], out[1:]) == (want[0], list(want[1]))
    tq = check_record(oracle, top, ot, rec, f, 4)
    assert tq.counts()["sends"] > 100


# ---- the product's capture against the restatement -----------------------------------------------------------------
def _cross_check_all(oracle, top, entry, blob):
    ot = oracle.OracleToplevel(top.serialize())
    nchips = int(blob[1])
    assert nchips == 2 + len(top.funcs) + 6          # build_lair_chip_vector, lair_chip.rs:196-212
    total = np.zeros(3, dtype=np.int64)
    for ci in range(nchips):
        total += oracle.air_cross_check(ot, blob, ci, nrows=3, seed=11 + ci)
    return total


@pytest.mark.parametrize("tname,entry", [("demo", "fib"), ("memory", "test"), ("range", "range_test"), ("match_many", "match_many"),
                                         ("partial", "pmain"), ("hash", "cons_chain"), ("u64", "divrem"), ("lurk_misc", "u64_demo"),
                                         ("lurk_data", None), ("lurk_core", "lurk_eval_demo")])
def test_captured_chip_vector_equals_air_eval(oracle, tname, entry):
    from lurk_b200 import host
    from lurk_b200.air import build_chip_vector, serialize_machine
    top = goldens.toplevel_for(tname)
    f = top.func_by_name(entry) if entry else top.funcs[-1]
    num_pv = f.input_size + f.output_size + (DEPTH_W if f.partial else 0)
    cpp = host.chip_vector_blob(top.surface_image(), f.index)                    # csrc/host_air.cpp
    py = serialize_machine(build_chip_vector(FuncChip(f, top)), num_pv)          # lurk_b200/air.py
    for blob in (cpp, py):
        total = _cross_check_all(oracle, top, f, np.asarray(blob, dtype=np.uint32))
        assert total[0] > 0 and total[1] == total[2] > 0


def test_captured_lurk_machine_equals_air_eval(oracle):
    # the 48 chips of the machine the benches prove on: Entrypoint, 39 funcs (hash3/4/5 with the 449/603/755-column
    # Poseidon2 witness, the six u64 chips, big_num_lessthan, partial funcs with depth provenance), 6 memories, bytes
    from lurk_b200 import host
    from lurk_b200.lurk_eval import build_lurk_toplevel
    top = build_lurk_toplevel()
    f = top.func_by_name("lurk_main")
    blob = np.asarray(host.chip_vector_blob(top.surface_image(), f.index), dtype=np.uint32)
    total = _cross_check_all(oracle, top, f, blob)
    assert int(blob[1]) == 48
    assert total[0] > 4000 and total[1] == total[2] > 500


def test_a_dropped_constraint_is_noticed(oracle):
    # flip one constraint-program word of the captured blob: the cross-check must fail
    from lurk_b200 import host
    top = demo_toplevel()
    f = top.func_by_name("fib")
    blob = np.asarray(host.chip_vector_blob(top.surface_image(), f.index), dtype=np.uint32).copy()
    ot = oracle.OracleToplevel(top.serialize())
    oracle.air_cross_check(ot, blob, 2)
    hits = 0
    for i in range(len(blob)):
        if blob[i] == 5 and i + 4 < len(blob):           # an ASSERT opcode: turn it into a NEG into register 0
            bad = blob.copy()
            bad[i] = 4
            bad[i + 1] = 0
            try:
                for ci in range(int(blob[1])):
                    oracle.air_cross_check(ot, bad, ci)
            except (oracle.AirCheckError, RuntimeError):
                hits += 1
            if hits >= 3:
                break
    assert hits >= 3
