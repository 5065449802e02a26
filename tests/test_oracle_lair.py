"""Pins the CPU oracle's Lair executor + trace generation against every golden trace matrix the
reference's tests hold."""
import numpy as np
import pytest

import goldens
from lurk_b200.lair import FuncChip, LayoutSizes, demo_toplevel


def _run(oracle, top, fname, arg_list):
    ot = oracle.OracleToplevel(top.serialize())
    rec = oracle.OracleRecord(ot)
    f = top.func_by_name(fname)
    for args in arg_list:
        rec.execute(f.index, list(args), f.output_size)
    return rec, f


@pytest.mark.parametrize("case", goldens.CASES, ids=[c[0] + ":" + c[1] for c in goldens.CASES])
def test_golden_func_trace(oracle, case):
    tname, fname, arg_list, golden = case
    top = goldens.toplevel_for(tname)
    rec, f = _run(oracle, top, fname, arg_list)
    trace = rec.func_trace(f.index)
    assert trace.shape == golden.shape
    assert np.array_equal(trace, golden)


def test_golden_mem_trace(oracle):
    # src/lair/memory.rs:157-169
    top = goldens.toplevel_for("memory")
    rec, _ = _run(oracle, top, "test", [()])
    assert np.array_equal(rec.mem_trace(1), goldens.MEMORY_MEM3)
    for idx in (0, 2, 3, 4, 5):  # empty tables: height max(4, ..) all zero
        t = rec.mem_trace(idx)
        assert t.shape[0] == 4 and not t.any()


def test_layout_sizes():
    # src/lair/trace.rs:446-459, :545-552, :612-619
    top = demo_toplevel()
    assert top.compute_layout_sizes(top.func_by_name("factorial")) == LayoutSizes(1, 1, 1, 8, 2, 6)
    ls = FuncChip.from_name("test", goldens.toplevel_for("match")).layout_sizes
    assert (ls.nonce, ls.input, ls.aux, ls.sel, ls.output) == (1, 2, 10, 5, 1)
    ls = FuncChip.from_name("test", goldens.toplevel_for("inner_match")).layout_sizes
    assert (ls.nonce, ls.input, ls.aux, ls.sel, ls.output) == (1, 2, 2, 4, 1)
    assert FuncChip.from_name("fib", top).width() == 18


def test_fib_100000(oracle):
    # src/lair/execute.rs:826-834 (lair_execute_iter_test)
    top = demo_toplevel()
    rec, f = _run(oracle, top, "fib", [])
    out = rec.execute(f.index, [100000], 1)
    assert int(out[0]) == 1123328132
    assert rec.num_queries(f.index) == 100001
    assert list(rec.public_values()) == [100000, 1123328132]


def test_ackermann_small(oracle):
    # src/lair/trace.rs:683-687: A(3, n) = 2^(n+3) - 3
    top = goldens.toplevel_for("ackermann")
    rec, f = _run(oracle, top, "ackermann", [])
    assert int(rec.execute(f.index, [3, 5], 1)[0]) == 2 ** 8 - 3


def test_sharding_ranges(oracle):
    # src/lair/execute.rs:106-111,186-216 with the reference's test shard size 4 (core/tests/mod.rs:57-63)
    top = demo_toplevel()
    rec, f = _run(oracle, top, "fib", [(9,)])
    assert rec.num_queries(f.index) == 10
    assert rec.num_shards(4) == 3
    full = rec.func_trace(f.index)
    parts = [rec.func_trace(f.index, shard=s, shard_size=4) for s in range(3)]
    assert [p.shape[0] for p in parts] == [4, 4, 2]
    assert np.array_equal(np.concatenate(parts)[:10], full[:10])


def test_bytes_traces(oracle):
    # src/gadgets/bytes/trace.rs:49-67,75-101; record from src/lair/air.rs:943-977
    top = goldens.toplevel_for("range")
    rec, _ = _run(oracle, top, "range_test", [(100, 12, 64)])
    main = rec.bytes_trace()
    assert main.shape == (1 << 16, 13) and (main[:, 0] == 1).all()
    assert list(main[100 | (12 << 8), 1:3]) == [0, 1]   # range_u8 provide record (nonce 0, count 1)
    assert list(main[64, 1:3]) == [0, 1]
    assert main[:, 1:].sum() == 2
    assert not rec.bytes_trace(shard=1).any()
    prep = oracle.bytes_preprocessed()
    i = 0x3c | (0xa5 << 8)
    assert list(prep[i]) == [0x3c, 0xa5, 1, 0x3c & 0xa5, 0x3c ^ 0xa5, 0x3c | 0xa5]


def test_lurk_misc_funcs_have_the_reference_widths():
    # src/core/misc.rs:5-128 restated (lurk_b200/programs.py::lurk_misc_toplevel): the widths the reference's own
    # test pins for these eleven funcs (eval_direct.rs:2053-2063) -- load / store / extern_call / `if` on an array /
    # invertible funcs through compute_layout_sizes, in the Python specification and in the C++ host layer
    import goldens
    from lurk_b200 import host
    from lurk_b200.lair import FuncChip
    top = goldens.toplevel_for("lurk_misc")
    want = goldens.KNOWN_ANSWERS["lurk_misc_widths"]["value"]
    assert len(want) == 11
    for name, width in want.items():
        assert FuncChip.from_name(name, top).width() == width, name
    # the C++ layer compiles the same surface to the same image (header words aux / sel give the same widths)
    assert np.array_equal(host.compile_toplevel(top.surface_image()), top.serialize_py())


def test_lurk_evaluator_funcs_have_the_reference_widths():
    # four funcs of the Lurk evaluator restated statement for statement (tests/lair_programs.py): a partial func with
    # depth columns (car_cdr), a preimage of an invertible func (open_comm), multi-key and exhaustive matches, negated
    # ifs, loads into several variables -- the widths the reference's own test pins (eval_direct.rs:2044-2048)
    import goldens
    from lurk_b200 import host
    from lurk_b200.lair import FuncChip
    top = goldens.toplevel_for("lurk_eval_subset")
    for name, width in goldens.KNOWN_ANSWERS["lurk_eval_subset_widths"]["value"].items():
        assert FuncChip.from_name(name, top).width() == width, name
    assert np.array_equal(host.compile_toplevel(top.surface_image()), top.serialize_py())


def test_the_assembled_lurk_toplevel_has_every_reference_width():
    # build_lurk_toplevel (lurk_b200/lurk_eval.py; src/core/eval_direct.rs:29-118) assembled from the real funcs only --
    # no stubs -- gives all 39 widths of the reference's test_widths (eval_direct.rs:2025-2063) in ONE toplevel, in the
    # Python specification and through the C++ host layer alike
    from lurk_b200 import host
    from lurk_b200.lair import FuncChip
    from lurk_b200.lurk_eval import build_lurk_toplevel
    top = build_lurk_toplevel()
    want = dict(lurk_main=97, preallocate_symbols=188, eval_coroutine_expr=10, eval=78, eval_builtin_expr=148,
                eval_bind_builtin=110, eval_env_builtin=81, eval_apply_builtin=79, eval_opening_unop=97, eval_hide=115,
                eval_unop=78, eval_binop_num=107, eval_binop_misc=70, eval_begin=68, eval_list=72, eval_let=94,
                eval_letrec=66, extend_env_with_mutuals=54, eval_letrec_bindings=66, coerce_if_sym=9, open_comm=50,
                equal=86, equal_inner=58, car_cdr=61, apply=114, env_lookup=52, ingress=104, egress=81, hash3=493,
                hash4=655, hash5=815, u64_add=53, u64_sub=53, u64_mul=85, u64_divrem=166, u64_lessthan=44,
                u64_iszero=26, digest_equal=38, big_num_lessthan=78)
    assert len(want) == 39
    assert {f.name for f in top.funcs_e} == set(want) | {"lurk_eval_demo"}
    assert top.funcs_e[0].name == "lurk_main"
    for name, width in want.items():
        assert FuncChip.from_name(name, top).width() == width, name
    assert np.array_equal(host.compile_toplevel(top.surface_image()), top.serialize_py())
