"""ZStore (lurk_b200/zstore.py, host hashing through lsk_poseidon2_hash_host) against the reference's known answers, and
data going into and out of the Lair memory through `ingress` / `egress` restated from src/core/ingress.rs."""
import numpy as np
import pytest

import goldens
from lurk_b200.execute import QueryRecord, execute_by_name
from lurk_b200.lair import FuncChip
from lurk_b200.zstore import TAG, ZStore


def test_zstore_known_answers():
    z = ZStore()
    # zstore.rs:1008-1019 and cli/tests/prove.lurk:12-17 / meta.rs:1040
    assert z.hide([0] * 8, z.num(123)).to_biguint() == int(goldens.KNOWN_ANSWERS["KAT-1"]["digest"], 16)
    assert z.commit(z.cons(z.u64(13), z.u64(17))).to_biguint() == int(goldens.KNOWN_ANSWERS["KAT-2"]["digest"], 16)
    # interning is content addressing: equal data, equal pointers; the hash caches remember every preimage
    assert z.string("abc") == z.string("abc") and z.string("abc") != z.string("abd")
    assert z.lurk_sym("nil") != z.lurk_sym("t") and z.builtin("cons").tag == TAG["Builtin"]
    assert len(z.hashes[24]) == 2 and len(z.hashes[32]) > 10


def test_ingress_egress_have_the_reference_widths():
    top = goldens.toplevel_for("lurk_data")
    for name, width in goldens.KNOWN_ANSWERS["lurk_data_widths"]["value"].items():   # incl. preallocate_symbols = 188
        assert FuncChip.from_name(name, top).width() == width, name


def _samples(z):
    return {
        "num": z.num(123),
        "u64": z.u64(2 ** 63 + 5),
        "char": z.char("é"),
        "nil": z.nil(),
        "t": z.t(),
        "symbol": z.lurk_sym("foo"),
        "builtin": z.builtin("cons"),
        "empty string": z.string(""),
        "string": z.string("hi"),
        "cons": z.cons(z.u64(13), z.u64(17)),
        "list": z.list([z.num(1), z.string("ab"), z.t(), z.list([z.u64(7), z.char("x")])]),
        "comm": z.commit(z.num(5)),
        "fun": z.tuple110(TAG["Fun"], z.list([z.lurk_sym("x")]), z.lurk_sym("x"),
                          z.tuple110(TAG["Env"], z.lurk_sym("y"), z.num(3), z.null(TAG["Env"]))),
    }


@pytest.mark.parametrize("which", ["num", "u64", "char", "nil", "t", "symbol", "builtin", "empty string", "string", "cons",
                                   "list", "comm", "fun"])
def test_data_round_trips_through_the_lair_memory(oracle, which):
    # the reference's test_ingress_egress (eval_direct.rs:2067-2120) in spirit: intern, inject the preimages, ingress,
    # egress, and the same ZPtr comes back -- through the native executor AND the oracle's
    top = goldens.toplevel_for("lurk_data")
    z = ZStore()
    x = _samples(z)[which]
    args = [x.tag] + [0] * 7 + list(x.digest)
    q = QueryRecord.new(top)
    z.inject(q, top)
    out = [int(v) for v in execute_by_name(top, "data_roundtrip", args, q)]
    assert out == [x.tag] + list(x.digest)
    ot = oracle.OracleToplevel(top.serialize())
    rec = oracle.OracleRecord(ot)
    for name, width in (("hash3", 24), ("hash4", 32), ("hash5", 40)):
        if z.hashes[width]:
            rec.inject_inv_queries(top.func_by_name(name).index, [list(k) for k in z.hashes[width]],
                                   [list(v) for v in z.hashes[width].values()])
    f = top.func_by_name("data_roundtrip")
    assert [int(v) for v in rec.execute(f.index, args, f.output_size)] == out


def test_ingress_rejects_a_digest_without_preimage():
    from lurk_b200._native import LurkStarkError
    top = goldens.toplevel_for("lurk_data")
    q = QueryRecord.new(top)
    with pytest.raises(LurkStarkError, match="Preimg not found"):
        execute_by_name(top, "data_roundtrip", [TAG["Cons"]] + [0] * 7 + [1, 2, 3, 4, 5, 6, 7, 8], q)


def test_extend_env_with_mutuals_builds_the_letrec_environment(oracle):
    # src/core/eval_direct.rs:1674-1725 restated (width pinned at 54, eval_direct.rs:2041): for bindings
    # ((f e1) (g e2)) the result is Env(f, Fix(e1, binds, env0), Env(g, Fix(e2, binds, env0), env0)) -- the ZStore
    # predicts its digest; the func runs on data ingressed from content-addressed form
    top = goldens.toplevel_for("lurk_data")
    assert FuncChip.from_name("extend_env_with_mutuals", top).width() == 54
    z = ZStore()
    f, g = z.lurk_sym("f"), z.lurk_sym("g")
    e1, e2 = z.list([z.builtin("cons"), z.num(1), g]), z.num(42)
    binds = z.list([z.list([f, e1]), z.list([g, e2])])
    env0 = z.null(TAG["Env"])
    fix = lambda e: z.tuple110(TAG["Fix"], e, binds, env0)
    want = z.tuple110(TAG["Env"], f, fix(e1), z.tuple110(TAG["Env"], g, fix(e2), env0))
    args = [binds.tag] + [0] * 7 + list(binds.digest)
    q = QueryRecord.new(top)
    z.inject(q, top)
    out = [int(v) for v in execute_by_name(top, "mutuals_demo", args, q)]
    assert out == [TAG["Env"]] + list(want.digest)
    # malformed bindings: (f) has no expression -> (Err, InvalidForm); a number as variable -> (Err, IllegalBindingVar)
    for bad, err in ((z.list([z.list([f])]), 1), (z.list([z.list([z.num(3), e2])]), 2)):
        q = QueryRecord.new(top)
        z.inject(q, top)
        out = [int(v) for v in execute_by_name(top, "mutuals_demo", [bad.tag] + [0] * 7 + list(bad.digest), q)]
        assert out == [TAG["Err"], err] + [0] * 7


# ---- batch interning (zstore.rs:305-349 hashed level by level) -------------------------------------
def _intern_directly(z, obj):
    if isinstance(obj, int):
        return z.u64(obj)
    if isinstance(obj, str):
        return z.string(obj)
    return z.list([_intern_directly(z, x) for x in obj])


def _wide_data(n):
    # n sublists: every level of the cons DAG is ~n nodes wide; shared tails and repeated strings exercise the caches
    return [[i, i % 7, [3 * i, "ab"], "shared"] for i in range(n)]


def test_deferred_interning_equals_one_by_one_interning():
    a, b = ZStore(), ZStore()
    want = _intern_directly(a, _wide_data(40))
    got = b.resolve([b.defer_data(_wide_data(40))])[0]
    assert got == want
    assert a.hashes == b.hashes                     # the same preimages are remembered, each once
    assert b.gpu_batches == 0                       # no context given: host hasher only
    # functions / environments (hash5) and commitments (hash3), resolved together with their children
    env = b.defer_tuple110(TAG["Env"], b.lurk_sym("x"), b.defer_list([b.u64(1), b.u64(2)]), b.null(TAG["Env"]))
    comm = b.defer_hide([0] * 8, env)
    zenv, zcomm = b.resolve([env, comm])
    lst = a.list([a.u64(1), a.u64(2)])
    assert zenv == a.tuple110(TAG["Env"], a.lurk_sym("x"), lst, a.null(TAG["Env"]))
    assert zcomm == a.commit(zenv)
    assert b.resolve([]) == [] and b.resolve([a.u64(5)]) == [a.u64(5)]


def test_deferred_levels_are_the_dependency_depth():
    z = ZStore()
    chain = z.defer_list([z.u64(i) for i in range(10)])
    assert chain.level == 10                        # a list is a chain: ten levels, one node each
    wide = z.defer_list([z.defer_list([z.u64(i)]) for i in range(10)])
    assert wide.children[0].level == 1 and wide.level == 11   # the ten sublists are one level; the spine is a chain


@pytest.mark.gpu
def test_deferred_interning_on_the_gpu_equals_the_host():
    from lurk_b200 import _native as N
    ctx = N.default_context()
    a, b = ZStore(), ZStore()
    want = _intern_directly(a, _wide_data(600))
    got = b.resolve([b.defer_data(_wide_data(600))], ctx=ctx, gpu_min_batch=64)[0]
    assert got == want and a.hashes == b.hashes
    assert b.gpu_batches >= 3                       # the wide levels went through lsk_poseidon2_hash
    # KAT-1 (zstore.rs:1008-1019) through the batch path: commitments of 300 numbers, 123 among them
    c = ZStore()
    comms = c.resolve([c.defer_hide([0] * 8, c.num(i)) for i in range(300)], ctx=ctx, gpu_min_batch=64)
    assert c.gpu_batches == 1
    assert comms[123].to_biguint() == 0xaa8db8504fa55b480f3da7a75f3480174f28d683f4c3ac451b7cee488d2fe
