"""CPU-side tests of the native host logic: the C-ABI library loads and exports every declared
symbol, and the flattening executor agrees with the oracle's restatement of execute.rs (no GPU
compute is called here)."""
import ctypes

import numpy as np
import pytest

import goldens
from lurk_b200 import _native as N
from lurk_b200.execute import QueryRecord, Shard, ShardingConfig, execute_by_name
from lurk_b200.lair import demo_toplevel


@pytest.fixture(scope="module", autouse=True)
def _built():
    N.build()


def test_library_exports_every_declared_symbol():
    L = N.lib()
    syms = N.declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_ctx_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(N.LurkStarkError) as ei:
        N.Context(0)
    assert "no CPU fallback" in str(ei.value)


@pytest.mark.parametrize("case", goldens.CASES, ids=[c[0] + ":" + c[1] for c in goldens.CASES])
def test_executor_matches_oracle(oracle, case):
    tname, fname, arg_list, golden = case
    top = goldens.toplevel_for(tname)
    f = top.func_by_name(fname)
    q = QueryRecord.new(top)
    ot = oracle.OracleToplevel(top.serialize())
    orec = oracle.OracleRecord(ot)
    for args in arg_list:
        got = execute_by_name(top, fname, args, q)
        want = orec.execute(f.index, list(args), f.output_size)
        assert np.array_equal(got, want)
    for fi in range(top.num_funcs()):
        assert q.num_queries(fi) == orec.num_queries(fi)
    for mi in range(6):
        assert q.num_mem(mi) == orec.num_mem(mi)
    assert np.array_equal(q.expect_public_values(), orec.public_values())
    # outputs column of the golden equals what the executor memoised
    assert q.num_queries(f.index) <= golden.shape[0]


def test_partial_funcs_depth_provenance(oracle):
    # src/lair/execute.rs:387-389,510-512,702-711: depth = 1 + max over partial callees, public values
    # gain the depth's 4 little-endian bytes; the native executor and the oracle agree
    top = goldens.toplevel_for("partial")
    f = top.func_by_name("pmain")
    q = QueryRecord.new(top)
    got = execute_by_name(top, "pmain", [300, 21], q)
    ot = oracle.OracleToplevel(top.serialize())
    orec = oracle.OracleRecord(ot)
    want = orec.execute(f.index, [300, 21], 2)
    assert np.array_equal(got, want) and int(got[1]) == 21
    pv = q.expect_public_values()
    assert np.array_equal(pv, orec.public_values())
    assert list(pv[-4:]) == [300 & 255, 300 >> 8, 0, 0]       # pfib(300) has depth 299
    for fi in range(top.num_funcs()):
        assert q.num_queries(fi) == orec.num_queries(fi)


def test_u64_chipsets_executor_matches_oracle(oracle):
    # src/core/u64.rs:76-128: native executor == oracle, incl. the byte-lookup bookkeeping (public values, counts)
    from test_oracle_stark import u64_bytes
    top = goldens.toplevel_for("u64")
    ot = oracle.OracleToplevel(top.serialize())
    for fname, args in [("add", u64_bytes(2**64 - 3) + u64_bytes(77)), ("sub", u64_bytes(5) + u64_bytes(2**40)),
                        ("less_than", u64_bytes(2**56) + u64_bytes(2**56 + 1)), ("fibu", [120]),
                        ("maxu", u64_bytes(3) + u64_bytes(3)), ("factu", u64_bytes(30)),
                        ("gcdu", u64_bytes(10**18) + u64_bytes(6**20)), ("divrem", u64_bytes(2**64 - 1) + u64_bytes(2**32 + 1)),
                        ("mul", u64_bytes(2**63 + 12345) + u64_bytes(2**33 + 99)), ("is_zero", u64_bytes(256)),
                        ("big_num_less_than", [9, 8, 7, 6, 5, 4, 3, 2013265920] + [9, 8, 7, 6, 5, 4, 3, 77])]:
        f = top.func_by_name(fname)
        q = QueryRecord.new(top)
        got = execute_by_name(top, fname, args, q)
        orec = oracle.OracleRecord(ot)
        want = orec.execute(f.index, args, f.output_size)
        assert np.array_equal(got, want), fname
        assert np.array_equal(q.expect_public_values(), orec.public_values())
        for fi in range(top.num_funcs()):
            assert q.num_queries(fi) == orec.num_queries(fi)
    q = QueryRecord.new(top)
    with pytest.raises(N.LurkStarkError, match="byte"):
        execute_by_name(top, "add", [256] + [0] * 15, q)       # into_u64 panics on a non-byte limb
    with pytest.raises(N.LurkStarkError, match="non-zero"):
        execute_by_name(top, "divrem", u64_bytes(7) + u64_bytes(0), QueryRecord.new(top))   # division by zero


def test_total_func_may_not_call_partial():
    # src/lair/execute.rs:1005-1023 (nonpartial_calls_partial: "assertion failed: ctx.partial")
    from lurk_b200.lair import Toplevel, call, func, ret
    foo = func("foo", ["a"], 1, [ret("a")], partial=True)
    bar = func("bar", ["a"], 1, [call(["b"], "foo", ["a"]), ret("b")])
    with pytest.raises(AssertionError, match="ctx.partial"):
        Toplevel.new_pure([foo, bar])


def test_executor_fib_100000():
    # src/lair/execute.rs:826-834
    top = demo_toplevel()
    q = QueryRecord.new(top)
    out = execute_by_name(top, "fib", [100000], q)
    assert int(out[0]) == 1123328132
    assert q.num_queries(1) == 100001
    assert list(q.expect_public_values()) == [100000, 1123328132]


def test_executor_errors_are_reported_not_thrown():
    top = goldens.toplevel_for("assert")
    q = QueryRecord.new(top)
    with pytest.raises(N.LurkStarkError):
        execute_by_name(top, "assert", [2, 4, 5, 8], q)   # assert_ne!(a, arr1) fails


def test_sharding_policy():
    # src/lair/execute.rs:106-111,186-216
    top = demo_toplevel()
    q = QueryRecord.new(top)
    execute_by_name(top, "fib", [9], q)
    shards = Shard.new(q).shard(ShardingConfig(4))
    assert len(shards) == 3
    assert [len(s.get_func_range(1)) for s in shards] == [4, 4, 2]
    assert len(Shard.new(q).shard(ShardingConfig.default())) == 1


def test_loaders_reject_malformed_images_without_crashing():
    # lsk_host_compile_toplevel (surface image), lsk_program_load (program image: every walk of the executor and of
    # the trace kernel must stay inside the image and the value stack) and lsk_vk_new ('MACH' blob): truncations,
    # absurd counts at any position and random words end in an error code or in another well-formed object
    import ctypes
    from lurk_b200 import host
    from lurk_b200.programs import hash_toplevel
    L = N.lib()
    rng = np.random.default_rng(3)

    def try_surface(w):
        out, n = ctypes.POINTER(ctypes.c_uint32)(), ctypes.c_size_t()
        rc = L.lsk_host_compile_toplevel(N.vp(w), ctypes.c_size_t(len(w)), ctypes.byref(out), ctypes.byref(n))
        if rc == 0:
            L.lsk_words_free(out)
        return rc

    def try_image(w):
        h = ctypes.c_void_p()
        rc = L.lsk_program_load(N.vp(w), ctypes.c_size_t(len(w)), ctypes.byref(h))
        if rc == 0:
            L.lsk_program_free(h)
        return rc

    def try_blob(w):
        h = ctypes.c_void_p()
        rc = L.lsk_vk_new(N.vp(w), ctypes.c_size_t(len(w)), N.vp(N.u32(np.zeros(8))), 1, 6, 6, ctypes.byref(h))
        if rc == 0:
            L.lsk_vk_free(h)
        return rc

    for top, entry in ((demo_toplevel(), "fib"), (hash_toplevel(), "cons_chain")):
        surface = N.u32(top.surface_image())
        image = N.u32(top.serialize())
        blob = N.u32(host.chip_vector_blob(surface, top.func_by_name(entry).index))
        for fn, base in ((try_surface, surface), (try_image, image), (try_blob, blob)):
            assert fn(base) == 0
            rejected = 0
            for k in list(range(12)) + list(rng.integers(0, len(base), 40)):
                rejected += fn(N.u32(base[:k])) != 0
            assert rejected >= 45                       # a prefix is (almost) never a whole image
            for pos in list(range(min(len(base), 150))) + list(rng.integers(0, len(base), 150)):
                for val in (0xffffffff, 1 << 24):
                    w = base.copy()
                    w[pos] = val
                    fn(w)                               # must return, whatever it answers
            for _ in range(30):
                assert fn(rng.integers(0, 2 ** 32, size=int(rng.integers(1, 2000)), dtype=np.uint64).astype(np.uint32)) != 0


def test_jit_precompile_fills_and_reuses_the_cubin_cache(tmp_path, monkeypatch):
    # csrc/jit.cu: NVRTC needs no GPU.  With the threshold at 1024 instructions a small program has nothing to
    # compile; the u64 chipset programs (2 chips above it) compile into an empty cache and are found there again.
    from lurk_b200 import jit_warm
    from lurk_b200.programs import u64_toplevel
    monkeypatch.setenv("LSK_JIT_CACHE", str(tmp_path / "cache"))
    monkeypatch.setenv("LSK_JIT_MIN_INSTR", "1024")
    assert jit_warm.precompile(demo_toplevel(), "fib") == {"compiled": 0, "cached": 0}
    assert jit_warm.precompile(u64_toplevel(), "factu") == {"compiled": 2, "cached": 0}
    assert jit_warm.precompile(u64_toplevel(), "factu") == {"compiled": 0, "cached": 2}
    files = sorted(p.name for p in (tmp_path / "cache").iterdir())
    assert len(files) == 2 and all(f.startswith("q_") and f.endswith("_sm100a.cubin") for f in files)


def test_lurk_misc_u64_funcs_native_executor(oracle):
    # src/core/misc.rs funcs through the native executor: same outputs as the oracle and as Python integers
    from test_oracle_stark import _u64_demo_expected, u64_bytes
    top = goldens.toplevel_for("lurk_misc")
    for a, b in [(10**18 + 7, 12345), (5, 2**64 - 1), (77, 77), (123456789, 0)]:
        q = QueryRecord.new(top)
        out = execute_by_name(top, "u64_demo", u64_bytes(a) + u64_bytes(b), q)
        assert [int(x) for x in out] == _u64_demo_expected(a, b)


def test_the_package_never_imports_the_checker():
    # oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it
    import os
    import re
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lurk_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, name), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), name
                assert "liboracle" not in src and "pyoracle" not in src, name


def _roundtrip_tables(top, entry, args, setup=None):
    import numpy as np
    from lurk_b200.execute import QueryRecord, execute_by_name
    q = QueryRecord.new(top)
    inv = setup(q) if setup else ()
    execute_by_name(top, entry, args, q)
    tables = q.export_tables()
    q2 = QueryRecord.from_tables(top, tables, inv)
    t2 = q2.export_tables()
    assert set(tables["funcs"]) == set(t2["funcs"]) and set(tables["mems"]) == set(t2["mems"])
    for idx, t in tables["funcs"].items():
        for k, v in t.items():
            assert np.array_equal(v, t2["funcs"][idx][k]), (idx, k)      # the HINTS included: resolve recomputed them
    for i, m in tables["mems"].items():
        for k, v in m.items():
            assert np.array_equal(v, t2["mems"][i][k]), (i, k)
    assert np.array_equal(tables["bytes"], t2["bytes"])
    assert np.array_equal(tables["public_values"], t2["public_values"])
    return tables, q2


@pytest.mark.parametrize("tname,fname,args", [("demo", "fib", [200]), ("demo", "even", [33]), ("memory", "test", []),
                                              ("partial", "pmain", [300, 21]), ("range", "range_test", [100, 12, 64]),
                                              ("u64", "divrem", [255] * 8 + [1, 1] + [0] * 6), ("hash", "cons_chain", [9]),
                                              ("ackermann", "ackermann", [2, 12])])
def test_an_imported_record_equals_the_executed_one(tname, fname, args):
    # lsk_record_import_* + lsk_record_resolve: the drop-in path for a QueryRecord made by the reference's own executor
    import goldens
    tables, _ = _roundtrip_tables(goldens.toplevel_for(tname), fname, args)
    assert any(t["hints"].size for t in tables["funcs"].values()) or tname in ("range", "u64")


def test_an_imported_lurk_record_equals_the_executed_one():
    # a Lurk evaluation through lurk_main: 17 func tables, 2 memories, byte lookups, injected hash preimages
    from lurk_b200.bench_programs import source
    from lurk_b200.lurk_eval import build_lurk_toplevel
    from lurk_b200.reader import read
    from lurk_b200.zstore import ZStore
    top = build_lurk_toplevel()
    z = ZStore()
    e = read(z, source("fib", 25))
    inv = [(name, list(z.hashes[w].items())) for name, w in (("hash3", 24), ("hash4", 32), ("hash5", 40)) if z.hashes[w]]

    def setup(q):
        z.inject(q, top)
        return inv
    tables, _ = _roundtrip_tables(top, "lurk_main", [e.tag] + [0] * 7 + list(e.digest) + [0] * 8, setup)
    assert len(tables["funcs"]) >= 15


def test_import_refuses_inconsistent_tables():
    import numpy as np
    from lurk_b200._native import LurkStarkError
    from lurk_b200.execute import QueryRecord, execute_by_name
    from lurk_b200.lair import demo_toplevel
    top = demo_toplevel()
    q = QueryRecord.new(top)
    execute_by_name(top, "fib", [30], q)
    tables = q.export_tables()
    fib = top.func_by_name("fib").index
    for mutate in ("output", "missing-row", "duplicate", "non-canonical", "offsets"):
        t = {k: v.copy() for k, v in tables["funcs"][fib].items()}
        if mutate == "output":
            t["outs"][7, 0] = (int(t["outs"][7, 0]) + 1) % 2013265921      # resolve: the row does not return that
        elif mutate == "missing-row":
            n = len(t["provide"]) - 1                                        # drop fib(0): fib(2) calls a query that is not there
            t = {"args": t["args"][:n], "outs": t["outs"][:n], "provide": t["provide"][:n], "req_off": t["req_off"][:n + 1],
                 "reqs": t["reqs"][:int(t["req_off"][n])]}
        elif mutate == "duplicate":
            t["args"][3] = t["args"][4]
        elif mutate == "non-canonical":
            t["args"][3, 0] = 2013265921
        else:
            t["req_off"][5] = t["req_off"][6] + 1
        funcs = dict(tables["funcs"])
        funcs[fib] = t
        bad = dict(tables, funcs=funcs)
        with pytest.raises(LurkStarkError):
            QueryRecord.from_tables(top, bad)
