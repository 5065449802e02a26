"""GPU parity for the prover stages and for whole proofs: every test calls the CUDA library through
its C ABI and compares, bit for bit, with the CPU oracle (oracle/stark.h, oracle/stark_prove.h).
The oracle's verifier must also accept every GPU proof and reject a tampered one."""
import ctypes
import json
import os

import numpy as np
import pytest

import goldens
from lurk_b200 import _native as N
from lurk_b200.execute import QueryRecord, Shard, ShardingConfig, execute_by_name
from lurk_b200.lair import demo_toplevel
from lurk_b200.machine import SphinxCoreOpts, assemble_proof, new_machine

pytestmark = pytest.mark.gpu
P = 2013265921


def two_adic_generator(bits):
    return pow(0x1A427A41, 1 << (27 - bits), P)


def gpu_coset_lde(rows, kind):
    rows = N.u32(rows)
    h, w = rows.shape
    out = np.zeros((2 * h, w), dtype=np.uint32)
    N.check(N.lib().lsk_coset_lde(N.default_context().h, N.vp(rows), ctypes.c_size_t(h), ctypes.c_size_t(w), kind,
                                  N.vp(out)))
    return out


def gpu_merkle_root(mats):
    ms = [N.u32(m) for m in mats]
    n = len(ms)
    ptrs = (ctypes.c_void_p * n)(*[m.ctypes.data for m in ms])
    hs = (ctypes.c_size_t * n)(*[m.shape[0] for m in ms])
    ws = (ctypes.c_size_t * n)(*[m.shape[1] for m in ms])
    out = np.zeros(8, dtype=np.uint32)
    N.check(N.lib().lsk_merkle_root(N.default_context().h, ptrs, hs, ws, ctypes.c_size_t(n), N.vp(out)))
    return out


@pytest.mark.parametrize("logh,w", [(0, 3), (1, 1), (2, 6), (5, 18), (10, 7), (12, 5), (13, 3), (14, 18), (17, 4),
                                    (21, 2)])
def test_coset_lde_matches_oracle(oracle, logh, w):
    rng = np.random.default_rng(logh * 100 + w)
    rows = rng.integers(0, P, size=(1 << logh, w), dtype=np.uint32)
    assert np.array_equal(gpu_coset_lde(rows, 0), oracle.coset_lde(rows, 1))


@pytest.mark.parametrize("logh", [0, 3, 12, 14])
def test_quotient_chunk_lde_matches_oracle(oracle, logh):
    rng = np.random.default_rng(7 + logh)
    rows = rng.integers(0, P, size=(1 << logh, 4), dtype=np.uint32)
    assert np.array_equal(gpu_coset_lde(rows, 1), oracle.coset_lde(rows, 31))
    shift = 31 * two_adic_generator(logh + 1) % P
    assert np.array_equal(gpu_coset_lde(rows, 2), oracle.coset_lde(rows, shift))


def test_lde_is_linear_and_extends_constants():
    # size-independent properties at a size the oracle is too slow for
    rng = np.random.default_rng(1)
    h, w = 1 << 20, 3
    a = rng.integers(0, P, size=(h, w), dtype=np.uint32)
    b = rng.integers(0, P, size=(h, w), dtype=np.uint32)
    s = ((a.astype(np.uint64) + b) % P).astype(np.uint32)
    la, lb, ls = gpu_coset_lde(a, 0), gpu_coset_lde(b, 0), gpu_coset_lde(s, 0)
    assert np.array_equal(((la.astype(np.uint64) + lb) % P).astype(np.uint32), ls)
    c = np.full((h, 1), 12345, dtype=np.uint32)
    assert (gpu_coset_lde(c, 0) == 12345).all()


@pytest.mark.parametrize("shapes", [
    [(1, 3)], [(2, 8)], [(8, 5), (8, 9)], [(64, 18), (8, 3), (64, 2), (1, 7)], [(4096, 16), (2048, 1), (2, 40)],
    [(1 << 15, 8), (1 << 15, 13), (1 << 11, 655), (4, 6)],
])
def test_merkle_root_matches_oracle(oracle, shapes):
    rng = np.random.default_rng(len(shapes))
    mats = [rng.integers(0, P, size=s, dtype=np.uint32) for s in shapes]
    assert np.array_equal(gpu_merkle_root(mats), oracle.merkle_root(mats))


def pk_of(m):
    return m.setup()[0]


def _prove_both(oracle, top, entry, arg_list, shard_size=1 << 22, nq=6, pow_bits=6):
    opts = SphinxCoreOpts(num_queries=nq, proof_of_work_bits=pow_bits)
    m = new_machine(top, entry, opts=opts)
    pk, vk = m.setup()
    q = QueryRecord.new(top)
    ot = oracle.OracleToplevel(top.serialize())
    om = oracle.OracleMachine(ot, m.entry.index, m.blob, nq, pow_bits)
    rec = oracle.OracleRecord(ot)
    for args in arg_list:
        execute_by_name(top, entry, args, q)
        rec.execute(m.entry.index, list(args), m.entry.output_size)
    assert np.array_equal(pk.commit, om.prep_commit())
    proof = m.prove(pk, Shard(q, 0, ShardingConfig(shard_size)))
    want = om.prove(rec, shard_size)
    return m, om, q, proof, want


@pytest.mark.parametrize("n", [0, 1, 7, 100, 5000])
def test_fib_proof_is_bit_exact(oracle, n):
    # src/lair/lair_chip.rs:246-276 / benches/fib.rs shape on the demo toplevel's fib
    m, om, q, proof, want = _prove_both(oracle, demo_toplevel(), "fib", [[n]])
    assert len(proof.words) == len(want)
    assert np.array_equal(proof.words, want)
    assert om.verify(proof.words), om.last_error
    bad = proof.words.copy()
    bad[40] = (int(bad[40]) + 1) % P
    assert not om.verify(bad)


def test_factorial_proof_default_parameters(oracle):
    # lair_chip.rs:246-276 (factorial), sphinx default FRI parameters: 100 queries, 16 PoW bits
    m, om, q, proof, want = _prove_both(oracle, demo_toplevel(), "factorial", [[5]], nq=100, pow_bits=16)
    assert np.array_equal(proof.words, want)
    assert om.verify(proof.words), om.last_error


@pytest.mark.parametrize("tname,fname,args", [
    ("range", "range_test", [100, 12, 64]),     # byte lookups -> non-trivial Bytes chip
    ("memory", "test", []),                      # store / load -> Mem chips
    ("match_many", "match_many", [1, 0]),
    ("assert", "assert", [2, 4, 6, 8]),
    ("not_eq", "eq", [4, 2]),
])
def test_reference_test_programs_prove_bit_exact(oracle, tname, fname, args):
    top = goldens.toplevel_for(tname)
    m, om, q, proof, want = _prove_both(oracle, top, fname, [args])
    assert np.array_equal(proof.words, want)
    assert om.verify(proof.words), om.last_error


@pytest.mark.parametrize("fname,args", [("commit_cons", "kat2"), ("cons_chain", [9]), ("env_chain", [4])])
def test_hasher_chip_programs_prove_bit_exact(oracle, fname, args):
    # hash3/hash4/hash5 funcs (widths 493/655/815): Poseidon2 wide witness on the GPU + the captured
    # Poseidon2Cols::eval constraints (src/poseidon/wide/air.rs:11-124) in the quotient kernel
    from test_oracle_stark import KAT2, big_uint, kat2_args
    top = goldens.toplevel_for("hash")
    a = kat2_args() if args == "kat2" else args
    m, om, q, proof, want = _prove_both(oracle, top, fname, [a])
    if args == "kat2":
        assert big_uint(q.expect_public_values()[-8:]) == KAT2
    assert np.array_equal(proof.words, want)
    assert om.verify(proof.words), om.last_error


@pytest.mark.parametrize("fname,args,ss", [("pfib", [12], 1 << 22), ("pmain", [300, 21], 1 << 22), ("pmain", [700, 5], 256)])
def test_partial_func_programs_prove_bit_exact(oracle, fname, args, ss):
    # depth provenance on the GPU: depth bytes + LessThanWitness in the trace kernel, byte less-than lookups
    top = goldens.toplevel_for("partial")
    m, om, q, proof, want = _prove_both(oracle, top, fname, [args], shard_size=ss)
    assert np.array_equal(proof.words, want)
    assert om.verify(proof.words), om.last_error
    if ss < 1 << 22:   # sharded: the sliced two-phase path (preimg / depth gathers cross shard boundaries)
        q.finalize()
        cfg = ShardingConfig(ss)
        shards = Shard(q, 0, cfg).shard(cfg)
        roots = np.stack([m.commit_shard(s) for s in shards])
        parts = [m.prove_shard(s, roots) for s in shards]
        m.release_record()
        assert np.array_equal(assemble_proof(parts).words, want)


@pytest.mark.parametrize("fname,args", [("add", "add"), ("sub", "sub"), ("less_than", "lt"), ("fibu", [200]), ("maxu", "max"),
                                        ("is_zero", "z"), ("mul", "mul"), ("divrem", "div"), ("factu", "fact"), ("gcdu", "gcd"),
                                        ("big_num_less_than", [5, 7, 0, 0, 0, 0, 0, 2013265920] + [6, 7, 0, 0, 0, 0, 0, 2013265920]),
                                        ("big_num_less_than", [4] * 16)])
def test_u64_chipset_programs_prove_bit_exact(oracle, fname, args):
    # U64::{Add, Sub, LessThan} witnesses on the GPU + their AIR in the quotient kernel + byte lookups
    from test_oracle_stark import u64_bytes
    a = {"add": u64_bytes(2**64 - 1) + u64_bytes(2), "sub": u64_bytes(0) + u64_bytes(1),
         "lt": u64_bytes(0x0102030405060708) + u64_bytes(0x0102030505060708),
         "max": u64_bytes(9) + u64_bytes(10**15), "z": u64_bytes(0), "mul": u64_bytes(2**64 - 1) + u64_bytes(2**64 - 1),
         "div": u64_bytes(10**18 + 7) + u64_bytes(12345), "fact": u64_bytes(25),
         "gcd": u64_bytes(2**40 * 3**5 * 7) + u64_bytes(2**20 * 3**9 * 11)}.get(args, args) if isinstance(args, str) else args
    top = goldens.toplevel_for("u64")
    m, om, q, proof, want = _prove_both(oracle, top, fname, [a])
    assert np.array_equal(proof.words, want)
    assert om.verify(proof.words), om.last_error
    m.verify(m.setup()[1], proof)


def test_cons_chain_2pow12_verifies(oracle):
    # Lurk-shaped load: the 655-wide hash4 chip dominates (2^12 x 655 main trace); acceptance by the verifier
    top = goldens.toplevel_for("hash")
    opts = SphinxCoreOpts(num_queries=10, proof_of_work_bits=8)
    m = new_machine(top, "cons_chain", opts=opts)
    pk, vk = m.setup()
    q = QueryRecord.new(top)
    execute_by_name(top, "cons_chain", [(1 << 12) - 1], q)
    proof = m.prove(pk, Shard.new(q))
    ot = oracle.OracleToplevel(top.serialize())
    om = oracle.OracleMachine(ot, m.entry.index, m.blob, 10, 8)
    assert om.verify(proof.words), om.last_error
    m.verify(vk, proof)


def test_multi_shard_proof_is_bit_exact(oracle):
    # the reference tests sharding with tiny shard sizes (src/air/debug.rs:119-158); ragged last shard
    top = goldens.toplevel_for("ackermann")
    m, om, q, proof, want = _prove_both(oracle, top, "ackermann", [[2, 30]], shard_size=64)
    assert proof.num_shards() == 3
    assert np.array_equal(proof.words, want)
    assert om.verify(proof.words), om.last_error
    # two-phase (shard-per-GPU) API gives the same bytes: whole-record uploads first, then (finalized record)
    # per-shard slices of the record tables
    shards = Shard(q, 0, ShardingConfig(64)).shard(ShardingConfig(64))
    for finalize in (False, True):
        if finalize:
            q.finalize()
        roots = np.stack([m.commit_shard(s) for s in shards])
        parts = [m.prove_shard(s, roots) for s in shards]
        h2d = m.last_transfer_bytes()[0]
        m.release_record()
        assert np.array_equal(assemble_proof(parts).words, want)
    assert np.array_equal(m.prove(pk_of(m), Shard(q, 0, ShardingConfig(64))).words, want)   # finalized + whole upload


def test_large_fib_proof_verifies(oracle):
    # 2^16 rows: too slow to re-prove on the CPU inside the test budget -> acceptance by the oracle verifier
    top = demo_toplevel()
    opts = SphinxCoreOpts(num_queries=20, proof_of_work_bits=12)
    m = new_machine(top, "fib", opts=opts)
    pk, vk = m.setup()
    q = QueryRecord.new(top)
    execute_by_name(top, "fib", [(1 << 16) - 1], q)
    proof = m.prove(pk, Shard.new(q))
    ot = oracle.OracleToplevel(top.serialize())
    om = oracle.OracleMachine(ot, m.entry.index, m.blob, 20, 12)
    assert om.verify(proof.words), om.last_error
    # machine.verify (the product's own host verifier, csrc/verify.cu) agrees, also on a tampered proof
    from lurk_b200.machine import VerificationError
    m.verify(vk, proof)
    bad = proof.words.copy()
    bad[len(bad) // 2] = (int(bad[len(bad) // 2]) + 1) % P
    with pytest.raises(VerificationError):
        m.verify(vk, bad)


def test_full_size_fib_2pow20_default_parameters_verifies():
    # BASELINE configs[1] at full size with sphinx's default FRI parameters (100 queries, 16 PoW bits): too big
    # for the CPU prover inside a test, so acceptance is by the product verifier (independent host code), which
    # must also reject a flipped word anywhere in the 1.1 MB proof
    from lurk_b200.machine import VerificationError, Verifier
    top = demo_toplevel()
    m = new_machine(top, "fib")
    pk, vk = m.setup()
    q = QueryRecord.new(top)
    out = execute_by_name(top, "fib", [(1 << 20) - 1], q)
    proof = m.prove(pk, Shard.new(q))
    assert list(q.expect_public_values()) == [(1 << 20) - 1, int(out[0])]
    m.verify(vk, proof)
    Verifier(top, "fib", vk).verify(proof)          # the device-free form agrees
    rng = np.random.default_rng(5)
    for pos in rng.integers(2, len(proof.words), 6):
        bad = proof.words.copy()
        bad[pos] = (int(bad[pos]) + 1) % P
        with pytest.raises(VerificationError):
            m.verify(vk, bad)
    # proving is deterministic (smallest PoW witness): a second prove gives the same bytes
    assert np.array_equal(m.prove(pk, Shard.new(q)).words, proof.words)


def test_default_shard_size_sharding_2pow22():
    # ShardingConfig::default() = 2^22 rows (src/lair/execute.rs:226-241): 2^22 + 5 rows -> 2 shards, ragged tail;
    # exercises the 3-pass (4,3,3) middle LDE kernel at t = 10 and the sliced record upload
    top = demo_toplevel()
    opts = SphinxCoreOpts(num_queries=12, proof_of_work_bits=10)
    m = new_machine(top, "fib", opts=opts)
    pk, vk = m.setup()
    q = QueryRecord.new(top)
    execute_by_name(top, "fib", [(1 << 22) + 4], q)
    q.finalize()
    shards = Shard.new(q).shard(ShardingConfig.default())
    assert len(shards) == 2
    whole = m.prove(pk, Shard.new(q))
    assert whole.num_shards() == 2
    m.verify(vk, whole)
    roots = np.stack([m.commit_shard(s) for s in shards])
    parts = [m.prove_shard(s, roots) for s in shards]
    m.release_record()
    assert np.array_equal(assemble_proof(parts).words, whole.words)


def test_lde_2pow22_is_linear_and_extends_constants():
    rng = np.random.default_rng(22)
    h = 1 << 22
    a = rng.integers(0, P, size=(h, 1), dtype=np.uint32)
    b = rng.integers(0, P, size=(h, 1), dtype=np.uint32)
    s = ((a.astype(np.uint64) + b) % P).astype(np.uint32)
    la, lb, ls = gpu_coset_lde(a, 0), gpu_coset_lde(b, 0), gpu_coset_lde(s, 0)
    assert np.array_equal(((la.astype(np.uint64) + lb) % P).astype(np.uint32), ls)
    assert (gpu_coset_lde(np.full((h, 1), 7, dtype=np.uint32), 0) == 7).all()
    # the LDE restricted to the original domain's coset is NOT the input, but extending x -> x (degree 1) must give
    # the coset points themselves: row r of the output is g * w_{2N}^bitrev(r)
    w = pow(0x1A427A41, 1 << (27 - 22), P)
    xs = np.empty(h, dtype=np.uint64)
    acc = 1
    for i in range(h):      # natural-order domain H_N
        xs[i] = acc
        acc = acc * w % P
    lx = gpu_coset_lde(xs.astype(np.uint32).reshape(h, 1), 0)[:, 0]
    w2 = pow(0x1A427A41, 1 << (27 - 23), P)
    for r in (0, 1, 2, 12345, 2 * h - 1):
        br = int(format(r, "023b")[::-1], 2)
        assert int(lx[r]) == 31 * pow(w2, br, P) % P


def test_column_sharded_commit_tool_single_rank():
    # SURVEY 8e.2 pieces (lsk_lde_columns_dev, all-to-all, lsk_merkle_subtree_root_dev, lsk_merkle_join_roots)
    # through the real NCCL code path with a one-rank group; the tool asserts root == lsk_merkle_root.
    # (G = 2, 4, 8 are run with `gpurun --gpus G` and recorded in profiles/r01_column_sharded_commit.jsonl.)
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29577", RANK="0", LOCAL_RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "column_sharded_commit.py"), "--log-rows", "13",
                          "--width", "24", "--steps", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 1 and line["value"] > 0


def test_cpp_end_to_end():
    # tests/cpp/fib_prove.cpp: demo_toplevel built with the C++ builder, executed, proved on the GPU and verified
    # through the C ABI only -- the shape of src/lair/lair_chip.rs:246-276 with no Python in the loop
    import subprocess
    exe = N.build_cpp_tests()[1]
    out = subprocess.run([exe, "5000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "verified" in out.stdout, out.stderr + out.stdout


def test_lair_shard_test_ackermann_3_18():
    # the reference's heaviest prover test, `lair_shard_test` (src/lair/trace.rs:654-727, #[ignore]d there):
    # ackermann(3, 18) has 5 242 889 queries -> two shards at the default 2^22 shard size; prove, verify
    top = goldens.toplevel_for("ackermann")
    m = new_machine(top, "ackermann")          # sphinx default FRI parameters
    pk, vk = m.setup()
    q = QueryRecord.new(top)
    out = execute_by_name(top, "ackermann", [3, 18], q)
    assert int(out[0]) == 2097149                 # 2^(18+3) - 3
    assert q.num_queries(0) == 5242889            # trace.rs:685-687
    shard = Shard.new(q)
    assert len(shard.shard(ShardingConfig.default())) == 2
    proof = m.prove(pk, shard)
    assert proof.num_shards() == 2
    m.verify(vk, proof)


def _powers(w, n):
    """w^0 .. w^(n-1) mod P, vectorised by doubling"""
    out = np.ones(n, dtype=np.uint64)
    k, wk = 1, w % P
    while k < n:
        m = min(k, n - k)
        out[k:k + m] = out[:m] * np.uint64(wk) % np.uint64(P)
        wk = wk * wk % P
        k *= 2
    return out


@pytest.mark.parametrize("logh", [13, 21, 23, 24])
def test_lde_of_identity_polynomial_at_every_tile_shape(logh):
    # the trace x -> x on H_N extends to the coset points themselves: output row r = g * w_{2N}^bitrev(r).  Checks
    # every row, at the sizes that select each middle-pass tile shape of the LDE (t = 1, 9, 11, 12: c = 4, 4, 3, 2)
    # up to the largest supported height 2^24.
    h = 1 << logh
    xs = _powers(pow(0x1A427A41, 1 << (27 - logh), P), h).astype(np.uint32).reshape(h, 1)
    got = gpu_coset_lde(xs, 0)[:, 0]
    nat = _powers(pow(0x1A427A41, 1 << (27 - logh - 1), P), 2 * h) * np.uint64(31) % np.uint64(P)   # g * w_2N^i
    r = np.arange(2 * h, dtype=np.uint64)
    br = np.zeros(2 * h, dtype=np.uint64)
    for b in range(logh + 1):
        br |= ((r >> np.uint64(b)) & np.uint64(1)) << np.uint64(logh - b)
    assert np.array_equal(got, nat[br].astype(np.uint32))


def test_device_group_of_one_changes_nothing():
    # lsk_ctx_comm_init with a group of one (no NCCL involved): the column-sharded code path degenerates to the
    # single-GPU prover
    from lurk_b200 import _native as N
    from lurk_b200.execute import QueryRecord, Shard, execute_by_name
    from lurk_b200.lair import demo_toplevel
    from lurk_b200.machine import new_machine
    ctx = N.Context(0)
    top = demo_toplevel()
    m = new_machine(top, "fib", ctx=ctx)
    pk, vk = m.setup()
    q = QueryRecord.new(top)
    execute_by_name(top, "fib", [(1 << 16) - 1], q)
    alone = m.prove(pk, Shard.new(q)).words
    ctx.comm_init(0, 1, bytes(128))
    assert ctx.comm_info()["size"] == 1
    together = m.prove(pk, Shard.new(q))
    assert np.array_equal(alone, together.words)
    m.verify(vk, together)
    ctx.comm_destroy()


def test_column_sharded_prove_two_gpus():
    # two ranks prove each shard together; tools/column_sharded_prove.py asserts that the group's proof
    # equals the single-GPU proof on every rank (fib one shard, fib multi-shard, the hash4 chip) and verifies it.
    # (G = 2, 4, 8 timings: profiles/r01_column_sharded_prove.jsonl)
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(root, "tools", "column_sharded_prove.py"), "--log-rows", "17", "--steps", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert all(line["parity_vs_single_gpu_proof"].values())


def test_compiled_constraint_programs_match_the_interpreter(monkeypatch):
    # csrc/jit.cu: the hasher chips' constraint programs compiled by NVRTC (cubins cached in lurk_b200/jit_cache by
    # `python -m lurk_b200.jit_warm`, part of build()).  The same record proved with the interpreter and with the
    # compiled programs must give the same words; the verifier accepts them.
    from lurk_b200.programs import hash_toplevel
    top = hash_toplevel()
    q = QueryRecord.new(top)
    execute_by_name(top, "cons_chain", [3000], q)
    monkeypatch.setenv("LSK_JIT", "0")
    m0 = new_machine(top, "cons_chain")
    assert N.lib().lsk_machine_jit_chips(m0.h) == 0
    interpreted = m0.prove(m0.setup()[0], Shard.new(q))
    monkeypatch.delenv("LSK_JIT")
    m1 = new_machine(top, "cons_chain")
    if N.lib().lsk_machine_jit_chips(m1.h) == 0:      # cold cache (no build() step ran): compile now
        monkeypatch.setenv("LSK_JIT", "1")
        m1 = new_machine(top, "cons_chain")
    assert N.lib().lsk_machine_jit_chips(m1.h) >= 3   # hash3, hash4, hash5 (+ the short programs above 16 instructions)
    compiled = m1.prove(m1.setup()[0], Shard.new(q))
    assert np.array_equal(interpreted.words, compiled.words)
    m1.verify(m1.setup()[1], compiled)


def test_side_stream_does_not_change_the_proof(monkeypatch):
    # prover.cu / trace_kernels.cu: the record tables travel on the context's second stream in waves (trace kernels wait on
    # per-func events), and the permutation-row / quotient launches of chips with a handful of rows run there beside the
    # tall chips.  Same record, second stream on and off (LSK_NO_SIDE_STREAM=1: everything on one stream): the same words,
    # repeatedly (a race would show up as a differing or rejected proof).
    from lurk_b200.programs import hash_toplevel
    top = hash_toplevel()
    q = QueryRecord.new(top)
    execute_by_name(top, "cons_chain", [700], q)          # hash3/4/5 and helper chips of a few rows next to a tall chip
    m = new_machine(top, "cons_chain")
    pk, vk = m.setup()
    monkeypatch.setenv("LSK_NO_SIDE_STREAM", "1")
    one_stream = m.prove(pk, Shard.new(q))
    monkeypatch.delenv("LSK_NO_SIDE_STREAM")
    for _ in range(4):
        two_streams = m.prove(pk, Shard.new(q))
        assert np.array_equal(one_stream.words, two_streams.words)
    m.verify(vk, two_streams)


@pytest.mark.parametrize("a,b", [(10**18 + 7, 12345), (123456789, 0)])
def test_lurk_misc_u64_funcs_prove_bit_exact(oracle, a, b):
    # src/core/misc.rs:32-128 (pointer-taking u64 / big-num / digest_equal funcs; widths pinned by the reference,
    # tests/golden/known_answers.json) driven by u64_demo: executor, GPU witness (memory chips, byte lookups,
    # extern chips), AIR, whole proof equal to the oracle's
    from test_oracle_stark import _u64_demo_expected, u64_bytes
    top = goldens.toplevel_for("lurk_misc")
    m, om, q, proof, want = _prove_both(oracle, top, "u64_demo", [u64_bytes(a) + u64_bytes(b)])
    assert list(q.expect_public_values())[16:] == _u64_demo_expected(a, b)
    assert np.array_equal(proof.words, want)
    assert om.verify(proof.words), om.last_error
    m.verify(m.setup()[1], proof)


@pytest.mark.parametrize("x_digest", [list(range(1, 9)), [5] * 8])
def test_lurk_env_lookup_proves_bit_exact(oracle, x_digest):
    # env_lookup of the Lurk evaluator (eval_direct.rs:1937-1956) through the whole GPU path, proof equal to the oracle's
    top = goldens.toplevel_for("lurk_eval_subset")
    m, om, q, proof, want = _prove_both(oracle, top, "env_demo", [[10] + x_digest])
    assert np.array_equal(proof.words, want)
    assert om.verify(proof.words), om.last_error
    m.verify(m.setup()[1], proof)


@pytest.mark.parametrize("which", ["list", "fun"])
def test_lurk_ingress_egress_proves_bit_exact(oracle, which):
    # src/core/ingress.rs restated (widths pinned: 104 / 81): content-addressed data interned by the ZStore, its
    # preimages injected (inject_inv_queries), ingress + egress executed, the whole thing proved on the GPU -- hash4 /
    # hash5 chips with preimage lookups, memory chips of four lengths, byte range checks -- equal to the oracle's proof
    from lurk_b200.zstore import ZStore
    from test_zstore import _samples
    top = goldens.toplevel_for("lurk_data")
    z = ZStore()
    x = _samples(z)[which]
    args = [x.tag] + [0] * 7 + list(x.digest)
    opts = SphinxCoreOpts(num_queries=6, proof_of_work_bits=6)
    m = new_machine(top, "data_roundtrip", opts=opts)
    pk, vk = m.setup()
    q = QueryRecord.new(top)
    z.inject(q, top)
    out = execute_by_name(top, "data_roundtrip", args, q)
    assert [int(v) for v in out] == [x.tag] + list(x.digest)
    ot = oracle.OracleToplevel(top.serialize())
    om = oracle.OracleMachine(ot, m.entry.index, m.blob, 6, 6)
    rec = oracle.OracleRecord(ot)
    for name, width in (("hash3", 24), ("hash4", 32), ("hash5", 40)):
        if z.hashes[width]:
            rec.inject_inv_queries(top.func_by_name(name).index, [list(k) for k in z.hashes[width]],
                                   [list(v) for v in z.hashes[width].values()])
    rec.execute(m.entry.index, args, m.entry.output_size)
    proof = m.prove(pk, Shard.new(q))
    assert np.array_equal(proof.words, om.prove(rec))
    m.verify(vk, proof)


def test_fib_2pow20_proof_is_bit_exact_default_parameters(oracle):
    # BASELINE configs[1] on the Lair fib at FULL size and sphinx's default FRI parameters: the GPU proof is word for
    # word the CPU oracle's proof (the oracle needs ~6 s on 16 cores for it)
    m, om, q, proof, want = _prove_both(oracle, demo_toplevel(), "fib", [[(1 << 20) - 1]], nq=100, pow_bits=16)
    assert len(proof.words) == len(want)
    assert np.array_equal(proof.words, want)


@pytest.mark.parametrize("prog,arg", [("fib", 400), ("sum", 300), ("lcs", 12)])
def test_lurk_bench_programs_prove_bit_exact(oracle, prog, arg):
    # the reference's three benches (benches/fib.rs, sum.rs, lcs.rs) from source text through lurk_main on the 48-chip
    # machine, thousands of `eval` rows, default FRI parameters: GPU proof == oracle proof, product verifier accepts
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import CpuProver, Workload
    wl = Workload("lurk-" + prog)
    cp = CpuProver(wl)
    _, want = cp.prove(cp.record(arg))
    m = new_machine(wl.top, wl.entry)
    pk, vk = m.setup()
    q, _ = wl.record(arg)
    proof = m.prove(pk, Shard.new(q))
    assert np.array_equal(proof.words, want)
    m.verify(vk, proof)


def test_an_imported_record_proves_like_the_executed_one():
    # the drop-in path behind an unchanged evaluator: the record crosses the C ABI as tables (lsk_record_import_*),
    # lsk_record_resolve finds the hints, and the proof is the proof of the natively executed record
    from lurk_b200.bench_programs import source
    from lurk_b200.lurk_eval import build_lurk_toplevel
    from lurk_b200.reader import read
    from lurk_b200.zstore import ZStore
    top = build_lurk_toplevel()
    z = ZStore()
    e = read(z, source("fib", 30))
    q = QueryRecord.new(top)
    z.inject(q, top)
    execute_by_name(top, "lurk_main", [e.tag] + [0] * 7 + list(e.digest) + [0] * 8, q)
    inv = [(name, list(z.hashes[w].items())) for name, w in (("hash3", 24), ("hash4", 32), ("hash5", 40)) if z.hashes[w]]
    q2 = QueryRecord.from_tables(top, q.export_tables(), inv)
    m = new_machine(top, "lurk_main", opts=SphinxCoreOpts(num_queries=6, proof_of_work_bits=6))
    pk, vk = m.setup()
    a = m.prove(pk, Shard.new(q))
    b = m.prove(pk, Shard.new(q2))
    assert np.array_equal(a.words, b.words)
    m.verify(vk, b)
