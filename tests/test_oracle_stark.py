"""CPU oracle of the STARK machine (oracle/stark.h, stark_prove.h): prove -> verify round trips on
the reference's own test programs (the only pin the reference offers below the trace level is
`verify() == Ok`), multi-shard, and rejection of tampered proofs."""
import numpy as np
import pytest

import goldens
from lurk_b200.air import build_chip_vector, serialize_machine
from lurk_b200.lair import DEPTH_W, FuncChip, demo_toplevel

P = 2013265921


def oracle_machine(oracle, top, entry, nq=6, pow_bits=6):
    fc = FuncChip.from_name(entry, top)
    f = fc.func
    num_pv = f.input_size + f.output_size + (DEPTH_W if f.partial else 0)
    blob = serialize_machine(build_chip_vector(fc), num_pv)
    ot = oracle.OracleToplevel(top.serialize())
    return ot, oracle.OracleMachine(ot, f.index, blob, nq, pow_bits), f


@pytest.mark.parametrize("tname,fname,args", [
    ("demo", "fib", [100]),              # benches/fib.rs shape, configs[0] substitute
    ("demo", "factorial", [5]),          # src/lair/lair_chip.rs:246-276
    ("range", "range_test", [100, 12, 64]),
    ("memory", "test", []),
    ("match_many", "match_many", [1, 0]),
    ("assert", "assert", [2, 4, 6, 8]),
])
def test_prove_verify_roundtrip(oracle, tname, fname, args):
    top = goldens.toplevel_for(tname)
    ot, om, f = oracle_machine(oracle, top, fname)
    rec = oracle.OracleRecord(ot)
    rec.execute(f.index, args, f.output_size)
    proof = om.prove(rec)
    assert om.verify(proof), om.last_error
    for pos in (12, len(proof) // 3, len(proof) - 5):
        bad = proof.copy()
        bad[pos] = (int(bad[pos]) + 1) % P
        assert not om.verify(bad)
    assert not om.verify(proof[:-1])


KAT2 = 0x955f855f302a30ed988cc48685c442ebd98c8711e989fc64df8f27f52e1350   # src/core/cli/tests/prove.lurk:12-17


def big_uint(limbs):
    v = 0
    for x in reversed(limbs):     # little-endian base-p limbs, src/core/big_num.rs:101-108
        v = v * P + int(x)
    return v


def kat2_args():
    a = [0] * 8 + [13] + [0] * 7
    b = [0] * 8 + [17] + [0] * 7
    return a + b + [0] * 8 + [11] + [0] * 7


@pytest.mark.parametrize("fname,args", [("commit_cons", None), ("cons_chain", [5]), ("env_chain", [3])])
def test_hasher_chips_prove_verify(oracle, fname, args):
    # Poseidon2Cols::eval (src/poseidon/wide/air.rs:11-124) captured into the constraint program: the proof
    # only verifies if every wide-witness cell satisfies the hasher AIR
    top = goldens.toplevel_for("hash")
    ot, om, f = oracle_machine(oracle, top, fname)
    rec = oracle.OracleRecord(ot)
    out = rec.execute(f.index, args if args is not None else kat2_args(), f.output_size)
    if args is None:
        assert big_uint(out) == KAT2
    proof = om.prove(rec)
    assert om.verify(proof), om.last_error
    bad = proof.copy()
    bad[len(bad) // 2] = (int(bad[len(bad) // 2]) + 1) % P
    assert not om.verify(bad)


def test_hasher_air_shape():
    # 611 hasher constraints + outputs... the chip's constraint count and degree (sphinx: max degree 3)
    top = goldens.toplevel_for("hash")
    chips = {c.name: c for c in build_chip_vector(FuncChip.from_name("commit_cons", top))}
    for name, t, rp in (("Func[hash3]", 24, 21), ("Func[hash4]", 32, 30), ("Func[hash5]", 40, 38)):
        c = chips[name]
        # 16t (external state + sbox) + t (internal init) + (rp - 1) + rp + 8 outputs; + nonce, sel-bool, 8 returns
        assert c.n_constraints == 17 * t + 2 * rp - 1 + 8 + 10
        assert c.max_degree == 3 and c.log_quotient_degree == 1
        assert len(c.sends) == 1 and len(c.receives) == 1


@pytest.mark.parametrize("fname,args", [("pfib", [12]), ("pmain", [300, 21])])
def test_partial_funcs_prove_verify(oracle, fname, args):
    # depth provenance: eval_depth + LessThanWitness::assert_less_than + byte less-than lookups
    # (src/lair/air.rs:105-131,210-224, src/gadgets/unsigned/less_than.rs:49-103); pmain's depth 300 needs
    # a second byte, so the comparison limb moves
    top = goldens.toplevel_for("partial")
    ot, om, f = oracle_machine(oracle, top, fname)
    rec = oracle.OracleRecord(ot)
    rec.execute(f.index, args, f.output_size)
    proof = om.prove(rec)
    assert om.verify(proof), om.last_error
    bad = proof.copy()
    bad[len(bad) // 2] = (int(bad[len(bad) // 2]) + 1) % P
    assert not om.verify(bad)


def u64_bytes(x):
    return [(x >> (8 * i)) & 255 for i in range(8)]


@pytest.mark.parametrize("fname,args,want", [
    ("add", u64_bytes(200) + u64_bytes(56), u64_bytes(256)),                      # u64_add_test, src/core/u64.rs:237-290
    ("add", u64_bytes(2**64 - 1) + u64_bytes(2), u64_bytes(1)),                   # wraps
    ("sub", u64_bytes(256) + u64_bytes(1), u64_bytes(255)),                       # u64_sub_test
    ("sub", u64_bytes(0) + u64_bytes(1), u64_bytes(2**64 - 1)),
    ("less_than", u64_bytes(0x0102030405060708) + u64_bytes(0x0102030505060708), [1]),
    ("less_than", u64_bytes(77) + u64_bytes(77), [0]),                            # equal: the (0, 0) byte lookup
    ("less_than", u64_bytes(2**63) + u64_bytes(5), [0]),
    ("fibu", [90], u64_bytes(2880067194370816120)),
    ("fibu", [100], u64_bytes(354224848179261915075 % 2**64)),                    # wrapping u64 arithmetic
    ("maxu", u64_bytes(9) + u64_bytes(10**15), u64_bytes(10**15)),
    ("is_zero", u64_bytes(0), [1]),
    ("is_zero", u64_bytes(2**40), [0]),
    ("mul", u64_bytes(0xFFFFFFFFFFFFFFFF) + u64_bytes(0xFFFFFFFFFFFFFFFF), u64_bytes(1)),      # wraps
    ("mul", u64_bytes(123456789) + u64_bytes(987654321), u64_bytes(123456789 * 987654321)),
    ("divrem", u64_bytes(10**18 + 7) + u64_bytes(12345), u64_bytes((10**18 + 7) // 12345) + u64_bytes((10**18 + 7) % 12345)),
    ("divrem", u64_bytes(5) + u64_bytes(2**63), u64_bytes(0) + u64_bytes(5)),
    ("divrem", u64_bytes(2**64 - 1) + u64_bytes(1), u64_bytes(2**64 - 1) + u64_bytes(0)),
    ("factu", u64_bytes(25), u64_bytes(15511210043330985984000000 % 2**64)),
    ("big_num_less_than", [5, 7, 0, 0, 0, 0, 0, 9] + [6, 7, 0, 0, 0, 0, 0, 9], [1]),          # differ in the lowest limb
    ("big_num_less_than", [1, 2, 3, 4, 5, 6, 7, 2013265920] + [1, 2, 3, 4, 5, 6, 7, 2013265919], [0]),   # p - 1 vs p - 2: msb = 0x78 path
    ("big_num_less_than", [3] * 8 + [3] * 8, [0]),                                             # equal
    ("big_num_less_than", [0, 0, 0, 0x01000000, 9, 9, 9, 0] + [0, 0, 0, 0x00FFFFFF, 9, 9, 9, 0], [0]),   # byte 3 decides
    ("gcdu", u64_bytes(2**40 * 3**5 * 7) + u64_bytes(2**20 * 3**9 * 11), u64_bytes(2**20 * 3**5)),
])
def test_u64_chipsets_prove_verify(oracle, fname, args, want):
    # U64::{Add, Sub, LessThan} (src/core/u64.rs): execute, witness, AIR (assert_add carries, CompareWitness) and
    # their byte lookups; the proof only verifies if all of them are consistent
    top = goldens.toplevel_for("u64")
    ot, om, f = oracle_machine(oracle, top, fname)
    rec = oracle.OracleRecord(ot)
    out = rec.execute(f.index, args, f.output_size)
    assert list(out) == want
    proof = om.prove(rec)
    assert om.verify(proof), om.last_error


def test_multi_shard_roundtrip(oracle):
    # sharding tested with tiny shards as in src/air/debug.rs:119-158; cross-shard lookups must cancel
    top = goldens.toplevel_for("ackermann")
    ot, om, f = oracle_machine(oracle, top, "ackermann")
    rec = oracle.OracleRecord(ot)
    rec.execute(f.index, [2, 30], 1)
    assert rec.num_shards(64) == 3
    proof = om.prove(rec, 64)
    assert int(proof[1]) == 3
    assert om.verify(proof), om.last_error


def test_proof_is_deterministic(oracle):
    top = demo_toplevel()
    ot, om, f = oracle_machine(oracle, top, "fib")
    proofs = []
    for _ in range(2):
        rec = oracle.OracleRecord(ot)
        rec.execute(f.index, [50], 1)
        proofs.append(om.prove(rec))
    assert np.array_equal(proofs[0], proofs[1])


def test_stage_helpers(oracle):
    # LDE of a constant column is constant; LDE is linear; Merkle root depends on every cell
    rng = np.random.default_rng(0)
    a = rng.integers(0, P, size=(64, 3), dtype=np.uint32)
    b = rng.integers(0, P, size=(64, 3), dtype=np.uint32)
    s = ((a.astype(np.uint64) + b) % P).astype(np.uint32)
    la, lb, ls = oracle.coset_lde(a), oracle.coset_lde(b), oracle.coset_lde(s)
    assert np.array_equal(((la.astype(np.uint64) + lb) % P).astype(np.uint32), ls)
    assert (oracle.coset_lde(np.full((16, 1), 9, dtype=np.uint32)) == 9).all()
    r0 = oracle.merkle_root([la, lb[:32]])
    lb2 = lb.copy()
    lb2[31, 2] ^= 1
    assert not np.array_equal(r0, oracle.merkle_root([la, lb2[:32]]))


def _u64_demo_expected(a, b):
    M = 1 << 64
    q, r = (a // b, a % b) if b else (0, 0)
    pa = sum(x * P ** i for i, x in enumerate(u64_bytes(a)))     # the same 8 limbs read as a base-p integer
    pb = sum(x * P ** i for i, x in enumerate(u64_bytes(b)))
    return (u64_bytes((a + b) % M) + u64_bytes((a - b) % M) + u64_bytes(a * b % M) + u64_bytes(q) + u64_bytes(r)
            + [int(a < b), int(b == 0), int(a == b), int(pa < pb)])


@pytest.mark.parametrize("a,b", [(10**18 + 7, 12345), (5, 2**64 - 1), (77, 77), (123456789, 0)])
def test_lurk_misc_u64_funcs_execute_prove_verify(oracle, a, b):
    # the pointer-taking u64 / big-num / digest_equal funcs of the Lurk toplevel (src/core/misc.rs:32-128) driven the
    # way eval_binop_num drives them (eval_direct.rs:1205-1340): store the operands, call, load the results
    top = goldens.toplevel_for("lurk_misc")
    ot, om, f = oracle_machine(oracle, top, "u64_demo")
    rec = oracle.OracleRecord(ot)
    out = rec.execute(f.index, u64_bytes(a) + u64_bytes(b), f.output_size)
    assert list(out) == _u64_demo_expected(a, b)
    proof = om.prove(rec)
    assert om.verify(proof), om.last_error


@pytest.mark.parametrize("x_tag,x_digest,want", [
    (10, list(range(9, 17)), [1, 9]),          # b: the head entry
    (10, list(range(1, 9)), [1, 7]),           # a: one recursion
    (10, [5] * 8, [14, 0]),                    # unbound: (Tag::Err, EvalErr::UnboundVar)
    (6, list(range(1, 9)), [14, 0]),           # same digest under another tag (Key) is a different variable
])
def test_lurk_env_lookup_executes_proves_verifies(oracle, x_tag, x_digest, want):
    # env_lookup (src/core/eval_direct.rs:1937-1956, width pinned at 52) on an environment laid out in memory the way
    # the evaluator does it: execution result, then a whole proof (memory chips of lengths 5 and 8, recursion)
    top = goldens.toplevel_for("lurk_eval_subset")
    ot, om, f = oracle_machine(oracle, top, "env_demo")
    rec = oracle.OracleRecord(ot)
    out = rec.execute(f.index, [x_tag] + x_digest, f.output_size)
    assert list(out) == want
    proof = om.prove(rec)
    assert om.verify(proof), om.last_error
