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
