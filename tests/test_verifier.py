"""The product verifier (lurk_b200/csrc/verify.cu behind lsk_vk_verify / lsk_verify; machine.verify,
src/core/cli/repl.rs:186,198) against proofs of the CPU oracle prover: it is independent code (the
oracle's own verifier lives in oracle/stark_prove.h), needs no GPU, must accept every honest proof and
reject every tampered one, for every kind of chip the machine has."""
import numpy as np
import pytest

import goldens
from lurk_b200.machine import SphinxCoreOpts, VerificationError, Verifier, VerifyingKey
from test_oracle_stark import kat2_args, oracle_machine

P = 2013265921
OPTS = SphinxCoreOpts(num_queries=6, proof_of_work_bits=6)


def _proof(oracle, tname, fname, args, shard_size=None):
    top = goldens.toplevel_for(tname)
    ot, om, f = oracle_machine(oracle, top, fname)
    rec = oracle.OracleRecord(ot)
    rec.execute(f.index, args, f.output_size)
    proof = om.prove(rec, shard_size) if shard_size else om.prove(rec)
    v = Verifier(top, fname, VerifyingKey(om.prep_commit()), OPTS)
    return om, v, proof


@pytest.mark.parametrize("tname,fname,args,ss", [
    ("demo", "fib", [100], None),
    ("demo", "factorial", [5], None),
    ("range", "range_test", [100, 12, 64], None),
    ("memory", "test", [], None),
    ("match_many", "match_many", [1, 0], None),
    ("ackermann", "ackermann", [2, 30], 64),        # 3 shards
    ("hash", "commit_cons", kat2_args(), None),     # hasher chips
    ("partial", "pmain", [300, 21], None),          # depth provenance
])
def test_verifier_accepts_honest_and_rejects_tampered(oracle, tname, fname, args, ss):
    om, v, proof = _proof(oracle, tname, fname, args, ss)
    v.verify(proof)
    rng = np.random.default_rng(len(proof))
    positions = [2, 12, 30, len(proof) // 3, len(proof) // 2, len(proof) - 5] + list(rng.integers(2, len(proof), 12))
    for pos in positions:
        bad = proof.copy()
        bad[pos] = (int(bad[pos]) + 1) % P
        with pytest.raises(VerificationError):
            v.verify(bad)
        assert not om.verify(bad)       # the two verifiers agree
    with pytest.raises(VerificationError):
        v.verify(proof[:-1])
    with pytest.raises(VerificationError):
        v.verify(np.concatenate([proof, [0]]).astype(np.uint32))


def test_verifier_rejects_wrong_key_and_wrong_parameters(oracle):
    om, v, proof = _proof(oracle, "demo", "fib", [20])
    top = goldens.toplevel_for("demo")
    bad_commit = om.prep_commit().copy()
    bad_commit[0] ^= 1
    with pytest.raises(VerificationError):
        Verifier(top, "fib", VerifyingKey(bad_commit), OPTS).verify(proof)
    with pytest.raises(VerificationError):
        Verifier(top, "fib", VerifyingKey(om.prep_commit()), SphinxCoreOpts(7, 6)).verify(proof)
    with pytest.raises(VerificationError):      # another entry point: different Entrypoint chip
        Verifier(top, "factorial", VerifyingKey(om.prep_commit()), OPTS).verify(proof)


def test_verifier_rejects_swapped_shards(oracle):
    om, v, proof = _proof(oracle, "ackermann", "ackermann", [2, 30], 64)
    assert int(proof[1]) == 3
    fewer = proof.copy()
    fewer[1] = 2
    with pytest.raises(VerificationError):
        v.verify(fewer)


def test_verifier_survives_hostile_input(oracle):
    # the verifier parses untrusted words: truncations, absurd counts / lengths / non-canonical residues at any
    # position, random garbage and garbage behind a valid prefix must all end in VerificationError (never a crash,
    # never acceptance); a mutation that leaves the words unchanged is the only thing accepted
    om, v, proof = _proof(oracle, "demo", "fib", [50])
    v.verify(proof)
    n = len(proof)
    rng = np.random.default_rng(1)

    def rejected(words):
        words = np.asarray(words, dtype=np.uint32)
        if len(words) == n and np.array_equal(words, proof):
            return
        with pytest.raises(VerificationError):
            v.verify(words)

    for k in list(range(40)) + list(rng.integers(0, n, 60)):
        rejected(proof[:k])
    for pos in list(range(120)) + list(rng.integers(0, n, 150)):
        for val in (0xffffffff, 1 << 20, 0, P):
            bad = proof.copy()
            bad[pos] = val
            rejected(bad)
    for _ in range(60):
        rejected(rng.integers(0, 2 ** 32, size=int(rng.integers(0, 3000)), dtype=np.uint64).astype(np.uint32))
    for _ in range(60):
        bad = proof.copy()
        k = int(rng.integers(2, n))
        bad[k:] = rng.integers(0, 2 ** 32, size=n - k, dtype=np.uint64).astype(np.uint32)
        rejected(bad)


@pytest.mark.parametrize("omit", [0b1011, 0b1000, 0b0010, 0b0100], ids=["only-bytes", "no-entrypoint", "no-memory", "no-bytes"])
def test_verifier_rejects_a_proof_without_the_chips_included_demands(oracle, omit):
    # A dishonest prover that leaves out the Entrypoint (the only chip binding the public values, src/lair/lair_chip.rs:
    # 166-191) can claim ANY public values: with an empty record and only the byte table, every LogUp sum is 0 and the
    # quotient / FRI checks pass.  Both verifiers must insist on MachineAir::included (lair_chip.rs:124-139).
    top = goldens.toplevel_for("demo")
    ot, om, f = oracle_machine(oracle, top, "fib")
    rec = oracle.OracleRecord(ot)
    if omit == 0b1011:
        rec.set_public_values([20, 19110])          # fib(20) is 6765
    else:
        rec.execute(f.index, [20], f.output_size)
    om.forge_omit_kinds(omit)
    proof = om.prove(rec)
    om.forge_omit_kinds(0)
    v = Verifier(top, "fib", VerifyingKey(om.prep_commit()), OPTS)
    with pytest.raises(VerificationError, match="lacks the"):
        v.verify(proof)
    assert not om.verify(proof)
    assert "included" in om.last_error


def test_vk_new_validates_fri_parameters(oracle):
    top = goldens.toplevel_for("demo")
    ot, om, f = oracle_machine(oracle, top, "fib")
    from lurk_b200._native import LurkStarkError
    for nq, pb in ((0, 6), (2000, 6), (6, 31), (6, 40)):
        with pytest.raises(LurkStarkError):
            Verifier(top, "fib", VerifyingKey(om.prep_commit()), SphinxCoreOpts(nq, pb))


def test_machine_blob_with_an_out_of_range_interaction_column_is_refused(oracle):
    from lurk_b200 import host
    from lurk_b200._native import LurkStarkError
    top = goldens.toplevel_for("demo")
    f = top.func_by_name("fib")
    blob = np.asarray(host.chip_vector_blob(top.surface_image(), f.index), dtype=np.uint32)
    _, om, _ = oracle_machine(oracle, top, "fib")
    refused = 0
    # an affine term is (is_prep, column, coefficient): find (0, col < width, coeff) triples of the entrypoint's
    # interactions at the tail of its descriptor and push the column out of range
    for i in range(3, len(blob) - 2):
        bad = blob.copy()
        if blob[i] == 0 and blob[i + 1] < 18 and blob[i + 2] == 1:
            bad[i + 1] = 1 << 19
            try:
                Verifier(top, "fib", VerifyingKey(om.prep_commit()), OPTS, blob=bad)
            except LurkStarkError:
                refused += 1
    assert refused > 10
