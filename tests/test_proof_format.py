"""The proof as the reference's types (lurk_b200/proof_format.py): the word image of lsk_prove decodes into the fields of
sphinx's ShardProof / the reference's CryptoShardProof (src/core/cli/proofs.rs:22-35) and back, word for word; the
public values are stripped and recomposed the way proofs.rs:46-77,87-131 do.  Proofs come from the CPU oracle prover (same
word image as the GPU prover's: tests/test_gpu_prover.py holds the two equal)."""
import numpy as np
import pytest

import goldens
from lurk_b200 import proof_format as PF
from test_oracle_stark import oracle_machine


def _lurk_proof(oracle):
    from lurk_b200.bench_programs import source
    from lurk_b200.lurk_eval import build_lurk_toplevel
    from lurk_b200.reader import read
    from lurk_b200.zstore import ZStore
    from test_lurk_reference_eval_cases import _oracle_main
    top = build_lurk_toplevel()
    z = ZStore()
    e = read(z, source("fib", 6))
    ot, rec, out = _oracle_main(oracle, top, z, e)
    _, om, _ = oracle_machine(oracle, top, "lurk_main")
    return top, e, out, om, om.prove(rec)


def test_word_image_roundtrips_through_the_reference_structure(oracle):
    top = goldens.toplevel_for("ackermann")
    ot, om, f = oracle_machine(oracle, top, "ackermann")
    rec = oracle.OracleRecord(ot)
    rec.execute(f.index, [2, 30], 1)
    words = om.prove(rec, 64)                                  # 3 shards
    proof = PF.decode_machine_proof(words, PF.chip_names(top, "ackermann"))
    assert len(proof["shard_proofs"]) == 3
    assert np.array_equal(PF.encode_machine_proof(proof), words)
    sp = proof["shard_proofs"][0]
    assert set(sp) >= {"commitment", "opened_values", "opening_proof", "chip_ordering", "public_values"}
    assert set(sp["commitment"]) == {"main_commit", "permutation_commit", "quotient_commit"}
    assert set(sp["opening_proof"]) == {"fri_proof", "query_openings"}
    assert set(sp["opening_proof"]["fri_proof"]) == {"commit_phase_commits", "query_proofs", "final_poly", "pow_witness"}
    # chip_ordering: name -> position in the shard (height descending); shard 0 has the Entrypoint, the memories and "CPU"
    names = PF.chip_names(top, "ackermann")
    assert names[0] == "Entrypoint[0]" and names[-1] == "CPU" and "Func[ackermann]" in names
    order = sp["chip_ordering"]
    assert sorted(order.values()) == list(range(len(order)))
    assert {"Entrypoint[0]", "CPU", "Func[ackermann]", "Mem[2-wide]", "Mem[8-wide]"} <= set(order)
    degs = [c["log_degree"] for c in sp["opened_values"]["chips"]]
    assert degs == sorted(degs, reverse=True) and degs[order["CPU"]] == 16 and degs[order["Entrypoint[0]"]] == 0
    assert set(proof["shard_proofs"][1]["chip_ordering"]) == {"Func[ackermann]", "CPU"}
    # a tampered structure no longer verifies once re-encoded
    sp["opened_values"]["chips"][0]["main"]["local"][0][0] ^= 1
    assert not om.verify(PF.encode_machine_proof(proof))


def test_crypto_proof_strips_and_recomposes_the_public_values(oracle):
    top, expr, out, om, words = _lurk_proof(oracle)
    proof = PF.decode_machine_proof(words, PF.chip_names(top, "lurk_main"))
    pv = proof["shard_proofs"][0]["public_values"]
    assert len(pv) == 44                                        # 24 inputs + 16 outputs + 4 depth bytes
    cp = PF.to_crypto_proof(proof, verifier_version="322297f")
    assert "public_values" not in cp["shard_proofs"][0] and cp["depth"] == int.from_bytes(bytes(pv[-4:]), "little") > 0
    # proofs.rs:46-77: expr.flatten() ++ env.digest ++ result.flatten() ++ depth bytes
    expr_flat = [expr.tag] + [0] * 7 + list(expr.digest)
    back = PF.from_crypto_proof(cp, expr_flat, [0] * 8, out)
    assert back["shard_proofs"][0]["public_values"] == pv
    again = PF.encode_machine_proof(back)
    assert np.array_equal(again, words) and om.verify(again)
    wrong = PF.from_crypto_proof(cp, expr_flat, [0] * 8, out[:-1] + [(out[-1] + 1) % 2013265921])
    assert not om.verify(PF.encode_machine_proof(wrong))       # a different claimed result does not verify
    blob = PF.bincode_crypto_proof(cp)
    nfelts = len(words)
    assert 3 * nfelts < len(blob) < 6 * nfelts                 # 4 bytes per field element + u64 length prefixes
    assert blob.endswith(b"322297f" + cp["depth"].to_bytes(4, "little"))


def test_decoder_refuses_garbage():
    with pytest.raises(ValueError):
        PF.decode_machine_proof([1, 2, 3], ["a"])
    with pytest.raises((ValueError, IndexError)):
        PF.decode_machine_proof([PF.PROOF_MAGIC, 1, 0] + [0] * 10, ["a"])
