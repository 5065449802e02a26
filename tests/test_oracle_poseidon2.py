"""Pins the CPU oracle's Poseidon2 against the reference's known-answer digests."""
import numpy as np

import goldens

P = 2013265921


def big(digest):
    # field_elts_to_biguint, src/core/big_num.rs:101-108 (little-endian limbs, base p)
    n = 0
    for limb in reversed([int(x) for x in digest]):
        n = n * P + limb
    return n


def test_kat1_hash3(oracle):
    # src/core/zstore.rs:1008-1019: hash3(0^8 || ZPtr::num(123).flatten())
    pre = [0] * 8 + [1] + [0] * 7 + [123] + [0] * 7
    d = oracle.poseidon2_hash(24, [pre])[0]
    assert big(d) == int(goldens.KNOWN_ANSWERS["KAT-1"]["digest"], 16)


def test_kat2_commit_cons(oracle):
    # src/core/cli/tests/prove.lurk:12-17: (commit '(13 . 17)) = hash3(0^8 || [Cons] || hash4(u64 13, u64 17))
    h4 = oracle.poseidon2_hash(32, [[0] * 8 + [13] + [0] * 7 + [0] * 8 + [17] + [0] * 7])[0]
    d = oracle.poseidon2_hash(24, [[0] * 8 + [11] + [0] * 7 + list(h4)])[0]
    assert big(d) == int(goldens.KNOWN_ANSWERS["KAT-2"]["digest"], 16)


def test_witness_consistent_with_permutation(oracle):
    # src/poseidon/wide/mod.rs:94-118 (test_trace_eq_hash): populate() output == permute()
    rng = np.random.default_rng(1)
    for w, rp in ((24, 21), (32, 30), (40, 38)):
        x = rng.integers(0, P, size=(5, w), dtype=np.uint32)
        x[0] = np.arange(w)
        full = oracle.poseidon2_permute(w, x)
        wit = oracle.poseidon2_witness(w, x)
        assert wit.shape[1] == 8 + 17 * w + 2 * rp - 1      # 457 / 611 / 763
        assert np.array_equal(wit[:, :8], full[:, :8])
        assert np.array_equal(oracle.poseidon2_hash(w, x), full[:, :8])


def test_sponge16_matches_manual_absorb(oracle):
    rng = np.random.default_rng(2)
    rows = rng.integers(0, P, size=(3, 18), dtype=np.uint32)
    got = oracle.sponge16_hash_rows(rows)
    for r, g in zip(rows, got):
        st = np.zeros(16, dtype=np.uint32)
        for off in range(0, 18, 8):
            chunk = r[off:off + 8]
            st[:len(chunk)] = chunk
            st = oracle.poseidon2_permute(16, [st])[0]
        assert np.array_equal(st[:8], g)
