"""GPU parity: Poseidon2 batch hashing through the C ABI vs the oracle and the reference KATs."""
import numpy as np

import goldens
import pytest

pytestmark = pytest.mark.gpu
P = 2013265921


def splitmix_inputs(n, w, seed=1):
    # in[i][j] = splitmix64(seed, i*W + j) mod p
    idx = np.arange(n * w, dtype=np.uint64)
    z = idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z % np.uint64(P)).astype(np.uint32).reshape(n, w)


def big(d):
    n = 0
    for limb in reversed([int(x) for x in d]):
        n = n * P + limb
    return n


def test_kats_on_gpu():
    from lurk_b200.poseidon import hash3, hash4
    d = hash3([[0] * 8 + [1] + [0] * 7 + [123] + [0] * 7])[0]      # src/core/zstore.rs:1008-1019
    assert big(d) == int(goldens.KNOWN_ANSWERS["KAT-1"]["digest"], 16)
    h4 = hash4([[0] * 8 + [13] + [0] * 7 + [0] * 8 + [17] + [0] * 7])[0]   # cli/tests/prove.lurk:12-17
    d = hash3([[0] * 8 + [11] + [0] * 7 + list(h4)])[0]
    assert big(d) == int(goldens.KNOWN_ANSWERS["KAT-2"]["digest"], 16)


@pytest.mark.parametrize("w", [16, 24, 32, 40])
@pytest.mark.parametrize("n", [0, 1, 127, 128, 129, 4099])
def test_hash_matches_oracle(oracle, w, n):
    from lurk_b200.poseidon import PoseidonChipset
    x = splitmix_inputs(n, w)
    if n:
        x[0] = np.arange(w)                     # src/poseidon/wide/mod.rs:98 input [0..W)
        x[-1] = P - 1                           # maximum canonical value
    got = PoseidonChipset(w).hash(x)
    assert np.array_equal(got, oracle.poseidon2_hash(w, x).reshape(-1, 8))
    gotp = PoseidonChipset(w).permute(x)
    assert np.array_equal(gotp, oracle.poseidon2_permute(w, x).reshape(-1, w))


@pytest.mark.parametrize("w", [24, 32, 40])
def test_wide_witness_matches_oracle(oracle, w):
    from lurk_b200.poseidon import PoseidonChipset
    x = splitmix_inputs(300, w, seed=7)
    got = PoseidonChipset(w).populate_witness(x)
    assert got.shape[1] == {24: 457, 32: 611, 40: 763}[w]
    assert np.array_equal(got, oracle.poseidon2_witness(w, x))


def test_large_batch_checksum(oracle):
    # size-independent property at bench scale: digest of 2^20 hashes, sampled against the oracle
    from lurk_b200.poseidon import PoseidonChipset
    n = 1 << 20
    x = splitmix_inputs(n, 32)
    got = PoseidonChipset(32).hash(x)
    sel = np.arange(0, n, 4099)
    assert np.array_equal(got[sel], oracle.poseidon2_hash(32, x[sel]))
    assert len(np.unique(got.view([('', got.dtype)] * 8))) == n      # digests of distinct preimages are distinct
