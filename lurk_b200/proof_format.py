"""The proof as the reference's types: liblurkstark's word image <-> `MachineProof` / `CryptoProof`.

    src/core/cli/proofs.rs:22-35    CryptoShardProof { commitment, opened_values, opening_proof, chip_ordering }
                                    CryptoProof { shard_proofs, verifier_version, depth }
    src/core/cli/proofs.rs:46-77    into_machine_proof: public values = expr.flatten() ++ env.digest ++ result.flatten()
                                    ++ depth.to_le_bytes(), the same vector for every shard
    src/core/cli/proofs.rs:87-131   From<MachineProof>: strips the public values, keeps the depth (their last 4 bytes)

`decode_machine_proof` turns the canonical-u32 word image of lsk_prove (layout in DESIGN.md section 4) into a nested
dict with exactly the fields of sphinx's `ShardProof` (sphinx-core 1.0.0 @ 8a39b95 -- absent here, field names as in
the SP1 v1 lineage it derives from):

    ShardProof { commitment: ShardCommitment { main_commit, permutation_commit, quotient_commit },
                 opened_values: ShardOpenedValues { chips: [ChipOpenedValues { preprocessed: {local, next},
                     main: {local, next}, permutation: {local, next}, quotient: [[Challenge; 4]; 2],
                     cumulative_sum, log_degree }] },
                 opening_proof: TwoAdicFriPcsProof { fri_proof: FriProof { commit_phase_commits, query_proofs:
                     [QueryProof { commit_phase_openings: [CommitPhaseProofStep { sibling_value, opening_proof }] }],
                     final_poly, pow_witness }, query_openings: [[BatchOpening { opened_values, opening_proof }]] },
                 chip_ordering: HashMap<String, usize>, public_values }

`encode_machine_proof` is the inverse (word for word), `to_crypto_proof` / `from_crypto_proof` are proofs.rs:87-131 and
:46-77, and `bincode_crypto_proof` writes the CryptoProof the way serde + bincode 1 (fixint, little-endian: u64
lengths, u32 words) would.  NOT byte-pinned: the reference holds no proof bytes, `HashMap` iteration order
is unspecified in Rust (two runs of the reference itself need not produce the same bytes), and whether BabyBear
serialises as its canonical or Montgomery word is decided in Plonky3 sources that are not in this tree; canonical is
written here.  What IS pinned: the structure, the chip_ordering map (chip name -> position, byte chip named "CPU",
src/lair/lair_chip.rs:90-92) and the public-value recomposition.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Sequence

import numpy as np

PROOF_MAGIC = 0x4C505246
DEPTH_W = 4                      # src/lair/provenance.rs:11


class _Reader:
    def __init__(self, words):
        self.w = [int(x) for x in words]
        self.p = 0

    def u(self) -> int:
        v = self.w[self.p]
        self.p += 1
        return v

    def take(self, n: int) -> List[int]:
        v = self.w[self.p:self.p + n]
        if len(v) != n:
            raise ValueError("proof truncated")
        self.p += n
        return v

    def ext_vec(self) -> List[List[int]]:
        n = self.u()
        return [self.take(4) for _ in range(n)]

    def path(self) -> List[List[int]]:
        n = self.u()
        return [self.take(8) for _ in range(n)]


def decode_machine_proof(words: Sequence[int], chip_names: Sequence[str]) -> dict:
    """words: lsk_prove's output; chip_names: the machine's chips in chip-vector order (MachineAir::name)."""
    r = _Reader(words)
    if r.u() != PROOF_MAGIC:
        raise ValueError("not a liblurkstark proof")
    shard_proofs = []
    for _ in range(r.u()):
        index = r.u()
        commitment = {"main_commit": r.take(8), "permutation_commit": r.take(8), "quotient_commit": r.take(8)}
        chips, ordering = [], {}
        has_prep = False
        for pos in range(r.u()):
            chip_id, log_degree = r.u(), r.u()
            cum = r.take(4)
            prep = {"local": r.ext_vec(), "next": r.ext_vec()}
            main = {"local": r.ext_vec(), "next": r.ext_vec()}
            perm = {"local": r.ext_vec(), "next": r.ext_vec()}
            quotient = [r.ext_vec(), r.ext_vec()]
            has_prep |= bool(prep["local"])
            ordering[chip_names[chip_id]] = pos
            chips.append({"preprocessed": prep, "main": main, "permutation": perm, "quotient": quotient,
                          "cumulative_sum": cum, "log_degree": log_degree, "_chip_id": chip_id})
        public_values = r.take(r.u())
        commit_phase_commits = [r.take(8) for _ in range(r.u())]
        final_poly = r.take(4)
        pow_witness = r.u()
        query_proofs, query_openings = [], []
        for _ in range(r.u()):
            rounds = []
            for _ in range(r.u()):
                opened = [r.take(r.u()) for _ in range(r.u())]
                rounds.append({"opened_values": opened, "opening_proof": r.path()})
            steps = [{"sibling_value": r.take(4), "opening_proof": r.path()} for _ in range(r.u())]
            query_openings.append(rounds)
            query_proofs.append({"commit_phase_openings": steps})
        shard_proofs.append({
            "commitment": commitment, "opened_values": {"chips": chips},
            "opening_proof": {"fri_proof": {"commit_phase_commits": commit_phase_commits, "query_proofs": query_proofs,
                                            "final_poly": final_poly, "pow_witness": pow_witness},
                              "query_openings": query_openings},
            "chip_ordering": ordering, "public_values": public_values, "_shard_index": index})
    if r.p != len(r.w):
        raise ValueError("trailing words after the last shard proof")
    return {"shard_proofs": shard_proofs}


def encode_machine_proof(proof: dict) -> np.ndarray:
    """The inverse of decode_machine_proof, word for word."""
    w: List[int] = [PROOF_MAGIC, len(proof["shard_proofs"])]

    def ext_vec(v):
        w.append(len(v))
        for e in v:
            w.extend(e)

    def path(p):
        w.append(len(p))
        for d in p:
            w.extend(d)
    for sp in proof["shard_proofs"]:
        w.append(sp["_shard_index"])
        for k in ("main_commit", "permutation_commit", "quotient_commit"):
            w.extend(sp["commitment"][k])
        chips = sp["opened_values"]["chips"]
        w.append(len(chips))
        for c in chips:
            w.extend([c["_chip_id"], c["log_degree"]])
            w.extend(c["cumulative_sum"])
            for part in ("preprocessed", "main", "permutation"):
                ext_vec(c[part]["local"])
                ext_vec(c[part]["next"])
            ext_vec(c["quotient"][0])
            ext_vec(c["quotient"][1])
        w.append(len(sp["public_values"]))
        w.extend(sp["public_values"])
        fri = sp["opening_proof"]["fri_proof"]
        w.append(len(fri["commit_phase_commits"]))
        for d in fri["commit_phase_commits"]:
            w.extend(d)
        w.extend(fri["final_poly"])
        w.append(fri["pow_witness"])
        w.append(len(fri["query_proofs"]))
        for qp, rounds in zip(fri["query_proofs"], sp["opening_proof"]["query_openings"]):
            w.append(len(rounds))
            for bo in rounds:
                w.append(len(bo["opened_values"]))
                for row in bo["opened_values"]:
                    w.append(len(row))
                    w.extend(row)
                path(bo["opening_proof"])
            w.append(len(qp["commit_phase_openings"]))
            for st in qp["commit_phase_openings"]:
                w.extend(st["sibling_value"])
                path(st["opening_proof"])
    return np.asarray(w, dtype=np.uint32)


def to_crypto_proof(proof: dict, verifier_version: str = "") -> dict:
    """From<MachineProof> for CryptoProof (proofs.rs:87-131): every shard must carry the same public values; they are
    dropped and only the depth (their last DEPTH_W values, bytes) is kept."""
    pvs = [sp["public_values"] for sp in proof["shard_proofs"]]
    if not pvs or any(pv != pvs[0] for pv in pvs):
        raise ValueError("all shards must have the same public values")
    depth_bytes = pvs[0][-DEPTH_W:]
    if any(b > 255 for b in depth_bytes):
        raise ValueError("depth limbs must be bytes")
    shards = [{k: sp[k] for k in ("commitment", "opened_values", "opening_proof", "chip_ordering", "_shard_index")}
              for sp in proof["shard_proofs"]]
    return {"shard_proofs": shards, "verifier_version": verifier_version,
            "depth": int.from_bytes(bytes(depth_bytes), "little")}


def from_crypto_proof(cp: dict, expr: Sequence[int], env_digest: Sequence[int], result: Sequence[int]) -> dict:
    """CryptoProof::into_machine_proof (proofs.rs:46-77).  expr / result: ZPtr::flatten() = [tag, 0 x 7, digest x 8]
    (src/core/zstore.rs:184-204); env_digest: 8 words."""
    if len(expr) != 16 or len(env_digest) != 8 or len(result) != 16:
        raise ValueError("expr and result are flattened ZPtrs (16 words), env a digest (8 words)")
    pv = list(expr) + list(env_digest) + list(result) + list(cp["depth"].to_bytes(4, "little"))
    return {"shard_proofs": [dict(sp, public_values=list(pv)) for sp in cp["shard_proofs"]]}


# ---- serde + bincode 1 (fixint, little-endian) -------------------------------------------------------------------
def _u64(n: int) -> bytes:
    return struct.pack("<Q", n)


def _felts(v: Sequence[int]) -> bytes:
    return struct.pack("<%dI" % len(v), *v)


def _vec(items, enc) -> bytes:
    return _u64(len(items)) + b"".join(enc(x) for x in items)


def bincode_crypto_proof(cp: dict) -> bytes:
    """bincode::serialize(&CryptoProof): Vec / String / HashMap lengths as u64, BabyBear as one u32, [T; N] arrays and
    tuples without a length, usize as u64.  See the module docstring for what is and is not pinned."""
    ext = _felts                                            # BinomialExtensionField<BabyBear, 4> = [BabyBear; 4]
    digest = _felts                                         # Hash<BabyBear, BabyBear, 8> = [BabyBear; 8]

    def air_opened(a):
        return _vec(a["local"], ext) + _vec(a["next"], ext)

    def chip(c):
        return (air_opened(c["preprocessed"]) + air_opened(c["main"]) + air_opened(c["permutation"]) +
                _vec(c["quotient"], lambda chunk: _vec(chunk, ext)) + ext(c["cumulative_sum"]) + _u64(c["log_degree"]))

    def shard(sp):
        com = sp["commitment"]
        out = digest(com["main_commit"]) + digest(com["permutation_commit"]) + digest(com["quotient_commit"])
        out += _vec(sp["opened_values"]["chips"], chip)
        fri = sp["opening_proof"]["fri_proof"]
        out += _vec(fri["commit_phase_commits"], digest)
        out += _vec(fri["query_proofs"], lambda qp: _vec(qp["commit_phase_openings"],
                                                         lambda st: ext(st["sibling_value"]) + _vec(st["opening_proof"], digest)))
        out += ext(fri["final_poly"]) + _felts([fri["pow_witness"]])
        out += _vec(sp["opening_proof"]["query_openings"],
                    lambda rounds: _vec(rounds, lambda bo: _vec(bo["opened_values"], lambda row: _vec(row, lambda x: _felts([x]))) +
                                        _vec(bo["opening_proof"], digest)))
        # HashMap<String, usize>: iteration order is unspecified in Rust; written sorted by position here
        items = sorted(sp["chip_ordering"].items(), key=lambda kv: kv[1])
        out += _u64(len(items)) + b"".join(_u64(len(k.encode())) + k.encode() + _u64(v) for k, v in items)
        return out
    ver = cp["verifier_version"].encode()
    return _vec(cp["shard_proofs"], shard) + _u64(len(ver)) + ver + struct.pack("<I", cp["depth"])


def chip_names(toplevel, entry: str) -> List[str]:
    """MachineAir::name of every chip of build_lair_chip_vector (src/lair/lair_chip.rs:79-95,196-212)."""
    from .lair import MEM_TABLE_SIZES
    f = toplevel.func_by_name(entry)
    return (["Entrypoint[%d]" % f.index] + ["Func[%s]" % g.name for g in toplevel.funcs] +
            ["Mem[%d-wide]" % n for n in MEM_TABLE_SIZES] + ["CPU"])
