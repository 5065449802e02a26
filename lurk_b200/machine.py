"""Host-side mirror of the reference's STARK machine interface for the proving hot path.

  new_machine(toplevel)                    src/core/stark_machine.rs:20-30
  machine.setup(&LairMachineProgram)       src/core/cli/repl.rs:175-176, benches/fib.rs:142
  machine.prove::<LocalProver>(&pk, shard, &mut challenger, opts)
                                           src/core/cli/repl.rs:194-196, benches/fib.rs:154-158,
                                           src/lair/lair_chip.rs:258-275, tests/fib.rs:90-94
  MachineProof / ShardProof                src/core/cli/proofs.rs:5-10,61-76

Same names and argument meaning; the prover behind `prove` is liblurkstark (CUDA, no CPU
fallback).  A proof is the canonical-u32 word image documented in DESIGN.md.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _native as N
from .air import ChipDesc, build_chip_vector, serialize_machine
from .execute import DEFAULT_SHARD_SIZE, Shard, ShardingConfig, program_of
from .lair import DEPTH_W, Func, FuncChip, Toplevel

PROOF_MAGIC = 0x4C505246


@dataclass
class SphinxCoreOpts:
    """FRI parameters the reference takes from SphinxCoreOpts::default() / the FRI_QUERIES env."""
    num_queries: int = 100
    proof_of_work_bits: int = 16

    @classmethod
    def default(cls) -> "SphinxCoreOpts":
        return cls()


@dataclass
class MachineProof:
    """MachineProof<BabyBearPoseidon2> as a flat word image: MAGIC, nshards, ShardProof*."""
    words: np.ndarray

    def num_shards(self) -> int:
        return int(self.words[1])

    def to_bytes(self) -> bytes:
        return self.words.astype("<u4").tobytes()


@dataclass
class ProvingKey:
    commit: np.ndarray     # preprocessed commitment (byte table)


@dataclass
class VerifyingKey:
    commit: np.ndarray
    pc_start: int = 0      # LairMachineProgram::pc_start, src/lair/lair_chip.rs:71-77


class VerificationError(Exception):
    """machine.verify(...) returned Err: the reason is the message."""


class Verifier:
    """`machine.verify` without a machine or a GPU (the reference verifies on the CPU): the chip vector
    of `entry`, the preprocessed commitment from `setup`, and the FRI parameters.  Host code in
    liblurkstark (csrc/verify.cu); never touches a device."""

    def __init__(self, toplevel: Toplevel, entry: str, vk: VerifyingKey, opts: Optional[SphinxCoreOpts] = None,
                 blob: Optional[np.ndarray] = None):
        opts = opts or SphinxCoreOpts.default()
        if blob is None:
            from . import host
            blob = host.chip_vector_blob(toplevel.surface_image(), toplevel.func_by_name(entry).index)
        self.blob = N.u32(blob)
        self.h = ctypes.c_void_p()
        commit = N.u32(vk.commit)
        N.check(N.lib().lsk_vk_new(N.vp(self.blob), ctypes.c_size_t(len(self.blob)), N.vp(commit), 1,
                                   opts.num_queries, opts.proof_of_work_bits, ctypes.byref(self.h)))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                N.lib().lsk_vk_free(self.h)
                self.h = None
        except Exception:
            pass

    def verify(self, proof) -> None:
        words = N.u32(proof.words if isinstance(proof, MachineProof) else proof)
        rc = N.lib().lsk_vk_verify(self.h, N.vp(words), ctypes.c_size_t(len(words)))
        if rc == -4:
            raise VerificationError(N.lib().lsk_last_error().decode())
        N.check(rc)


class LairMachineProgram:
    pass


class StarkMachine:
    """sphinx_core::stark::StarkMachine<BabyBearPoseidon2, LairChip> for one Toplevel + entry func."""

    def __init__(self, chips: Optional[List[ChipDesc]], num_pv: int, toplevel: Toplevel, entry: Func,
                 ctx: Optional[N.Context] = None, opts: Optional[SphinxCoreOpts] = None,
                 blob: Optional[np.ndarray] = None):
        self._chips = chips
        self.num_pv = num_pv
        self.toplevel = toplevel
        self.entry = entry
        self.ctx = ctx or N.default_context()
        self.opts = opts or SphinxCoreOpts.default()
        self.program = program_of(toplevel)
        self.blob = blob if blob is not None else serialize_machine(chips, num_pv)
        self.h = ctypes.c_void_p()
        N.check(N.lib().lsk_machine_new(self.ctx.h, self.program.h, entry.index, N.vp(self.blob),
                                        ctypes.c_size_t(len(self.blob)), self.opts.num_queries,
                                        self.opts.proof_of_work_bits, ctypes.byref(self.h)))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                N.lib().lsk_machine_free(self.h)
                self.h = None
        except Exception:
            pass

    def chips(self) -> List[ChipDesc]:
        if self._chips is None:     # introspection only: the Python restatement of the capture
            self._chips = build_chip_vector(FuncChip(self.entry, self.toplevel))
        return self._chips

    def setup(self, program: Optional[LairMachineProgram] = None):
        c = np.zeros(8, dtype=np.uint32)
        N.check(N.lib().lsk_machine_prep_commit(self.h, N.vp(c)))
        return ProvingKey(c), VerifyingKey(c.copy())

    # -- whole proof on this GPU ---------------------------------------------------------------
    def prove(self, pk: ProvingKey, shard: Shard, challenger=None, opts=None) -> MachineProof:
        """`shard` is Shard::new(&record): the whole record; it is cut into shards of
        shard.shard_config.max_shard_size rows (reference default 2^22, env SHARD_SIZE)."""
        out = ctypes.POINTER(ctypes.c_uint32)()
        n = ctypes.c_size_t()
        N.check(N.lib().lsk_prove(self.h, shard.queries.h, shard.shard_config.max_shard_size, ctypes.byref(out),
                                  ctypes.byref(n)))
        words = np.ctypeslib.as_array(out, shape=(n.value,)).copy()
        N.lib().lsk_proof_free(out)
        return MachineProof(words)

    def verify(self, vk: VerifyingKey, proof: MachineProof, challenger=None) -> None:
        """machine.verify(&vk, &proof, &mut challenger) (src/core/cli/repl.rs:186,198): raises
        VerificationError when the proof is rejected."""
        if not np.array_equal(N.u32(vk.commit), self.setup()[1].commit):
            raise VerificationError("verifying key does not belong to this machine")
        words = N.u32(proof.words if isinstance(proof, MachineProof) else proof)
        rc = N.lib().lsk_verify(self.h, N.vp(words), ctypes.c_size_t(len(words)))
        if rc == -4:
            raise VerificationError(N.lib().lsk_last_error().decode())
        N.check(rc)

    def last_transfer_bytes(self):
        return N.lib().lsk_machine_last_h2d_bytes(self.h), N.lib().lsk_machine_last_d2h_bytes(self.h)

    # -- shard-per-GPU form (one process per GPU; see parallel.py) ---------------------------------
    def commit_shard(self, shard: Shard) -> np.ndarray:
        root = np.zeros(8, dtype=np.uint32)
        N.check(N.lib().lsk_shard_commit(self.h, shard.queries.h, shard.index, shard.shard_config.max_shard_size,
                                         N.vp(root)))
        return root

    def prove_shard(self, shard: Shard, all_roots: np.ndarray) -> np.ndarray:
        roots = N.u32(all_roots).reshape(-1, 8)
        out = ctypes.POINTER(ctypes.c_uint32)()
        n = ctypes.c_size_t()
        N.check(N.lib().lsk_shard_prove(self.h, shard.queries.h, shard.index, N.vp(roots), len(roots),
                                        ctypes.byref(out), ctypes.byref(n)))
        words = np.ctypeslib.as_array(out, shape=(n.value,)).copy()
        N.lib().lsk_proof_free(out)
        return words

    def release_record(self):
        N.lib().lsk_release_record(self.h)


def new_machine(toplevel: Toplevel, entry: Optional[str] = None, ctx: Optional[N.Context] = None,
                opts: Optional[SphinxCoreOpts] = None) -> StarkMachine:
    """new_machine (src/core/stark_machine.rs:20-30): chip vector of `entry` (reference: lurk_main;
    here any func of the toplevel, e.g. "fib" of demo_toplevel as in src/lair/lair_chip.rs:246-276)."""
    if entry is None:
        entry = "lurk_main"
    from . import host
    f = toplevel.func_by_name(entry)
    num_pv = f.input_size + f.output_size + (DEPTH_W if f.partial else 0)
    # Chip::new for the whole chip vector, in the C++ host layer (csrc/host_air.cpp)
    blob = host.chip_vector_blob(toplevel.surface_image(), f.index)
    return StarkMachine(None, num_pv, toplevel, f, ctx, opts, blob=blob)


def assemble_proof(shard_words: List[np.ndarray]) -> MachineProof:
    head = np.asarray([PROOF_MAGIC, len(shard_words)], dtype=np.uint32)
    return MachineProof(np.concatenate([head] + [N.u32(w) for w in shard_words]))
