"""Host-side mirror of the reference's Poseidon2 hashing interface for content-addressed data.

  PoseidonChipset<C, W>::hash            src/core/poseidon.rs:29-39
  lurk_hasher / hash3 / hash4 / hash5    src/core/chipset.rs:173-182, src/core/zstore.rs:305-333
  populate_witness                       src/core/poseidon.rs:65-72

All functions run on the GPU through liblurkstark; inputs/outputs are canonical u32 numpy arrays.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as N


class PoseidonChipset:
    def __init__(self, width: int, ctx: N.Context | None = None):
        assert width in (16, 24, 32, 40)
        self.width = width
        self.ctx = ctx or N.default_context()

    def input_size(self):
        return self.width

    def output_size(self):
        return 8

    def witness_size(self):
        return N.lib().lsk_poseidon2_witness_cols(self.width)

    def hash(self, preimages) -> np.ndarray:
        """Batch `hash`: (n, W) -> (n, 8)."""
        x = N.u32(preimages).reshape(-1, self.width)
        out = np.empty((len(x), 8), dtype=np.uint32)
        N.check(N.lib().lsk_poseidon2_hash(self.ctx.h, self.width, N.vp(x), ctypes.c_size_t(len(x)), N.vp(out)))
        return out

    def permute(self, states) -> np.ndarray:
        x = N.u32(states).reshape(-1, self.width)
        out = np.empty_like(x)
        N.check(N.lib().lsk_poseidon2_permute(self.ctx.h, self.width, N.vp(x), ctypes.c_size_t(len(x)), N.vp(out)))
        return out

    def populate_witness(self, preimages) -> np.ndarray:
        x = N.u32(preimages).reshape(-1, self.width)
        out = np.empty((len(x), self.witness_size()), dtype=np.uint32)
        N.check(N.lib().lsk_poseidon2_witness(self.ctx.h, self.width, N.vp(x), ctypes.c_size_t(len(x)), N.vp(out)))
        return out

    def hash_dev(self, in_ptr: int, n: int, out_ptr: int):
        """Device-pointer variant (inputs already resident in HBM); asynchronous on ctx.stream."""
        N.check(N.lib().lsk_poseidon2_hash_dev(self.ctx.h, self.width, ctypes.c_void_p(in_ptr), ctypes.c_size_t(n),
                                               ctypes.c_void_p(out_ptr)))


def constants(width: int):
    """(diag[W], full_rc[8][W], part_rc[R_P]) as canonical residues -- PoseidonConfig::{matrix_diag,
    external_constants, internal_constants}, src/poseidon/config.rs:20-75.  Host-only (no GPU needed)."""
    diag = np.empty(width, dtype=np.uint32)
    full = np.empty((8, width), dtype=np.uint32)
    part = np.empty(64, dtype=np.uint32)
    rp = N.lib().lsk_poseidon2_constants(int(width), N.vp(diag), N.vp(full), N.vp(part), ctypes.c_size_t(64))
    if rp < 0:
        N.check(rp)
    return diag, full, part[:rp].copy()


def hash3(preimages, ctx=None):
    return PoseidonChipset(24, ctx).hash(preimages)


def hash4(preimages, ctx=None):
    return PoseidonChipset(32, ctx).hash(preimages)


def hash5(preimages, ctx=None):
    return PoseidonChipset(40, ctx).hash(preimages)
