"""ZStore: content-addressing of Lurk data (src/core/zstore.rs), host side.

A `ZPtr` is (tag, 8-word digest).  Atoms carry their value in the digest (`ZPtr::num / char / u64`,
zstore.rs:112-133); compound data is hashed with the Poseidon2 hashers (zstore.rs:305-333):
  tuple11(a, b)     digest = hash4(a.flatten() ++ b.flatten())                 conses, strings, symbols
  tuple110(a, b, c) digest = hash5(a.flatten() ++ b.flatten() ++ c.digest)     functions, environments
  hide(secret, x)   digest = hash3(secret ++ x.flatten())                      commitments
where flatten() = [tag, 0 x 7] ++ digest.  The store remembers every preimage it hashed (`hashes3/4/5`): the
Lair funcs `ingress` / `open_comm` recover data from digests through `preimg(hash4, ..)`, so the preimages are
injected into the QueryRecord as inverse queries before execution (`QueryRecord::inject_inv_queries`,
src/lair/execute.rs; `inject` below).  Hashing runs on the host, one preimage after another, exactly where the
reference does it (`lsk_poseidon2_hash_host`); the GPU batch hash (`lsk_poseidon2_hash`) is for bulk data.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

from . import _native as N

P = 2013265921
# src/core/tag.rs:23-39
TAG = dict(U64=0, Num=1, BigNum=2, Comm=3, Char=4, Str=5, Key=6, Fun=7, Builtin=8, Coroutine=9, Sym=10, Cons=11, Env=12,
           Fix=13, Err=14)
LURK_PACKAGE = "lurk"            # src/core/state.rs:265
BUILTIN_PACKAGE = "builtin"


@dataclass(frozen=True)
class ZPtr:
    tag: int
    digest: Tuple[int, ...]      # 8 canonical field elements

    def flatten(self) -> List[int]:                      # zstore.rs:179-184
        return [self.tag] + [0] * 7 + list(self.digest)

    def to_biguint(self) -> int:                         # field_elts_to_biguint, src/core/big_num.rs:101-108
        n = 0
        for limb in reversed(self.digest):
            n = n * P + limb
        return n


def _digest_from_field(f: int) -> Tuple[int, ...]:
    return (f % P,) + (0,) * 7


class ZStore:
    def __init__(self):
        self.hashes: Dict[int, Dict[Tuple[int, ...], Tuple[int, ...]]] = {24: {}, 32: {}, 40: {}}
        self._str: Dict[str, ZPtr] = {}
        self._sym: Dict[Tuple[str, ...], ZPtr] = {}

    # ---- hashing (zstore.rs:305-333) ---------------------------------------------------------------------------
    def hash(self, preimg: Sequence[int]) -> Tuple[int, ...]:
        pre = tuple(int(x) for x in preimg)
        table = self.hashes[len(pre)]
        if pre not in table:
            out = np.zeros(8, dtype=np.uint32)
            N.check(N.lib().lsk_poseidon2_hash_host(len(pre), N.vp(N.u32(pre)), ctypes.c_size_t(1), N.vp(out)))
            table[pre] = tuple(int(x) for x in out)
        return table[pre]

    # ---- atoms ----------------------------------------------------------------------------------------------------
    @staticmethod
    def null(tag: int) -> ZPtr: return ZPtr(tag, (0,) * 8)
    @staticmethod
    def num(f: int) -> ZPtr: return ZPtr(TAG["Num"], _digest_from_field(f))
    @staticmethod
    def u64(u: int) -> ZPtr: return ZPtr(TAG["U64"], tuple((u >> (8 * i)) & 255 for i in range(8)))
    @staticmethod
    def char(c: str) -> ZPtr:
        b = c.encode("utf-8")
        return ZPtr(TAG["Char"], tuple(b) + (0,) * (8 - len(b)))

    # ---- compound data ------------------------------------------------------------------------------------------------
    def tuple11(self, tag: int, a: ZPtr, b: ZPtr) -> ZPtr:                    # zstore.rs:335-341
        return ZPtr(tag, self.hash(a.flatten() + b.flatten()))

    def tuple110(self, tag: int, a: ZPtr, b: ZPtr, c: ZPtr) -> ZPtr:          # zstore.rs:343-349
        return ZPtr(tag, self.hash(a.flatten() + b.flatten() + list(c.digest)))

    def cons(self, a: ZPtr, b: ZPtr) -> ZPtr:
        return self.tuple11(TAG["Cons"], a, b)

    def string(self, s: str) -> ZPtr:                                          # zstore.rs:396-406
        if s not in self._str:
            z = self.null(TAG["Str"])
            for c in reversed(s):
                z = self.tuple11(TAG["Str"], self.char(c), z)
            self._str[s] = z
        return self._str[s]

    def symbol(self, path: Sequence[str], last_tag: int = TAG["Sym"]) -> ZPtr:   # zstore.rs:408-443
        key = tuple(path) + (str(last_tag),)
        if key not in self._sym:
            z = self.null(TAG["Sym"])
            for i, part in enumerate(path):
                z = self.tuple11(last_tag if i + 1 == len(path) else TAG["Sym"], self.string(part), z)
            self._sym[key] = z
        return self._sym[key]

    def lurk_sym(self, name: str) -> ZPtr:                                     # state.rs:230-232
        return self.symbol([LURK_PACKAGE, name])

    def builtin(self, name: str) -> ZPtr:                                      # state.rs:236-238, Tag::Builtin
        return self.symbol([LURK_PACKAGE, BUILTIN_PACKAGE, name], TAG["Builtin"])

    def nil(self) -> ZPtr: return self.lurk_sym("nil")
    def t(self) -> ZPtr: return self.lurk_sym("t")

    def list(self, xs: Iterable[ZPtr]) -> ZPtr:                                 # intern_list: cons chain ending in nil
        z = self.nil()
        for x in reversed(list(xs)):
            z = self.cons(x, z)
        return z

    def hide(self, secret: Sequence[int], payload: ZPtr) -> ZPtr:               # zstore.rs `hide`: Tag::Comm
        assert len(secret) == 8
        return ZPtr(TAG["Comm"], self.hash(list(secret) + payload.flatten()))

    def commit(self, payload: ZPtr) -> ZPtr:
        return self.hide([0] * 8, payload)

    # ---- preimages for the Lair side -----------------------------------------------------------------------------
    def inject(self, record, toplevel, names=("hash3", "hash4", "hash5")) -> None:
        """QueryRecord::inject_inv_queries for every hash this store computed: func hashN gets (preimage, digest)."""
        funcs = {f.name for f in toplevel.funcs}
        for name, width in zip(names, (24, 32, 40)):
            if self.hashes[width] and name in funcs:
                record.inject_inv_queries(name, toplevel, list(self.hashes[width].items()))
