"""ZStore: content-addressing of Lurk data (src/core/zstore.rs), host side.

A `ZPtr` is (tag, 8-word digest).  Atoms carry their value in the digest (`ZPtr::num / char / u64`,
zstore.rs:112-133); compound data is hashed with the Poseidon2 hashers (zstore.rs:305-333):
  tuple11(a, b)     digest = hash4(a.flatten() ++ b.flatten())                 conses, strings, symbols
  tuple110(a, b, c) digest = hash5(a.flatten() ++ b.flatten() ++ c.digest)     functions, environments
  hide(secret, x)   digest = hash3(secret ++ x.flatten())                      commitments
where flatten() = [tag, 0 x 7] ++ digest.  The store remembers every preimage it hashed (`hashes3/4/5`): the
Lair funcs `ingress` / `open_comm` recover data from digests through `preimg(hash4, ..)`, so the preimages are
injected into the QueryRecord as inverse queries before execution (`QueryRecord::inject_inv_queries`,
src/lair/execute.rs; `inject` below).  Interning one datum hashes on the host, one preimage after another, exactly where
the reference does it (`lsk_poseidon2_hash_host`).

Batch interning: `defer_*` builds the same data as a DAG of pending nodes without hashing anything,
and `resolve` hashes it level by level -- level k holds the nodes whose children are all known after level k - 1, i.e.
one batch of independent preimages per hasher width.  A batch of at least `gpu_min_batch` preimages goes through the
GPU kernel (`lsk_poseidon2_hash`, one launch per level and width), smaller ones through the host hasher: a list is a
chain of dependent conses (every level one node wide -- the GPU's ~25 us per dependent hash loses to 0.7 us on a CPU
core), a wide structure (many sublists, a tree, a batch of independent expressions) has levels thousands wide.  The
digests and the preimage tables are identical to one-by-one interning either way (tests/test_zstore.py).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

from . import _native as N
from typing import Optional, Union

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


class Deferred:
    """A compound datum whose digest is not known yet: `kind` 11 (tuple11), 110 (tuple110) or 3 (hide); children are
    ZPtrs or other Deferred nodes.  `ZStore.resolve` fills in `zptr`."""
    __slots__ = ("tag", "kind", "children", "secret", "level", "zptr")

    def __init__(self, tag: int, kind: int, children, secret=None):
        self.tag, self.kind, self.children, self.secret = tag, kind, tuple(children), secret
        self.level = 1 + max((c.level for c in self.children if isinstance(c, Deferred)), default=0)
        self.zptr: Optional[ZPtr] = None


Node = Union[ZPtr, Deferred]


class ZStore:
    def __init__(self):
        self.hashes: Dict[int, Dict[Tuple[int, ...], Tuple[int, ...]]] = {24: {}, 32: {}, 40: {}}
        self._str: Dict[str, ZPtr] = {}
        self._sym: Dict[Tuple[str, ...], ZPtr] = {}
        self.gpu_batches = 0         # launches of the GPU batch hasher issued by hash_many / resolve

    # ---- hashing (zstore.rs:305-333) ---------------------------------------------------------------------------
    def hash(self, preimg: Sequence[int]) -> Tuple[int, ...]:
        pre = tuple(int(x) for x in preimg)
        table = self.hashes[len(pre)]
        if pre not in table:
            out = np.zeros(8, dtype=np.uint32)
            N.check(N.lib().lsk_poseidon2_hash_host(len(pre), N.vp(N.u32(pre)), ctypes.c_size_t(1), N.vp(out)))
            table[pre] = tuple(int(x) for x in out)
        return table[pre]

    def hash_many(self, width: int, preimgs: np.ndarray, ctx=None, gpu_min_batch: int = 256) -> np.ndarray:
        """Digests of a batch of independent preimages, (n, width) -> (n, 8); remembered like `hash` remembers them.
        n >= gpu_min_batch with a context: one GPU launch (lsk_poseidon2_hash); otherwise the host hasher."""
        pre = np.ascontiguousarray(preimgs, dtype=np.uint32).reshape(-1, width)
        out = np.empty((len(pre), 8), dtype=np.uint32)
        if len(pre) == 0:
            return out
        if ctx is not None and len(pre) >= gpu_min_batch:
            N.check(N.lib().lsk_poseidon2_hash(ctx.h, width, N.vp(pre), ctypes.c_size_t(len(pre)), N.vp(out)))
            self.gpu_batches += 1
        else:
            N.check(N.lib().lsk_poseidon2_hash_host(width, N.vp(pre), ctypes.c_size_t(len(pre)), N.vp(out)))
        table = self.hashes[width]
        for row, dig in zip(pre.tolist(), out.tolist()):
            table[tuple(row)] = tuple(dig)
        return out

    # ---- batch interning (zstore.rs:305-349 level by level) ------------------------------------
    def defer_tuple11(self, tag: int, a: Node, b: Node) -> Deferred: return Deferred(tag, 11, (a, b))
    def defer_tuple110(self, tag: int, a: Node, b: Node, c: Node) -> Deferred: return Deferred(tag, 110, (a, b, c))
    def defer_cons(self, a: Node, b: Node) -> Deferred: return Deferred(TAG["Cons"], 11, (a, b))
    def defer_hide(self, secret: Sequence[int], payload: Node) -> Deferred:
        assert len(secret) == 8
        return Deferred(TAG["Comm"], 3, (payload,), tuple(int(x) for x in secret))

    def defer_list(self, xs: Iterable[Node]) -> Node:
        z: Node = self.nil()
        for x in reversed(list(xs)):
            z = self.defer_cons(x, z)
        return z

    def defer_data(self, obj) -> Node:
        """Nested Python lists / ints / strs / ZPtrs as Lurk data: a list becomes a proper list, an int a u64, a str a
        string (interned at once: a string is a chain)."""
        if isinstance(obj, (ZPtr, Deferred)):
            return obj
        if isinstance(obj, int):
            return self.u64(obj)
        if isinstance(obj, str):
            return self.string(obj)
        return self.defer_list([self.defer_data(x) for x in obj])

    def resolve(self, roots: Sequence[Node], ctx=None, gpu_min_batch: int = 256) -> List[ZPtr]:
        """Hash every pending node under `roots`, one batch per (level, width); -> the roots as ZPtrs."""
        pending: Dict[int, Deferred] = {}
        stack = [r for r in roots if isinstance(r, Deferred)]
        while stack:
            d = stack.pop()
            if d.zptr is not None or id(d) in pending:
                continue
            pending[id(d)] = d
            stack.extend(c for c in d.children if isinstance(c, Deferred))
        by_level: Dict[int, List[Deferred]] = {}
        for d in pending.values():
            by_level.setdefault(d.level, []).append(d)
        flat = lambda c: (c if isinstance(c, ZPtr) else c.zptr)
        for level in sorted(by_level):
            groups: Dict[int, List[Deferred]] = {24: [], 32: [], 40: []}
            for d in by_level[level]:
                groups[{3: 24, 11: 32, 110: 40}[d.kind]].append(d)
            for width, nodes in groups.items():
                if not nodes:
                    continue
                rows = []
                for d in nodes:
                    ch = [flat(c) for c in d.children]
                    if d.kind == 3:
                        rows.append(list(d.secret) + ch[0].flatten())
                    elif d.kind == 11:
                        rows.append(ch[0].flatten() + ch[1].flatten())
                    else:
                        rows.append(ch[0].flatten() + ch[1].flatten() + list(ch[2].digest))
                pre = np.asarray(rows, dtype=np.uint32)
                table = self.hashes[width]
                # what the store already knows is not hashed again (the reference's hashes3/4/5 caches); duplicates
                # inside the batch are hashed once
                uniq: Dict[Tuple[int, ...], int] = {}
                todo = []
                for row in rows:
                    t = tuple(row)
                    if t not in table and t not in uniq:
                        uniq[t] = len(todo)
                        todo.append(row)
                if todo:
                    self.hash_many(width, np.asarray(todo, dtype=np.uint32), ctx, gpu_min_batch)
                for d, row in zip(nodes, rows):
                    d.zptr = ZPtr(d.tag, table[tuple(row)])
        return [flat(r) for r in roots]

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
