"""Host-side mirror of the reference's execution record types.

  QueryRecord          src/lair/execute.rs:67-75,259-373   (held natively, already flattened)
  Shard / ShardingConfig   src/lair/execute.rs:77-130,226-241
  Toplevel::execute*   src/lair/execute.rs:375-420
  FuncChip / MemChip / BytesChip ::generate_trace   src/lair/trace.rs:72-135, memory.rs:30-68,
                                                    gadgets/bytes/trace.rs:75-101
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

from . import _native as N
from .lair import MEM_TABLE_SIZES, Func, FuncChip, Toplevel

DEFAULT_SHARD_SIZE = 1 << 22


class Program:
    """Native handle of a serialised Toplevel."""

    def __init__(self, toplevel: Toplevel):
        self.toplevel = toplevel
        words = toplevel.serialize()
        self.h = ctypes.c_void_p()
        N.check(N.lib().lsk_program_load(N.vp(words), ctypes.c_size_t(len(words)), ctypes.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None):
            N.lib().lsk_program_free(self.h)


def program_of(toplevel: Toplevel) -> Program:
    p = getattr(toplevel, "_program", None)
    if p is None:
        p = Program(toplevel)
        toplevel._program = p
    return p


class QueryRecord:
    def __init__(self, toplevel: Toplevel):
        self.toplevel = toplevel
        self.program = program_of(toplevel)
        self.h = ctypes.c_void_p()
        N.check(N.lib().lsk_record_new(self.program.h, ctypes.byref(self.h)))

    @classmethod
    def new(cls, toplevel: Toplevel) -> "QueryRecord":
        return cls(toplevel)

    def __del__(self):
        if getattr(self, "h", None):
            N.lib().lsk_record_free(self.h)

    def num_queries(self, func_index: int) -> int:
        return N.lib().lsk_record_num_queries(self.h, func_index)

    def num_mem(self, mem_idx: int) -> int:
        return N.lib().lsk_record_num_mem(self.h, mem_idx)

    def expect_public_values(self) -> np.ndarray:
        out = np.zeros(64, dtype=np.uint32)
        n = N.lib().lsk_record_public_values(self.h, N.vp(out), ctypes.c_size_t(64))
        return out[:n].copy()

    def inject_inv_queries(self, name: str, toplevel: Toplevel, pairs):
        f = toplevel.func_by_name(name)
        inps = N.u32([p[0] for p in pairs]).reshape(-1, f.input_size)
        outs = N.u32([p[1] for p in pairs]).reshape(-1, f.output_size)
        N.check(N.lib().lsk_inject_inv_queries(self.program.h, self.h, f.index, N.vp(inps), N.vp(outs),
                                               ctypes.c_size_t(len(inps))))

    def finalize(self):
        """Re-pack the per-row require / hint arrays in row order so that shard-per-GPU proving uploads only
        each shard's slice of the record (lsk_record_finalize); call once after the last execute."""
        N.check(N.lib().lsk_record_finalize(self.h))

    def pin(self):
        """Page-lock the record's tables so that prove uploads them at PCIe speed."""
        N.check(N.lib().lsk_record_pin(self.h))

    def device_bytes(self) -> int:
        return N.lib().lsk_record_device_bytes(self.h)

    # -- tables in and out: the form in which a record made by another executor (the reference's own
    #    `toplevel.execute`, src/lair/execute.rs:375-783) crosses the C ABI ---------------------------------------
    def export_tables(self) -> dict:
        """Every table of the record as numpy arrays: {"funcs": {index: {...}}, "mems": {idx: {...}}, "bytes", "public_values"}.
        Row-ordered (finalize() is called)."""
        self.finalize()
        L = N.lib()
        funcs = {}
        for f in self.toplevel.funcs:
            sz = (ctypes.c_size_t * 4)()
            N.check(L.lsk_record_func_sizes(self.h, f.index, sz))
            n, nreq, nhint, ndreq = (int(x) for x in sz)
            if not n:
                continue
            t = {"args": np.zeros((n, f.input_size), np.uint32), "outs": np.zeros((n, f.output_size), np.uint32),
                 "provide": np.zeros((n, 2), np.uint32), "req_off": np.zeros(n + 1, np.uint32),
                 "reqs": np.zeros((nreq, 2), np.uint32), "hint_off": np.zeros(n + 1, np.uint32),
                 "hints": np.zeros(nhint, np.uint32)}
            if f.partial:
                t.update(depth=np.zeros(n, np.uint32), dreq_off=np.zeros(n + 1, np.uint32), dreqs=np.zeros((ndreq, 2), np.uint32))
            p = lambda k: N.vp(t[k]) if k in t and t[k].size else None
            N.check(L.lsk_record_export_func(self.h, f.index, p("args"), p("outs"), p("provide"), p("req_off"), p("reqs"),
                                             p("hint_off"), p("hints"), p("depth"), p("dreq_off"), p("dreqs")))
            funcs[f.index] = t
        mems = {}
        for i, ln in enumerate(MEM_TABLE_SIZES):
            n = self.num_mem(i)
            if n:
                m = {"vals": np.zeros((n, ln), np.uint32), "provide": np.zeros((n, 2), np.uint32)}
                N.check(L.lsk_record_export_mem(self.h, i, N.vp(m["vals"]), N.vp(m["provide"])))
                mems[i] = m
        b = np.zeros((65536, 6, 2), np.uint32)
        N.check(L.lsk_record_export_bytes(self.h, N.vp(b)))
        return {"funcs": funcs, "mems": mems, "bytes": b, "public_values": self.expect_public_values()}

    @classmethod
    def from_tables(cls, toplevel: Toplevel, tables: dict, inv_queries=()) -> "QueryRecord":
        """A record imported table by table (lsk_record_import_* + lsk_record_resolve): what a Rust shim does with the
        reference's QueryRecord.  The "hint" tables are NOT read: resolve recomputes them.  inv_queries: [(func name,
        [(preimage, image), ...])] for preimages that no imported query produced (inject_inv_queries)."""
        q = cls(toplevel)
        L = N.lib()
        for idx, t in tables["funcs"].items():
            g = lambda k: N.vp(N.u32(t[k])) if k in t and np.size(t[k]) else None
            keep = {k: N.u32(v) for k, v in t.items()}      # keep the contiguous copies alive across the call
            p = lambda k: N.vp(keep[k]) if k in keep and keep[k].size else None
            N.check(L.lsk_record_import_func(q.program.h, q.h, idx, ctypes.c_size_t(len(keep["provide"])), p("args"), p("outs"),
                                             p("provide"), N.vp(keep["req_off"]), p("reqs"), p("depth"),
                                             N.vp(keep["dreq_off"]) if "dreq_off" in keep else None, p("dreqs")))
        for i, m in tables["mems"].items():
            v, pr = N.u32(m["vals"]), N.u32(m["provide"])
            N.check(L.lsk_record_import_mem(q.h, i, ctypes.c_size_t(len(pr)), N.vp(v), N.vp(pr)))
        b = N.u32(tables["bytes"])
        N.check(L.lsk_record_import_bytes(q.h, N.vp(b)))
        pv = N.u32(tables["public_values"])
        N.check(L.lsk_record_import_public_values(q.h, N.vp(pv), ctypes.c_size_t(len(pv))))
        for name, pairs in inv_queries:
            q.inject_inv_queries(name, toplevel, pairs)
        N.check(L.lsk_record_resolve(q.program.h, q.h))
        return q


def execute(toplevel: Toplevel, func: Func, args, queries: QueryRecord) -> np.ndarray:
    """Toplevel::execute (src/lair/execute.rs:376-392)"""
    a = N.u32(list(args))
    out = np.zeros(func.output_size, dtype=np.uint32)
    N.check(N.lib().lsk_execute(queries.program.h, queries.h, func.index, N.vp(a), ctypes.c_size_t(len(a)), N.vp(out),
                                ctypes.c_size_t(len(out))))
    return out


def execute_by_name(toplevel: Toplevel, name: str, args, queries: QueryRecord) -> np.ndarray:
    return execute(toplevel, toplevel.func_by_name(name), args, queries)


@dataclass(frozen=True)
class ShardingConfig:
    max_shard_size: int = DEFAULT_SHARD_SIZE

    @classmethod
    def default(cls) -> "ShardingConfig":
        # src/lair/execute.rs:231-241: env SHARD_SIZE, default 1 << 22
        try:
            return cls(int(os.environ["SHARD_SIZE"]))
        except (KeyError, ValueError):
            return cls(DEFAULT_SHARD_SIZE)


class Shard:
    def __init__(self, queries: QueryRecord, index: int = 0, shard_config: ShardingConfig | None = None):
        self.queries = queries
        self.index = index
        self.shard_config = shard_config or ShardingConfig.default()

    @classmethod
    def new(cls, queries: QueryRecord) -> "Shard":
        return cls(queries)

    def get_func_range(self, func_index: int) -> range:
        n = self.queries.num_queries(func_index)
        s = self.shard_config.max_shard_size
        return range(self.index * s, min((self.index + 1) * s, n))

    def shard(self, config: ShardingConfig) -> list:
        n = N.lib().lsk_record_num_shards(self.queries.h, config.max_shard_size)
        return [Shard(self.queries, i, config) for i in range(n)]

    def expect_public_values(self):
        return self.queries.expect_public_values()


def func_generate_trace(chip: FuncChip, shard: Shard, ctx: N.Context | None = None) -> np.ndarray:
    """FuncChip::generate_trace -> RowMajorMatrix values, shape (height, width)."""
    ctx = ctx or N.default_context()
    q = shard.queries
    ss = shard.shard_config.max_shard_size
    h = N.lib().lsk_func_trace_height(q.h, chip.func.index, shard.index, ss)
    out = np.zeros((h, chip.width()), dtype=np.uint32)
    N.check(N.lib().lsk_func_trace(ctx.h, q.program.h, q.h, chip.func.index, shard.index, ss, N.vp(out)))
    return out


FuncChip.generate_trace = func_generate_trace  # same spelling as the reference


class MemChip:
    """src/lair/memory.rs:15-68"""

    def __init__(self, length: int):
        self.len = length

    def width(self) -> int:
        return 4 + self.len

    def generate_trace(self, shard: Shard, ctx: N.Context | None = None) -> np.ndarray:
        ctx = ctx or N.default_context()
        idx = MEM_TABLE_SIZES.index(self.len)
        h = N.lib().lsk_mem_trace_height(shard.queries.h, idx)
        out = np.zeros((h, self.width()), dtype=np.uint32)
        N.check(N.lib().lsk_mem_trace(ctx.h, shard.queries.h, idx, N.vp(out)))
        return out


class BytesChip:
    """src/gadgets/bytes/trace.rs:16-101"""

    def width(self) -> int:
        return 13

    def preprocessed_width(self) -> int:
        return 6

    def generate_trace(self, shard: Shard, ctx: N.Context | None = None) -> np.ndarray:
        ctx = ctx or N.default_context()
        out = np.zeros((1 << 16, 13), dtype=np.uint32)
        N.check(N.lib().lsk_bytes_trace(ctx.h, shard.queries.h, shard.index, N.vp(out)))
        return out

    def generate_preprocessed_trace(self, ctx: N.Context | None = None) -> np.ndarray:
        ctx = ctx or N.default_context()
        out = np.zeros((1 << 16, 6), dtype=np.uint32)
        N.check(N.lib().lsk_bytes_preprocessed(ctx.h, N.vp(out)))
        return out
