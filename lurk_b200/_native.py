"""ctypes binding of liblurkstark.so (the C ABI in include/lurk_stark.h).

There is no CPU fallback: if the library is missing, or no CUDA device is usable when a context
is created, this module raises.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "liblurkstark.so")
HEADER = os.path.join(os.path.dirname(_PKG), "include", "lurk_stark.h")
_lib = None


class LurkStarkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liblurkstark error {code}: {msg}")
        self.code = code


def build(jobs: int = 8):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", f"-j{jobs}", "-C", os.path.join(_PKG, "csrc")])


def build_cpp_tests():
    """Compile the C++ programs under tests/cpp against the in-tree library (host builder API + C ABI)."""
    root = os.path.dirname(_PKG)
    outs = []
    for name in ("host_layout_test", "fib_prove"):
        src = os.path.join(root, "tests", "cpp", name + ".cpp")
        out = os.path.join(root, "tests", "cpp", name)
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(LIB_PATH)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(_PKG, "csrc"), "-I",
                                   os.path.join(root, "include"), src, "-o", out, "-L", _PKG, "-llurkstark",
                                   "-Wl,-rpath," + _PKG])
        outs.append(out)
    return outs


def declared_symbols():
    """Every `lsk_*` function declared in include/lurk_stark.h."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lsk_[a-z0-9_]+)\s*\(", src)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(liblurkstark has no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.lsk_last_error.restype = ctypes.c_char_p
        L.lsk_ctx_stream.restype = ctypes.c_void_p
        for f in ("lsk_record_num_queries", "lsk_record_num_mem", "lsk_record_public_values",
                  "lsk_func_trace_height", "lsk_mem_trace_height", "lsk_record_device_bytes",
                  "lsk_machine_last_h2d_bytes", "lsk_machine_last_d2h_bytes"):
            getattr(L, f).restype = ctypes.c_size_t
        L.lsk_record_num_shards.restype = ctypes.c_uint32
        L.lsk_profile_tag_name.restype = ctypes.c_char_p
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise LurkStarkError(rc, lib().lsk_last_error().decode())


def u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Context:
    """lsk_ctx: one CUDA device + stream."""

    def __init__(self, device: int = 0):
        self.h = ctypes.c_void_p()
        check(lib().lsk_ctx_create(int(device), ctypes.byref(self.h)))
        self.device = device

    def close(self):
        if self.h:
            lib().lsk_ctx_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return lib().lsk_ctx_stream(self.h)

    def sync(self):
        check(lib().lsk_ctx_sync(self.h))

    # ---- device group of a column-sharded prove (include/lurk_stark.h, lsk_ctx_comm_*) ----
    def comm_init(self, rank: int, size: int, unique_id: bytes):
        """Collective: join a group of `size` contexts (one per GPU / process) that prove each shard together."""
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        check(lib().lsk_ctx_comm_init(self.h, int(rank), int(size), buf))

    def comm_destroy(self):
        check(lib().lsk_ctx_comm_destroy(self.h))

    def comm_info(self) -> dict:
        r, s, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
        check(lib().lsk_ctx_comm_info(self.h, ctypes.byref(r), ctypes.byref(s), ctypes.byref(b)))
        return {"rank": r.value, "size": s.value, "bytes_received": b.value}

    def profile_enable(self, on: bool = True):
        """Reset the per-kernel counters; with on=True also time every kernel family with CUDA events."""
        check(lib().lsk_profile_enable(self.h, 1 if on else 0))

    def profile_read(self) -> dict:
        """{kernel family: {"ms", "bytes" (algorithmic), "launches"}} since the last profile_enable."""
        cap = 64
        ms = (ctypes.c_double * cap)()
        by = (ctypes.c_double * cap)()
        ln = (ctypes.c_uint64 * cap)()
        n = lib().lsk_profile_read(self.h, ms, by, ln, cap)
        if n < 0:
            check(n)
        wk = (ctypes.c_double * cap)()
        lib().lsk_profile_read_work(self.h, wk, cap)
        return {lib().lsk_profile_tag_name(i).decode(): {"ms": ms[i], "bytes": by[i], "launches": int(ln[i]), "work": wk[i]}
                for i in range(n) if ln[i]}


def comm_unique_id() -> bytes:
    """The 128-byte id rank 0 creates and hands to every member of a device group (any transport)."""
    buf = (ctypes.c_uint8 * 128)()
    check(lib().lsk_comm_unique_id(buf))
    return bytes(buf)


def comm_partition(n: int, rank: int, size: int) -> range:
    """The share of n columns / rows / Merkle nodes that `rank` of `size` computes (host only)."""
    b, e = ctypes.c_uint64(), ctypes.c_uint64()
    check(lib().lsk_comm_partition(ctypes.c_uint64(n), int(rank), int(size), ctypes.byref(b), ctypes.byref(e)))
    return range(b.value, e.value)


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx
