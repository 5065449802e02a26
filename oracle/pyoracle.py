"""ctypes loader for the CPU oracle (test infrastructure; see oracle/bb.h header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def poseidon2_permute(width, states):
    s = _u32(states).reshape(-1, width)
    out = np.empty_like(s)
    rc = lib().orc_poseidon2_permute(width, s.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(s)),
                                     out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def poseidon2_hash(width, preimages):
    s = _u32(preimages).reshape(-1, width)
    out = np.empty((len(s), 8), dtype=np.uint32)
    rc = lib().orc_poseidon2_hash(width, s.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(s)),
                                  out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def poseidon2_witness(width, preimages):
    s = _u32(preimages).reshape(-1, width)
    wc = lib().orc_poseidon2_witness_cols(width)
    out = np.empty((len(s), wc), dtype=np.uint32)
    rc = lib().orc_poseidon2_witness(width, s.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(s)),
                                     out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def sponge16_hash_rows(rows):
    r = _u32(rows)
    assert r.ndim == 2
    out = np.empty((len(r), 8), dtype=np.uint32)
    rc = lib().orc_sponge16_hash_rows(r.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(r.shape[0]),
                                      ctypes.c_size_t(r.shape[1]), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out
