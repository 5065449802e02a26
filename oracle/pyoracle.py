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


# ---- Lair executor / traces (oracle/lair.h) -----------------------------------------------------
def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _err():
    lib().orc_last_error.restype = ctypes.c_char_p
    return lib().orc_last_error().decode()


class OracleToplevel:
    def __init__(self, program_words):
        L = lib()
        L.orc_toplevel_new.restype = ctypes.c_void_p
        w = _u32(program_words)
        self.h = L.orc_toplevel_new(_vp(w), ctypes.c_size_t(len(w)))
        if not self.h:
            raise RuntimeError(_err())

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_toplevel_free(ctypes.c_void_p(self.h))

    def width(self, func):
        return lib().orc_func_width(ctypes.c_void_p(self.h), func)


class OracleRecord:
    """QueryRecord (src/lair/execute.rs:67-75) held by the oracle."""

    def __init__(self, top: OracleToplevel):
        L = lib()
        L.orc_record_new.restype = ctypes.c_void_p
        for f in ("orc_num_queries", "orc_num_mem", "orc_public_values", "orc_func_trace_height", "orc_mem_trace_height"):
            getattr(L, f).restype = ctypes.c_size_t
        self.top = top
        self.h = L.orc_record_new(ctypes.c_void_p(top.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_record_free(ctypes.c_void_p(self.h))

    def execute(self, func, args, nout):
        a = _u32(args)
        out = np.zeros(nout, dtype=np.uint32)
        rc = lib().orc_execute(ctypes.c_void_p(self.top.h), ctypes.c_void_p(self.h), func, _vp(a),
                               ctypes.c_size_t(len(a)), _vp(out), ctypes.c_size_t(nout))
        if rc:
            raise RuntimeError(_err())
        return out

    def inject_inv_queries(self, func, inps, outs):
        i, o = _u32(inps), _u32(outs)
        rc = lib().orc_inject_inv_queries(ctypes.c_void_p(self.top.h), ctypes.c_void_p(self.h), func, _vp(i), _vp(o),
                                          ctypes.c_size_t(len(i)))
        if rc:
            raise RuntimeError(_err())

    def set_public_values(self, pv):
        """Negative tests only: claim public values the record did not compute."""
        v = _u32(list(pv))
        lib().orc_record_set_public_values(ctypes.c_void_p(self.h), _vp(v), ctypes.c_size_t(len(v)))

    def num_queries(self, func):
        return lib().orc_num_queries(ctypes.c_void_p(self.h), func)

    def num_mem(self, mem_idx):
        return lib().orc_num_mem(ctypes.c_void_p(self.h), mem_idx)

    def num_shards(self, shard_size):
        return lib().orc_num_shards(ctypes.c_void_p(self.h), shard_size)

    def public_values(self):
        out = np.zeros(64, dtype=np.uint32)
        n = lib().orc_public_values(ctypes.c_void_p(self.h), _vp(out), ctypes.c_size_t(64))
        return out[:n].copy()

    def func_trace(self, func, shard=0, shard_size=1 << 22):
        h = lib().orc_func_trace_height(ctypes.c_void_p(self.h), func, shard, shard_size)
        w = self.top.width(func)
        out = np.zeros((h, w), dtype=np.uint32)
        rc = lib().orc_func_trace(ctypes.c_void_p(self.top.h), ctypes.c_void_p(self.h), func, shard, shard_size, _vp(out))
        if rc:
            raise RuntimeError(_err())
        return out

    def mem_trace(self, mem_idx):
        h = lib().orc_mem_trace_height(ctypes.c_void_p(self.h), mem_idx)
        w = 4 + [2, 3, 4, 5, 6, 8][mem_idx]
        out = np.zeros((h, w), dtype=np.uint32)
        rc = lib().orc_mem_trace(ctypes.c_void_p(self.h), mem_idx, _vp(out))
        if rc:
            raise RuntimeError(_err())
        return out

    def bytes_trace(self, shard=0):
        out = np.zeros((1 << 16, 13), dtype=np.uint32)
        rc = lib().orc_bytes_trace(ctypes.c_void_p(self.h), shard, _vp(out))
        if rc:
            raise RuntimeError(_err())
        return out


def bytes_preprocessed():
    out = np.zeros((1 << 16, 6), dtype=np.uint32)
    lib().orc_bytes_preprocessed(_vp(out))
    return out


# ---- STARK machine (oracle/stark.h, oracle/stark_prove.h) ---------------------------------------
class OracleMachine:
    """CPU restatement of sphinx's StarkMachine for a Lair chip vector (parity unpinned, see stark.h)."""

    def __init__(self, top: OracleToplevel, entry_func: int, machine_blob, num_queries=100, pow_bits=16):
        L = lib()
        L.orc_machine_new.restype = ctypes.c_void_p
        L.orc_stark_last_error.restype = ctypes.c_char_p
        blob = _u32(machine_blob)
        self.top = top
        self.h = L.orc_machine_new(ctypes.c_void_p(top.h), entry_func, _vp(blob), ctypes.c_size_t(len(blob)),
                                   num_queries, pow_bits)
        if not self.h:
            raise RuntimeError(L.orc_stark_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_machine_free(ctypes.c_void_p(self.h))

    def forge_omit_kinds(self, kind_mask):
        """Negative tests only: prove as a dishonest prover that leaves out the chips of these kinds
        (bit 0 func, 1 memory, 2 bytes, 3 entrypoint)."""
        lib().orc_machine_forge_omit_kinds(ctypes.c_void_p(self.h), kind_mask)

    def prep_commit(self):
        out = np.zeros(8, dtype=np.uint32)
        lib().orc_machine_prep_commit(ctypes.c_void_p(self.h), _vp(out))
        return out

    def prove(self, rec: OracleRecord, shard_size=1 << 22) -> np.ndarray:
        L = lib()
        out = ctypes.POINTER(ctypes.c_uint32)()
        n = ctypes.c_size_t()
        rc = L.orc_prove(ctypes.c_void_p(self.h), ctypes.c_void_p(rec.h), shard_size, ctypes.byref(out), ctypes.byref(n))
        if rc:
            raise RuntimeError(L.orc_stark_last_error().decode())
        arr = np.ctypeslib.as_array(out, shape=(n.value,)).copy()
        L.orc_free(out)
        return arr

    def verify(self, proof) -> bool:
        p = _u32(proof)
        rc = lib().orc_verify(ctypes.c_void_p(self.h), _vp(p), ctypes.c_size_t(len(p)))
        self.last_error = lib().orc_stark_last_error().decode() if rc else ""
        return rc == 0


def coset_lde(rows, in_shift=1):
    r = _u32(rows)
    out = np.zeros((2 * r.shape[0], r.shape[1]), dtype=np.uint32)
    rc = lib().orc_coset_lde(_vp(r), ctypes.c_size_t(r.shape[0]), ctypes.c_size_t(r.shape[1]), in_shift, _vp(out))
    assert rc == 0
    return out


def merkle_root(mats):
    ms = [_u32(m) for m in mats]
    n = len(ms)
    ptrs = (ctypes.c_void_p * n)(*[m.ctypes.data for m in ms])
    hs = (ctypes.c_size_t * n)(*[m.shape[0] for m in ms])
    ws = (ctypes.c_size_t * n)(*[m.shape[1] for m in ms])
    out = np.zeros(8, dtype=np.uint32)
    rc = lib().orc_merkle_root(ptrs, hs, ws, ctypes.c_size_t(n), _vp(out))
    assert rc == 0
    return out


# ---- independent AIR / lookup checker (oracle/air_check.h) ----------------------------------------
AIR_FUNC, AIR_MEM, AIR_BYTES, AIR_ENTRY = 0, 1, 2, 3


class AirCheckError(AssertionError):
    pass


class TraceQueries:
    """TraceQueries (src/air/debug.rs:26-116): the messages and memoset accesses of the traces checked so far."""

    def __init__(self, top: OracleToplevel):
        L = lib()
        L.orc_tq_new.restype = ctypes.c_void_p
        L.orc_air_last_error.restype = ctypes.c_char_p
        self.top = top
        self.h = L.orc_tq_new()

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tq_free(ctypes.c_void_p(self.h))

    def check(self, kind, index, main, prep=None, public_values=()):
        """debug_constraints_collecting_queries (debug.rs:161-206): every constraint of every row must hold;
        raises AirCheckError naming the first failing constraint and row."""
        m = _u32(main)
        assert m.ndim == 2
        p = _u32(prep) if prep is not None else None
        pv = _u32(list(public_values))
        rc = lib().orc_air_check(ctypes.c_void_p(self.top.h), ctypes.c_void_p(self.h), kind, index, _vp(m),
                                 ctypes.c_size_t(m.shape[0]), ctypes.c_size_t(m.shape[1]),
                                 _vp(p) if p is not None else None, ctypes.c_size_t(p.shape[1] if p is not None else 0),
                                 _vp(pv), ctypes.c_size_t(len(pv)))
        if rc:
            raise AirCheckError(lib().orc_air_last_error().decode())

    def verify(self):
        """TraceQueries::verify (debug.rs:91-107): sends == receives, every memoset chain closes."""
        if lib().orc_tq_verify(ctypes.c_void_p(self.h)):
            raise AirCheckError(lib().orc_air_last_error().decode())

    def has_memoset_key(self, key) -> bool:
        k = _u32(list(key))
        return bool(lib().orc_tq_has_memoset_key(ctypes.c_void_p(self.h), _vp(k), ctypes.c_size_t(len(k))))

    def counts(self):
        out = (ctypes.c_uint64 * 3)()
        lib().orc_tq_counts(ctypes.c_void_p(self.h), out)
        return {"sends": out[0], "receives": out[1], "memoset_keys": out[2]}


def air_cross_check(top: OracleToplevel, machine_blob, chip: int, nrows=3, seed=1):
    """The captured chip vector ('MACH' blob) against the oracle's direct restatement of Air::eval on random rows:
    -> (constraints, sends, receives) compared; raises AirCheckError on the first difference."""
    L = lib()
    L.orc_air_last_error.restype = ctypes.c_char_p
    blob = _u32(machine_blob)
    counts = (ctypes.c_uint32 * 3)()
    rc = L.orc_air_cross_check(ctypes.c_void_p(top.h), _vp(blob), ctypes.c_size_t(len(blob)), chip, nrows,
                               ctypes.c_uint64(seed), counts)
    if rc:
        raise AirCheckError(L.orc_air_last_error().decode())
    return counts[0], counts[1], counts[2]


# ---- stage-level counterparts of lsk_perm_trace / lsk_quotient / lsk_fri_prove ----------------------------------
def chip_perm_width(machine: OracleMachine, chip: int) -> int:
    return lib().orc_chip_perm_width(ctypes.c_void_p(machine.h), chip)


def perm_trace(machine: OracleMachine, chip: int, main, alpha, beta):
    """sphinx generate_permutation_trace for one chip: -> (h x 4*pw matrix, cumulative sum[4])"""
    m = _u32(main)
    pw = chip_perm_width(machine, chip)
    out = np.zeros((m.shape[0], 4 * pw), dtype=np.uint32)
    cum = np.zeros(4, dtype=np.uint32)
    a, b = _u32(alpha), _u32(beta)
    rc = lib().orc_perm_trace(ctypes.c_void_p(machine.h), chip, _vp(m), ctypes.c_size_t(m.shape[0]), _vp(a), _vp(b), _vp(out), _vp(cum))
    if rc:
        raise RuntimeError(lib().orc_stark_last_error().decode())
    return out, cum


def quotient(machine: OracleMachine, chip: int, main, perm, pv, cum, chal_alpha, chal_beta, alpha):
    """quotient values on g*H_2h in commitment order: (2h x 4)"""
    m, p = _u32(main), _u32(perm)
    out = np.zeros((2 * m.shape[0], 4), dtype=np.uint32)
    args = [_u32(x) for x in (pv if len(pv) else [0], cum, chal_alpha, chal_beta, alpha)]
    rc = lib().orc_quotient(ctypes.c_void_p(machine.h), chip, _vp(m), _vp(p), ctypes.c_size_t(m.shape[0]), *[_vp(x) for x in args], _vp(out))
    if rc:
        raise RuntimeError(lib().orc_stark_last_error().decode())
    return out


def fri_prove(vals, log_n, state, num_queries, pow_bits):
    v, s = _u32(vals), _u32(state)
    out = ctypes.POINTER(ctypes.c_uint32)()
    n = ctypes.c_size_t()
    rc = lib().orc_fri_prove(_vp(v), log_n, _vp(s), num_queries, pow_bits, ctypes.byref(out), ctypes.byref(n))
    if rc:
        raise RuntimeError(lib().orc_stark_last_error().decode())
    arr = np.ctypeslib.as_array(out, shape=(n.value,)).copy()
    lib().orc_free(out)
    return arr
