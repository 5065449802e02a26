"""GPU parity of the prover's stages, one at a time, through the C ABI: lsk_perm_trace (A8),
lsk_quotient (A6 + A10) and lsk_fri_prove (A12) against the oracle's restatement of the same stage on the same inputs --
so that a mismatch between two whole proofs can be localised."""
import ctypes

import numpy as np
import pytest

import goldens
from lurk_b200 import _native as N
from lurk_b200.execute import BytesChip, MemChip, QueryRecord, Shard, execute_by_name
from lurk_b200.lair import FuncChip
from lurk_b200.machine import SphinxCoreOpts, new_machine

pytestmark = pytest.mark.gpu
P = 2013265921


def _ext(rng):
    return rng.integers(1, P, size=4, dtype=np.uint32)


def _machines(oracle, top, entry):
    m = new_machine(top, entry, opts=SphinxCoreOpts(6, 6))
    ot = oracle.OracleToplevel(top.serialize())
    om = oracle.OracleMachine(ot, m.entry.index, m.blob, 6, 6)
    return m, om


def _chip_traces(top, q, m):
    """(chip index in the chip vector, main trace) for every chip with rows: Entrypoint, funcs, memories, bytes"""
    shard = Shard.new(q)
    out = [(0, np.asarray([list(q.expect_public_values())], dtype=np.uint32))]
    for f in top.funcs:
        if q.num_queries(f.index):
            out.append((1 + f.index, FuncChip(f, top).generate_trace(shard)))
    base = 1 + len(top.funcs)
    for i, ln in enumerate([2, 3, 4, 5, 6, 8]):
        out.append((base + i, MemChip(ln).generate_trace(shard)))
    out.append((base + 6, BytesChip().generate_trace(shard)))
    return out


def gpu_perm_trace(m, chip, main, alpha, beta):
    main = N.u32(main)
    pw = N.lib().lsk_machine_chip_perm_width(m.h, chip)
    out = np.zeros((main.shape[0], 4 * pw), dtype=np.uint32)
    cum = np.zeros(4, dtype=np.uint32)
    N.check(N.lib().lsk_perm_trace(m.h, chip, N.vp(main), ctypes.c_size_t(main.shape[0]), N.vp(N.u32(alpha)), N.vp(N.u32(beta)),
                                   N.vp(out), N.vp(cum)))
    return out, cum


def gpu_quotient(m, chip, main, perm, pv, cum, ca, cb, alpha):
    main, perm = N.u32(main), N.u32(perm)
    out = np.zeros((2 * main.shape[0], 4), dtype=np.uint32)
    pv = N.u32(pv if len(pv) else [0])
    N.check(N.lib().lsk_quotient(m.h, chip, N.vp(main), N.vp(perm), ctypes.c_size_t(main.shape[0]), N.vp(pv), N.vp(N.u32(cum)),
                                 N.vp(N.u32(ca)), N.vp(N.u32(cb)), N.vp(N.u32(alpha)), N.vp(out)))
    return out


@pytest.mark.parametrize("tname,entry,args", [("demo", "fib", [300]), ("partial", "pmain", [300, 21]), ("hash", "cons_chain", [9]),
                                               ("u64", "divrem", [255] * 8 + [1, 1] + [0] * 6), ("memory", "test", [])])
def test_perm_trace_and_quotient_match_the_oracle_chip_by_chip(oracle, tname, entry, args):
    top = goldens.toplevel_for(tname)
    m, om = _machines(oracle, top, entry)
    q = QueryRecord.new(top)
    execute_by_name(top, entry, args, q)
    pv = list(q.expect_public_values())
    rng = np.random.default_rng(7)
    ca, cb, alpha = _ext(rng), _ext(rng), _ext(rng)
    total = np.zeros(4, dtype=np.uint64)
    for chip, main in _chip_traces(top, q, m):
        perm, cum = gpu_perm_trace(m, chip, main, ca, cb)
        operm, ocum = oracle.perm_trace(om, chip, main, ca, cb)
        assert np.array_equal(perm, operm), chip
        assert np.array_equal(cum, ocum), chip
        total = (total + cum) % P
        quot = gpu_quotient(m, chip, main, perm, pv, cum, ca, cb, alpha)
        assert np.array_equal(quot, oracle.quotient(om, chip, main, perm, pv, cum, ca, cb, alpha)), chip
    assert not total.any()          # the LogUp argument: the cumulative sums of all chips cancel


def test_quotient_of_a_bad_trace_is_not_a_polynomial_of_the_right_degree(oracle):
    # the quotient stage divides by the vanishing polynomial whatever the trace: for an honest trace the 2h values are
    # the evaluations of a polynomial of degree < 2h - ... that FRI accepts; here only the stage parity matters, also
    # for a trace that violates its constraints (GPU and oracle must still agree word for word)
    top = goldens.toplevel_for("demo")
    m, om = _machines(oracle, top, "fib")
    q = QueryRecord.new(top)
    execute_by_name(top, "fib", [100], q)
    chip = 1 + top.func_by_name("fib").index
    main = FuncChip.from_name("fib", top).generate_trace(Shard.new(q)).copy()
    main[5, 2] = (int(main[5, 2]) + 1) % P
    rng = np.random.default_rng(9)
    ca, cb, alpha = _ext(rng), _ext(rng), _ext(rng)
    perm, cum = gpu_perm_trace(m, chip, main, ca, cb)
    pv = list(q.expect_public_values())
    assert np.array_equal(gpu_quotient(m, chip, main, perm, pv, cum, ca, cb, alpha),
                          oracle.quotient(om, chip, main, perm, pv, cum, ca, cb, alpha))


@pytest.mark.parametrize("log_n,nq,pow_bits", [(3, 4, 4), (9, 8, 8), (10, 8, 6), (11, 10, 8), (16, 20, 12)])
def test_fri_prove_matches_the_oracle(oracle, log_n, nq, pow_bits):
    # a low-degree codeword: evaluations over g*H_n (bit-reversed) of a polynomial of degree < n/2, extension-valued
    rng = np.random.default_rng(log_n)
    n = 1 << log_n
    coeffs = rng.integers(0, P, size=(n // 2, 4), dtype=np.uint32)
    vals = oracle.coset_lde(coeffs_to_evals(oracle, coeffs))
    state = rng.integers(0, P, size=16, dtype=np.uint32)
    want = oracle.fri_prove(vals, log_n, state, nq, pow_bits)
    out = ctypes.POINTER(ctypes.c_uint32)()
    ln = ctypes.c_size_t()
    v = N.u32(vals)
    N.check(N.lib().lsk_fri_prove(N.default_context().h, N.vp(v), log_n, N.vp(N.u32(state)), nq, pow_bits, ctypes.byref(out),
                                  ctypes.byref(ln)))
    got = np.ctypeslib.as_array(out, shape=(ln.value,)).copy()
    N.lib().lsk_proof_free(out)
    assert np.array_equal(got, want)
    # a vector that is not a codeword is refused (the folded vector does not end constant)
    bad = v.copy()
    bad[3, 1] = (int(bad[3, 1]) + 1) % P
    rc = N.lib().lsk_fri_prove(N.default_context().h, N.vp(bad), log_n, N.vp(N.u32(state)), nq, pow_bits, ctypes.byref(out), ctypes.byref(ln))
    assert rc != 0


def coeffs_to_evals(oracle, coeffs):
    """(n/2 x 4) arbitrary values taken as the evaluations over H_{n/2} of 4 base polynomials of degree < n/2: their coset
    LDE (oracle.coset_lde) is then a codeword of rate 1/2 -- each of the 4 columns, hence the extension-valued vector."""
    return coeffs
