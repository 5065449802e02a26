"""The reference's benchmark programs (benches/fib.rs, lcs.rs, sum.rs), read from their source text and run through
`lurk_main` of the packaged toplevel; results checked against the same computation in Python."""
import sys

import pytest

from lurk_b200.bench_programs import LCS_ARGS, setup
from lurk_b200.execute import execute_by_name
from lurk_b200.zstore import TAG
from test_lurk_reference_eval_cases import toplevel


def run(name, arg):
    top = toplevel()
    args, q, z = setup(name, arg, top)
    out = [int(v) for v in execute_by_name(top, "lurk_main", args, q)]
    return z, out[0], tuple(out[8:])


def lurk_lcs(a, b):
    # the program's own tie-breaking: lte(x, y) is len(x) <= len(y)
    sys.setrecursionlimit(10000)
    from functools import lru_cache

    @lru_cache(None)
    def f(i, j):
        if i == len(a) or j == len(b):
            return ""
        if a[i] == b[j]:
            return a[i] + f(i + 1, j + 1)
        x, y = f(i, j + 1), f(i + 1, j)
        return y if len(x) <= len(y) else x
    return f(0, 0)


@pytest.mark.parametrize("n,want", [(0, 0), (1, 1), (10, 55), (15, 610)])
def test_fib_bench_program(n, want):
    z, tag, digest = run("fib", n)
    assert (tag, digest) == (TAG["U64"], z.u64(want).digest)


@pytest.mark.parametrize("n", [0, 1, 100])
def test_sum_bench_program(n):
    z, tag, digest = run("sum", n)
    assert (tag, digest) == (TAG["U64"], z.u64(n * (n - 1) // 2).digest)


@pytest.mark.parametrize("n", [0, 3, 7])
def test_lcs_bench_program(n):
    z, tag, digest = run("lcs", n)
    want = z.string(lurk_lcs(LCS_ARGS[0][:n], LCS_ARGS[1][:n]))
    assert (tag, digest) == (want.tag, want.digest)
