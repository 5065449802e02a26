"""Lair programs used by the reference's own tests, re-expressed with the lurk_b200.lair builder.
Each cites the `func!` it mirrors."""
from lurk_b200.lair import (Toplevel, add, array, assert_eq, assert_ne, call, const, contains, eq, func, if_, load,
                            match, match_many, mul, not_, range_u8, ret, store, sub)


def match_toplevel():
    """src/lair/trace.rs:517-543 (lair_match_trace_test)"""
    f = func("test", ["n", "m"], 1, [
        const("one", 1),
        match("n", [
            (0, [ret("one")]),
            (1, [ret("m")]),
            (2, [mul("res", "m", "m"), ret("res")]),
            (3, [mul("res", "m", "m"), mul("res2", "res", "res"), ret("res2")]),
        ]),
        sub("pred", "n", "one"),
        call(["res"], "test", ["pred", "m"]),
        ret("res"),
    ])
    return Toplevel.new_pure([f])


def inner_match_toplevel():
    """src/lair/trace.rs:580-610 (lair_inner_match_trace_test)"""
    f = func("test", ["n", "m"], 1, [
        const("zero", 0), const("one", 1), const("two", 2), const("three", 3),
        match("n", [
            (0, [match("m", [(0, [ret("zero")]), (1, [ret("one")])], exhaustive=True)]),
            (1, [match("m", [(0, [ret("two")]), (1, [ret("three")])], exhaustive=True)]),
        ], exhaustive=True),
    ])
    return Toplevel.new_pure([f])


def not_eq_toplevel():
    """src/lair/air.rs:627-637 (lair_not_eq_test); note the order [eq, not]"""
    not_f = func("not", ["a"], 1, [not_("x", "a"), ret("x")])
    eq_f = func("eq", ["a", "b"], 1, [eq("x", "a", "b"), ret("x")])
    return Toplevel.new_pure([eq_f, not_f])


def if_many_toplevel():
    """src/lair/air.rs:716-725"""
    f = func("if_many", [("a", 4)], 1, [
        if_("a", [const("one", 1), ret("one")]),
        const("zero", 0),
        ret("zero"),
    ])
    return Toplevel.new_pure([f])


def match_many_toplevel():
    """src/lair/air.rs:771-794"""
    f = func("match_many", [("a", 2)], 2, [
        match_many("a", [
            ([0, 0], [array("res", [1, 0]), ret("res")]),
            ([0, 1], [array("res", [1, 1]), ret("res")]),
            ([1, 0], [array("res", [1, 2]), ret("res")]),
            ([1, 1], [array("res", [1, 3]), ret("res")]),
        ]),
        array("fail", [0, 0]),
        ret("fail"),
    ])
    return Toplevel.new_pure([f])


def assert_toplevel():
    """src/lair/air.rs:849-861"""
    f = func("assert", [("a", 4)], 4, [
        array("arr1", [2, 4, 5, 8]),
        array("arr2", [2, 4, 6, 8]),
        assert_ne("a", "arr1"),
        const("two", 2),
        const("four", 4),
        contains("a", "two"),
        contains("a", "four"),
        assert_eq("a", "arr2"),
        ret("a"),
    ])
    return Toplevel.new_pure([f])


def equal_branch_toplevel():
    """src/lair/air.rs:888-898"""
    f = func("test", ["a"], 1, [
        match("a", [((2, 3, 4), [const("one", 1), ret("one")])]),
        ret("a"),
    ])
    return Toplevel.new_pure([f])


def range_toplevel():
    """src/lair/air.rs:943-948"""
    f = func("range_test", [("x", 3)], 0, [range_u8(["x"]), ret()])
    return Toplevel.new_pure([f])


def memory_toplevel():
    """src/lair/memory.rs:133-142"""
    f = func("test", [], 2, [
        const("one", 1), const("two", 2), const("three", 3),
        store("ptr1", ["one", "two", "three"]),
        store("ptr2", ["one", "one", "one"]),
        load(["_x", "y", "_z"], "ptr1"),
        ret("ptr2", "y"),
    ])
    return Toplevel.new_pure([f])


def ackermann_toplevel():
    """src/lair/trace.rs:659-680 (lair_shard_test)"""
    f = func("ackermann", ["m", "n"], 1, [
        const("one", 1),
        match("m", [(0, [add("ret", "n", "one"), ret("ret")])]),
        sub("m_minus_one", "m", "one"),
        match("n", [(0, [call(["ret"], "ackermann", ["m_minus_one", "one"]), ret("ret")])]),
        sub("n_minus_one", "n", "one"),
        call(["inner"], "ackermann", ["m", "n_minus_one"]),
        call(["ret"], "ackermann", ["m_minus_one", "inner"]),
        ret("ret"),
    ])
    return Toplevel.new_pure([f])


def hash_toplevel():
    """lurk_b200.programs.hash_toplevel (the bench's Lurk-shaped workload lives in the package)."""
    from lurk_b200.programs import hash_toplevel as h
    return h()


def partial_toplevel():
    """Partial funcs (depth provenance, src/lair/provenance.rs; `partial fn` in src/lair/macros.rs:14-47):
    every row carries its call depth as 4 range-checked bytes, every call to a partial func proves
    callee depth < own depth with a LessThanWitness (src/gadgets/unsigned/less_than.rs).
      pfib(n)      partial fib: depth(n) = n - 1, so n > 257 exercises multi-byte depth comparisons
      pdouble(x)   invertible partial;  phalf(y) = preimg pdouble(y)
      pmain(n, x)  calls all of them (a total helper `twice` in between is allowed: partial -> total)"""
    from lurk_b200.lair import preimg
    pfib = func("pfib", ["n"], 1, [
        const("one", 1),
        match("n", [((0, 1), [ret("one")])]),
        sub("n1", "n", "one"),
        call(["a"], "pfib", ["n1"]),
        sub("n2", "n1", "one"),
        call(["b"], "pfib", ["n2"]),
        add("r", "a", "b"),
        ret("r"),
    ], partial=True)
    twice = func("twice", ["x"], 1, [add("y", "x", "x"), ret("y")])
    pdouble = func("pdouble", ["x"], 1, [call(["y"], "twice", ["x"]), ret("y")], invertible=True, partial=True)
    phalf = func("phalf", ["y"], 1, [preimg(["x"], "pdouble", ["y"]), ret("x")], partial=True)
    pmain = func("pmain", ["n", "x"], 2, [
        call(["f"], "pfib", ["n"]),
        call(["d"], "pdouble", ["x"]),
        call(["h"], "phalf", ["d"]),
        ret("f", "h"),
    ], partial=True)
    return Toplevel.new_pure([pmain, pfib, twice, pdouble, phalf])


def u64_toplevel():
    from lurk_b200.programs import u64_toplevel as u
    return u()


def lurk_misc_toplevel():
    from lurk_b200.programs import lurk_misc_toplevel as u
    return u()


# ---- the Lurk evaluator, restated from src/core/eval_direct.rs: now in the package (lurk_b200/lurk_eval.py) -------------
from lurk_b200.lurk_eval import (TAG, ERR, lurk_eval_subset_toplevel, lurk_core_toplevel,   # noqa: E402,F401
                                 build_lurk_toplevel as lurk_full_toplevel)


def lurk_data_toplevel():
    from lurk_b200.programs import lurk_data_toplevel as u
    return u()
