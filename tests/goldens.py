"""Golden trace matrices restated as *data* from the reference's own tests: tests/golden/lair_traces.json holds the
matrices with the reference file:line of each; this module loads them and lists the cases."""
import json
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(_DIR, "lair_traces.json")) as _f:
    _TRACES = json.load(_f)
with open(os.path.join(_DIR, "known_answers.json")) as _f:
    KNOWN_ANSWERS = json.load(_f)


def _load(name):
    t = _TRACES[name]
    return np.asarray(t["rows"], dtype=np.uint32).reshape(-1, t["width"])


ASSERT = _load("ASSERT")   # src/lair/air.rs:875-880
EQ = _load("EQ")   # src/lair/air.rs:697-708
EQUAL_BRANCH = _load("EQUAL_BRANCH")   # src/lair/air.rs:923-935
FACTORIAL_5 = _load("FACTORIAL_5")   # src/lair/trace.rs:472-489
FIB_7 = _load("FIB_7")   # src/lair/trace.rs:498-513
IF_MANY = _load("IF_MANY")   # src/lair/air.rs:752-763
INNER_MATCH = _load("INNER_MATCH")   # src/lair/trace.rs:640-651
MATCH_5_2 = _load("MATCH_5_2")   # src/lair/trace.rs:560-574
MATCH_MANY = _load("MATCH_MANY")   # src/lair/air.rs:826-842
MEMORY_FUNC = _load("MEMORY_FUNC")   # src/lair/memory.rs:150-153
MEMORY_MEM3 = _load("MEMORY_MEM3")   # src/lair/memory.rs:163-168
NOT = _load("NOT")   # src/lair/air.rs:662-674
RANGE = _load("RANGE")   # src/lair/air.rs:960-962

# (toplevel factory name, func name, list of argument tuples executed in order, golden)
CASES = [
    ("demo", "factorial", [(5,)], FACTORIAL_5),
    ("demo", "fib", [(7,)], FIB_7),
    ("match", "test", [(5, 2)], MATCH_5_2),
    ("inner_match", "test", [(0, 0), (0, 1), (1, 0), (1, 1)], INNER_MATCH),
    ("not_eq", "not", [(4,), (8,), (0,), (1,)], NOT),
    ("not_eq", "eq", [(4, 2), (4, 4), (0, 3), (0, 0)], EQ),
    ("if_many", "if_many", [(0, 0, 0, 0), (1, 3, 8, 2), (0, 0, 4, 1), (0, 0, 0, 9)], IF_MANY),
    ("match_many", "match_many", [(0, 0), (0, 1), (1, 0), (1, 1), (0, 8)], MATCH_MANY),
    ("assert", "assert", [(2, 4, 6, 8)], ASSERT),
    ("equal_branch", "test", [(1,), (2,), (3,), (4,)], EQUAL_BRANCH),
    ("range", "range_test", [(100, 12, 64)], RANGE),
    ("memory", "test", [()], MEMORY_FUNC),
]


def toplevel_for(name):
    from lurk_b200.lair import demo_toplevel
    import lair_programs as lp
    if name == "demo":
        return demo_toplevel()
    return getattr(lp, name + "_toplevel")()
