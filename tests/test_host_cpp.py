"""The C++ host layer (csrc/host_lair.cpp, host_air.cpp -- the reference's Toplevel::new / FuncChip layout /
Chip::new in the language class the reference is written in) against the Python executable specification
(lurk_b200/lair.py, air.py): identical program images and identical chip-vector blobs, word for word, on every
test program.  Host code only, no GPU."""
import numpy as np
import pytest

import goldens
from lurk_b200 import _native as N
from lurk_b200 import host
from lurk_b200.air import build_chip_vector, serialize_machine
from lurk_b200.lair import DEPTH_W, FuncChip, Toplevel, call, func, ret

TOPLEVELS = ["demo", "match", "inner_match", "not_eq", "if_many", "match_many", "assert", "equal_branch", "range",
             "memory", "ackermann", "hash", "partial", "u64", "lurk_misc", "lurk_eval_subset", "lurk_data", "lurk_core"]


@pytest.fixture(scope="module", autouse=True)
def _built():
    N.build()


@pytest.mark.parametrize("tname", TOPLEVELS)
def test_program_image_is_identical(tname):
    top = goldens.toplevel_for(tname)
    want = top.serialize_py()
    got = host.compile_toplevel(top.surface_image())
    assert np.array_equal(got, want)


@pytest.mark.parametrize("tname", TOPLEVELS)
def test_chip_vector_blob_is_identical(tname):
    top = goldens.toplevel_for(tname)
    surface = top.surface_image()
    for f in top.funcs:
        fc = FuncChip(f, top)
        num_pv = f.input_size + f.output_size + (DEPTH_W if f.partial else 0)
        want = serialize_machine(build_chip_vector(fc), num_pv)
        got = host.chip_vector_blob(surface, f.index)
        assert np.array_equal(got, want), (tname, f.name)


def test_cpp_builder_demo_toplevel():
    # the builder API of host_lair.h (func / match / call ...) used directly from C++
    assert np.array_equal(host.demo_toplevel_image(), goldens.toplevel_for("demo").serialize_py())


def test_cpp_compile_errors_are_reported():
    foo = func("foo", ["a"], 1, [ret("a")], partial=True)
    bar = func("bar", ["a"], 1, [call(["b"], "foo", ["a"]), ret("b")])
    top = Toplevel.__new__(Toplevel)          # bypass the Python-side check: serialise the surface only
    top.chips, top.funcs_e = [], [foo, bar]
    with pytest.raises(N.LurkStarkError, match="ctx.partial"):
        host.compile_toplevel(top.surface_image())
    with pytest.raises(N.LurkStarkError):
        host.compile_toplevel(np.asarray([1, 2, 3], dtype=np.uint32))


def test_cpp_layout_tests():
    # tests/cpp/host_layout_test.cpp: the reference's layout-size tests (src/lair/trace.rs:445-459,545-552,612-619)
    # and chip widths (src/core/eval_direct.rs:2053-2055) written in C++ against the builder API
    import subprocess
    exe = N.build_cpp_tests()[0]
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
