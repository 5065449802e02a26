"""The C++ host layer (csrc/host_lair.{h,cpp}, host_air.cpp) behind ctypes: compiling a Toplevel and capturing
its chip vector's AIR are done by liblurkstark's host code -- the reference does them in Rust
(Toplevel::new, src/lair/toplevel.rs; Chip::new, src/lair/lair_chip.rs:196-231) -- and `lair.py` / `air.py`
keep the same passes in Python as the executable specification the tests hold the C++ to."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as N


def _take(fn, *args) -> np.ndarray:
    out = ctypes.POINTER(ctypes.c_uint32)()
    n = ctypes.c_size_t()
    N.check(fn(*args, ctypes.byref(out), ctypes.byref(n)))
    words = np.ctypeslib.as_array(out, shape=(n.value,)).copy()
    N.lib().lsk_words_free(out)
    return words


def compile_toplevel(surface: np.ndarray) -> np.ndarray:
    """surface image (Toplevel.surface_image) -> program image for lsk_program_load"""
    surface = N.u32(surface)
    return _take(N.lib().lsk_host_compile_toplevel, N.vp(surface), ctypes.c_size_t(len(surface)))


def chip_vector_blob(surface: np.ndarray, entry_func: int) -> np.ndarray:
    """surface image + entry func index -> 'MACH' blob for lsk_machine_new / lsk_vk_new"""
    surface = N.u32(surface)
    return _take(N.lib().lsk_host_chip_vector, N.vp(surface), ctypes.c_size_t(len(surface)), int(entry_func))


def demo_toplevel_image() -> np.ndarray:
    """program image of the C++ builder's demo_toplevel() (src/lair/mod.rs:53-112)"""
    return _take(N.lib().lsk_host_demo_toplevel)
