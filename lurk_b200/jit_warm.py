"""Fill the cubin cache of the compiled constraint programs (csrc/jit.cu) for the stock programs (hash / u64 / fib demos and the full Lurk toplevel).

    python -m lurk_b200.jit_warm

NVRTC runs without a GPU, so `__graft_entry__.build()` calls this after compiling the library: a machine created
later (here or on a GPU box that received this tree) finds the cubins and never waits for the ~30 s per chip.
Programs outside this list are interpreted unless LSK_JIT=1 asks for compilation on a miss.
"""
from __future__ import annotations

import ctypes
import time

from . import _native as N
from . import host


def precompile(toplevel, entry: str) -> dict:
    """lsk_jit_precompile for the chip vector of `entry`: {"compiled": n, "cached": m}."""
    blob = N.u32(host.chip_vector_blob(toplevel.surface_image(), toplevel.func_by_name(entry).index))
    c, h = ctypes.c_uint32(), ctypes.c_uint32()
    N.check(N.lib().lsk_jit_precompile(N.vp(blob), ctypes.c_size_t(len(blob)), ctypes.byref(c), ctypes.byref(h)))
    return {"compiled": c.value, "cached": h.value}


def stock_programs():
    from .lair import demo_toplevel
    from .lurk_eval import build_lurk_toplevel
    from .programs import hash_toplevel, u64_toplevel
    return [(hash_toplevel(), "cons_chain"), (u64_toplevel(), "factu"), (demo_toplevel(), "fib"),
            (build_lurk_toplevel(), "lurk_main")]      # the 41 chips behind tools/lurk_prove.py (benches/fib, lcs, sum)


def main():
    for top, entry in stock_programs():
        t0 = time.time()
        r = precompile(top, entry)
        print(f"jit_warm {entry}: {r} in {time.time() - t0:.1f} s", flush=True)


if __name__ == "__main__":
    main()
