"""Host-side mirror of the reference's Lair front half: the `func!` surface (as a small
builder API), expansion of `if`/`match` into `Choose` blocks, compilation to index-based
bytecode, `Toplevel`, `FuncChip` column layout, and the flat u32 serialisation that crosses the
C ABI (the "serialised Toplevel" of SURVEY.md 8f.2).

Reference files mirrored (behaviour, not text):
  src/lair/expr.rs, src/lair/macros.rs        surface IR (`FuncE`, `OpE`, `CtrlE`)
  src/lair/toplevel.rs:423-573                `expand` (If -> Choose+AssertNe/AssertEq, Match -> Choose+Contains/AssertNe)
  src/lair/toplevel.rs:255-330,506-573,739-880  `compile` (stack-index linking, return idents)
  src/lair/bytecode.rs:12-146                 `Op`, `Block`, `Ctrl`, `Func`
  src/lair/func_chip.rs:90-277                `compute_layout_sizes`
  src/lair/mod.rs:53-112                      `demo_toplevel`
This module is static, once-per-program host logic (the analogue of building a `Toplevel` and
running `Chip::new`); the per-proof work is native (csrc/).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

P = 2013265921  # BabyBear


def F(x: int) -> int:
    """field_from_i32 (src/lair/mod.rs:33-40)"""
    return x % P


# ------------------------------------------------------------------------------------------------
# Surface language (what `func!` produces).  Variables are names; a name may denote an array.
# ------------------------------------------------------------------------------------------------

VarRef = Union[str, Tuple[str, int]]  # "x" (size 1) or ("x", size)


def _vname(v: VarRef) -> str:
    return v if isinstance(v, str) else v[0]


def _vsize(v: VarRef) -> int:
    return 1 if isinstance(v, str) else v[1]


@dataclass
class St:
    """One surface statement: kind + payload (see the helpers below)."""
    kind: str
    args: tuple


def const(tgt: str, value: int) -> St: return St("const", (tgt, F(value)))
def array(tgt: str, values: Sequence[int]) -> St: return St("array", (tgt, [F(v) for v in values]))
def add(tgt: VarRef, a: str, b: str) -> St: return St("add", (tgt, a, b))
def sub(tgt: VarRef, a: str, b: str) -> St: return St("sub", (tgt, a, b))
def mul(tgt: VarRef, a: str, b: str) -> St: return St("mul", (tgt, a, b))
def div(tgt: VarRef, a: str, b: str) -> St: return St("div", (tgt, a, b))
def inv(tgt: VarRef, a: str) -> St: return St("inv", (tgt, a))
def not_(tgt: str, a: str) -> St: return St("not", (tgt, a))
def eq(tgt: str, a: str, b: str) -> St: return St("eq", (tgt, a, b))
def assert_eq(a: str, b: str) -> St: return St("assert_eq", (a, b))
def assert_ne(a: str, b: str) -> St: return St("assert_ne", (a, b))
def contains(a: str, b: str) -> St: return St("contains", (a, b))
def call(outs: Sequence[VarRef], name: str, inps: Sequence[str]) -> St: return St("call", (list(outs), name, list(inps)))
def preimg(outs: Sequence[VarRef], name: str, inps: Sequence[str]) -> St: return St("preimg", (list(outs), name, list(inps)))
def store(ptr: str, vals: Sequence[str]) -> St: return St("store", (ptr, list(vals)))
def load(vals: Sequence[VarRef], ptr: str) -> St: return St("load", (list(vals), ptr))
def extern_call(outs: Sequence[VarRef], name: str, inps: Sequence[str]) -> St: return St("extern_call", (list(outs), name, list(inps)))
def range_u8(xs: Sequence[str]) -> St: return St("range_u8", (list(xs),))
def emit(xs: Sequence[str]) -> St:
    """`emit(x, ...)` -- OpE::Emit (src/lair/expr.rs:147-148): pushes values to QueryRecord::emitted while executing.  It has
    no columns, no constraints and no place in the layout (air.rs:493, func_chip.rs:274, trace.rs:415), so traces and
    proofs do not depend on it; here it is checked (the variables must be bound) and compiles to nothing -- the
    `emitted` side channel of the REPL is not kept."""
    return St("emit", (list(xs),))
def slice_(outs: Sequence[VarRef], inps: Sequence[str]) -> St:
    """`let (x, y: [7]) = (a, b);` -- OpE::Slice (src/lair/expr.rs:141-143, macros.rs:247-268): rebinds the values of
    `inps`, concatenated, to the names `outs`; emits no bytecode."""
    return St("slice", (list(outs), list(inps)))
def ret(*vs: str) -> St: return St("return", (list(vs),))


def if_(x: str, then: List[St], else_: Optional[List[St]] = None) -> St:
    """`if x { then }` -- when `else_` is None the statements that follow form the else block."""
    return St("if", (x, then, else_))


def match(x: str, arms: Sequence[Tuple[Union[int, Sequence[int]], List[St]]], default: Optional[List[St]] = None,
          exhaustive: bool = False) -> St:
    """`match x { k => {..}, k1 | k2 => {..} }` -- following statements form the default arm
    unless `default` is given or `exhaustive` is set (no default, src/lair/macros.rs)."""
    norm = []
    for keys, blk in arms:
        ks = [keys] if isinstance(keys, int) else list(keys)
        norm.append(([F(k) for k in ks], blk))
    return St("match", (x, norm, default, exhaustive))


def match_many(x: str, arms: Sequence[Tuple[Sequence[int], List[St]]], default: Optional[List[St]] = None,
               exhaustive: bool = False) -> St:
    """`match [x] { [a, b] => {..} }` on an array variable (CtrlE::MatchMany)."""
    return St("match_many", (x, [([F(k) for k in ks], blk) for ks, blk in arms], default, exhaustive))


@dataclass
class FuncE:
    name: str
    input_params: List[VarRef]
    output_size: int
    body: List[St]
    invertible: bool = False
    partial: bool = False

    @property
    def input_size(self) -> int:
        return sum(_vsize(v) for v in self.input_params)


def func(name: str, params: Sequence[VarRef], output_size: int, body: List[St], *, invertible=False,
         partial=False) -> FuncE:
    return FuncE(name, list(params), output_size, body, invertible, partial)


# ------------------------------------------------------------------------------------------------
# Bytecode (src/lair/bytecode.rs)
# ------------------------------------------------------------------------------------------------

@dataclass
class Op:
    kind: str          # const add sub mul inv not assert_eq assert_ne contains call preimg store load extern_call range_u8
    args: tuple


@dataclass
class Ctrl:
    kind: str          # return | choose | choose_many
    args: tuple
    # return:      (ident, [vars])
    # choose:      (var, [(keys, Block)] unique branches in order, default Block | None)
    # choose_many: (vars, [(key_tuple, Block)], default Block | None)


@dataclass
class Block:
    ops: List[Op]
    ctrl: Ctrl
    return_idents: List[int]


@dataclass
class Func:
    name: str
    invertible: bool
    partial: bool
    index: int
    input_size: int
    output_size: int
    body: Block


@dataclass
class ChipInfo:
    """What Lair needs to know about an extern chip (src/lair/chipset.rs:9-47)."""
    name: str
    input_size: int
    output_size: int
    witness_size: int
    require_size: int
    kind: int = 0        # native chipset id (1/2/3 = Poseidon2 hasher W=24/32/40)


class _Linker:
    def __init__(self, info: Dict[str, Tuple[int, int, int, bool]], chips: Dict[str, Tuple[int, ChipInfo]]):
        self.var_index = 0
        self.return_ident = 0
        self.return_idents: List[int] = []
        self.link: Dict[str, List[int]] = {}
        self.info = info
        self.chips = chips
        self.partial = False
        self.uniq = 0

    def new(self, v: VarRef) -> List[int]:
        n = _vsize(v)
        idxs = list(range(self.var_index, self.var_index + n))
        self.var_index += n
        self.link[_vname(v)] = idxs
        return idxs

    def get(self, name: str) -> List[int]:
        if name not in self.link:
            raise KeyError(f"unbound variable {name}")
        return self.link[name]

    def save(self):
        return self.var_index, dict(self.link)

    def restore(self, st):
        self.var_index, self.link = st[0], dict(st[1])

    def fresh(self) -> str:
        self.uniq += 1
        return f"${self.uniq}"


def _compile_block(stmts: List[St], pre_ops: List[St], ctx: _Linker) -> Block:
    ops: List[Op] = []
    stmts = list(pre_ops) + list(stmts)
    i = 0
    while i < len(stmts):
        s = stmts[i]
        k, a = s.kind, s.args
        last = False
        if k == "const":
            ops.append(Op("const", (a[1],))); ctx.new(a[0])
        elif k == "array":
            for f in a[1]:
                ops.append(Op("const", (f,)))
            ctx.new((a[0], len(a[1])))
        elif k in ("add", "sub", "mul"):
            xa, xb = list(ctx.get(a[1])), list(ctx.get(a[2]))
            assert len(xa) == len(xb), f"Var mismatch on {k}"
            for u, v in zip(xa, xb):
                ops.append(Op(k, (u, v)))
            ctx.new((_vname(a[0]), len(xa)))
        elif k == "div":  # expand: inv + mul (toplevel.rs OpE::expand)
            t = ctx.fresh()
            stmts[i:i + 1] = [inv((t, len(ctx.get(a[2]))), a[2]), mul(a[0], a[1], t)]
            continue
        elif k == "eq":   # expand: sub + not
            t = ctx.fresh()
            stmts[i:i + 1] = [sub(t, a[1], a[2]), not_(a[0], t)]
            continue
        elif k == "inv":
            xa = list(ctx.get(a[1]))
            for u in xa:
                ops.append(Op("inv", (u,)))
            ctx.new((_vname(a[0]), len(xa)))
        elif k == "not":
            ops.append(Op("not", (ctx.get(a[1])[0],))); ctx.new(a[0])
        elif k == "assert_eq":
            ops.append(Op("assert_eq", (list(ctx.get(a[0])), list(ctx.get(a[1])))))
        elif k == "assert_ne":
            ops.append(Op("assert_ne", (list(ctx.get(a[0])), list(ctx.get(a[1])))))
        elif k == "contains":
            ops.append(Op("contains", (list(ctx.get(a[0])), ctx.get(a[1])[0])))
        elif k in ("call", "preimg"):
            outs, name, inps = a
            idx, in_sz, out_sz, callee_partial = ctx.info[name]
            if callee_partial:   # src/lair/toplevel.rs:640-641
                assert ctx.partial, "assertion failed: ctx.partial"
            flat = [j for v in inps for j in ctx.get(v)]
            want_in, want_out = (in_sz, out_sz) if k == "call" else (out_sz, in_sz)
            assert len(flat) == want_in, f"Input mismatch on {k} {name}"
            assert sum(_vsize(o) for o in outs) == want_out, f"Output mismatch on {k} {name}"
            ops.append(Op(k, (idx, flat)))
            for o in outs:
                ctx.new(o)
        elif k == "store":
            ops.append(Op("store", ([j for v in a[1] for j in ctx.get(v)],))); ctx.new(a[0])
        elif k == "load":
            ops.append(Op("load", (sum(_vsize(v) for v in a[0]), ctx.get(a[1])[0])))
            for v in a[0]:
                ctx.new(v)
        elif k == "extern_call":
            outs, name, inps = a
            cidx, chip = ctx.chips[name]
            flat = [j for v in inps for j in ctx.get(v)]
            assert len(flat) == chip.input_size and sum(_vsize(o) for o in outs) == chip.output_size
            ops.append(Op("extern_call", (cidx, flat)))
            for o in outs:
                ctx.new(o)
        elif k == "range_u8":
            ops.append(Op("range_u8", ([j for v in a[0] for j in ctx.get(v)],)))
        elif k == "emit":
            for v in a[0]:
                ctx.get(v)
        elif k == "slice":   # toplevel.rs OpE::Slice: pure re-linking
            flat = [j for v in a[1] for j in ctx.get(v)]
            assert sum(_vsize(o) for o in a[0]) == len(flat), "Var mismatch on slice"
            pos = 0
            for o in a[0]:
                ctx.link[_vname(o)] = flat[pos:pos + _vsize(o)]
                pos += _vsize(o)
        elif k in ("return", "if", "match", "match_many"):
            last = True
        else:
            raise ValueError(k)
        if last:
            break
        i += 1
    assert i < len(stmts), "block must end in a control statement"
    s, rest = stmts[i], stmts[i + 1:]
    saved = ctx.return_idents
    ctx.return_idents = []
    ctrl = _compile_ctrl(s, rest, ctx)
    block_idents = ctx.return_idents
    assert block_idents, "A block must have at least one return ident"
    ctx.return_idents = saved + block_idents
    return Block(ops, ctrl, list(block_idents))


def _compile_ctrl(s: St, rest: List[St], ctx: _Linker) -> Ctrl:
    k, a = s.kind, s.args
    if k == "return":
        assert not rest, "statements after return"
        vs = [j for v in a[0] for j in ctx.get(v)]
        c = Ctrl("return", (ctx.return_ident, vs))
        ctx.return_idents.append(ctx.return_ident)
        ctx.return_ident += 1
        return c
    if k == "if":
        x, then, else_ = a
        else_ = rest if else_ is None else else_
        n = len(ctx.get(x))
        z = ctx.fresh()
        zeros = [0] * n
        # CtrlE::If expansion (toplevel.rs:426-441): branch [0..] = else with AssertEq, default = then with AssertNe
        t_pre = [array(z, zeros), assert_ne(x, z)]
        f_pre = [array(z, zeros), assert_eq(x, z)]
        if n == 1:
            return _choose(x, [([0], else_, f_pre)], (then, t_pre), ctx)
        return _choose_many(x, [(zeros, else_, f_pre)], (then, t_pre), ctx)
    if k == "match":
        x, arms, default, exhaustive = a
        if default is None and not exhaustive:
            default = rest
        branches = []
        for keys, blk in arms:
            arr = ctx.fresh()
            branches.append((keys, blk, [array(arr, keys), contains(arr, x)]))
        dflt = None
        if default is not None:
            pre = []
            for keys, _ in arms:
                for f in keys:
                    fv = ctx.fresh()
                    pre += [const(fv, f), assert_ne(x, fv)]
            dflt = (default, pre)
        return _choose(x, branches, dflt, ctx)
    if k == "match_many":
        x, arms, default, exhaustive = a
        if default is None and not exhaustive:
            default = rest
        branches = []
        for keys, blk in arms:
            arr = ctx.fresh()
            branches.append((keys, blk, [array(arr, keys), assert_eq(x, arr)]))
        dflt = None
        if default is not None:
            pre = []
            for keys, _ in arms:
                arr = ctx.fresh()
                pre += [array(arr, keys), assert_ne(x, arr)]
            dflt = (default, pre)
        return _choose_many(x, branches, dflt, ctx)
    raise ValueError(k)


def _choose(x, branches, dflt, ctx: _Linker) -> Ctrl:
    var = ctx.get(x)[0]
    uniq = []
    for keys, blk, pre in branches:
        st = ctx.save()
        b = _compile_block(blk, pre, ctx)
        ctx.restore(st)
        uniq.append((list(keys), b))
    d = _compile_block(dflt[0], dflt[1], ctx) if dflt is not None else None
    return Ctrl("choose", (var, uniq, d))


def _choose_many(x, branches, dflt, ctx: _Linker) -> Ctrl:
    vars_ = list(ctx.get(x))
    brs = []
    for keys, blk, pre in branches:
        st = ctx.save()
        b = _compile_block(blk, pre, ctx)
        ctx.restore(st)
        brs.append((list(keys), b))
    d = _compile_block(dflt[0], dflt[1], ctx) if dflt is not None else None
    return Ctrl("choose_many", (vars_, brs, d))


MEM_TABLE_SIZES = [2, 3, 4, 5, 6, 8]   # src/lair/execute.rs:243-244
DEPTH_W = 4                            # src/lair/provenance.rs
DEPTH_LESS_THAN_SIZE = 6               # src/gadgets/unsigned/less_than.rs:132-135 (u32)
DEPTH_LESS_THAN_REQUIRES = 1


def mem_index_from_len(n: int) -> int:
    return MEM_TABLE_SIZES.index(n)


class Toplevel:
    """src/lair/toplevel.rs:25-56 -- ordered map name -> Func plus the extern chip map."""

    def __init__(self, funcs_e: Sequence[FuncE], chips: Sequence[ChipInfo] = ()):
        self.chips: List[ChipInfo] = list(chips)
        chip_map = {c.name: (i, c) for i, c in enumerate(self.chips)}
        info = {f.name: (i, f.input_size, f.output_size, f.partial) for i, f in enumerate(funcs_e)}
        assert len(info) == len(funcs_e), "duplicate function name"
        self.funcs: List[Func] = []
        for i, fe in enumerate(funcs_e):
            ctx = _Linker(info, chip_map)
            ctx.partial = fe.partial
            for v in fe.input_params:
                ctx.new(v)
            body = _compile_block(fe.body, [], ctx)
            self.funcs.append(Func(fe.name, fe.invertible, fe.partial, i, fe.input_size, fe.output_size, body))
        self._by_name = {f.name: f for f in self.funcs}
        self._serial: Optional[np.ndarray] = None
        self.funcs_e: List[FuncE] = list(funcs_e)

    @classmethod
    def new_pure(cls, funcs_e: Sequence[FuncE]) -> "Toplevel":
        return cls(funcs_e, ())

    def func_by_name(self, name: str) -> Func:
        return self._by_name[name]

    def func_by_index(self, i: int) -> Func:
        return self.funcs[i]

    def num_funcs(self) -> int:
        return len(self.funcs)

    # ---- column layout (src/lair/func_chip.rs:90-277) ------------------------------------------
    def compute_layout_sizes(self, f: Func) -> "LayoutSizes":
        aux = 2
        if f.partial:
            aux += DEPTH_W + 3 * ((DEPTH_W + 1) // 2)
        st = _LayoutState(aux=aux, sel=0, max_map=f.input_size)
        self._layout_block(f.body, [1] * f.input_size, st)
        return LayoutSizes(1, f.input_size, f.output_size, st.aux, st.sel, st.max_map)

    def _layout_block(self, b: Block, degrees: List[int], st: "_LayoutState"):
        for op in b.ops:
            self._layout_op(op, degrees, st)
        st.max_map = max(st.max_map, len(degrees))
        c = b.ctrl
        if c.kind == "return":
            st.sel += 1
            return
        blocks = [blk for _, blk in c.args[1]] + ([c.args[2]] if c.args[2] is not None else [])
        n = len(degrees)
        base_aux, max_aux = st.aux, st.aux
        for blk in blocks:
            st.aux = base_aux
            self._layout_block(blk, degrees, st)
            del degrees[n:]
            max_aux = max(max_aux, st.aux)
        st.aux = max_aux

    def _layout_op(self, op: Op, degrees: List[int], st: "_LayoutState"):
        k, a = op.kind, op.args
        op_aux = 0
        if k == "assert_ne":
            op_aux = len(a[0])
        elif k == "contains":
            op_aux = len(a[0]) - 1
        elif k == "const":
            degrees.append(0)
        elif k in ("add", "sub"):
            degrees.append(max(degrees[a[0]], degrees[a[1]]))
        elif k == "mul":
            d = degrees[a[0]] + degrees[a[1]]
            if d < 2:
                degrees.append(d)
            else:
                degrees.append(1); op_aux = 1
        elif k == "inv":
            if degrees[a[0]] == 0:
                degrees.append(0)
            else:
                degrees.append(1); op_aux = 1
        elif k == "not":
            if degrees[a[0]] == 0:
                degrees.append(0)
            else:
                degrees.append(1); op_aux = 2
        elif k in ("call", "preimg"):
            callee = self.funcs[a[0]]
            n = callee.output_size if k == "call" else callee.input_size
            op_aux = n + 3
            if callee.partial:
                op_aux += DEPTH_W + DEPTH_LESS_THAN_SIZE + 3 * DEPTH_LESS_THAN_REQUIRES
            degrees.extend([1] * n)
        elif k == "store":
            op_aux = 4; degrees.append(1)
        elif k == "load":
            op_aux = a[0] + 3; degrees.extend([1] * a[0])
        elif k == "extern_call":
            chip = self.chips[a[0]]
            op_aux = chip.witness_size + 3 * chip.require_size
            # NB the reference extends `degrees` by aux_size here (func_chip.rs:262), while the
            # executor/trace push `output_size` values; we track degrees for the outputs only.
            degrees.extend([1] * chip.output_size)
        elif k == "range_u8":
            op_aux = 3 * ((len(a[0]) + 1) // 2)
        st.aux += op_aux

    # ---- surface image: the FuncE list as words, for the C++ host layer (csrc/host_lair.cpp) -----------
    def surface_image(self) -> np.ndarray:
        strs: Dict[str, int] = {}

        def sid(name: str) -> int:
            if name not in strs:
                strs[name] = len(strs)
            return strs[name]

        KIND = {"const": 0, "array": 1, "add": 2, "sub": 3, "mul": 4, "div": 5, "inv": 6, "not": 7, "eq": 8,
                "assert_eq": 9, "assert_ne": 10, "contains": 11, "call": 12, "preimg": 13, "store": 14, "load": 15,
                "extern_call": 16, "range_u8": 17, "return": 18, "if": 19, "match": 20, "match_many": 21, "slice": 22, "emit": 23}

        def vars_(vs) -> List[int]:
            out = [len(vs)]
            for v in vs:
                out += [sid(_vname(v)), _vsize(v)]
            return out

        def names(ns) -> List[int]:
            return [len(ns)] + [sid(n) for n in ns]

        def block(stmts) -> List[int]:
            out = [len(stmts)]
            for st in stmts:
                out += stmt(st)
            return out

        def stmt(st: St) -> List[int]:
            k, a = st.kind, st.args
            out = [KIND[k]]
            if k == "const":
                out += [sid(a[0]), a[1]]
            elif k == "array":
                out += [sid(a[0]), len(a[1])] + list(a[1])
            elif k in ("add", "sub", "mul", "div", "eq"):
                out += [sid(_vname(a[0])), sid(a[1]), sid(a[2])]
            elif k in ("inv", "not"):
                out += [sid(_vname(a[0])), sid(a[1])]
            elif k in ("assert_eq", "assert_ne", "contains"):
                out += [sid(a[0]), sid(a[1])]
            elif k in ("call", "preimg", "extern_call"):
                out += vars_(a[0]) + [sid(a[1])] + names(a[2])
            elif k == "store":
                out += [sid(a[0])] + names(a[1])
            elif k == "load":
                out += vars_(a[0]) + [sid(a[1])]
            elif k in ("range_u8", "return", "emit"):
                out += names(a[0])
            elif k == "slice":
                out += vars_(a[0]) + names(a[1])
            elif k == "if":
                out += [sid(a[0])] + block(a[1]) + ([1] + block(a[2]) if a[2] is not None else [0])
            elif k in ("match", "match_many"):
                out += [sid(a[0]), len(a[1])]
                for keys, blk in a[1]:
                    out += [len(keys)] + list(keys) + block(blk)
                out += ([1] + block(a[2]) if a[2] is not None else [0]) + [int(a[3])]
            else:
                raise ValueError(k)
            return out

        body: List[int] = [len(self.chips)]
        for c in self.chips:
            body += [sid(c.name), c.kind, c.input_size, c.output_size, c.witness_size, c.require_size]
        body.append(len(self.funcs_e))
        for fe in self.funcs_e:
            body += [sid(fe.name)] + vars_(fe.input_params) + [fe.output_size, int(fe.invertible), int(fe.partial)]
            body += block(fe.body)
        head = [0x4C535246, len(strs)]
        for name in strs:                    # insertion order = index order
            raw = name.encode()
            head.append(len(raw))
            head += list(np.frombuffer(raw + b"\0" * (-len(raw) % 4), dtype=np.uint32)) if raw else []
        return np.asarray(head + body, dtype=np.uint32)

    # ---- flat serialisation ---------------------------------------------------------------------
    def serialize(self) -> np.ndarray:
        """Flat u32 program image consumed by the native library (layout in csrc/program.h), compiled by the
        C++ host layer (csrc/host_lair.cpp); `serialize_py` is the same pass in Python."""
        if self._serial is None:
            from . import host
            self._serial = host.compile_toplevel(self.surface_image())
        return self._serial

    def serialize_py(self) -> np.ndarray:
        out: List[int] = [0x4c414952, len(self.funcs), len(self.chips)]  # 'LAIR'
        for c in self.chips:
            out += [c.kind, c.input_size, c.output_size, c.witness_size, c.require_size]
        table_at = len(out)
        out += [0] * len(self.funcs)
        for f in self.funcs:
            out[table_at + f.index] = len(out)
            ls = self.compute_layout_sizes(f)
            out += [f.input_size, f.output_size, int(f.partial), int(f.invertible), ls.aux, ls.sel, ls.max_map]
            self._ser_block(f.body, [1] * f.input_size, out)
        return np.asarray(out, dtype=np.uint32)

    def _ser_block(self, b: Block, degrees: List[int], out: List[int]):
        start = len(out)
        out.append(0)  # block length in words (filled below)
        out.append(len(b.ops))
        for op in b.ops:
            k, a = op.kind, op.args
            code = OPCODES[k]
            if k == "const":
                out += [code, a[0]]; degrees.append(0)
            elif k in ("add", "sub"):
                out += [code, a[0], a[1]]; degrees.append(max(degrees[a[0]], degrees[a[1]]))
            elif k == "mul":
                d = degrees[a[0]] + degrees[a[1]]
                has_aux = int(d >= 2)
                out += [code, a[0], a[1], has_aux]; degrees.append(1 if has_aux else d)
            elif k in ("inv", "not"):
                has_aux = int(degrees[a[0]] != 0)
                out += [code, a[0], has_aux]; degrees.append(has_aux)
            elif k in ("assert_eq", "assert_ne"):
                out += [code, len(a[0])] + list(a[0]) + list(a[1])
            elif k == "contains":
                out += [code, len(a[0])] + list(a[0]) + [a[1]]
            elif k in ("call", "preimg"):
                callee = self.funcs[a[0]]
                n = callee.output_size if k == "call" else callee.input_size
                out += [code, a[0], len(a[1])] + list(a[1]) + [n, int(callee.partial)]
                degrees.extend([1] * n)
            elif k == "store":
                out += [code, len(a[0])] + list(a[0]); degrees.append(1)
            elif k == "load":
                out += [code, a[0], a[1]]; degrees.extend([1] * a[0])
            elif k == "extern_call":
                chip = self.chips[a[0]]
                out += [code, a[0], len(a[1])] + list(a[1])
                degrees.extend([1] * chip.output_size)
            elif k == "range_u8":
                out += [code, len(a[0])] + list(a[0])
            else:
                raise ValueError(k)
        c = b.ctrl
        if c.kind == "return":
            out += [CTRL_RETURN, c.args[0], len(c.args[1])] + list(c.args[1])
        else:
            many = c.kind == "choose_many"
            vars_ = list(c.args[0]) if many else [c.args[0]]
            out += [CTRL_CHOOSE_MANY if many else CTRL_CHOOSE, len(vars_)] + vars_
            out += [len(c.args[1]), int(c.args[2] is not None)]
            n = len(degrees)
            for keys, blk in c.args[1]:
                if many:
                    out += [1] + list(keys)            # one key tuple of len(vars)
                else:
                    out += [len(keys)] + list(keys)    # several scalar keys share one block
                self._ser_block(blk, degrees, out)
                del degrees[n:]
            if c.args[2] is not None:
                self._ser_block(c.args[2], degrees, out)
                del degrees[n:]
        out += [len(b.return_idents)] + list(b.return_idents)
        out[start] = len(out) - start


OPCODES = {"const": 1, "add": 2, "sub": 3, "mul": 4, "inv": 5, "not": 6, "assert_eq": 7, "assert_ne": 8,
           "contains": 9, "call": 10, "preimg": 11, "store": 12, "load": 13, "extern_call": 14, "range_u8": 15}
CTRL_RETURN, CTRL_CHOOSE, CTRL_CHOOSE_MANY = 100, 101, 102


@dataclass
class _LayoutState:
    aux: int
    sel: int
    max_map: int


@dataclass(frozen=True)
class LayoutSizes:
    """src/lair/func_chip.rs:11-26 (`ColumnLayout<usize, usize>`); max_map is ours (stack bound)."""
    nonce: int
    input: int
    output: int
    aux: int
    sel: int
    max_map: int = 0

    def total(self) -> int:
        return self.nonce + self.input + self.aux + self.sel + self.output


class FuncChip:
    """src/lair/func_chip.rs:28-80"""

    def __init__(self, func: Func, toplevel: Toplevel):
        self.func = func
        self.toplevel = toplevel
        self.layout_sizes = toplevel.compute_layout_sizes(func)

    @classmethod
    def from_name(cls, name: str, toplevel: Toplevel) -> "FuncChip":
        return cls(toplevel.func_by_name(name), toplevel)

    @classmethod
    def from_index(cls, idx: int, toplevel: Toplevel) -> "FuncChip":
        return cls(toplevel.func_by_index(idx), toplevel)

    @classmethod
    def from_toplevel(cls, toplevel: Toplevel) -> List["FuncChip"]:
        return [cls(f, toplevel) for f in toplevel.funcs]

    def width(self) -> int:
        return self.layout_sizes.total()


def demo_toplevel() -> Toplevel:
    """src/lair/mod.rs:53-112"""
    factorial = func("factorial", ["n"], 1, [
        const("one", 1),
        if_("n", [
            sub("pred", "n", "one"),
            call(["m"], "factorial", ["pred"]),
            mul("res", "n", "m"),
            ret("res"),
        ]),
        ret("one"),
    ])
    fib = func("fib", ["n"], 1, [
        const("one", 1),
        match("n", [
            (0, [const("zero", 0), ret("zero")]),
            (1, [ret("one")]),
        ]),
        sub("n_1", "n", "one"),
        call(["a"], "fib", ["n_1"]),
        sub("n_2", "n_1", "one"),
        call(["b"], "fib", ["n_2"]),
        add("res", "a", "b"),
        ret("res"),
    ])
    even = func("even", ["n"], 1, [
        const("one", 1),
        match("n", [(0, [ret("one")])]),
        sub("pred", "n", "one"),
        call(["res"], "odd", ["pred"]),
        ret("res"),
    ])
    odd = func("odd", ["n"], 1, [
        const("one", 1),
        match("n", [(0, [const("zero", 0), ret("zero")])]),
        sub("pred", "n", "one"),
        call(["res"], "even", ["pred"]),
        ret("res"),
    ])
    return Toplevel.new_pure([factorial, fib, even, odd])
