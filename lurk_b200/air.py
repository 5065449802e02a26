"""Symbolic capture of every chip's AIR into the flat "chip descriptor" that crosses the C ABI.

In the reference each chip's `Air::eval` is executed by sphinx-core with different builders: once
symbolically in `Chip::new` to collect the lookup interactions (call sites:
src/lair/lair_chip.rs:222-231), and then per row of the quotient domain with a folding builder.
Here `eval` is run once, symbolically, in this module (static per Toplevel), and the result is
compiled into (a) a register-allocated straight-line constraint program over base-field values
and (b) the send/receive interaction list with affine column combinations -- both interpreted by
the CUDA quotient / LogUp kernels (and by the CPU oracle).

Mirrored reference code:
  Func::eval / Op::eval / Ctrl::eval      src/lair/air.rs:133-566
  MemChip::eval                           src/lair/memory.rs:71-109
  BytesChip::eval                         src/gadgets/bytes/trace.rs:116-143
  Entrypoint eval                         src/lair/lair_chip.rs:166-191
  LookupBuilder::{provide, require}       src/air/builder.rs:42-104
  relations                               src/lair/relations.rs:6-59, src/gadgets/bytes/relation.rs:7-135
  build_lair_chip_vector                  src/lair/lair_chip.rs:196-215
Third-party behaviour assumed (sphinx-core, not in the tree): interactions are
`InteractionKind::Memory` (argument index 1), `when(c).assert_zero(x)` is the constraint `c * x`,
`log_quotient_degree = log2_ceil(max(max_constraint_degree, 3) - 1)`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .lair import (DEPTH_LESS_THAN_REQUIRES, DEPTH_LESS_THAN_SIZE, DEPTH_W, MEM_TABLE_SIZES, Block, Ctrl, Func,
                   FuncChip, Op, P, Toplevel)

CALL_TAG, MEMORY_TAG, BYTE_TAG = 0, 1, 3
INTERACTION_KIND_MEMORY = 1  # sphinx_core::lookup::InteractionKind::Memory as usize

# leaf kinds
L_CONST, L_MAIN, L_MAIN_NEXT, L_PREP, L_PREP_NEXT, L_PV, L_SPECIAL = range(7)
SP_IS_FIRST, SP_IS_LAST, SP_IS_TRANSITION = 0, 1, 2


class Expr:
    """Symbolic polynomial over trace cells.  Affine expressions are kept in normal form
    (`lin`: {leaf -> coeff}, `const`), everything else as a DAG node."""
    __slots__ = ("g", "kind", "a", "b", "lin", "const", "degree", "id")

    def __init__(self, g: "Graph", kind: str, a=None, b=None, lin=None, const=0, degree=0):
        self.g, self.kind, self.a, self.b = g, kind, a, b
        self.lin, self.const, self.degree = lin, const % P, degree
        self.id = g.fresh()

    # -- construction helpers --------------------------------------------------------------------
    def is_affine(self) -> bool:
        return self.kind == "aff"

    def is_const(self) -> bool:
        return self.kind == "aff" and not self.lin

    def _coerce(self, o) -> "Expr":
        return o if isinstance(o, Expr) else self.g.const(int(o))

    def __add__(self, o):
        o = self._coerce(o)
        if self.is_affine() and o.is_affine():
            lin = dict(self.lin)
            for k, c in o.lin.items():
                v = (lin.get(k, 0) + c) % P
                if v:
                    lin[k] = v
                else:
                    lin.pop(k, None)
            return self.g.affine(lin, self.const + o.const)
        return Expr(self.g, "add", self, o, degree=max(self.degree, o.degree))

    __radd__ = __add__

    def __neg__(self):
        if self.is_affine():
            return self.g.affine({k: (-c) % P for k, c in self.lin.items()}, -self.const)
        return Expr(self.g, "neg", self, degree=self.degree)

    def __sub__(self, o):
        o = self._coerce(o)
        if self.is_affine() and o.is_affine():
            return self + (-o)
        return Expr(self.g, "sub", self, o, degree=max(self.degree, o.degree))

    def __rsub__(self, o):
        return self._coerce(o) - self

    def __mul__(self, o):
        o = self._coerce(o)
        if o.is_const():
            self, o = o, self
        if self.is_const():
            c = self.const
            if c == 0:
                return self.g.const(0)
            if c == 1:
                return o
            if o.is_affine():
                return self.g.affine({k: v * c % P for k, v in o.lin.items()}, o.const * c)
        return Expr(self.g, "mul", self, o, degree=self.degree + o.degree)

    __rmul__ = __mul__


class Graph:
    def __init__(self):
        self.n = 0

    def fresh(self) -> int:
        self.n += 1
        return self.n

    def affine(self, lin: Dict[Tuple[int, int], int], const: int) -> Expr:
        return Expr(self, "aff", lin=lin, const=const, degree=1 if lin else 0)

    def const(self, c: int) -> Expr:
        return self.affine({}, c)

    def leaf(self, kind: int, idx: int) -> Expr:
        return self.affine({(kind, idx): 1}, 0)


@dataclass
class Interaction:
    values: List[Expr]      # affine
    multiplicity: Expr      # affine
    kind: int = INTERACTION_KIND_MEMORY


class SymbolicBuilder:
    """Plays the role of sphinx's interaction-collecting symbolic builder (`Chip::new`) and of the
    constraint folder at the same time: it records constraints (in evaluation order) and
    sends / receives."""

    def __init__(self, width: int, prep_width: int = 0, num_pv: int = 0):
        self.g = Graph()
        self.width, self.prep_width, self.num_pv = width, prep_width, num_pv
        self.constraints: List[Expr] = []
        self.sends: List[Interaction] = []
        self.receives: List[Interaction] = []
        self._cond: List[Expr] = []

    # main / preprocessed windows
    def local(self, c: int) -> Expr: return self.g.leaf(L_MAIN, c)
    def next(self, c: int) -> Expr: return self.g.leaf(L_MAIN_NEXT, c)
    def prep(self, c: int) -> Expr: return self.g.leaf(L_PREP, c)
    def public_value(self, i: int) -> Expr: return self.g.leaf(L_PV, i)
    def is_first_row(self) -> Expr: return self.g.leaf(L_SPECIAL, SP_IS_FIRST)
    def is_last_row(self) -> Expr: return self.g.leaf(L_SPECIAL, SP_IS_LAST)
    def is_transition(self) -> Expr: return self.g.leaf(L_SPECIAL, SP_IS_TRANSITION)
    def const(self, c: int) -> Expr: return self.g.const(c)

    class _When:
        def __init__(self, b, cond):
            self.b, self.cond = b, cond

        def __enter__(self):
            self.b._cond.append(self.cond)
            return self.b

        def __exit__(self, *a):
            self.b._cond.pop()

    def when(self, cond) -> "_When":
        return SymbolicBuilder._When(self, cond if isinstance(cond, Expr) else self.const(cond))

    def when_transition(self): return self.when(self.is_transition())
    def when_first_row(self): return self.when(self.is_first_row())

    def assert_zero(self, x):
        x = x if isinstance(x, Expr) else self.const(x)
        for c in reversed(self._cond):
            x = c * x
        self.constraints.append(x)

    def assert_eq(self, a, b): self.assert_zero((a if isinstance(a, Expr) else self.const(a)) - b)
    def assert_one(self, a): self.assert_zero((a if isinstance(a, Expr) else self.const(a)) - 1)
    def assert_bool(self, a): self.assert_zero(a * (a - 1))

    # LookupBuilder (src/air/builder.rs:34-133)
    def receive(self, values: Sequence[Expr], is_real: Expr):
        self.receives.append(Interaction(list(values), is_real))

    def send(self, values: Sequence[Expr], is_real: Expr):
        self.sends.append(Interaction(list(values), is_real))

    def provide(self, relation: Sequence[Expr], last_nonce: Expr, last_count: Expr, is_real):
        is_real = is_real if isinstance(is_real, Expr) else self.const(is_real)
        self.receive([last_nonce, last_count] + list(relation), is_real)
        self.send([self.const(0), self.const(0)] + list(relation), is_real)

    def require(self, relation: Sequence[Expr], nonce, prev_nonce, prev_count, count_inv, is_real):
        cv = lambda x: x if isinstance(x, Expr) else self.const(x)
        is_real, nonce = cv(is_real), cv(nonce)
        prev_nonce, prev_count, count_inv = cv(prev_nonce), cv(prev_count), cv(count_inv)
        count = prev_count + 1
        with self.when(is_real):
            self.assert_one(count * count_inv)
        self.receive([prev_nonce, prev_count] + list(relation), is_real)
        self.send([nonce, count] + list(relation), is_real)


# ------------------------------------------------------------------------------------------------
# Chip evals
# ------------------------------------------------------------------------------------------------

class _Cols:
    """ColumnSlice + ColumnIndex (src/lair/air.rs:24-103)"""

    def __init__(self, b: SymbolicBuilder, ls):
        self.b = b
        self.in0 = 1
        self.out0 = 1 + ls.input
        self.aux0 = self.out0 + ls.output
        self.sel0 = self.aux0 + ls.aux
        self.i_in = self.i_aux = self.i_out = 0

    def nonce(self): return self.b.local(0)
    def next_input(self):
        e = self.b.local(self.in0 + self.i_in); self.i_in += 1; return e
    def next_aux(self):
        e = self.b.local(self.aux0 + self.i_aux); self.i_aux += 1; return e
    def next_output(self):
        e = self.b.local(self.out0 + self.i_out); self.i_out += 1; return e
    def output(self, i): return self.b.local(self.out0 + i)
    def sel(self, i): return self.b.local(self.sel0 + i)
    def next_require(self): return self.next_aux(), self.next_aux(), self.next_aux()
    def save(self): return self.i_aux, self.i_out
    def restore(self, s): self.i_aux, self.i_out = s


def _return_sel(block: Block, cols: _Cols) -> Expr:
    e = cols.b.const(0)
    for i in block.return_idents:
        e = e + cols.sel(i)
    return e


def eval_func(b: SymbolicBuilder, f: Func, top: Toplevel):
    """Func::eval (src/lair/air.rs:158-235)"""
    ls = top.compute_layout_sizes(f)
    cols = _Cols(b, ls)
    nonce = cols.nonce()
    with b.when_transition():
        b.assert_eq(b.next(0), nonce + 1)
    vmap: List[Tuple[Expr, int]] = []   # (expr, layout degree as in func_chip.rs)
    call_inp = []
    for _ in range(f.input_size):
        i = cols.next_input()
        vmap.append((i, 1))
        call_inp.append(i)
    toplevel_sel = _return_sel(f.body, cols)
    b.assert_bool(toplevel_sel)
    last_nonce, last_count = cols.next_aux(), cols.next_aux()
    out = [cols.output(i) for i in range(f.output_size)]
    depth: List[Expr] = []
    if f.partial:
        # src/lair/air.rs:210-224: the row's depth as DEPTH_W range-checked bytes, part of the provided relation
        depth = [cols.next_aux() for _ in range(DEPTH_W)]
        requires = [cols.next_require() for _ in range((DEPTH_W + 1) // 2)]
        rec = _ByteAirRecord()
        rec.range_check_u8_iter(depth, toplevel_sel)
        rec.require_all(b, nonce, requires)
        out = out + depth
    b.provide([b.const(CALL_TAG), b.const(f.index)] + call_inp + out, last_nonce, last_count, toplevel_sel)
    _eval_block(b, f.body, cols, toplevel_sel, vmap, top, depth)


class _ByteAirRecord:
    """BytesAirRecordWithContext (src/gadgets/bytes/builder.rs:18-89): byte relations logged while a
    gadget evaluates, then required in order against the row's require triples."""

    def __init__(self):
        self.records: List[Tuple[List, Expr]] = []

    def range_check_u8_iter(self, xs, is_real):     # src/gadgets/bytes/mod.rs:127-138
        xs = list(xs)
        for i in range(0, len(xs), 2):
            self.records.append(([1, xs[i], xs[i + 1] if i + 1 < len(xs) else 0], is_real))

    def less_than(self, i1, i2, r, is_real):
        self.records.append(([3, i1, i2, r], is_real))

    def range_check_u16(self, i, is_real):
        self.records.append(([2, i], is_real))

    def require_all(self, b: SymbolicBuilder, nonce, requires):
        assert len(requires) == len(self.records)
        cv = lambda x: x if isinstance(x, Expr) else b.const(x)
        for (rel, is_real), (pn, pc, ci) in zip(self.records, requires):
            b.require([b.const(BYTE_TAG)] + [cv(x) for x in rel], nonce, pn, pc, ci, is_real)


def _assert_less_than(b: SymbolicBuilder, wit: List[Expr], lhs: List[Expr], rhs: List[Expr], rec: _ByteAirRecord,
                      is_real: Expr):
    """LessThanWitness::assert_less_than (src/gadgets/unsigned/less_than.rs:49-103);
    wit = is_comp[W] ++ [lhs_comp_limb, rhs_comp_limb]."""
    w = len(lhs)
    is_comp, lhs_limb, rhs_limb = wit[:w], wit[w], wit[w + 1]
    with b.when(is_real):
        is_equal = b.const(0)
        for i in range(w):
            if i > 0:
                with b.when(is_equal):
                    b.assert_eq(lhs[i], rhs[i])
            b.assert_bool(is_comp[i])
            is_equal = is_equal + is_comp[i]
        b.assert_one(is_equal)
        sel_l, sel_r = b.const(0), b.const(0)
        for i in range(w):
            sel_l = sel_l + lhs[i] * is_comp[i]
            sel_r = sel_r + rhs[i] * is_comp[i]
        b.assert_eq(sel_l, lhs_limb)
        b.assert_eq(sel_r, rhs_limb)
    rec.less_than(lhs_limb, rhs_limb, 1, is_real)


def _eval_depth(b: SymbolicBuilder, cols: _Cols, depth: List[Expr], sel: Expr, out: List[Expr]):
    """eval_depth (src/lair/air.rs:105-131): the callee's depth bytes, proven smaller than the row's own."""
    dep_depth = [cols.next_aux() for _ in range(DEPTH_W)]
    witness = [cols.next_aux() for _ in range(DEPTH_LESS_THAN_SIZE)]
    rec = _ByteAirRecord()
    _assert_less_than(b, witness, dep_depth, depth, rec, sel)
    requires = [cols.next_require() for _ in range(DEPTH_LESS_THAN_REQUIRES)]
    rec.require_all(b, cols.nonce(), requires)
    out.extend(dep_depth)


def _eval_block(b, block: Block, cols: _Cols, sel: Expr, vmap, top: Toplevel, depth=()):
    for op in block.ops:
        _eval_op(b, op, cols, sel, vmap, top, depth)
    c = block.ctrl
    if c.kind == "return":
        s = cols.sel(c.args[0])
        for v in c.args[1]:
            with b.when(s):
                b.assert_eq(vmap[v][0], cols.next_output())
        return
    blocks = [blk for _, blk in c.args[1]] + ([c.args[2]] if c.args[2] is not None else [])
    n, st = len(vmap), cols.save()
    for blk in blocks:
        _eval_block(b, blk, cols, _return_sel(blk, cols), vmap, top, depth)
        del vmap[n:]
        cols.restore(st)


def _eval_op(b: SymbolicBuilder, op: Op, cols: _Cols, sel: Expr, vmap, top: Toplevel, depth=()):
    k, a = op.kind, op.args
    nonce = cols.nonce()
    if k == "assert_ne":
        coeffs = [cols.next_aux() for _ in a[0]]
        acc = b.const(0)
        for cf, x, y in zip(coeffs, a[0], a[1]):
            acc = acc + cf * (vmap[x][0] - vmap[y][0])
        with b.when(sel):
            b.assert_one(acc)
    elif k == "assert_eq":
        for x, y in zip(a[0], a[1]):
            with b.when(sel):
                b.assert_eq(vmap[x][0], vmap[y][0])
    elif k == "contains":
        bv = vmap[a[1]][0]
        acc = vmap[a[0][0]][0] - bv
        for x in a[0][1:]:
            aux = cols.next_aux()
            with b.when(sel):
                b.assert_eq(acc * (vmap[x][0] - bv), aux)
            acc = aux
        with b.when(sel):
            b.assert_zero(acc)
    elif k == "const":
        vmap.append((b.const(a[0]), 0))
    elif k in ("add", "sub"):
        (x, dx), (y, dy) = vmap[a[0]], vmap[a[1]]
        vmap.append((x + y if k == "add" else x - y, max(dx, dy)))
    elif k == "mul":
        (x, dx), (y, dy) = vmap[a[0]], vmap[a[1]]
        # a witness column exists iff the layout allocated one (func_chip.rs:199-209); the
        # reference's eval keys on const-ness instead (air.rs:345-358), which coincides whenever
        # no operand is a non-constant of degree 0.
        if dx + dy < 2:
            vmap.append((x * y, dx + dy))
        else:
            c = cols.next_aux()
            with b.when(sel):
                b.assert_eq(x * y, c)
            vmap.append((c, 1))
    elif k == "inv":
        x, dx = vmap[a[0]]
        if dx == 0:
            vmap.append((b.const(pow(x.const, P - 2, P)), 0))
        else:
            c = cols.next_aux()
            with b.when(sel):
                b.assert_one(x * c)
            vmap.append((c, 1))
    elif k == "not":
        x, dx = vmap[a[0]]
        if dx == 0:
            vmap.append((b.const(1 if x.const == 0 else 0), 0))
        else:
            d, r = cols.next_aux(), cols.next_aux()
            with b.when(sel):
                b.assert_zero(x * r)
            with b.when(sel):
                b.assert_one(x * d + r)
            vmap.append((r, 1))
    elif k == "call":
        callee = top.funcs[a[0]]
        out = []
        for _ in range(callee.output_size):
            o = cols.next_aux()
            vmap.append((o, 1))
            out.append(o)
        inp = [vmap[i][0] for i in a[1]]
        pn, pc, ci = cols.next_require()
        if callee.partial:
            _eval_depth(b, cols, list(depth), sel, out)
        b.require([b.const(CALL_TAG), b.const(a[0])] + inp + out, nonce, pn, pc, ci, sel)
    elif k == "preimg":
        callee = top.funcs[a[0]]
        inp = []
        for _ in range(callee.input_size):
            i = cols.next_aux()
            vmap.append((i, 1))
            inp.append(i)
        out = [vmap[o][0] for o in a[1]]
        pn, pc, ci = cols.next_require()
        if callee.partial:
            _eval_depth(b, cols, list(depth), sel, out)
        b.require([b.const(CALL_TAG), b.const(a[0])] + inp + out, nonce, pn, pc, ci, sel)
    elif k == "store":
        ptr = cols.next_aux()
        vmap.append((ptr, 1))
        vals = [vmap[i][0] for i in a[0]]
        pn, pc, ci = cols.next_require()
        b.require([b.const(MEMORY_TAG), ptr] + vals, nonce, pn, pc, ci, sel)
    elif k == "load":
        ptr = vmap[a[1]][0]
        vals = []
        for _ in range(a[0]):
            o = cols.next_aux()
            vmap.append((o, 1))
            vals.append(o)
        pn, pc, ci = cols.next_require()
        b.require([b.const(MEMORY_TAG), ptr] + vals, nonce, pn, pc, ci, sel)
    elif k == "range_u8":
        xs = [vmap[x][0] for x in a[0]]
        n_req = (len(xs) + 1) // 2
        reqs = [cols.next_require() for _ in range(n_req)]
        for j in range(n_req):
            i1 = xs[2 * j]
            i2 = xs[2 * j + 1] if 2 * j + 1 < len(xs) else b.const(0)
            pn, pc, ci = reqs[j]
            b.require([b.const(BYTE_TAG), b.const(1), i1, i2], nonce, pn, pc, ci, sel)
    elif k == "extern_call":
        # Op::ExternCall (src/lair/air.rs:461-482): witness columns first, then the require triples
        chip = top.chips[a[0]]
        inp = [vmap[i][0] for i in a[1]]
        witness = [cols.next_aux() for _ in range(chip.witness_size)]
        requires = [cols.next_require() for _ in range(chip.require_size)]
        for o in _eval_chip(b, chip, sel, inp, witness, nonce, requires):
            vmap.append((o, 1))
    else:
        raise ValueError(k)


def _eval_chip(b: SymbolicBuilder, chip, is_real: Expr, inp, witness, nonce, requires) -> List[Expr]:
    """Chipset::eval dispatch (src/core/chipset.rs:133-171)."""
    if chip.kind in (1, 2, 3):
        return eval_poseidon_chipset(b, {1: 24, 2: 32, 3: 40}[chip.kind], is_real, inp, witness)
    if chip.kind in (4, 5, 6, 7, 8, 9):
        return eval_u64_chipset(b, chip.kind, is_real, inp, witness, nonce, requires)
    if chip.kind == 10:
        return eval_big_num_less_than(b, is_real, inp, witness, nonce, requires)
    raise NotImplementedError(f"AIR capture of extern chip {chip.name!r} (kind {chip.kind})")


def _assert_add(b: SymbolicBuilder, lhs, rhs, out, is_real: Expr):
    """AddWitness::assert_add (src/gadgets/unsigned/add.rs:19-60): byte-wise lhs + rhs = out with boolean carries"""
    base_inv = pow(256, P - 2, P)
    with b.when(is_real):
        carry = b.const(0)
        for o, x, y in zip(out, lhs, rhs):
            total = x + y + carry
            carry = (total - o) * base_inv
            b.assert_bool(carry)
    return carry


def _eval_compare(b: SymbolicBuilder, wit, lhs, rhs, rec: "_ByteAirRecord", is_real: Expr):
    """CompareWitness::eval (src/gadgets/unsigned/cmp.rs:49-128) -> (is_less_than, is_equal);
    wit = is_comp[W] ++ [lhs_comp_limb, rhs_comp_limb, comp_diff_inv, is_less_than]"""
    w = len(lhs)
    is_comp, lhs_limb, rhs_limb, diff_inv, is_lt = wit[:w], wit[w], wit[w + 1], wit[w + 2], wit[w + 3]
    with b.when(is_real):
        is_equal = b.const(1)
        for i in reversed(range(w)):
            b.assert_bool(is_comp[i])
            is_equal = is_equal - is_comp[i]
            with b.when(is_equal):
                b.assert_eq(lhs[i], rhs[i])
        b.assert_bool(is_equal)
        sel_l, sel_r = b.const(0), b.const(0)
        for i in range(w):
            sel_l = sel_l + lhs[i] * is_comp[i]
            sel_r = sel_r + rhs[i] * is_comp[i]
        b.assert_eq(sel_l, lhs_limb)
        b.assert_eq(sel_r, rhs_limb)
        is_different = 1 - is_equal
        b.assert_eq((lhs_limb - rhs_limb) * diff_inv, is_different)
    rec.less_than(lhs_limb, rhs_limb, is_lt, is_real)
    return is_lt, is_equal


def _eval_product(b: SymbolicBuilder, wit, lhs, rhs, rec: "_ByteAirRecord", is_real: Expr):
    """Product::eval = MulWitness::assert_mul + range-checked result (src/gadgets/unsigned/mul.rs:68-164);
    wit = carry[W] ++ result[W]"""
    w = len(lhs)
    carry, result = wit[:w], list(wit[w:2 * w])
    products = [b.const(0) for _ in range(w)]
    for i in range(w):
        for j in range(w):
            if i + j < w:
                products[i + j] = products[i + j] + lhs[i] * rhs[j]
    carry_prev = b.const(0)
    for k in range(w):
        rec.range_check_u16(carry[k], is_real)
        out = result[k] + carry[k] * 256
        with b.when(is_real):
            b.assert_eq(products[k] + carry_prev, out)
        carry_prev = carry[k]
    rec.range_check_u8_iter(result, is_real)
    return result


def _eval_field_to_word32(b: SymbolicBuilder, wit, field: Expr, rec: "_ByteAirRecord", is_real: Expr):
    """FieldToWord32::eval (src/gadgets/unsigned/field.rs:36-140): wit = [is_msb_less_than] ++ result[4]; the
    bytes recompose to the field element and are its canonical (< p = 0x78000001) representation"""
    is_lt, word = wit[0], list(wit[1:5])
    with b.when(is_real):
        b.assert_bool(is_lt)
    recomposed = b.const(0)
    for i in reversed(range(4)):
        recomposed = recomposed * 256 + word[i]
    with b.when(is_real):
        b.assert_eq(field, recomposed)
    rec.less_than(word[3], 0x78, is_lt, is_real)
    with b.when(is_real * (1 - is_lt)):
        b.assert_eq(word[3], b.const(0x78))
        for i in range(3):
            b.assert_eq(word[i], b.const(0))
    rec.range_check_u8_iter(word, is_real)
    return word


def eval_big_num_less_than(b: SymbolicBuilder, is_real: Expr, inp, witness, nonce, requires) -> List[Expr]:
    """BigNum::LessThan (src/core/big_num.rs:75-97) = BigNumCompareWitness::eval (src/gadgets/big_num/cmp.rs:48-124):
    witness = is_comp[8], lhs_comp_limb, rhs_comp_limb, lhs_comp_word (5), rhs_comp_word (5), CompareWitness<4> (8)"""
    lhs, rhs = list(inp[:8]), list(inp[8:16])
    is_comp, lhs_limb, rhs_limb = witness[:8], witness[8], witness[9]
    rec = _ByteAirRecord()
    with b.when(is_real):
        is_equal = b.const(1)
        for i in reversed(range(8)):
            b.assert_bool(is_comp[i])
            is_equal = is_equal - is_comp[i]
            with b.when(is_equal):
                b.assert_eq(lhs[i], rhs[i])
        b.assert_bool(is_equal)
        sel_l, sel_r = b.const(0), b.const(0)
        for i in range(8):
            sel_l = sel_l + lhs[i] * is_comp[i]
            sel_r = sel_r + rhs[i] * is_comp[i]
        b.assert_eq(sel_l, lhs_limb)
        b.assert_eq(sel_r, rhs_limb)
    lhs_word = _eval_field_to_word32(b, witness[10:15], lhs_limb, rec, is_real)
    rhs_word = _eval_field_to_word32(b, witness[15:20], rhs_limb, rec, is_real)
    lt, word_equal = _eval_compare(b, witness[20:28], lhs_word, rhs_word, rec, is_real)
    with b.when(is_real):
        b.assert_eq(is_equal, word_equal)
    rec.require_all(b, nonce, requires)
    return [lt]


def eval_u64_chipset(b: SymbolicBuilder, kind: int, is_real: Expr, inp, witness, nonce, requires) -> List[Expr]:
    """U64::eval (src/core/u64.rs:170-222), kinds 4..9 = Add, Sub, LessThan, IsZero, Mul, DivRem: Sum64 / Diff64
    (src/gadgets/unsigned/add.rs:63-175), Compare64 (cmp.rs:49-128), IsZero64 (is_zero.rs:46-156), Product64
    (mul.rs:68-164), DivRem64 (div_rem.rs:62-103)."""
    in1, in2 = list(inp[:8]), list(inp[8:16])
    rec = _ByteAirRecord()
    if kind in (4, 5):
        result = list(witness[:8])
        rec.range_check_u8_iter(result, is_real)            # UncheckedWord::into_checked
        if kind == 4:
            _assert_add(b, in1, in2, result, is_real)
        else:                                               # result + rhs = lhs
            _assert_add(b, result, in2, in1, is_real)
        out = result
    elif kind == 6:
        out = [_eval_compare(b, witness[:12], in1, in2, rec, is_real)[0]]
    elif kind == 7:                                         # IsZeroOrEqual::eval_is_zero
        inverses, is_zero = witness[:8], witness[8]
        with b.when(is_real):
            b.assert_bool(is_zero)
        with b.when(is_real):
            lc = b.const(0)
            for x, inv in zip(in1, inverses):
                with b.when(is_zero):
                    b.assert_zero(x)
                lc = lc + x * inv
            b.assert_eq(lc, 1 - is_zero)
        out = [is_zero]
    elif kind == 8:
        out = _eval_product(b, witness[:16], in1, in2, rec, is_real)
    else:                                                   # DivRem::eval, (a, b) = (in1, in2)
        b_inv, q = witness[:8], list(witness[8:16])
        w_qb, w_r, w_lt, w_cmp = witness[16:32], list(witness[32:40]), witness[40:50], witness[50:62]
        with b.when(is_real):                               # IsZeroWitness::assert_non_zero(b)
            lc = b.const(0)
            for x, inv in zip(in2, b_inv):
                lc = lc + x * inv
            b.assert_one(lc)
        rec.range_check_u8_iter(q, is_real)
        qb = _eval_product(b, w_qb, q, in2, rec, is_real)
        rec.range_check_u8_iter(w_r, is_real)               # Diff::eval: r + qb = a
        _assert_add(b, w_r, qb, in1, is_real)
        _assert_less_than(b, w_lt, w_r, in2, rec, is_real)
        lt, eq_ = _eval_compare(b, w_cmp, qb, in1, rec, is_real)
        with b.when(is_real):
            b.assert_one(lt + eq_)
        out = q + w_r
    rec.require_all(b, nonce, requires)
    return out


def _external_linear(s: List[Expr]) -> List[Expr]:
    """Poseidon2ExternalMatrixGeneral (third-party p3-poseidon2): M4 = circ(2,3,1,1) on each 4-chunk,
    then every element gains the sum of its residue class mod 4 (structure: src/poseidon/air.rs:133-138)."""
    out = []
    for i in range(0, len(s), 4):
        x0, x1, x2, x3 = s[i:i + 4]
        t01, t23 = x0 + x1, x2 + x3
        t = t01 + t23
        t1, t3 = t + x1, t + x3
        out += [t1 + t01, t1 + 2 * x2, t3 + t23, t3 + 2 * x0]
    sums = [None] * 4
    for i, x in enumerate(out):
        sums[i & 3] = x if sums[i & 3] is None else sums[i & 3] + x
    return [x + sums[i & 3] for i, x in enumerate(out)]


def eval_poseidon_chipset(b: SymbolicBuilder, width: int, is_real: Expr, inp, witness) -> List[Expr]:
    """PoseidonChipset::eval (src/core/poseidon.rs:78-93) + Poseidon2Cols::eval
    (src/poseidon/wide/air.rs:11-124); witness = output[8] ++ Poseidon2Cols (wide/columns.rs:17-32)."""
    from .poseidon import constants
    diag, full_rc, part_rc = constants(width)
    rp = len(part_rc)
    output, w = witness[:8], witness[8:]
    assert len(inp) == width and len(w) == 17 * width + 2 * rp - 1
    ext_state = [w[r * width:(r + 1) * width] for r in range(8)]
    ext_sbox = [w[(8 + r) * width:(9 + r) * width] for r in range(8)]
    o = 16 * width
    int_init, int_state0, int_sbox = w[o:o + width], w[o + width:o + width + rp - 1], w[o + width + rp - 1:]
    assert len(int_sbox) == rp

    state = _external_linear([is_real * x for x in inp])

    def external_round(state, r):
        for i in range(width):
            b.assert_eq(state[i], ext_state[r][i])
        state = [ext_state[r][i] + is_real * int(full_rc[r][i]) for i in range(width)]
        for i in range(width):
            b.assert_eq(state[i] * state[i] * state[i], ext_sbox[r][i])
            state[i] = state[i] * (ext_sbox[r][i] * ext_sbox[r][i])
        return _external_linear(state)

    for r in range(4):
        state = external_round(state, r)
    for r in range(rp):
        if r == 0:
            for i in range(width):
                b.assert_eq(state[i], int_init[i])
            state = list(int_init)
        else:
            b.assert_eq(state[0], int_state0[r - 1])
            state[0] = int_state0[r - 1]
        x = state[0] + is_real * int(part_rc[r])
        b.assert_eq(x * x * x, int_sbox[r])
        state[0] = x * (int_sbox[r] * int_sbox[r])
        total = state[0]
        for v in state[1:]:
            total = total + v
        state = [v * int(diag[i]) + total for i, v in enumerate(state)]     # config.rs:104-119
    for r in range(4, 8):
        state = external_round(state, r)
    for st, out in zip(state, output):
        b.assert_eq(st, is_real * out)
    return list(output)


def eval_mem(b: SymbolicBuilder, length: int):
    """MemChip::eval (src/lair/memory.rs:71-109)"""
    is_real, ptr, last_nonce, last_count = b.local(0), b.local(1), b.local(2), b.local(3)
    values = [b.local(4 + i) for i in range(length)]
    is_real_next, ptr_next = b.next(0), b.next(1)
    b.assert_bool(is_real)
    is_real_transition = is_real_next * b.is_transition()
    with b.when(is_real_transition):
        b.assert_one(is_real)
    with b.when_first_row():
        with b.when(is_real):
            b.assert_one(ptr)
    with b.when(is_real_transition):
        b.assert_eq(ptr + 1, ptr_next)
    b.provide([b.const(MEMORY_TAG), ptr] + values, last_nonce, last_count, is_real)


def eval_bytes(b: SymbolicBuilder):
    """BytesChip::eval (src/gadgets/bytes/trace.rs:116-143)"""
    i1, i2, lt, and_, xor, or_ = (b.prep(i) for i in range(6))
    is_real = b.local(0)
    b.assert_bool(is_real)
    input_u16 = i1 + i2 * 256
    tag = b.const(BYTE_TAG)
    relations = [
        [tag, b.const(1), i1, i2],
        [tag, b.const(2), input_u16],
        [tag, b.const(3), i1, i2, lt],
        [tag, b.const(4), i1, i2, and_],
        [tag, b.const(5), i1, i2, xor],
        [tag, b.const(6), i1, i2, or_],
    ]
    for k, rel in enumerate(relations):
        b.provide(rel, b.local(1 + 2 * k), b.local(2 + 2 * k), is_real)


def eval_entrypoint(b: SymbolicBuilder, func_idx: int, num_pv: int):
    """LairChip::Entrypoint eval (src/lair/lair_chip.rs:166-191)"""
    pvs = [b.local(i) for i in range(num_pv)]
    for i, a in enumerate(pvs):
        b.assert_eq(a, b.public_value(i))
    b.require([b.const(CALL_TAG), b.const(func_idx)] + pvs, 0, 0, 0, 1, 1)


# ------------------------------------------------------------------------------------------------
# Compilation of the captured AIR into the flat descriptor
# ------------------------------------------------------------------------------------------------

I_ADD, I_SUB, I_MUL, I_NEG, I_ASSERT = 1, 2, 3, 4, 5
OPND_REG, OPND_LEAF = 0, 1   # operand = (tag << 28) | payload


def _encode_leaf(kind: int, idx: int) -> int:
    return ((1 + kind) << 28) | idx


class _Compiler:
    def __init__(self):
        self.consts: Dict[int, int] = {}
        self.code: List[int] = []
        self.node_reg: Dict[int, int] = {}
        self.n_instr = 0

    def const_idx(self, c: int) -> int:
        c %= P
        if c not in self.consts:
            self.consts[c] = len(self.consts)
        return self.consts[c]


def _compile_constraints(constraints: List[Expr]) -> Tuple[List[int], List[int], int]:
    """-> (code words, const table, n_regs).  Straight-line program, one pass for use counts,
    one pass emitting with a free-list register allocator."""
    uses: Dict[int, int] = {}
    nodes: Dict[int, Expr] = {}

    def visit(e: Expr):
        uses[e.id] = uses.get(e.id, 0) + 1
        if e.id in nodes:
            return
        nodes[e.id] = e
        if e.kind != "aff":
            visit(e.a)
            if e.b is not None:
                visit(e.b)

    for c in constraints:
        visit(c)

    comp = _Compiler()
    code = comp.code
    free: List[int] = []
    n_regs = 0
    val: Dict[int, int] = {}      # node id -> operand word
    remaining = dict(uses)

    def alloc() -> int:
        nonlocal n_regs
        if free:
            return free.pop()
        n_regs += 1
        return n_regs - 1

    def release(e: Expr):
        remaining[e.id] -= 1
        if remaining[e.id] == 0:
            w = val.pop(e.id)
            if (w >> 28) == 0:
                free.append(w)

    def emit_affine(e: Expr) -> int:
        # single leaf with coeff 1 and no constant -> direct operand
        items = sorted(e.lin.items())
        if len(items) == 1 and items[0][1] == 1 and e.const == 0:
            (kind, idx), _ = items[0]
            return _encode_leaf(kind, idx)
        if not items:
            return _encode_leaf(L_CONST, comp.const_idx(e.const))
        r = alloc()
        first = True
        for (kind, idx), cf in items:
            leaf = _encode_leaf(kind, idx)
            if cf == 1:
                term = leaf
                if first:
                    if e.const:
                        code.extend([I_ADD, r, leaf, _encode_leaf(L_CONST, comp.const_idx(e.const))])
                    else:
                        code.extend([I_ADD, r, leaf, _encode_leaf(L_CONST, comp.const_idx(0))])
                else:
                    code.extend([I_ADD, r, r, term])
            elif cf == P - 1 and not first:
                code.extend([I_SUB, r, r, leaf])
            else:
                if first:
                    code.extend([I_MUL, r, leaf, _encode_leaf(L_CONST, comp.const_idx(cf))])
                    if e.const:
                        code.extend([I_ADD, r, r, _encode_leaf(L_CONST, comp.const_idx(e.const))])
                else:
                    t = alloc()
                    code.extend([I_MUL, t, leaf, _encode_leaf(L_CONST, comp.const_idx(cf))])
                    code.extend([I_ADD, r, r, t])
                    free.append(t)
            first = False
        return r

    def gen(e: Expr) -> int:
        if e.id in val:
            return val[e.id]
        if e.kind == "aff":
            w = emit_affine(e)
        else:
            wa = gen(e.a)
            wb = gen(e.b) if e.b is not None else 0
            release(e.a)
            if e.b is not None:
                release(e.b)
            r = alloc()
            op = {"add": I_ADD, "sub": I_SUB, "mul": I_MUL, "neg": I_NEG}[e.kind]
            code.extend([op, r, wa, wb])
            w = r
        val[e.id] = w
        return w

    import sys
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
    for c in constraints:
        w = gen(c)
        code.extend([I_ASSERT, 0, w, 0])
        release(c)
    consts = [0] * len(comp.consts)
    for c, i in comp.consts.items():
        consts[i] = c
    return code, consts, max(n_regs, 1)


def _encode_affine(e: Expr) -> List[int]:
    assert e.is_affine(), "interaction values must be affine in the trace columns"
    items = sorted(e.lin.items())
    out = [len(items), e.const % P]
    for (kind, idx), cf in items:
        assert kind in (L_MAIN, L_PREP), "interactions may only reference local main/preprocessed cells"
        out += [0 if kind == L_MAIN else 1, idx, cf]
    return out


KIND_FUNC, KIND_MEM, KIND_BYTES, KIND_ENTRY = 0, 1, 2, 3


@dataclass
class ChipDesc:
    name: str
    kind: int
    index: int            # func index / mem table index / entry func index
    width: int
    prep_width: int
    n_constraints: int
    max_degree: int
    code: List[int]
    consts: List[int]
    n_regs: int
    sends: List[Interaction]
    receives: List[Interaction]

    @property
    def log_quotient_degree(self) -> int:
        d = max(self.max_degree, 3)
        return (d - 2).bit_length()      # log2_ceil(d - 1)

    def perm_width(self) -> int:
        n = len(self.sends) + len(self.receives)
        batch = 1 << self.log_quotient_degree
        return (n + batch - 1) // batch + 1

    def serialize(self) -> List[int]:
        name = self.name.encode()
        name_words = list(np.frombuffer(name + b"\0" * (-len(name) % 4), dtype=np.uint32)) if name else []
        out = [self.kind, self.index, self.width, self.prep_width, self.n_constraints, self.log_quotient_degree,
               self.n_regs, len(name), len(name_words)] + [int(x) for x in name_words]
        out += [len(self.consts)] + self.consts
        out += [len(self.code) // 4] + self.code
        for lst in (self.sends, self.receives):
            out.append(len(lst))
            for it in lst:
                out += [it.kind, len(it.values)]
                for v in it.values:
                    out += _encode_affine(v)
                out += _encode_affine(it.multiplicity)
        return out


def _finish(name, kind, index, b: SymbolicBuilder) -> ChipDesc:
    code, consts, n_regs = _compile_constraints(b.constraints)
    max_deg = max([c.degree for c in b.constraints] + [0])
    return ChipDesc(name, kind, index, b.width, b.prep_width, len(b.constraints), max_deg, code, consts, n_regs,
                    b.sends, b.receives)


def build_chip_vector(entry: FuncChip) -> List[ChipDesc]:
    """build_lair_chip_vector + Chip::new (src/lair/lair_chip.rs:196-231): Entrypoint, every func
    of the toplevel, the six memory tables, the byte table (named "CPU", lair_chip.rs:90-92)."""
    top, f = entry.toplevel, entry.func
    num_pv = f.input_size + f.output_size + (DEPTH_W if f.partial else 0)
    chips = []
    b = SymbolicBuilder(num_pv, 0, num_pv)
    eval_entrypoint(b, f.index, num_pv)
    chips.append(_finish(f"Entrypoint[{f.index}]", KIND_ENTRY, f.index, b))
    for fc in FuncChip.from_toplevel(top):
        b = SymbolicBuilder(fc.width(), 0, num_pv)
        eval_func(b, fc.func, top)
        chips.append(_finish(f"Func[{fc.func.name}]", KIND_FUNC, fc.func.index, b))
    for i, ln in enumerate(MEM_TABLE_SIZES):
        b = SymbolicBuilder(4 + ln, 0, num_pv)
        eval_mem(b, ln)
        chips.append(_finish(f"Mem[{ln}-wide]", KIND_MEM, i, b))
    b = SymbolicBuilder(13, 6, num_pv)
    eval_bytes(b)
    chips.append(_finish("CPU", KIND_BYTES, 0, b))
    return chips


def serialize_machine(chips: List[ChipDesc], num_pv: int) -> np.ndarray:
    out = [0x4d414348, len(chips), num_pv]   # 'MACH'
    for c in chips:
        body = c.serialize()
        out += [len(body)] + body
    return np.asarray(out, dtype=np.uint32)
