"""Builtin-free Lurk expressions evaluated by the restated evaluator (lurk_b200/lurk_eval.py::lurk_core_toplevel: the
reference's eval / apply / eval_begin / eval_list / env_lookup / ingress / egress, statement for statement): the
expression is interned by the ZStore, ingressed, evaluated in the empty environment, egressed, and the result must be
the ZPtr Lurk's semantics give.  Functions are written as literal Fun objects (the `lambda` builtin is not restated)."""
import pytest

import goldens
from lurk_b200.execute import QueryRecord, execute_by_name
from lurk_b200.zstore import TAG, ZStore

ERR = dict(UnboundVar=0, ApplyNonFunc=3, ArgsNotList=7, Todo=21)


def _cases(z):
    x, y, r, rest = z.lurk_sym("x"), z.lurk_sym("y"), z.lurk_sym("r"), z.lurk_sym("&rest")
    env0 = z.null(TAG["Env"])
    fun = lambda params, body, env=env0: z.tuple110(TAG["Fun"], z.list(params), z.list(body), env)
    second = fun([x, y], [y])                 # (lambda (x y) y)
    err = lambda e: (TAG["Err"], [e] + [0] * 7)
    zp = lambda p: (p.tag, list(p.digest))
    return {
        "a number evaluates to itself": (z.num(5), zp(z.num(5))),
        "a string evaluates to itself": (z.string("hi"), zp(z.string("hi"))),
        "nil evaluates to itself": (z.nil(), zp(z.nil())),
        "unbound variable": (x, err(ERR["UnboundVar"])),
        "saturated application": (z.list([second, z.num(5), z.u64(7)]), zp(z.u64(7))),
        "the body is a sequence": (z.list([fun([x], [z.num(1), z.num(2), x]), z.num(9)]), zp(z.num(9))),
        "undersaturated application returns a closure":
            (z.list([second, z.num(5)]), zp(fun([y], [y], z.tuple110(TAG["Env"], x, z.num(5), env0)))),
        "oversaturated application":                      # ((lambda (x) (lambda (y) y)) 1 2) with the inner one literal
            (z.list([fun([x], [fun([y], [y])]), z.num(1), z.num(2)]), zp(z.num(2))),
        "&rest collects the remaining arguments":
            (z.list([fun([rest, r], [r]), z.num(1), z.num(2), z.num(3)]), zp(z.list([z.num(1), z.num(2), z.num(3)]))),
        "nested application in an argument":
            (z.list([second, z.num(0), z.list([second, z.num(1), z.string("ok")])]), zp(z.string("ok"))),
        "applying a number": (z.list([z.num(5), z.num(1)]), err(ERR["ApplyNonFunc"])),
        "an error in an argument propagates": (z.list([second, x, z.num(1)]), err(ERR["UnboundVar"])),
        "a fixed point evaluates its body in the extended environment":    # Fix(body = f, binds = ((f 42)), env0)
            (z.tuple110(TAG["Fix"], z.lurk_sym("f"), z.list([z.list([z.lurk_sym("f"), z.num(42)])]), env0), zp(z.num(42))),
        "cons": (z.list([z.builtin("cons"), z.num(1), z.num(2)]), zp(z.cons(z.num(1), z.num(2)))),
        "atom": (z.list([z.builtin("atom"), z.num(1)]), zp(z.t())),
    }


CASE_NAMES = list(_cases(ZStore()).keys())


def run(top, z, expr):
    q = QueryRecord.new(top)
    z.inject(q, top)
    out = [int(v) for v in execute_by_name(top, "lurk_eval_demo", [expr.tag] + [0] * 7 + list(expr.digest), q)]
    return q, (out[0], out[1:])


@pytest.mark.parametrize("name", CASE_NAMES)
def test_builtin_free_expressions_evaluate(name):
    top = goldens.toplevel_for("lurk_core")
    z = ZStore()
    expr, want = _cases(z)[name]
    _, got = run(top, z, expr)
    assert got == (want[0], list(want[1])), name


def _oracle_record(oracle, top, z, expr):
    ot = oracle.OracleToplevel(top.serialize())
    rec = oracle.OracleRecord(ot)
    for name, width in (("hash3", 24), ("hash4", 32), ("hash5", 40)):
        if z.hashes[width]:
            rec.inject_inv_queries(top.func_by_name(name).index, [list(k) for k in z.hashes[width]],
                                   [list(v) for v in z.hashes[width].values()])
    f = top.func_by_name("lurk_eval_demo")
    out = rec.execute(f.index, [expr.tag] + [0] * 7 + list(expr.digest), f.output_size)
    return ot, rec, f, [int(v) for v in out]


@pytest.mark.parametrize("name", ["saturated application", "&rest collects the remaining arguments"])
def test_evaluation_proves_and_verifies_on_the_oracle(oracle, name):
    from test_oracle_stark import oracle_machine
    top = goldens.toplevel_for("lurk_core")
    z = ZStore()
    expr, want = _cases(z)[name]
    ot, rec, f, out = _oracle_record(oracle, top, z, expr)
    assert (out[0], out[1:]) == (want[0], list(want[1]))
    _, om, _ = oracle_machine(oracle, top, "lurk_eval_demo")
    proof = om.prove(rec)
    assert om.verify(proof), om.last_error


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["saturated application", "&rest collects the remaining arguments",
                                  "a fixed point evaluates its body in the extended environment"])
def test_evaluation_proves_bit_exact_on_the_gpu(oracle, name):
    # the whole path on a real Lurk evaluation: 23 func chips (partial funcs with depth provenance, preimages of
    # hash4 / hash5, memory of four lengths), the GPU proof equal to the oracle's, accepted by the product verifier
    import numpy as np
    from lurk_b200.execute import Shard
    from lurk_b200.machine import SphinxCoreOpts, new_machine
    top = goldens.toplevel_for("lurk_core")
    z = ZStore()
    expr, want = _cases(z)[name]
    m = new_machine(top, "lurk_eval_demo", opts=SphinxCoreOpts(num_queries=6, proof_of_work_bits=6))
    pk, vk = m.setup()
    q, got = run(top, z, expr)
    assert got == (want[0], list(want[1]))
    ot, rec, f, _ = _oracle_record(oracle, top, z, expr)
    om = oracle.OracleMachine(ot, m.entry.index, m.blob, 6, 6)
    proof = m.prove(pk, Shard.new(q))
    assert np.array_equal(proof.words, om.prove(rec))
    m.verify(vk, proof)


def _programs(z):
    """Lurk programs as S-expressions (lists of ZPtrs): (name, expression, expected ZPtr)"""
    B, L, U, N = z.builtin, z.list, z.u64, z.num
    sym = lambda s: z.symbol(["lurk-user", s])
    n, x, fib, fact, acc = sym("n"), sym("x"), sym("fib"), sym("fact"), sym("acc")
    fib_def = L([B("letrec"), L([L([fib, L([B("lambda"), L([n]),
                L([B("if"), L([B("<"), n, U(2)]), n,
                   L([B("+"), L([fib, L([B("-"), n, U(1)])]), L([fib, L([B("-"), n, U(2)])])])])])])]),
                L([fib, U(10)])])
    return [
        ("(+ 1 2) on field elements", L([B("+"), N(1), N(2)]), N(3)),
        ("(* 6 7) on u64", L([B("*"), U(6), U(7)]), U(42)),
        ("(/ 7 0) on u64", L([B("/"), U(7), U(0)]), ZPtr_err(9)),                      # DivByZero
        ("(if (< 1 2) 10 20)", L([B("if"), L([B("<"), U(1), U(2)]), N(10), N(20)]), N(10)),
        ("(if (> 1 2) 10 20)", L([B("if"), L([B(">"), U(1), U(2)]), N(10), N(20)]), N(20)),
        ("((lambda (x) (* x x)) 7)", L([L([B("lambda"), L([x]), L([B("*"), x, x])]), N(7)]), N(49)),
        ("(let ((x 5)) (+ x 1))", L([B("let"), L([L([x, N(5)])]), L([B("+"), x, N(1)])]), N(6)),
        ("(car (cdr (list 1 2 3)))", L([B("car"), L([B("cdr"), L([B("list"), N(1), N(2), N(3)])])]), N(2)),
        ("(quote (1 2))", L([B("quote"), L([N(1), N(2)])]), L([N(1), N(2)])),
        ("(eq (cons 1 2) (cons 1 2))", L([B("eq"), L([B("cons"), N(1), N(2)]), L([B("cons"), N(1), N(2)])]), z.t()),
        ("(open (commit 123))", L([B("open"), L([B("commit"), N(123)])]), N(123)),
        ("(commit 123) is the reference's KAT-1", L([B("commit"), N(123)]), z.commit(N(123))),
        ("(atom (cons 1 2))", L([B("atom"), L([B("cons"), N(1), N(2)])]), z.nil()),
        ("(u64 #\\a)", L([B("u64"), z.char("a")]), U(97)),
        ("(char 98)", L([B("char"), U(98)]), z.char("b")),
        ("(emit 5)", L([B("emit"), N(5)]), N(5)),
        ("(bignum (commit 123))", L([B("bignum"), L([B("commit"), N(123)])]), ZPtr_tag(TAG["BigNum"], z.commit(N(123)))),
        ("(letrec ((fib ...)) (fib 10)) = 55  [benches/fib.rs shape]", fib_def, U(55)),
    ]


def ZPtr_tag(tag, z):
    from lurk_b200.zstore import ZPtr
    return ZPtr(tag, z.digest)


def ZPtr_err(code):
    from lurk_b200.zstore import ZPtr
    return ZPtr(TAG["Err"], (code,) + (0,) * 7)


PROGRAM_NAMES = [p[0] for p in _programs(ZStore())]


@pytest.mark.parametrize("name", PROGRAM_NAMES)
def test_lurk_programs_evaluate(name):
    # the restated evaluator (all 39 funcs of the reference's native toplevel) on real Lurk programs, incl. the recursive fib of
    # benches/fib.rs; (commit 123) must give the digest the reference's own test pins (KAT-1, zstore.rs:1008-1019)
    top = goldens.toplevel_for("lurk_core")
    z = ZStore()
    expr, want = next((e, w) for n, e, w in _programs(z) if n == name)
    _, got = run(top, z, expr)
    assert got == (want.tag, list(want.digest)), name
    if name.startswith("(commit 123)"):
        assert want.to_biguint() == int(goldens.KNOWN_ANSWERS["KAT-1"]["digest"], 16)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["(letrec ((fib ...)) (fib 10)) = 55  [benches/fib.rs shape]", "(open (commit 123))"])
def test_lurk_programs_prove_bit_exact_on_the_gpu(oracle, name):
    # benches/fib.rs in miniature: the Lurk program is interned, ingressed, evaluated by the restated evaluator (39 func
    # chips, the u64 chipsets with their byte lookups, hash3/4/5 with preimages, partial funcs with depth provenance),
    # and the GPU proof equals the oracle's word for word and verifies
    import numpy as np
    from lurk_b200.execute import Shard
    from lurk_b200.machine import SphinxCoreOpts, new_machine
    top = goldens.toplevel_for("lurk_core")
    z = ZStore()
    expr, want = next((e, w) for n, e, w in _programs(z) if n == name)
    m = new_machine(top, "lurk_eval_demo", opts=SphinxCoreOpts(num_queries=6, proof_of_work_bits=6))
    pk, vk = m.setup()
    q, got = run(top, z, expr)
    assert got == (want.tag, list(want.digest))
    ot, rec, f, out = _oracle_record(oracle, top, z, expr)
    assert out == [want.tag] + list(want.digest)
    om = oracle.OracleMachine(ot, m.entry.index, m.blob, 6, 6)
    proof = m.prove(pk, Shard.new(q))
    assert np.array_equal(proof.words, om.prove(rec))
    m.verify(vk, proof)


def _run_main(top, z, expr, env):
    # lurk_main(full_expr_tag: [8], expr_digest: [8], env_digest: [8]): [16]  (eval_direct.rs:120-138)
    q = QueryRecord.new(top)
    z.inject(q, top)
    args = [expr.tag] + [0] * 7 + list(expr.digest) + list(env.digest)
    out = [int(v) for v in execute_by_name(top, "lurk_main", args, q)]
    assert out[1:8] == [0] * 7
    return q, (out[0], out[8:])


def test_lurk_main_is_the_entry_point_of_the_packaged_toplevel():
    # build_lurk_toplevel() (lurk_b200/lurk_eval.py): the reference's entry point takes the expression AND the environment
    # by digest; run fib(10) in the empty environment and a variable in a two-binding environment through it
    from lurk_b200.lurk_eval import build_lurk_toplevel
    top = build_lurk_toplevel()
    z = ZStore()
    progs = {n: (e, w) for n, e, w in _programs(z)}
    expr, want = progs["(letrec ((fib ...)) (fib 10)) = 55  [benches/fib.rs shape]"]
    _, got = _run_main(top, z, expr, z.null(TAG["Env"]))
    assert got == (want.tag, list(want.digest))
    x, y = z.lurk_sym("x"), z.lurk_sym("y")
    env = z.tuple110(TAG["Env"], x, z.num(7), z.tuple110(TAG["Env"], y, z.u64(9), z.null(TAG["Env"])))
    for sym, val in ((x, z.num(7)), (y, z.u64(9))):
        _, got = _run_main(top, z, sym, env)
        assert got == (val.tag, list(val.digest))
    plus = z.list([z.builtin("+"), x, z.num(35)])
    _, got = _run_main(top, z, plus, env)
    assert got == (TAG["Num"], list(z.num(42).digest))
