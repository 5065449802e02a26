"""Lair programs shipped with the package (workloads of bench.py and the smoke test).

`hash_toplevel` wires the Poseidon2 hasher chips the way the Lurk toplevel does (`hash3`/`hash4`/`hash5`
funcs around `extern_call`s, src/core/misc.rs:5-128 and src/core/chipset.rs:20-63); `cons_chain(n)` is
the Lurk-shaped load of bench.py --workload cons: one 655-wide hash4 row per list element, so the hash
chip dominates the trace exactly as it does when Lurk ingresses data."""
from .lair import (ChipInfo, Toplevel, array, call, const, extern_call, func, load, match, ret, store, sub)  # noqa: E501


def hash_toplevel():
    """Content-addressed hashing through the hasher chips, the way the Lurk toplevel wires them:
    `hash3`/`hash4`/`hash5` funcs wrapping `extern_call`s (src/core/misc.rs:5-128 hash funcs,
    chips src/core/chipset.rs:20-63; widths 493/655/815 pinned at src/core/eval_direct.rs:2053-2055).
      commit_cons(a, b, secret, tag) = hash3(secret ++ tag ++ hash4(a ++ b))   -- KAT-2's shape
                                                     (src/core/cli/tests/prove.lurk:12-17)
      cons_chain(n) = digest of the list (n n-1 ... 1): one hash4 row per element
      env_chain(n)  = the same through hash5 (fun/env shape: 16 + 16 + 8 words, zstore.rs:195-204)"""
    chips = [ChipInfo("hasher3", 24, 8, 457, 0, kind=1), ChipInfo("hasher4", 32, 8, 611, 0, kind=2),
             ChipInfo("hasher5", 40, 8, 763, 0, kind=3)]
    h3 = func("hash3", [("preimg", 24)], 8, [extern_call([("img", 8)], "hasher3", ["preimg"]), ret("img")])
    h4 = func("hash4", [("preimg", 32)], 8, [extern_call([("img", 8)], "hasher4", ["preimg"]), ret("img")])
    h5 = func("hash5", [("preimg", 40)], 8, [extern_call([("img", 8)], "hasher5", ["preimg"]), ret("img")])
    commit_cons = func("commit_cons", [("a", 16), ("b", 16), ("secret", 8), ("tag", 8)], 8, [
        call([("d", 8)], "hash4", ["a", "b"]),
        call([("c", 8)], "hash3", ["secret", "tag", "d"]),
        ret("c"),
    ])
    cons_chain = func("cons_chain", ["n"], 8, [
        array("nil", [0] * 8),
        match("n", [(0, [ret("nil")])]),
        const("one", 1),
        sub("pred", "n", "one"),
        call([("tail", 8)], "cons_chain", ["pred"]),
        array("num_tag", [1, 0, 0, 0, 0, 0, 0, 0]),       # Tag::Num, src/core/tag.rs:23-39
        array("cons_tag", [11, 0, 0, 0, 0, 0, 0, 0]),     # Tag::Cons
        array("pad7", [0] * 7),
        call([("d", 8)], "hash4", ["num_tag", "n", "pad7", "cons_tag", "tail"]),
        ret("d"),
    ])
    env_chain = func("env_chain", ["n"], 8, [
        array("nil", [0] * 8),
        match("n", [(0, [ret("nil")])]),
        const("one", 1),
        sub("pred", "n", "one"),
        call([("tail", 8)], "env_chain", ["pred"]),
        array("sym_tag", [6, 0, 0, 0, 0, 0, 0, 0]),
        array("num_tag", [1, 0, 0, 0, 0, 0, 0, 0]),
        array("pad7", [0] * 7),
        call([("d", 8)], "hash5", ["sym_tag", "n", "pad7", "num_tag", "pred", "pad7", "tail"]),
        ret("d"),
    ])
    return Toplevel([commit_cons, cons_chain, env_chain, h3, h4, h5], chips)


U64_CHIPS = [ChipInfo("u64_add", 16, 8, 8, 4, kind=4), ChipInfo("u64_sub", 16, 8, 8, 4, kind=5),
             ChipInfo("u64_lessthan", 16, 1, 12, 1, kind=6), ChipInfo("u64_iszero", 8, 1, 9, 0, kind=7),
             ChipInfo("u64_mul", 16, 8, 16, 12, kind=8), ChipInfo("u64_divrem", 16, 16, 62, 22, kind=9),
             ChipInfo("big_num_lessthan", 16, 1, 28, 7, kind=10)]


def u64_toplevel():
    """The u64 chipsets Lurk's evaluator uses for fib / lcs / sum (src/core/chipset.rs:34-52, src/core/u64.rs):
    `add`, `sub`, `less_than` are the funcs of the reference's own u64_add_test / u64_sub_test /
    u64_lessthan tests (src/core/u64.rs:237-420); `fibu(n)` = the n-th Fibonacci number as a wrapping u64 (8
    little-endian bytes, one u64_add row per step) and `maxu(a, b)` exercise them inside ordinary funcs."""
    from .lair import add as add_, match as match_
    add = func("add", [("a", 8), ("b", 8)], 8, [extern_call([("c", 8)], "u64_add", ["a", "b"]), ret("c")])
    sub_f = func("sub", [("a", 8), ("b", 8)], 8, [extern_call([("c", 8)], "u64_sub", ["a", "b"]), ret("c")])
    less_than = func("less_than", [("a", 8), ("b", 8)], 1, [extern_call(["c"], "u64_lessthan", ["a", "b"]), ret("c")])
    fibu = func("fibu", ["n"], 8, [
        array("zero", [0] * 8),
        array("one", [1] + [0] * 7),
        match_("n", [(0, [ret("zero")]), (1, [ret("one")])]),
        const("c1", 1),
        sub("n1", "n", "c1"),
        call([("a", 8)], "fibu", ["n1"]),
        sub("n2", "n1", "c1"),
        call([("b", 8)], "fibu", ["n2"]),
        extern_call([("r", 8)], "u64_add", ["a", "b"]),
        ret("r"),
    ])
    maxu = func("maxu", [("a", 8), ("b", 8)], 8, [
        extern_call(["lt"], "u64_lessthan", ["a", "b"]),
        match_("lt", [(1, [ret("b")])]),
        ret("a"),
    ])
    is_zero = func("is_zero", [("a", 8)], 1, [extern_call(["c"], "u64_iszero", ["a"]), ret("c")])
    mul_f = func("mul", [("a", 8), ("b", 8)], 8, [extern_call([("c", 8)], "u64_mul", ["a", "b"]), ret("c")])
    divrem = func("divrem", [("a", 8), ("b", 8)], 16, [extern_call([("qr", 16)], "u64_divrem", ["a", "b"]), ret("qr")])
    # factu(n) = n! as a wrapping u64: the counter itself is a u64 (decremented with u64_sub, tested with u64_iszero)
    factu = func("factu", [("n", 8)], 8, [
        array("one", [1] + [0] * 7),
        extern_call(["z"], "u64_iszero", ["n"]),
        match_("z", [(1, [ret("one")])]),
        extern_call([("pred", 8)], "u64_sub", ["n", "one"]),
        call([("m", 8)], "factu", ["pred"]),
        extern_call([("r", 8)], "u64_mul", ["n", "m"]),
        ret("r"),
    ])
    # gcdu(a, b) by Euclid with u64_divrem
    gcdu = func("gcdu", [("a", 8), ("b", 8)], 8, [
        extern_call(["z"], "u64_iszero", ["b"]),
        match_("z", [(1, [ret("a")])]),
        extern_call([("q", 8), ("r", 8)], "u64_divrem", ["a", "b"]),
        call([("g", 8)], "gcdu", ["b", "r"]),
        ret("g"),
    ])
    # BigNum::LessThan on two 8-limb digests read as base-p integers (src/core/big_num.rs; Lurk compares
    # commitments / big nums this way)
    big_lt = func("big_num_less_than", [("a", 8), ("b", 8)], 1,
                  [extern_call(["c"], "big_num_lessthan", ["a", "b"]), ret("c")])
    return Toplevel([add, sub_f, less_than, fibu, maxu, is_zero, mul_f, divrem, factu, gcdu, big_lt], U64_CHIPS)


def lurk_misc_toplevel():
    """The eleven helper funcs of the Lurk toplevel that wrap the extern chips, restated from src/core/misc.rs:5-128
    (hash3 / hash4 / hash5 are `invertible`; the u64 and big-num funcs take and return POINTERS to 8-word values:
    load, extern_call, store), plus `u64_demo`, a driver that stores two u64 values, runs every u64 func on the
    pointers and loads the results back (the reference reaches these funcs from eval_binop_num,
    src/core/eval_direct.rs:1205-1340).  Their trace widths are pinned by the reference's own test
    (eval_direct.rs:2053-2063): 493, 655, 815, 53, 53, 85, 166, 44, 26, 38, 78."""
    from .lair import if_
    chips = [ChipInfo("hasher3", 24, 8, 457, 0, kind=1), ChipInfo("hasher4", 32, 8, 611, 0, kind=2),
             ChipInfo("hasher5", 40, 8, 763, 0, kind=3)] + U64_CHIPS
    h3 = func("hash3", [("preimg", 24)], 8, [extern_call([("img", 8)], "hasher3", ["preimg"]), ret("img")], invertible=True)
    h4 = func("hash4", [("preimg", 32)], 8, [extern_call([("img", 8)], "hasher4", ["preimg"]), ret("img")], invertible=True)
    h5 = func("hash5", [("preimg", 40)], 8, [extern_call([("img", 8)], "hasher5", ["preimg"]), ret("img")], invertible=True)

    def ptr_binop(name, chip):   # fn name(a, b): [1] { a = load(a); b = load(b); c = extern_call(chip, a, b); store(c) }
        return func(name, ["a", "b"], 1, [
            load([("av", 8)], "a"), load([("bv", 8)], "b"),
            extern_call([("cv", 8)], chip, ["av", "bv"]),
            store("c", ["cv"]), ret("c")])
    u64_add, u64_sub, u64_mul = ptr_binop("u64_add", "u64_add"), ptr_binop("u64_sub", "u64_sub"), ptr_binop("u64_mul", "u64_mul")
    u64_divrem = func("u64_divrem", ["a", "b"], 2, [
        load([("av", 8)], "a"), load([("bv", 8)], "b"),
        extern_call([("qv", 8), ("rv", 8)], "u64_divrem", ["av", "bv"]),
        store("q", ["qv"]), store("r", ["rv"]), ret("q", "r")])
    u64_lessthan = func("u64_lessthan", ["a", "b"], 1, [
        load([("av", 8)], "a"), load([("bv", 8)], "b"),
        extern_call(["c"], "u64_lessthan", ["av", "bv"]), ret("c")])
    u64_iszero = func("u64_iszero", ["a"], 1, [load([("av", 8)], "a"), extern_call(["b"], "u64_iszero", ["av"]), ret("b")])
    digest_equal = func("digest_equal", ["a", "b"], 1, [
        load([("av", 8)], "a"), load([("bv", 8)], "b"),
        sub(("diff", 8), "av", "bv"),
        if_("diff", [const("zero", 0), ret("zero")]),
        const("one", 1), ret("one")])
    big_num_lessthan = func("big_num_lessthan", ["a", "b"], 1, [
        load([("av", 8)], "a"), load([("bv", 8)], "b"),
        extern_call(["c"], "big_num_lessthan", ["av", "bv"]), ret("c")])
    # driver: (a, b) as 8 little-endian bytes each -> a+b, a-b, a*b, a/b, a%b (8 words each), a<b, b==0, a==b, a <_bignum b
    u64_demo = func("u64_demo", [("a", 8), ("b", 8)], 44, [
        store("pa", ["a"]), store("pb", ["b"]),
        call(["ps"], "u64_add", ["pa", "pb"]), load([("s", 8)], "ps"),
        call(["pd"], "u64_sub", ["pa", "pb"]), load([("d", 8)], "pd"),
        call(["pm"], "u64_mul", ["pa", "pb"]), load([("m", 8)], "pm"),
        call(["z"], "u64_iszero", ["pb"]),
        call(["lt"], "u64_lessthan", ["pa", "pb"]),
        call(["e"], "digest_equal", ["pa", "pb"]),
        call(["blt"], "big_num_lessthan", ["pa", "pb"]),
        if_("z", [array("zeros", [0] * 16), ret("s", "d", "m", "zeros", "lt", "z", "e", "blt")]),
        call(["pq", "pr"], "u64_divrem", ["pa", "pb"]), load([("q", 8)], "pq"), load([("r", 8)], "pr"),
        ret("s", "d", "m", "q", "r", "lt", "z", "e", "blt")])
    return Toplevel([h3, h4, h5, u64_add, u64_sub, u64_mul, u64_divrem, u64_lessthan, u64_iszero, digest_equal,
                     big_num_lessthan, u64_demo], chips)




LURK_MISC_WIDTHS = {"hash3": 493, "hash4": 655, "hash5": 815, "u64_add": 53, "u64_sub": 53, "u64_mul": 85,
                    "u64_divrem": 166, "u64_lessthan": 44, "u64_iszero": 26, "digest_equal": 38,
                    "big_num_lessthan": 78}     # src/core/eval_direct.rs:2053-2063


# src/core/state.rs:270-316: the native symbols, in the order SymbolsDigests::new interns them (pointer = index + 1)
LURK_SYMBOLS = ["nil", "t", "&rest"]
BUILTIN_SYMBOLS = ["atom", "apply", "begin", "car", "cdr", "char", "commit", "comm", "bignum", "cons", "empty-env",
                   "current-env", "bind", "env", "emit", "eval", "eq", "eqq", "type-eq", "type-eqq", "hide", "if", "lambda",
                   "let", "letrec", "u64", "open", "quote", "secret", "strcons", "list", "+", "-", "*", "/", "%", "=", "<", ">",
                   "<=", ">=", "breakpoint", "fail"]
LURK_SYMBOL_PTRS = {name: i + 1 for i, name in enumerate(LURK_SYMBOLS + BUILTIN_SYMBOLS)}


def lurk_data_toplevel():
    """Data in and out of the Lair memory, restated from src/core/ingress.rs:142-329: `ingress(tag_full, digest)`
    turns a content-addressed ZPtr into (tag, pointer) by walking the hash preimages (`preimg(hash4 / hash5, ..)`,
    fed by ZStore.inject) and storing the parts; `egress(tag, ptr)` hashes it back.  nil and t become the
    InternalTags 15 / 16 with the pointers of their preallocated digests (the reference's `preallocate_symbols`
    stores every native symbol first: restated here too, nil = pointer 1, t = 2).
    `data_roundtrip` = preallocate_symbols, ingress, egress.  Widths pinned by the reference (eval_direct.rs:2051-2052):
    ingress 104, egress 81."""
    from .lair import assert_eq, if_, preimg, range_u8, slice_
    from .zstore import TAG, ZStore
    T = dict(TAG, Nil=15, T=16)
    z = ZStore()
    nil_digest, t_digest = list(z.nil().digest), list(z.t().digest)
    NIL_PTR, T_PTR = LURK_SYMBOL_PTRS["nil"], LURK_SYMBOL_PTRS["t"]
    chips = [ChipInfo("hasher3", 24, 8, 457, 0, kind=1), ChipInfo("hasher4", 32, 8, 611, 0, kind=2),
             ChipInfo("hasher5", 40, 8, 763, 0, kind=3)]
    h3 = func("hash3", [("preimg", 24)], 8, [extern_call([("img", 8)], "hasher3", ["preimg"]), ret("img")], invertible=True)
    h4 = func("hash4", [("preimg", 32)], 8, [extern_call([("img", 8)], "hasher4", ["preimg"]), ret("img")], invertible=True)
    h5 = func("hash5", [("preimg", 40)], 8, [extern_call([("img", 8)], "hasher5", ["preimg"]), ret("img")], invertible=True)

    def pair_in():      # Str (non-empty) and Cons: two children behind hash4
        return [preimg([("fst_tag_full", 8), ("fst_digest", 8), ("snd_tag_full", 8), ("snd_digest", 8)], "hash4", ["digest"]),
                call(["fst_tag", "fst_ptr"], "ingress", ["fst_tag_full", "fst_digest"]),
                call(["snd_tag", "snd_ptr"], "ingress", ["snd_tag_full", "snd_digest"]),
                store("ptr", ["fst_tag", "fst_ptr", "snd_tag", "snd_ptr"]), ret("tag", "ptr")]

    def triple_in(env_tag_var, mk_env_tag):   # Fun / Fix and Env: two children + an environment behind hash5
        return ([preimg([("a_tag_full", 8), ("a_digest", 8), ("b_tag_full", 8), ("b_digest", 8), ("env_digest", 8)],
                        "hash5", ["digest"])] + mk_env_tag +
                [call(["a_tag", "a_ptr"], "ingress", ["a_tag_full", "a_digest"]),
                 call(["b_tag", "b_ptr"], "ingress", ["b_tag_full", "b_digest"]),
                 call(["_env_tag", "env_ptr"], "ingress", [env_tag_var, "zeros", "env_digest"]),
                 store("ptr", ["a_tag", "a_ptr", "b_tag", "b_ptr", "env_ptr"]), ret("tag", "ptr")])

    ingress = func("ingress", [("tag_full", 8), ("digest", 8)], 2, [
        array("zeros", [0] * 7),
        slice_(["tag", ("rest", 7)], ["tag_full"]),
        assert_eq("rest", "zeros"),
        match("tag", [
            (T["Num"], [slice_(["x", ("rest2", 7)], ["digest"]), assert_eq("rest2", "zeros"), ret("tag", "x")]),
            (T["Char"], [slice_([("bytes", 4), ("rest4", 4)], ["digest"]), range_u8(["bytes"]), array("zeros4", [0] * 4),
                         assert_eq("rest4", "zeros4"), store("ptr", ["bytes"]), ret("tag", "ptr")]),
            (T["U64"], [range_u8(["digest"]), store("ptr", ["digest"]), ret("tag", "ptr")]),
            (T["Sym"], [
                array("nil_digest", nil_digest),
                sub(("not_nil", 8), "digest", "nil_digest"),
                if_("not_nil", [
                    array("t_digest", t_digest),
                    sub(("not_t", 8), "digest", "t_digest"),
                    if_("not_t", [store("ptr", ["digest"]), ret("tag", "ptr")],
                        [const("t_tag", T["T"]), const("ptr", T_PTR), ret("t_tag", "ptr")])],
                    [const("nil_tag", T["Nil"]), const("ptr", NIL_PTR), ret("nil_tag", "ptr")])]),
            ((T["Builtin"], T["Coroutine"], T["Key"], T["BigNum"], T["Comm"]), [store("ptr", ["digest"]), ret("tag", "ptr")]),
            (T["Str"], [if_("digest", pair_in(), [const("zero", 0), ret("tag", "zero")])]),
            (T["Cons"], pair_in()),
            ((T["Fun"], T["Fix"]), triple_in("env_tag", [const("env_tag", T["Env"])])),
            (T["Env"], [if_("digest", triple_in("tag", []), [const("zero", 0), ret("tag", "zero")])]),
        ], exhaustive=True)])

    def pair_out():
        return [load(["fst_tag", "fst_ptr", "snd_tag", "snd_ptr"], "val"),
                call(["fst_tag2", ("fst_digest", 8)], "egress", ["fst_tag", "fst_ptr"]),
                call(["snd_tag2", ("snd_digest", 8)], "egress", ["snd_tag", "snd_ptr"]),
                array("padding", [0] * 7),
                slice_([("fst_tag_full", 8)], ["fst_tag2", "padding"]),
                slice_([("snd_tag_full", 8)], ["snd_tag2", "padding"]),
                call([("digest", 8)], "hash4", ["fst_tag_full", "fst_digest", "snd_tag_full", "snd_digest"]),
                ret("tag", "digest")]

    def triple_out(env_tag_var, mk_env_tag):
        return ([load(["a_tag", "a_ptr", "b_tag", "b_ptr", "env_ptr"], "val"),
                 call(["a_tag2", ("a_digest", 8)], "egress", ["a_tag", "a_ptr"]),
                 call(["b_tag2", ("b_digest", 8)], "egress", ["b_tag", "b_ptr"])] + mk_env_tag +
                [call(["_env_tag", ("env_digest", 8)], "egress", [env_tag_var, "env_ptr"]),
                 array("padding", [0] * 7),
                 slice_([("a_tag_full", 8)], ["a_tag2", "padding"]),
                 slice_([("b_tag_full", 8)], ["b_tag2", "padding"]),
                 call([("digest", 8)], "hash5", ["a_tag_full", "a_digest", "b_tag_full", "b_digest", "env_digest"]),
                 ret("tag", "digest")])

    egress = func("egress", ["tag", "val"], 9, [
        match("tag", [
            ((T["Num"], T["Err"]), [array("padding", [0] * 7), slice_([("digest", 8)], ["val", "padding"]), ret("tag", "digest")]),
            (T["Char"], [array("padding", [0] * 4), load([("bytes", 4)], "val"), ret("tag", "bytes", "padding")]),
            (T["Nil"], [const("sym_tag", T["Sym"]), array("digest", nil_digest), ret("sym_tag", "digest")]),
            (T["T"], [const("sym_tag", T["Sym"]), array("digest", t_digest), ret("sym_tag", "digest")]),
            ((T["Sym"], T["Builtin"], T["Coroutine"], T["Key"], T["U64"], T["BigNum"], T["Comm"]),
             [load([("digest", 8)], "val"), ret("tag", "digest")]),
            (T["Str"], [if_("val", pair_out(), [array("digest", [0] * 8), ret("tag", "digest")])]),
            (T["Cons"], pair_out()),
            ((T["Fun"], T["Fix"]), triple_out("env_tag", [const("env_tag", T["Env"])])),
            (T["Env"], [if_("val", triple_out("tag", []), [array("digest", [0] * 8), ret("tag", "digest")])]),
        ], exhaustive=True)])

    # preallocate_symbols (src/core/ingress.rs:114-140): every native symbol's digest is stored first, in the order of
    # SymbolsDigests::new (ingress.rs:35-58: LURK_SYMBOLS then BUILTIN_SYMBOLS), so that its pointer is its index + 1;
    # the func asserts exactly that.  No inputs, no outputs; width pinned at 188 (eval_direct.rs:2027).
    pre_body = []
    for name, ptr in LURK_SYMBOL_PTRS.items():
        digest = z.lurk_sym(name).digest if name in LURK_SYMBOLS else z.builtin(name).digest
        pre_body += [array("arr", list(digest)), store("ptr", ["arr"]), const("addr", ptr), assert_eq("ptr", "addr")]
    preallocate = func("preallocate_symbols", [], 0, pre_body + [ret()])
    data_roundtrip = func("data_roundtrip", [("tag_full", 8), ("digest", 8)], 9, [
        call([], "preallocate_symbols", []),
        call(["tag", "ptr"], "ingress", ["tag_full", "digest"]),
        call(["out_tag", ("out_digest", 8)], "egress", ["tag", "ptr"]),
        ret("out_tag", "out_digest")])
    # extend_env_with_mutuals (src/core/eval_direct.rs:1674-1725, width pinned at 54): the environment a `letrec`
    # evaluates its body in -- every binding (var expr) becomes var -> Fix(expr, all bindings, env).  Total and
    # self-contained, so it runs for real: `mutuals_demo` ingresses a binding list, extends the empty environment
    # and egresses the result, whose digest the ZStore can predict.
    inv = [ret("err_tag", "invalid_form_err")]
    extend_env_with_mutuals = func("extend_env_with_mutuals", ["binds_tag", "binds", "mutual_binds", "mutual_env"], 2, [
        const("err_tag", T["Err"]), const("env_tag", T["Env"]), const("invalid_form_err", 1),   # EvalErr::InvalidForm
        match("binds_tag", [
            (T["Nil"], [ret("env_tag", "mutual_env")]),
            (T["Cons"], [
                const("cons_tag", T["Cons"]),
                load(["binding_tag", "binding", "binds_tag", "binds"], "binds"),
                sub("binding_not_cons", "binding_tag", "cons_tag"),
                if_("binding_not_cons", inv),
                load(["var_tag", "var", "rest_tag", "rest"], "binding"),
                sub("rest_tag_not_cons", "rest_tag", "cons_tag"),
                if_("rest_tag_not_cons", inv),
                load(["expr_tag", "expr", "rest_tag", "_rest"], "rest"),
                const("nil_tag", T["Nil"]),
                sub("rest_tag_not_nil", "rest_tag", "nil_tag"),
                if_("rest_tag_not_nil", inv),
                match("var_tag", [
                    ((T["Sym"], T["Builtin"], T["Coroutine"]), [
                        call(["ext_env_tag", "ext_env"], "extend_env_with_mutuals",
                             ["binds_tag", "binds", "mutual_binds", "mutual_env"]),
                        match("ext_env_tag", [(T["Err"], [ret("ext_env_tag", "ext_env")])]),
                        const("fix_tag", T["Fix"]),
                        store("fix", ["expr_tag", "expr", "cons_tag", "mutual_binds", "mutual_env"]),
                        store("res_env", ["var_tag", "var", "fix_tag", "fix", "ext_env"]),
                        ret("env_tag", "res_env")])]),
                const("illegal_binding_var_err", 2),                                            # EvalErr::IllegalBindingVar
                ret("err_tag", "illegal_binding_var_err")])]),
        ret("err_tag", "invalid_form_err")])
    mutuals_demo = func("mutuals_demo", [("tag_full", 8), ("digest", 8)], 9, [
        call([], "preallocate_symbols", []),
        call(["binds_tag", "binds"], "ingress", ["tag_full", "digest"]),
        const("empty_env", 0),
        call(["env_tag", "env"], "extend_env_with_mutuals", ["binds_tag", "binds", "binds", "empty_env"]),
        call(["out_tag", ("out_digest", 8)], "egress", ["env_tag", "env"]),
        ret("out_tag", "out_digest")])
    return Toplevel([data_roundtrip, preallocate, ingress, egress, h3, h4, h5, extend_env_with_mutuals, mutuals_demo],
                    chips)
