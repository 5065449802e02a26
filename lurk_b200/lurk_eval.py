"""The Lurk evaluator as Lair data: every func of the reference's native Lurk toplevel (`build_lurk_toplevel`,
src/core/eval_direct.rs:29-118), restated statement for statement with the lurk_b200.lair builder.  The data path
(preallocate_symbols, ingress, egress, hash3/4/5, extend_env_with_mutuals, the u64 / big-num funcs) lives in
lurk_b200/programs.py; this module holds the funcs of eval_direct.rs and puts the whole toplevel together.

Pins: every func reproduces the trace width the reference's own test expects (eval_direct.rs:2025-2063), and the
evaluation results of real Lurk programs match Lurk's semantics (tests/test_lurk_eval.py)."""
from lurk_b200.lair import (Toplevel, add, array, assert_eq, assert_ne, call, const, contains, eq, func, if_, load,
                            match, match_many, mul, not_, range_u8, ret, store, sub)

# ---- a slice of the Lurk evaluator, restated from src/core/eval_direct.rs for its pinned widths ------------------------
# Tag / InternalTag / EvalErr values: src/core/tag.rs:23-39, src/core/ingress.rs:88-98, src/core/error.rs:7-30
TAG = dict(U64=0, Num=1, BigNum=2, Comm=3, Char=4, Str=5, Key=6, Fun=7, Builtin=8, Coroutine=9, Sym=10, Cons=11, Env=12,
           Fix=13, Err=14, Nil=15, T=16)
ERR = dict(UnboundVar=0, InvalidForm=1, NotCons=12)


def lurk_eval_subset_toplevel(nil_ptr=4242, t_ptr=4343, rest_ptr=4444):
    """coerce_if_sym (eval_direct.rs:903-915), open_comm (:917-930), car_cdr (:932-977) and equal_inner (:1032-1115),
    statement for statement; `ingress` and `eval_builtin_expr` are STUBS with the
    reference's signatures (e.g. fn ingress(tag_full: [8], digest: [8]): [2]) -- a func's layout depends on its own body and on
    its callees' signatures only.  The reference pins the widths 9, 50, 61, 58 and, for env_lookup (:1937-1956), 52
    (eval_direct.rs:2044-2050)."""
    from lurk_b200.lair import (ChipInfo, Toplevel, array, assert_eq, call, const, func, if_, load, match, mul, preimg, ret,
                                slice_, store, sub)
    from lurk_b200.programs import U64_CHIPS
    chips = [ChipInfo("hasher3", 24, 8, 457, 0, kind=1), ChipInfo("hasher4", 32, 8, 611, 0, kind=2),
             ChipInfo("hasher5", 40, 8, 763, 0, kind=3)] + U64_CHIPS
    from lurk_b200.lair import extern_call
    hash3 = func("hash3", [("preimg", 24)], 8, [extern_call([("img", 8)], "hasher3", ["preimg"]), ret("img")], invertible=True)
    # eval itself (eval_direct.rs:386-441), statement for statement; its three big callees are stubs with the
    # reference's signatures
    def stub(name, params):
        return func(name, params, 2, [ret(params[0], params[1])], partial=True)
    # eval_builtin_expr (eval_direct.rs:443-765): the dispatcher on the head symbol of a built-in form, statement for
    # statement (the `breakpoint` pseudo-op, which has no effect on the trace, is left out)
    from lurk_b200.programs import LURK_SYMBOL_PTRS as SYM
    inv_ = [ret("err_tag", "invalid_form")]
    def need_cons():
        return [sub("rest_not_cons", "rest_tag", "cons_tag"), if_("rest_not_cons", inv_)]
    def need_nil():
        return [sub("rest_not_nil", "rest_tag", "nil_tag"), if_("rest_not_nil", inv_)]
    def errp(tag, val):
        return match(tag, [(TAG["Err"], [ret(tag, val)])])
    def fwd(callee, args):
        return [call(["res_tag", "res"], callee, args), ret("res_tag", "res")]
    two_args = (need_cons() + [load(["fst_tag", "fst", "rest_tag", "rest"], "rest")] + need_cons()
                + [load(["snd_tag", "snd", "rest_tag", "_rest"], "rest")] + need_nil())
    nil_or_t = lambda cond: [if_(cond, [const("nil", nil_ptr), ret("nil_tag", "nil")]),
                             const("t_tag", TAG["T"]), const("t", t_ptr), ret("t_tag", "t")]
    S = lambda *names: tuple(SYM[n] for n in names) if len(names) > 1 else SYM[names[0]]
    builtin_stub = func("eval_builtin_expr", ["head", "rest_tag", "rest", "env"], 2, [
        const("nil_tag", TAG["Nil"]), const("cons_tag", TAG["Cons"]), const("err_tag", TAG["Err"]),
        const("invalid_form", ERR["InvalidForm"]),
        match("head", [
            (S("let", "letrec", "lambda"), need_cons() + [load(["fst_tag", "fst", "rest_tag", "rest"], "rest")] + need_cons() + [
                match("head", [
                    (S("let"), fwd("eval_let", ["fst_tag", "fst", "rest_tag", "rest", "env"])),
                    (S("letrec"), fwd("eval_letrec", ["fst_tag", "fst", "rest_tag", "rest", "env"])),
                    (S("lambda"), [const("res_tag", TAG["Fun"]), store("res", ["fst_tag", "fst", "rest_tag", "rest", "env"]),
                                   ret("res_tag", "res")])], exhaustive=True)]),
            (S("cons", "strcons", "type-eq", "type-eqq", "apply"), two_args + [
                match("head", [
                    (S("cons", "strcons"), fwd("eval_binop_misc", ["head", "fst_tag", "fst", "snd_tag", "snd", "env"])),
                    (S("type-eq"), [
                        call(["fst_tag", "fst"], "eval", ["fst_tag", "fst", "env"]), errp("fst_tag", "fst"),
                        call(["snd_tag", "snd"], "eval", ["snd_tag", "snd", "env"]), errp("snd_tag", "snd"),
                        call(["fst_tag"], "coerce_if_sym", ["fst_tag"]), call(["snd_tag"], "coerce_if_sym", ["snd_tag"]),
                        sub("type_not_eq", "fst_tag", "snd_tag")] + nil_or_t("type_not_eq")),
                    (S("type-eqq"), [
                        call(["snd_tag", "snd"], "eval", ["snd_tag", "snd", "env"]), errp("snd_tag", "snd"),
                        call(["fst_tag"], "coerce_if_sym", ["fst_tag"]), call(["snd_tag"], "coerce_if_sym", ["snd_tag"]),
                        sub("type_not_eqq", "fst_tag", "snd_tag")] + nil_or_t("type_not_eqq")),
                    (S("apply"), fwd("eval_apply_builtin", ["fst_tag", "fst", "snd_tag", "snd", "env"]))], exhaustive=True)]),
            (S("list"), [call(["expr_tag", "expr"], "eval_list", ["rest_tag", "rest", "env"]), ret("expr_tag", "expr")]),
            (S("+", "-", "*", "/", "%", "=", "<", ">", "<=", ">="),
             two_args + fwd("eval_binop_num", ["head", "fst_tag", "fst", "snd_tag", "snd", "env"])),
            (S("eval"), need_cons() + [
                load(["expr_tag", "expr", "rest_tag", "rest"], "rest"),
                match("rest_tag", [
                    (TAG["Nil"], [
                        call(["res_tag", "res"], "eval", ["expr_tag", "expr", "env"]), errp("res_tag", "res"),
                        const("env", 0),
                        call(["res_tag", "res"], "eval", ["res_tag", "res", "env"]), ret("res_tag", "res")]),
                    (TAG["Cons"], [load(["env_expr_tag", "env_expr", "rest_tag", "_rest"], "rest")] + need_nil() + [
                        call(["res_tag", "res"], "eval", ["expr_tag", "expr", "env"]), errp("res_tag", "res"),
                        call(["env_tag", "new_env"], "eval", ["env_expr_tag", "env_expr", "env"]),
                        match("env_tag", [
                            (TAG["Err"], [ret("env_tag", "new_env")]),
                            (TAG["Env"], [call(["res_tag", "res"], "eval", ["res_tag", "res", "new_env"]), ret("res_tag", "res")])]),
                        const("err", 10), ret("err_tag", "err")])]),                              # EvalErr::NotEnv
                const("not_env", 10), ret("err_tag", "not_env")]),
            (S("quote"), need_cons() + [load(["expr_tag", "expr", "rest_tag", "_rest"], "rest")] + need_nil()
             + [ret("expr_tag", "expr")]),
            (S("begin"), [call(["expr_tag", "expr"], "eval_begin", ["rest_tag", "rest", "env"]), ret("expr_tag", "expr")]),
            (S("current-env", "empty-env", "fail"), need_nil() + [
                const("env_tag", TAG["Env"]),
                match("head", [
                    (S("current-env"), [ret("env_tag", "env")]),
                    (S("empty-env"), [const("env", 0), ret("env_tag", "env")]),
                    (S("fail"), [const("zero", 0), const("one", 1), assert_eq("zero", "one"), ret("zero", "zero")])],
                      exhaustive=True)]),
            (S("bind"), fwd("eval_bind_builtin", ["rest_tag", "rest", "env"])),
            (S("env"), fwd("eval_env_builtin", ["rest_tag", "rest", "env"])),
            (S("breakpoint"), [match("rest_tag", [
                (TAG["Nil"], [const("nil", nil_ptr), ret("nil_tag", "nil")]),
                (TAG["Cons"], [load(["expr_tag", "expr", "rest_tag", "_rest"], "rest")] + need_nil()
                 + [call(["val_tag", "val"], "eval", ["expr_tag", "expr", "env"]), ret("val_tag", "val")])], exhaustive=True)]),
            (S("if"), need_cons() + [load(["expr_tag", "expr", "rest_tag", "rest"], "rest")] + need_cons() + [
                load(["t_branch_tag", "t_branch", "rest_tag", "rest"], "rest"),
                match("rest_tag", [
                    (TAG["Nil"], [
                        call(["val_tag", "val"], "eval", ["expr_tag", "expr", "env"]),
                        match("val_tag", [((TAG["Nil"], TAG["Err"]), [ret("val_tag", "val")])]),
                        call(["res_tag", "res"], "eval", ["t_branch_tag", "t_branch", "env"]), ret("res_tag", "res")]),
                    (TAG["Cons"], [load(["f_branch_tag", "f_branch", "rest_tag", "_rest"], "rest")] + need_nil() + [
                        call(["val_tag", "val"], "eval", ["expr_tag", "expr", "env"]),
                        match("val_tag", [
                            (TAG["Nil"], [call(["res_tag", "res"], "eval", ["f_branch_tag", "f_branch", "env"]),
                                          ret("res_tag", "res")]),
                            (TAG["Err"], [ret("val_tag", "val")])]),
                        call(["res_tag", "res"], "eval", ["t_branch_tag", "t_branch", "env"]), ret("res_tag", "res")])]),
                ret("err_tag", "invalid_form")]),
            (S("eq"), [const("one", 1), call([("res", 2)], "equal", ["rest_tag", "rest", "env", "one"]), ret("res")]),
            (S("eqq"), [const("zero", 0), call([("res", 2)], "equal", ["rest_tag", "rest", "env", "zero"]), ret("res")]),
            (S("hide"), fwd("eval_hide", ["rest_tag", "rest", "env"])),
            (S("car", "cdr"), [
                call(["car_tag", "car", "cdr_tag", "cdr"], "car_cdr", ["rest_tag", "rest", "env"]),
                match("head", [(S("car"), [ret("car_tag", "car")]), (S("cdr"), [ret("cdr_tag", "cdr")])], exhaustive=True)]),
            (S("u64", "char", "atom", "emit", "bignum", "comm"), fwd("eval_unop", ["head", "rest_tag", "rest", "env"])),
            (S("commit", "open", "secret"), fwd("eval_opening_unop", ["head", "rest_tag", "rest", "env"])),
        ], exhaustive=True)], partial=True)
    # eval_unop (eval_direct.rs:1402-1500): atom / emit / u64 / char / bignum / comm
    from lurk_b200.lair import emit
    def cast(keep, convert, errcode):
        return [match("val_tag", [(keep, [ret("val_tag", "val")]), convert]), const("err", errcode), ret("err_tag", "err")]
    unop_stub = func("eval_unop", ["head", "rest_tag", "rest", "env"], 2, [
        const("err_tag", TAG["Err"]), const("cons_tag", TAG["Cons"]), const("nil_tag", TAG["Nil"]),
        const("invalid_form", ERR["InvalidForm"]),
        sub("rest_not_cons", "rest_tag", "cons_tag"), if_("rest_not_cons", [ret("err_tag", "invalid_form")]),
        load(["expr_tag", "expr", "rest_tag", "_rest"], "rest"),
        sub("rest_not_nil", "rest_tag", "nil_tag"), if_("rest_not_nil", [ret("err_tag", "invalid_form")]),
        call(["val_tag", "val"], "eval", ["expr_tag", "expr", "env"]),
        match("val_tag", [(TAG["Err"], [ret("val_tag", "val")])]),
        match("head", [
            (SYM["atom"], [
                sub("val_not_cons", "val_tag", "cons_tag"),
                if_("val_not_cons", [const("t_tag", TAG["T"]), const("t", t_ptr), ret("t_tag", "t")]),
                const("nil", nil_ptr), ret("nil_tag", "nil")]),
            (SYM["emit"], [emit(["val_tag", "val"]), ret("val_tag", "val")]),
            (SYM["u64"], cast(TAG["U64"], (TAG["Char"], [
                load([("bytes", 4)], "val"), array("padding", [0] * 4), store("val", ["bytes", "padding"]),
                const("val_tag", TAG["U64"]), ret("val_tag", "val")]), 18)),            # EvalErr::CantCastToU64
            (SYM["char"], cast(TAG["Char"], (TAG["U64"], [
                load([("bytes", 4), ("_ignored", 4)], "val"), store("val", ["bytes"]),
                const("val_tag", TAG["Char"]), ret("val_tag", "val")]), 17)),           # EvalErr::CantCastToChar
            (SYM["bignum"], cast(TAG["BigNum"], (TAG["Comm"], [const("big_num_tag", TAG["BigNum"]), ret("big_num_tag", "val")]), 19)),
            (SYM["comm"], [match("val_tag", [
                (TAG["BigNum"], [const("comm_tag", TAG["Comm"]), ret("comm_tag", "val")]),
                (TAG["Comm"], [ret("val_tag", "val")])]), const("err", 20), ret("err_tag", "err")]),   # CantCastToComm
        ], exhaustive=True)], partial=True)
    # eval_coroutine_expr when the Lang has no coroutines (eval_direct.rs:196-222): never callable, not partial
    coroutine_stub = func("eval_coroutine_expr", ["_head", "_args_tag", "_args", "_env"], 2, [
        const("zero", 0), const("one", 1), assert_eq("zero", "one"), ret("zero", "zero")])
    # apply (eval_direct.rs:1777-1935), statement for statement ("&rest" and nil pointers: any constants here)
    E = dict(ApplyNonFunc=3, ParamsNotList=4, ParamInvalidRest=6, ArgsNotList=7, IllegalBindingVar=2)
    SYMS = (TAG["Sym"], TAG["Builtin"], TAG["Coroutine"])
    err = lambda name: [const("err", E[name]), ret("err_tag", "err")]
    err_if_ = lambda tag, val: match(tag, [(TAG["Err"], [ret(tag, val)])])

    def apply_block():      # the "DUPLICATED APPLY BLOCK"
        return [match("args_tag", [
            (TAG["Nil"], [ret("head_tag", "head")]),
            (TAG["Cons"], [
                load(["arg_tag", "arg", "rest_args_tag", "rest_args"], "args"),
                match("param_tag", [(SYMS, [
                    call(["arg_tag", "arg"], "eval", ["arg_tag", "arg", "args_env"]),
                    err_if_("arg_tag", "arg"),
                    store("ext_env", ["param_tag", "param", "arg_tag", "arg", "func_env"]),
                    store("ext_fun", ["rest_params_tag", "rest_params", "body_tag", "body", "ext_env"]),
                    call(["res_tag", "res"], "apply", ["fun_tag", "ext_fun", "rest_args_tag", "rest_args", "args_env"]),
                    ret("res_tag", "res")])])] + err("IllegalBindingVar"))])] + err("ArgsNotList")

    rest_var = [match("rest_params_tag", [
        (TAG["Nil"], err("ParamInvalidRest")),
        (TAG["Cons"], [
            load(["param_tag", "param", "rest_params_tag", "rest_params"], "rest_params"),
            match("param_tag", [(SYMS, [
                match("rest_params_tag", [(TAG["Nil"], [
                    call(["arg_tag", "arg"], "eval_list", ["args_tag", "args", "args_env"]),
                    err_if_("arg_tag", "arg"),
                    store("ext_env", ["param_tag", "param", "arg_tag", "arg", "func_env"]),
                    store("ext_fun", ["rest_params_tag", "rest_params", "body_tag", "body", "ext_env"]),
                    const("nil_tag", TAG["Nil"]), const("nil", nil_ptr),
                    call(["res_tag", "res"], "apply", ["fun_tag", "ext_fun", "nil_tag", "nil", "args_env"]),
                    ret("res_tag", "res")])])] + err("ParamInvalidRest"))])] + err("IllegalBindingVar"))])] + err("ParamsNotList")
    apply_stub = func("apply", ["head_tag", "head", "args_tag", "args", "args_env"], 2, [
        const("err_tag", TAG["Err"]), const("fun_tag", TAG["Fun"]),
        sub("head_not_fun", "head_tag", "fun_tag"),
        if_("head_not_fun", err("ApplyNonFunc")),
        load(["params_tag", "params", "body_tag", "body", "func_env"], "head"),
        match("params_tag", [
            (TAG["Nil"], [
                call(["res_tag", "res"], "eval_begin", ["body_tag", "body", "func_env"]),
                err_if_("res_tag", "res"),
                match("args_tag", [
                    (TAG["Nil"], [ret("res_tag", "res")]),
                    (TAG["Cons"], [call(["app_res_tag", "app_res"], "apply", ["res_tag", "res", "args_tag", "args", "args_env"]),
                                   ret("app_res_tag", "app_res")])])] + err("ArgsNotList")),
            (TAG["Cons"], [
                load(["param_tag", "param", "rest_params_tag", "rest_params"], "params"),
                match("param_tag", [(SYMS, [
                    const("rest_sym", rest_ptr),
                    sub("is_not_rest_sym", "param", "rest_sym"),
                    if_("is_not_rest_sym", apply_block(), rest_var)])])] + apply_block())])] + err("ParamsNotList"), partial=True)
    eval_stub = func("eval", ["expr_tag", "expr", "env"], 2, [
        match("expr_tag", [
            ((TAG["Builtin"], TAG["Sym"], TAG["Coroutine"]), [
                load([("expr_digest", 8)], "expr"),
                call(["res_tag", "res"], "env_lookup", ["expr_tag", "expr_digest", "env"]),
                match("res_tag", [(TAG["Fix"], [const("nil_env", 0),
                                                call(["res_tag", "res"], "eval", ["res_tag", "res", "nil_env"]),
                                                ret("res_tag", "res")])]),
                ret("res_tag", "res")]),
            (TAG["Cons"], [
                load(["head_tag", "head", "rest_tag", "rest"], "expr"),
                match("head_tag", [
                    (TAG["Builtin"], [call(["res_tag", "res"], "eval_builtin_expr", ["head", "rest_tag", "rest", "env"]),
                                      ret("res_tag", "res")]),
                    (TAG["Coroutine"], [call(["res_tag", "res"], "eval_coroutine_expr", ["head", "rest_tag", "rest", "env"]),
                                        ret("res_tag", "res")])]),
                call(["head_tag", "head"], "eval", ["head_tag", "head", "env"]),
                match("head_tag", [
                    ((TAG["BigNum"], TAG["Comm"]), [
                        call(["head_tag", "head"], "open_comm", ["head"]),
                        call(["res_tag", "res"], "apply", ["head_tag", "head", "rest_tag", "rest", "env"]),
                        ret("res_tag", "res")]),
                    (TAG["Err"], [ret("head_tag", "head")])]),
                call(["res_tag", "res"], "apply", ["head_tag", "head", "rest_tag", "rest", "env"]),
                ret("res_tag", "res")]),
            (TAG["Fix"], [
                load(["body_tag", "body", "binds_tag", "binds", "mutual_env"], "expr"),
                call(["_tag", "ext_env"], "extend_env_with_mutuals", ["binds_tag", "binds", "binds", "mutual_env"]),
                call(["res_tag", "res"], "eval", ["body_tag", "body", "ext_env"]),
                ret("res_tag", "res")])]),
        ret("expr_tag", "expr")], partial=True)
    ingress_stub = func("ingress", [("tag_full", 8), ("digest", 8)], 2, [const("z", 0), ret("z", "z")])

    coerce_if_sym = func("coerce_if_sym", ["tag"], 1, [
        match("tag", [((TAG["Nil"], TAG["T"]), [const("sym_tag", TAG["Sym"]), ret("sym_tag")])]),
        ret("tag")])

    open_comm = func("open_comm", ["hash_ptr"], 2, [
        load([("comm_hash", 8)], "hash_ptr"),
        preimg([("_secret", 8), "payload_tag", ("padding", 7), ("val_digest", 8)], "hash3", ["comm_hash"]),
        call(["payload_tag2", "payload_ptr"], "ingress", ["payload_tag", "padding", "val_digest"]),
        ret("payload_tag2", "payload_ptr")])

    invalid = [ret("err_tag", "invalid_form", "err_tag", "invalid_form")]
    car_cdr = func("car_cdr", ["rest_tag", "rest", "env"], 4, [
        const("nil", nil_ptr),                    # digests.lurk_symbol_ptr("nil"): the preallocated pointer
        const("nil_tag", TAG["Nil"]), const("err_tag", TAG["Err"]), const("cons_tag", TAG["Cons"]),
        const("invalid_form", ERR["InvalidForm"]),
        sub("rest_not_cons", "rest_tag", "cons_tag"),
        if_("rest_not_cons", invalid),
        load(["expr_tag", "expr", "rest_tag2", "_rest"], "rest"),
        sub("rest_not_nil", "rest_tag2", "nil_tag"),
        if_("rest_not_nil", invalid),
        call(["val_tag", "val"], "eval", ["expr_tag", "expr", "env"]),
        match("val_tag", [
            (TAG["Err"], [ret("val_tag", "val", "val_tag", "val")]),
            (TAG["Cons"], [load(["car_tag", "car", "cdr_tag", "cdr"], "val"), ret("car_tag", "car", "cdr_tag", "cdr")]),
            (TAG["Nil"], [ret("nil_tag", "nil", "nil_tag", "nil")]),
            (TAG["Str"], [
                const("empty", 0),
                sub("not_empty", "val", "empty"),
                if_("not_empty", [load(["car_tag", "car", "cdr_tag", "cdr"], "val"), ret("car_tag", "car", "cdr_tag", "cdr")]),
                const("str_tag", TAG["Str"]),
                ret("nil_tag", "nil", "str_tag", "empty")]),
        ]),
        const("not_cons", ERR["NotCons"]),
        ret("err_tag", "not_cons", "err_tag", "not_cons")], partial=True)

    def pair_eq(third=None):   # the Str / Cons / Fun / Env arms: load both, compare component-wise, multiply
        n3 = ["a_trd"] if third else []
        m3 = ["b_trd"] if third else []
        body = [load([("a_fst", 2), ("a_snd", 2)] + n3, "a"), load([("b_fst", 2), ("b_snd", 2)] + m3, "b"),
                call(["fst_eq"], "equal_inner", ["a_fst", "b_fst"]),
                call(["snd_eq"], "equal_inner", ["a_snd", "b_snd"])]
        if third:
            body += [call(["trd_eq"], "equal_inner", [third, "a_trd", third, "b_trd"]),
                     mul("eq", "fst_eq", "snd_eq"), mul("eq2", "eq", "trd_eq"), ret("eq2")]
        else:
            body += [mul("eq", "fst_eq", "snd_eq"), ret("eq")]
        return body

    def digest_eq(n):
        return [load([("a_d", n)], "a"), load([("b_d", n)], "b"), sub(("diff", n), "a_d", "b_d"),
                if_("diff", [ret("zero")]), ret("one")]

    def non_null(body):        # let a_and_b = mul(a, b); if !a_and_b { return zero } body
        return [mul("a_and_b", "a", "b"), if_("a_and_b", body, [ret("zero")])]

    equal_inner = func("equal_inner", ["a_tag", "a", "b_tag", "b"], 1, [
        sub("not_eq_tag", "a_tag", "b_tag"),
        const("zero", 0), const("one", 1),
        if_("not_eq_tag", [ret("zero")]),
        sub("not_eq", "a", "b"),
        if_("not_eq", [                       # `if !not_eq { return one } match ...`  ==  If(not_eq, match..., return one)
            match("a_tag", [
                (TAG["Num"], [ret("zero")]),
                (TAG["Char"], digest_eq(4)),
                ((TAG["Key"], TAG["Sym"], TAG["Builtin"], TAG["Coroutine"], TAG["U64"], TAG["BigNum"], TAG["Comm"]), digest_eq(8)),
                (TAG["Str"], non_null(pair_eq())),
                (TAG["Cons"], pair_eq()),
                ((TAG["Fun"], TAG["Fix"]), [const("trd_tag", TAG["Env"])] + pair_eq("trd_tag")),
                (TAG["Env"], non_null(pair_eq("a_tag"))),
            ], exhaustive=True)],
            [ret("one")])])
    # eval_direct.rs:1937-1956: `if !env {..}`, a 5-word load, `let y_tag_digest: [9] = (y_tag, y_digest)` (OpE::Slice)
    env_lookup = func("env_lookup", [("x_tag_digest", 9), "env"], 2, [
        if_("env", [
            load(["y_tag", "y", "val_tag", "val", "tail_env"], "env"),
            load([("y_digest", 8)], "y"),
            slice_([("y_tag_digest", 9)], ["y_tag", "y_digest"]),
            sub(("not_eq", 9), "x_tag_digest", "y_tag_digest"),
            if_("not_eq", [
                call(["res_tag", "res"], "env_lookup", ["x_tag_digest", "tail_env"]),
                ret("res_tag", "res")],
                [ret("val_tag", "val")])],
            [const("err_tag", TAG["Err"]), const("err", ERR["UnboundVar"]), ret("err_tag", "err")])])
    # driver: an environment ((a . 7) (b . 9)) built in memory the way the evaluator lays it out -- entries
    # (sym_tag, sym_ptr, val_tag, val, tail_env), symbols as pointers to 8-word digests -- then env_lookup
    env_demo = func("env_demo", ["x_tag", ("x_digest", 8)], 2, [
        array("da", list(range(1, 9))), store("pa", ["da"]),
        array("db", list(range(9, 17))), store("pb", ["db"]),
        const("sym", TAG["Sym"]), const("num", TAG["Num"]), const("nil_env", 0), const("v7", 7), const("v9", 9),
        store("env1", ["sym", "pa", "num", "v7", "nil_env"]),
        store("env2", ["sym", "pb", "num", "v9", "env1"]),
        slice_([("x_tag_digest", 9)], ["x_tag", "x_digest"]),
        call(["res_tag", "res"], "env_lookup", ["x_tag_digest", "env2"]),
        ret("res_tag", "res")])
    # eval_direct.rs:979-1030, 1118-1182: three partial funcs around `eval` (names are re-bound the way the reference's
    # `let (head_tag, head) = call(eval, head_tag, head, env)` shadows them)
    err_if = lambda tag, val: match(tag, [(TAG["Err"], [ret(tag, val)])])      # `match x_tag { Tag::Err => return } ;`
    t_or_nil = [if_("is_equal_inner", [const("t_tag", TAG["T"]), const("t", t_ptr), ret("t_tag", "t")]),
                ret("nil_tag", "is_equal_inner")]
    equal = func("equal", ["rest_tag", "rest", "env", "eval_first"], 2, [
        const("err_tag", TAG["Err"]), const("cons_tag", TAG["Cons"]), const("nil_tag", TAG["Nil"]),
        const("invalid_form", ERR["InvalidForm"]),
        sub("rest_not_cons", "rest_tag", "cons_tag"),
        if_("rest_not_cons", [ret("err_tag", "invalid_form")]),
        load(["exp1_tag", "exp1", "rest_tag", "rest"], "rest"),
        sub("rest_not_cons", "rest_tag", "cons_tag"),
        if_("rest_not_cons", [ret("err_tag", "invalid_form")]),
        load(["exp2_tag", "exp2", "rest_tag", "_rest"], "rest"),
        sub("rest_not_nil", "rest_tag", "nil_tag"),
        if_("rest_not_nil", [ret("err_tag", "invalid_form")]),
        call(["val2_tag", "val2"], "eval", ["exp2_tag", "exp2", "env"]),
        err_if("val2_tag", "val2"),
        if_("eval_first", [
            call(["val1_tag", "val1"], "eval", ["exp1_tag", "exp1", "env"]),
            err_if("val1_tag", "val1"),
            call(["is_equal_inner"], "equal_inner", ["val1_tag", "val1", "val2_tag", "val2"])] + t_or_nil),
        call(["is_equal_inner"], "equal_inner", ["exp1_tag", "exp1", "val2_tag", "val2"])] + t_or_nil, partial=True)
    invalid_tail = [const("err_tag", TAG["Err"]), const("err", ERR["InvalidForm"]), ret("err_tag", "err")]
    eval_list = func("eval_list", ["rest_tag", "rest", "env"], 2, [
        match("rest_tag", [
            (TAG["Nil"], [ret("rest_tag", "rest")]),
            (TAG["Cons"], [
                load(["head_tag", "head", "rest_tag", "rest"], "rest"),
                call(["head_tag", "head"], "eval", ["head_tag", "head", "env"]),
                err_if("head_tag", "head"),
                call(["rest_tag", "rest"], "eval_list", ["rest_tag", "rest", "env"]),
                err_if("rest_tag", "rest"),
                const("cons_tag", TAG["Cons"]),
                store("cons", ["head_tag", "head", "rest_tag", "rest"]),
                ret("cons_tag", "cons")])])] + invalid_tail, partial=True)
    eval_begin = func("eval_begin", ["rest_tag", "rest", "env"], 2, [
        match("rest_tag", [
            (TAG["Nil"], [ret("rest_tag", "rest")]),
            (TAG["Cons"], [
                load(["head_tag", "head", "rest_tag", "rest"], "rest"),
                call(["head_tag", "head"], "eval", ["head_tag", "head", "env"]),
                err_if("head_tag", "head"),
                const("nil_tag", TAG["Nil"]),
                sub("rest_not_nil", "nil_tag", "rest_tag"),
                if_("rest_not_nil", [call(["res_tag", "res"], "eval_begin", ["rest_tag", "rest", "env"]), ret("res_tag", "res")]),
                ret("head_tag", "head")])])] + invalid_tail, partial=True)
    # eval_direct.rs:1614-1672 (eval_let), 1727-1751 (eval_letrec_bindings), 1753-1775 (eval_letrec); the callee
    # extend_env_with_mutuals is a stub here (the real one lives in lurk_b200/programs.py::lurk_data_toplevel)
    mutuals_stub = func("extend_env_with_mutuals", ["binds_tag", "binds", "mutual_binds", "mutual_env"], 2,
                        [ret("binds_tag", "mutual_env")])
    inv2 = [ret("err_tag", "invalid_form")]
    eval_let = func("eval_let", ["binds_tag", "binds", "body_tag", "body", "env"], 2, [
        const("err_tag", TAG["Err"]), const("invalid_form", ERR["InvalidForm"]),
        match("binds_tag", [
            (TAG["Nil"], [call(["res_tag", "res"], "eval_begin", ["body_tag", "body", "env"]), ret("res_tag", "res")]),
            (TAG["Cons"], [
                const("cons_tag", TAG["Cons"]), const("nil_tag", TAG["Nil"]),
                load(["bind_tag", "bind", "rest_binds_tag", "rest_binds"], "binds"),
                sub("bind_not_cons", "bind_tag", "cons_tag"),
                if_("bind_not_cons", inv2),
                load(["param_tag", "param", "rest_tag", "rest"], "bind"),
                sub("rest_not_cons", "rest_tag", "cons_tag"),
                if_("rest_not_cons", inv2),
                match("param_tag", [
                    ((TAG["Sym"], TAG["Builtin"], TAG["Coroutine"]), [
                        load(["expr_tag", "expr", "rest_tag", "_rest"], "rest"),
                        sub("rest_not_nil", "rest_tag", "nil_tag"),
                        if_("rest_not_nil", inv2),
                        call(["val_tag", "val"], "eval", ["expr_tag", "expr", "env"]),
                        err_if("val_tag", "val"),
                        store("ext_env", ["param_tag", "param", "val_tag", "val", "env"]),
                        sub("rest_binds_not_nil", "nil_tag", "rest_binds_tag"),
                        if_("rest_binds_not_nil", [
                            call(["res_tag", "res"], "eval_let", ["rest_binds_tag", "rest_binds", "body_tag", "body", "ext_env"]),
                            ret("res_tag", "res")]),
                        call(["res_tag", "res"], "eval_begin", ["body_tag", "body", "ext_env"]),
                        ret("res_tag", "res")])]),
                const("err", 2),                                   # EvalErr::IllegalBindingVar
                ret("err_tag", "err")])]),
        ret("err_tag", "invalid_form")], partial=True)
    eval_letrec_bindings = func("eval_letrec_bindings", ["init_env", "ext_env"], 2, [
        sub("not_eq", "ext_env", "init_env"),
        if_("not_eq", [
            load(["_var_tag", "_var", "val_tag", "val", "ext_env"], "ext_env"),
            const("fix_tag", TAG["Fix"]),
            assert_eq("fix_tag", "val_tag"),
            const("nil_env", 0),
            call(["res_tag", "res"], "eval", ["val_tag", "val", "nil_env"]),
            err_if("res_tag", "res"),
            call(["res_tag", "res"], "eval_letrec_bindings", ["init_env", "ext_env"]),
            ret("res_tag", "res")],
            [const("env_tag", TAG["Env"]), ret("env_tag", "init_env")])], partial=True)
    eval_letrec = func("eval_letrec", ["binds_tag", "binds", "body_tag", "body", "env"], 2, [
        call(["ext_env_tag", "ext_env"], "extend_env_with_mutuals", ["binds_tag", "binds", "binds", "env"]),
        err_if("ext_env_tag", "ext_env"),
        call(["res_tag", "res"], "eval_letrec_bindings", ["env", "ext_env"]),
        err_if("res_tag", "res"),
        call(["res_tag", "res"], "eval_begin", ["body_tag", "body", "ext_env"]),
        ret("res_tag", "res")], partial=True)
    # eval_binop_misc (eval_direct.rs:1359-1400): `cons` and `strcons`, keyed by the pointers of the preallocated symbols
    eval_binop_misc = func("eval_binop_misc", ["head", "exp1_tag", "exp1", "exp2_tag", "exp2", "env"], 2, [
        const("err_tag", TAG["Err"]), const("cons_tag", TAG["Cons"]),
        call(["val1_tag", "val1"], "eval", ["exp1_tag", "exp1", "env"]),
        err_if("val1_tag", "val1"),
        call(["val2_tag", "val2"], "eval", ["exp2_tag", "exp2", "env"]),
        err_if("val2_tag", "val2"),
        match("head", [
            (SYM["cons"], [store("cons", ["val1_tag", "val1", "val2_tag", "val2"]), ret("cons_tag", "cons")]),
            (SYM["strcons"], [
                const("char_tag", TAG["Char"]), const("str_tag", TAG["Str"]),
                store("strcons", ["val1_tag", "val1", "val2_tag", "val2"]),
                sub("not_char", "val1_tag", "char_tag"),
                sub("not_str", "val2_tag", "str_tag"),
                if_("not_char", [const("err", 11), ret("err_tag", "err")]),       # EvalErr::NotChar
                if_("not_str", [const("err", 13), ret("err_tag", "err")]),        # EvalErr::NotString
                ret("str_tag", "strcons")])], exhaustive=True)], partial=True)
    # eval_env_builtin (eval_direct.rs:835-881) and eval_apply_builtin (:883-902)
    eval_env_builtin = func("eval_env_builtin", ["rest_tag", "rest", "env"], 2, [
        const("env_tag", TAG["Env"]), const("err_tag", TAG["Err"]),
        match("rest_tag", [
            (TAG["Nil"], [const("env", 0), ret("env_tag", "env")]),
            (TAG["Cons"], [
                load(["head_tag", "head", "rest_tag", "rest"], "rest"),
                call(["head_tag", "head"], "eval", ["head_tag", "head", "env"]),
                match("head_tag", [
                    (TAG["Cons"], [
                        load(["sym_tag", "sym", "val_tag", "val"], "head"),
                        match("sym_tag", [((TAG["Sym"], TAG["Builtin"], TAG["Coroutine"]), [
                            call(["tail_env_tag", "tail_env"], "eval_env_builtin", ["rest_tag", "rest", "env"]),
                            match("tail_env_tag", [
                                (TAG["Env"], [store("env", ["sym_tag", "sym", "val_tag", "val", "tail_env"]), ret("env_tag", "env")]),
                                (TAG["Err"], [ret("tail_env_tag", "tail_env")])], exhaustive=True)])]),
                        const("err", 2), ret("err_tag", "err")]),                  # EvalErr::IllegalBindingVar
                    (TAG["Err"], [ret("head_tag", "head")])]),
                const("err", ERR["NotCons"]), ret("err_tag", "err")])]),
        const("err", ERR["InvalidForm"]), ret("err_tag", "err")], partial=True)
    eval_apply_builtin = func("eval_apply_builtin", ["fst_tag", "fst", "snd_tag", "snd", "env"], 2, [
        call(["fst_tag", "fst"], "eval", ["fst_tag", "fst", "env"]),
        err_if("fst_tag", "fst"),
        call(["snd_tag", "snd"], "eval", ["snd_tag", "snd", "env"]),
        err_if("snd_tag", "snd"),
        call(["res_tag", "res"], "apply", ["fst_tag", "fst", "snd_tag", "snd", "env"]),
        ret("res_tag", "res")], partial=True)
    # eval_bind_builtin (eval_direct.rs:774-828)
    invf = [ret("err_tag", "invalid_form")]
    eval_bind_builtin = func("eval_bind_builtin", ["rest_tag", "rest", "env"], 2, [
        const("nil_tag", TAG["Nil"]), const("cons_tag", TAG["Cons"]), const("err_tag", TAG["Err"]),
        const("invalid_form", ERR["InvalidForm"]),
        sub("rest_not_cons", "rest_tag", "cons_tag"), if_("rest_not_cons", invf),
        load(["fst_tag", "fst", "rest_tag", "rest"], "rest"),
        sub("rest_not_cons", "rest_tag", "cons_tag"), if_("rest_not_cons", invf),
        load(["snd_tag", "snd", "rest_tag", "rest"], "rest"),
        sub("rest_not_cons", "rest_tag", "cons_tag"), if_("rest_not_cons", invf),
        load(["trd_tag", "trd", "rest_tag", "_rest"], "rest"),
        sub("rest_not_nil", "rest_tag", "nil_tag"), if_("rest_not_nil", invf),
        call(["fst_tag", "fst"], "eval", ["fst_tag", "fst", "env"]),
        match("fst_tag", [
            ((TAG["Sym"], TAG["Builtin"], TAG["Coroutine"]), [
                call(["snd_tag", "snd"], "eval", ["snd_tag", "snd", "env"]),
                sub("not_err", "snd_tag", "err_tag"),
                if_("not_err", [
                    call(["trd_tag", "trd"], "eval", ["trd_tag", "trd", "env"]),
                    match("trd_tag", [
                        (TAG["Env"], [store("env2", ["fst_tag", "fst", "snd_tag", "snd", "trd"]), ret("trd_tag", "env2")]),
                        (TAG["Err"], [ret("trd_tag", "trd")])]),
                    const("err", 10), ret("err_tag", "err")],                      # EvalErr::NotEnv
                    [ret("snd_tag", "snd")])]),
            (TAG["Err"], [ret("fst_tag", "fst")])]),
        const("err", 2), ret("err_tag", "err")], partial=True)                     # EvalErr::IllegalBindingVar
    # eval_hide (eval_direct.rs:1564-1612): (hide secret value) -> commitment = hash3(secret ++ value.flatten())
    eval_hide = func("eval_hide", ["rest_tag", "rest", "env"], 2, [
        const("err_tag", TAG["Err"]), const("cons_tag", TAG["Cons"]), const("nil_tag", TAG["Nil"]),
        const("invalid_form", ERR["InvalidForm"]),
        sub("rest_not_cons", "rest_tag", "cons_tag"), if_("rest_not_cons", invf),
        load(["exp1_tag", "exp1", "rest_tag", "rest"], "rest"),
        sub("rest_not_cons", "rest_tag", "cons_tag"), if_("rest_not_cons", invf),
        load(["exp2_tag", "exp2", "rest_tag", "_rest"], "rest"),
        sub("rest_not_nil", "rest_tag", "nil_tag"), if_("rest_not_nil", invf),
        call(["val1_tag", "val1"], "eval", ["exp1_tag", "exp1", "env"]),
        err_if("val1_tag", "val1"),
        call(["val2_tag", "val2"], "eval", ["exp2_tag", "exp2", "env"]),
        err_if("val2_tag", "val2"),
        match("val1_tag", [(TAG["BigNum"], [
            load([("secret", 8)], "val1"),
            call(["val2_tag", ("val2_digest", 8)], "egress", ["val2_tag", "val2"]),
            array("padding", [0] * 7),
            call([("comm_hash", 8)], "hash3", ["secret", "val2_tag", "padding", "val2_digest"]),
            store("comm_ptr", ["comm_hash"]),
            const("comm_tag", TAG["Comm"]),
            ret("comm_tag", "comm_ptr")])]),
        const("not_comm", 15), ret("err_tag", "not_comm")], partial=True)           # EvalErr::NotBigNum
    # eval_opening_unop (eval_direct.rs:1502-1562): commit / open / secret
    eval_opening_unop = func("eval_opening_unop", ["head", "rest_tag", "rest", "env"], 2, [
        const("err_tag", TAG["Err"]), const("cons_tag", TAG["Cons"]), const("nil_tag", TAG["Nil"]),
        const("invalid_form", ERR["InvalidForm"]),
        sub("rest_not_cons", "rest_tag", "cons_tag"), if_("rest_not_cons", invf),
        load(["expr_tag", "expr", "rest_tag", "_rest"], "rest"),
        sub("rest_not_nil", "rest_tag", "nil_tag"), if_("rest_not_nil", invf),
        call(["val_tag", "val"], "eval", ["expr_tag", "expr", "env"]),
        err_if("val_tag", "val"),
        match("head", [(SYM["commit"], [
            call(["val_tag", ("val_digest", 8)], "egress", ["val_tag", "val"]),
            array("padding", [0] * 7), const("zero", 0),
            call([("comm_hash", 8)], "hash3", ["zero", "padding", "val_tag", "padding", "val_digest"]),
            const("comm_tag", TAG["Comm"]),
            store("comm_ptr", ["comm_hash"]),
            ret("comm_tag", "comm_ptr")])]),
        match("val_tag", [((TAG["Comm"], TAG["BigNum"]), [
            load([("comm_hash", 8)], "val"),
            preimg([("secret", 8), "tag", ("padding", 7), ("val_digest", 8)], "hash3", ["comm_hash"]),
            match("head", [
                (SYM["open"], [call(["tag", "ptr"], "ingress", ["tag", "padding", "val_digest"]), ret("tag", "ptr")]),
                (SYM["secret"], [store("ptr", ["secret"]), const("big_num_tag", TAG["BigNum"]), ret("big_num_tag", "ptr")]),
            ], exhaustive=True)])]),
        const("cant_open", 16), ret("err_tag", "cant_open")], partial=True)          # EvalErr::CantOpen
    # eval_binop_num (eval_direct.rs:1184-1356): arithmetic and comparisons on u64 / num / big-num pairs; its callees
    # are the real pointer-taking funcs of src/core/misc.rs (lurk_b200/programs.py::lurk_misc_toplevel)
    from lurk_b200.lair import add, div, match_many
    from lurk_b200.programs import lurk_misc_toplevel
    misc = [f for f in lurk_misc_toplevel().funcs_e if f.name in {
        "u64_add", "u64_sub", "u64_mul", "u64_divrem", "u64_lessthan", "u64_iszero", "digest_equal", "big_num_lessthan"}]
    T_ = [ret("t_tag", "t")]
    NIL_ = [ret("nil_tag", "nil")]
    def cmp(callee, a, b, yes, no):
        return [call(["res"], callee, [a, b]), if_("res", yes), *no]
    def u64_arith(callee):
        return [call(["res"], callee, ["val1", "val2"]), ret("u64_tag", "res")]
    eval_binop_num = func("eval_binop_num", ["head", "exp1_tag", "exp1", "exp2_tag", "exp2", "env"], 2, [
        const("err_tag", TAG["Err"]), const("num_tag", TAG["Num"]), const("u64_tag", TAG["U64"]),
        const("nil_tag", TAG["Nil"]), const("err_div_zero", 9),                          # EvalErr::DivByZero
        const("t", t_ptr), const("nil", nil_ptr),
        call(["val1_tag", "val1"], "eval", ["exp1_tag", "exp1", "env"]),
        err_if("val1_tag", "val1"),
        call(["val2_tag", "val2"], "eval", ["exp2_tag", "exp2", "env"]),
        err_if("val2_tag", "val2"),
        const("t_tag", TAG["T"]),
        slice_([("tags", 2)], ["val1_tag", "val2_tag"]),
        match_many("tags", [
            ([TAG["U64"], TAG["U64"]], [match("head", [
                (SYM["+"], u64_arith("u64_add")),
                (SYM["-"], u64_arith("u64_sub")),
                (SYM["*"], u64_arith("u64_mul")),
                ((SYM["/"], SYM["%"]), [
                    call(["is_zero"], "u64_iszero", ["val2"]),
                    if_("is_zero", [ret("err_tag", "err_div_zero")]),
                    call(["quot", "rem"], "u64_divrem", ["val1", "val2"]),
                    match("head", [(SYM["/"], [ret("u64_tag", "quot")]), (SYM["%"], [ret("u64_tag", "rem")])], exhaustive=True)]),
                (SYM["<"], cmp("u64_lessthan", "val1", "val2", T_, NIL_)),
                (SYM[">="], cmp("u64_lessthan", "val1", "val2", NIL_, T_)),
                (SYM[">"], cmp("u64_lessthan", "val2", "val1", T_, NIL_)),
                (SYM["<="], cmp("u64_lessthan", "val2", "val1", NIL_, T_)),
                (SYM["="], cmp("digest_equal", "val1", "val2", T_, NIL_)),
            ], exhaustive=True)]),
            ([TAG["Num"], TAG["Num"]], [match("head", [
                (SYM["+"], [add("res", "val1", "val2"), ret("num_tag", "res")]),
                (SYM["-"], [sub("res", "val1", "val2"), ret("num_tag", "res")]),
                (SYM["*"], [mul("res", "val1", "val2"), ret("num_tag", "res")]),
                (SYM["/"], [if_("val2", [div("res", "val1", "val2"), ret("num_tag", "res")], [ret("err_tag", "err_div_zero")])]),
                (SYM["="], [sub("diff", "val1", "val2"), if_("diff", NIL_), *T_]),
                ((SYM["%"], SYM["<"], SYM[">"], SYM["<="], SYM[">="]), [const("err", 14), ret("err_tag", "err")]),   # NotU64
            ], exhaustive=True)]),
            ([TAG["BigNum"], TAG["BigNum"]], [match("head", [
                (SYM["<"], cmp("big_num_lessthan", "val1", "val2", T_, NIL_)),
                (SYM[">="], cmp("big_num_lessthan", "val1", "val2", NIL_, T_)),
                (SYM[">"], cmp("big_num_lessthan", "val2", "val1", T_, NIL_)),
                (SYM["<="], cmp("big_num_lessthan", "val2", "val1", NIL_, T_)),
                (SYM["="], cmp("digest_equal", "val2", "val1", T_, NIL_)),
                ((SYM["+"], SYM["-"], SYM["*"], SYM["/"], SYM["%"]), [const("err", 8), ret("err_tag", "err")]),      # InvalidArg
            ], exhaustive=True)]),
        ]),
        const("err", 8), ret("err_tag", "err")], partial=True)
    # lurk_main (eval_direct.rs:120-138): the entry point -- a call with no outputs, two ingresses, eval, egress
    pre_stub = func("preallocate_symbols", [], 0, [ret()])
    egress_stub = func("egress", ["tag", "val"], 9, [array("d", [0] * 8), ret("tag", "d")])
    lurk_main = func("lurk_main", [("full_expr_tag", 8), ("expr_digest", 8), ("env_digest", 8)], 16, [
        call([], "preallocate_symbols", []),
        call(["expr_tag", "expr"], "ingress", ["full_expr_tag", "expr_digest"]),
        array("padding", [0] * 7),
        const("env_tag", TAG["Env"]),
        slice_([("full_env_tag", 8)], ["env_tag", "padding"]),
        call(["_env_tag", "env"], "ingress", ["full_env_tag", "env_digest"]),
        call(["val_tag", "val"], "eval", ["expr_tag", "expr", "env"]),
        call(["val_tag2", ("val_digest", 8)], "egress", ["val_tag", "val"]),
        slice_([("full_val_tag", 8)], ["val_tag2", "padding"]),
        ret("full_val_tag", "val_digest")], partial=True)
    return Toplevel([hash3, eval_stub, ingress_stub, coerce_if_sym, open_comm, car_cdr, equal_inner, env_lookup,
                     env_demo, equal, eval_list, eval_begin, mutuals_stub, eval_let, eval_letrec_bindings, eval_letrec,
                     builtin_stub, coroutine_stub, apply_stub, pre_stub, egress_stub, lurk_main, eval_binop_misc,
                     eval_env_builtin, eval_apply_builtin, eval_bind_builtin, eval_hide,
                     eval_opening_unop, eval_binop_num, unop_stub] + misc, chips)


def lurk_core_toplevel():
    """The restated evaluator put together: data path (preallocate, ingress, egress, hash3/4/5, extend_env_with_mutuals:
    lurk_b200/programs.py::lurk_data_toplevel) + evaluator core (eval, apply, eval_begin, eval_list, env_lookup, eval_let,
    eval_letrec*, equal*, car_cdr, open_comm, coerce_if_sym, eval_builtin_expr and the built-in operator family: above,
    with the real preallocated pointers) -- all 39 funcs of the reference's native Lurk toplevel (`emit` is accepted and
    keeps no side channel; `breakpoint` is a no-op).  `lurk_eval_demo(expr)` = preallocate_symbols,
    ingress, eval in the empty environment, egress: Lurk programs evaluated for real."""
    from lurk_b200.lair import Toplevel, call, const, func, ret
    from lurk_b200.programs import LURK_SYMBOL_PTRS as PTR, lurk_data_toplevel as data_toplevel
    data = data_toplevel()
    ev = lurk_eval_subset_toplevel(PTR["nil"], PTR["t"], PTR["&rest"])
    take = lambda top, names: [f for f in top.funcs_e if f.name in names]
    stubs = []
    demo = func("lurk_eval_demo", [("tag_full", 8), ("digest", 8)], 9, [
        call([], "preallocate_symbols", []),
        call(["tag", "ptr"], "ingress", ["tag_full", "digest"]),
        const("empty_env", 0),
        call(["res_tag", "res"], "eval", ["tag", "ptr", "empty_env"]),
        call(["out_tag", ("out_digest", 8)], "egress", ["res_tag", "res"]),
        ret("out_tag", "out_digest")], partial=True)
    funcs = ([demo] + take(data, {"preallocate_symbols", "ingress", "egress", "hash3", "hash4", "hash5", "extend_env_with_mutuals"})
             + take(ev, {"eval", "apply", "env_lookup", "eval_begin", "eval_list", "eval_let", "eval_letrec",
                         "eval_letrec_bindings", "equal", "equal_inner", "car_cdr", "open_comm", "coerce_if_sym",
                         "eval_builtin_expr", "eval_coroutine_expr", "eval_binop_num", "eval_binop_misc", "eval_env_builtin",
                         "eval_apply_builtin", "eval_bind_builtin", "eval_hide", "eval_opening_unop", "eval_unop",
                         "u64_add", "u64_sub", "u64_mul", "u64_divrem", "u64_lessthan", "u64_iszero", "digest_equal",
                         "big_num_lessthan"}) + stubs)
    return Toplevel(funcs, ev.chips)


def build_lurk_toplevel():
    """`build_lurk_toplevel` (src/core/eval_direct.rs:29-118) for the native Lang (no coroutines): the 39 funcs of the
    reference's toplevel -- `lurk_main` first, as there -- with the native chip map.  `lurk_eval_demo` (preallocate,
    ingress, eval in the empty environment, egress) is kept as a 40th func for callers that hold no environment digest."""
    core = lurk_core_toplevel()
    ev = lurk_eval_subset_toplevel(*_real_ptrs())
    main = [f for f in ev.funcs_e if f.name == "lurk_main"]
    return Toplevel(main + list(core.funcs_e), core.chips)


def _real_ptrs():
    from lurk_b200.programs import LURK_SYMBOL_PTRS as PTR
    return PTR["nil"], PTR["t"], PTR["&rest"]
