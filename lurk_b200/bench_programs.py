"""The Lurk source of the reference's three benchmarks, argument for argument: benches/fib.rs:36-43, benches/lcs.rs:27-46,
benches/sum.rs:36-47.  `setup(name, arg)` does what their `setup` does: read the program, inject the store's hash
preimages into a fresh QueryRecord and build the 24-word input of `lurk_main` (expression by digest, empty environment)."""
from __future__ import annotations

LCS_ARGS = ("When in the Course of human events, it becomes necessary for one people to dissolve the political bands which "
            "have connected them with another",
            "There must be some kind of way outta here Said the joker to the thief. There's too much confusion. I can't get "
            "no relief.")                                               # benches/lcs.rs:27-30


def fib_source(n: int) -> str:
    return ("(letrec ((fib\n          (lambda (n)\n            (if (<= n 1) n\n"
            "              (+ (fib (- n 1)) (fib (- (- n 1) 1)))))))\n  (fib %d))" % n)


def lcs_source(a: str, b: str) -> str:
    esc = lambda s: s.replace("\\", "\\\\").replace('"', '\\"')
    return '''
(letrec ((lte (lambda (a b)
                (if (eq a "") t
                    (if (eq b "") nil
                        (lte (cdr a) (cdr b))))))
         (lcs (lambda (a b)
                (if (eq a "") ""
                    (if (eq b "") ""
                        (if (eq (car a) (car b)) (strcons (car a) (lcs (cdr a) (cdr b)))
                            (if (lte (lcs a (cdr b)) (lcs (cdr a) b)) (lcs (cdr a) b)
                                (lcs a (cdr b)))))))))
  (lcs "%s" "%s"))''' % (esc(a), esc(b))


def sum_source(n: int) -> str:
    return "\n(letrec ((sum (lambda (l) (if l (+ (car l) (sum (cdr l))) 0))))\n  (sum '(%s)))\n" % " ".join(map(str, range(n)))


def source(name: str, arg: int) -> str:
    if name == "fib":
        return fib_source(arg)
    if name == "sum":
        return sum_source(arg)
    if name == "lcs":        # arg = how many leading characters of the reference's two strings (repeated cyclically
        a, b = LCS_ARGS      # beyond their 142 / 121 characters: the bench needs inputs for taller traces than the reference's)
        return lcs_source((a * (arg // len(a) + 1))[:arg], (b * (arg // len(b) + 1))[:arg])
    raise ValueError(name)


# eval-chip rows (the tallest chip of all three programs) as a function of the argument, measured on the executor:
# fib(n) -> 10 n + 4, sum over n elements -> 7 n + 9, lcs of two k-character prefixes ~ 21.2 k^2 (data dependent).
LCS_ARG_FOR_LOG_ROWS = {14: 27, 16: 56, 18: 111, 19: 157, 20: 222, 21: 315, 22: 446}   # largest k whose trace fits 2^L rows (measured)


def arg_for_rows(name: str, log_rows: int, shards: int = 1) -> int:
    """The largest argument whose `eval` trace has at most shards * 2^log_rows rows."""
    rows = shards << log_rows
    if name == "fib":
        return max((rows - 4) // 10, 0)
    if name == "sum":
        return max((rows - 9) // 7, 0)
    if name == "lcs":
        if shards == 1 and log_rows in LCS_ARG_FOR_LOG_ROWS:
            return LCS_ARG_FOR_LOG_ROWS[log_rows]
        return max(int((rows / 21.6) ** 0.5) - 3, 1)
    raise ValueError(name)


def setup(name: str, arg: int, toplevel):
    """-> (args of lurk_main, QueryRecord with the preimages injected, ZStore)"""
    from .execute import QueryRecord
    from .reader import read
    from .zstore import ZStore
    z = ZStore()
    e = read(z, source(name, arg))
    q = QueryRecord.new(toplevel)
    z.inject(q, toplevel)
    return [e.tag] + [0] * 7 + list(e.digest) + [0] * 8, q, z
