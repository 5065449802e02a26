"""A reader for Lurk source text: `read(zstore, text) -> ZPtr`, the subset of src/core/parser/syntax.rs the reference's own
evaluation tests use (src/core/tests/eval_direct.rs), interned the way `ZStore::intern_syntax` does (zstore.rs:510-548).

  1 / 1u64       Tag::U64 (the default numeric type, syntax.rs:250-262)      1n        Tag::Num (syntax.rs:218-231)
  #0x..          Tag::BigNum, #c0x.. Tag::Comm: the hex number written in base p, little-endian limbs (syntax.rs:266-306)
  'a'            Tag::Char                      "abc"  Tag::Str             :foo      Tag::Key, path [foo]
  'x             (quote x), `quote` the builtin                              (a b . c) improper list
  symbols        resolved in the `lurk-user` package, which uses `lurk` and `lurk.builtin` (state.rs:176-216): nil / t / &rest
                 are lurk symbols, the 43 builtins are Tag::Builtin, everything else is .lurk-user.<name>
Not read: package paths (`.a.b`), `|quoted|` symbols, negative and non-decimal numerals, meta commands (`!(...)`)."""
from __future__ import annotations

import re
from typing import List, Tuple

from .programs import BUILTIN_SYMBOLS, LURK_SYMBOLS
from .zstore import P, TAG, ZPtr, ZStore

USER_PACKAGE = "lurk-user"          # src/core/state.rs:268
_NUM = re.compile(r"^([0-9]+)(u64|n)?$")
_DELIM = set(" \t\r\n()'\";")
_ESC = {"n": "\n", "t": "\t", "r": "\r", "0": "\0", "\\": "\\", "'": "'", '"': '"'}


class ReadError(ValueError):
    pass


def _digest_from_hex(h: str) -> Tuple[int, ...]:
    n = int(h, 16)
    out = []
    for _ in range(8):
        out.append(n % P)
        n //= P
    if n:
        raise ReadError("digest literal too big")        # ParseErrorKind::DigestLiteralTooBig
    return tuple(out)


def user_sym(z: ZStore, name: str) -> ZPtr:             # state.rs:254-256
    return z.symbol([USER_PACKAGE, name])


def key(z: ZStore, name: str) -> ZPtr:                  # Symbol::key(&[name])
    return z.symbol([name], TAG["Key"])


def _atom(z: ZStore, tok: str) -> ZPtr:
    m = _NUM.match(tok)
    if m:
        v = int(m.group(1))
        if m.group(2) == "n":
            return z.num(v % P)
        if v >= 1 << 64:
            raise ReadError("u64 literal out of range")
        return z.u64(v)
    if tok.startswith("#c0x"):
        return ZPtr(TAG["Comm"], _digest_from_hex(tok[4:]))
    if tok.startswith("#0x"):
        return ZPtr(TAG["BigNum"], _digest_from_hex(tok[3:]))
    if tok.startswith(":") and len(tok) > 1:
        return key(z, tok[1:])
    if tok in LURK_SYMBOLS:
        return z.lurk_sym(tok)
    if tok in BUILTIN_SYMBOLS:
        return z.builtin(tok)
    if tok[0] in ".|#" or tok[0] == "-" and tok[1:2].isdigit():
        raise ReadError(f"not read by this subset: {tok!r}")
    return user_sym(z, tok)


class _Reader:
    def __init__(self, z: ZStore, s: str):
        self.z, self.s, self.i = z, s, 0

    def ws(self) -> None:
        s = self.s
        while self.i < len(s):
            if s[self.i] == ";":
                while self.i < len(s) and s[self.i] != "\n":
                    self.i += 1
            elif s[self.i].isspace():
                self.i += 1
            else:
                break

    def esc(self) -> str:
        c = self.s[self.i]
        if c == "\\":
            self.i += 1
            c = _ESC.get(self.s[self.i])
            if c is None:
                raise ReadError("unknown escape")
        self.i += 1
        return c

    def expr(self) -> ZPtr:
        self.ws()
        s, z = self.s, self.z
        if self.i >= len(s):
            raise ReadError("unexpected end of input")
        c = s[self.i]
        if c == "(":
            self.i += 1
            xs: List[ZPtr] = []
            tail = None
            while True:
                self.ws()
                if self.i >= len(s):
                    raise ReadError("unclosed list")
                if s[self.i] == ")":
                    self.i += 1
                    break
                if s[self.i] == "." and self.i + 1 < len(s) and s[self.i + 1] in _DELIM and xs:
                    self.i += 1
                    tail = self.expr()
                    self.ws()
                    if self.i >= len(s) or s[self.i] != ")":
                        raise ReadError("malformed improper list")
                    self.i += 1
                    break
                xs.append(self.expr())
            out = z.nil() if tail is None else tail       # intern_list / intern_list_full
            for x in reversed(xs):
                out = z.cons(x, out)
            return out
        if c == ")":
            raise ReadError("unexpected )")
        if c == "'":
            j = self.i + 1
            if j < len(s) and (s[j] == "\\" and s[j + 2:j + 3] == "'" or s[j] != "\\" and s[j + 1:j + 2] == "'"):
                self.i = j
                ch = self.esc()
                self.i += 1
                return z.char(ch)
            self.i += 1
            return z.list([z.builtin("quote"), self.expr()])
        if c == '"':
            self.i += 1
            out = []
            while True:
                if self.i >= len(s):
                    raise ReadError("unclosed string")
                if s[self.i] == '"':
                    self.i += 1
                    return z.string("".join(out))
                out.append(self.esc())
        j = self.i
        while j < len(s) and s[j] not in _DELIM:
            j += 1
        tok, self.i = s[self.i:j], j
        return _atom(z, tok)


def read(z: ZStore, text: str) -> ZPtr:
    """`ZStore::read` (zstore.rs:551-575): exactly one expression."""
    r = _Reader(z, text)
    out = r.expr()
    r.ws()
    if r.i != len(text):
        raise ReadError("trailing input")
    return out
