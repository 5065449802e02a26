"""lurk_b200/reader.py: the expression subset of src/core/parser/syntax.rs, interned as ZStore::intern_syntax does."""
import pytest

from lurk_b200.reader import ReadError, key, read, user_sym
from lurk_b200.zstore import P, TAG, ZPtr, ZStore


def test_atoms():
    z = ZStore()
    assert read(z, "42") == z.u64(42) == read(z, "42u64")                    # the default numeric type is u64
    assert read(z, "18446744073709551615") == z.u64((1 << 64) - 1)
    assert read(z, "42n") == z.num(42) and read(z, str(P + 1) + "n") == z.num(1)
    assert read(z, "'a'") == z.char("a") and read(z, r"'\n'") == z.char("\n") and read(z, r"'\''") == z.char("'")
    assert read(z, '"a\\"b"') == z.string('a"b') and read(z, '""') == z.null(TAG["Str"])
    assert read(z, ":foo") == key(z, "foo") and read(z, ":foo").tag == TAG["Key"]
    assert read(z, "nil") == z.nil() and read(z, "t") == z.t() and read(z, "&rest") == z.lurk_sym("&rest")
    assert read(z, "cons") == z.builtin("cons") and read(z, "cons").tag == TAG["Builtin"]
    assert read(z, "foo") == user_sym(z, "foo") == z.symbol(["lurk-user", "foo"])
    assert read(z, "#0x0") == z.null(TAG["BigNum"]) and read(z, "#c0x0") == z.null(TAG["Comm"])
    # hex digests are the number written in base p, little-endian limbs (syntax.rs:266-292)
    assert read(z, "#0x%x" % (P + 5)) == ZPtr(TAG["BigNum"], (5, 1, 0, 0, 0, 0, 0, 0))


def test_lists_quotes_and_comments():
    z = ZStore()
    one, two, x = z.u64(1), z.u64(2), user_sym(z, "x")
    assert read(z, "()") == z.nil()
    assert read(z, "(1 2)") == z.list([one, two])
    assert read(z, "(1 . 2)") == z.cons(one, two) and read(z, "(1 2 . x)") == z.cons(one, z.cons(two, x))
    assert read(z, "'x") == z.list([z.builtin("quote"), x]) == read(z, "(quote x)")
    assert read(z, "'(1 'a')") == z.list([z.builtin("quote"), z.list([one, z.char("a")])])
    assert read(z, " ; leading comment\n (1 ; inside\n 2)\n; trailing\n") == z.list([one, two])
    assert read(z, "((1)(2))") == z.list([z.list([one]), z.list([two])])


@pytest.mark.parametrize("text", ["", "(1 2", ")", "1 2", '"abc', "(1 . 2 3)", "18446744073709551616", "#0x" + "f" * 64,
                                  ".lurk.nil", "-1", r"'\q'"])
def test_malformed_or_unsupported_input_is_an_error(text):
    with pytest.raises(ReadError):
        read(ZStore(), text)
