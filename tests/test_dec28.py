"""oracle/dec28.h (the Decimal(prec 28, ROUND_HALF_EVEN) arithmetic of the oracle's optional decimal ledger) against Python's own
`decimal` — the library the reference's ledger runs on (envs/account/*.py)."""
import ctypes
import random
from decimal import Decimal, getcontext

import pytest

from oracle import cda_oracle


@pytest.fixture(scope="module")
def L():
    lib = cda_oracle.lib()
    lib.orc_dec_op.argtypes = [ctypes.c_char, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    lib.orc_dec_to_double.argtypes = [ctypes.c_char_p]
    lib.orc_dec_to_double.restype = ctypes.c_double
    return lib


def rand_dec(rng):
    kind = rng.random()
    if kind < 0.25:                                   # ledger-like integers
        return Decimal(rng.randrange(-10**rng.randrange(1, 12), 10**rng.randrange(1, 12)))
    if kind < 0.5:                                    # integer plus a residue of the size the reference's VWAP division leaves
        return Decimal(rng.randrange(0, 10**9)) + Decimal(rng.randrange(-999, 999)).scaleb(-rng.randrange(18, 27))
    digits = rng.randrange(1, 29)
    return Decimal(rng.randrange(-10**digits, 10**digits)).scaleb(rng.randrange(-30, 12))


def test_four_operations_and_compare_match_python_decimal(L):
    assert getcontext().prec == 28
    rng = random.Random(12345)
    out = ctypes.create_string_buffer(96)
    n = 0
    for _ in range(6000):
        a, b = rand_dec(rng), rand_dec(rng)
        for op, f in (("+", lambda: a + b), ("-", lambda: a - b), ("*", lambda: a * b), ("/", lambda: a / b)):
            if op == "/" and b == 0:
                continue
            L.orc_dec_op(op.encode(), str(a).encode(), str(b).encode(), out, 96)
            assert Decimal(out.value.decode()) == f(), (op, a, b, out.value)
            n += 1
        L.orc_dec_op(b"c", str(a).encode(), str(b).encode(), out, 96)
        assert int(out.value) == (a > b) - (a < b)
    assert n > 20000


def test_rounding_corners(L):
    out = ctypes.create_string_buffer(96)
    cases = [("+", "9999999999999999999999999999", "0.5"), ("+", "9999999999999999999999999999", "0.4999999"),
             ("+", "1000000000000000000000000000.5", "0"), ("-", "1", "1e-40"), ("-", "1e10", "1e-30"),
             ("/", "1", "3"), ("/", "2", "3"), ("/", "129", "6"), ("/", "64", "3"), ("*", "21.33333333333333333333333333", "3"),
             ("*", "0.6666666666666666666666666667", "3"), ("+", "395.999999999999999999999999", "1e-24"),
             ("-", "1000000", "0.000000000000000000000049999"), ("/", "1e-30", "7"), ("*", "-99999999999999", "99999999999999"),
             ("+", "-11", "-7142857142857142857142857143e-54"), ("-", "10", "1e-40"), ("-", "1", "5e-29"), ("+", "1", "5e-28"),
             ("+", "1", "5.000000000000000000000000001e-28"), ("-", "1e10", "4.9999999999999999999999999999e-18")]
    L128 = cda_oracle.lib(dec128=True)
    L128.orc_dec_op.argtypes = L.orc_dec_op.argtypes
    for op, a, b in cases:
        A, B = Decimal(a), Decimal(b)
        want = A + B if op == "+" else A - B if op == "-" else A * B if op == "*" else A / B
        for lib in (L, L128):                           # digit arrays and the fixed-width form
            lib.orc_dec_op(op.encode(), a.encode(), b.encode(), out, 96)
            assert Decimal(out.value.decode()) == want, (op, a, b, out.value, want)


def test_to_float_is_correctly_rounded(L):
    rng = random.Random(7)
    for _ in range(3000):
        a = rand_dec(rng)
        assert L.orc_dec_to_double(str(a).encode()) == float(a), a


def test_fixed_width_form_matches_python_decimal_on_ledger_shaped_operands():
    """oracle/dec128.h (unsigned __int128 coefficients — the form meant for the device ledger): exact on the operand shapes the
    ledger produces (money +- residue, sizes and prices as the short factor / divisor), and it COUNTS what does not fit 128
    bits instead of returning a wrong digit."""
    lib = cda_oracle.lib(dec128=True)
    lib.orc_dec_op.argtypes = [ctypes.c_char, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    rng = random.Random(99)
    out = ctypes.create_string_buffer(96)

    def money():
        v = Decimal(rng.randrange(-10**rng.randrange(1, 11), 10**rng.randrange(1, 11)))
        if rng.random() < 0.7:                          # a full-precision value: integer part + residue down to 28 significant digits
            v = (v + Decimal(rng.randrange(-10**6, 10**6)).scaleb(-rng.randrange(16, 27))) * 1
        return v

    def small():
        return Decimal(rng.randrange(1, 10**rng.randrange(1, 7)))

    def tiny():                                          # far below the other operand's 28-digit window, down to irrelevance
        return Decimal(rng.randrange(-10**28, 10**28)).scaleb(-rng.randrange(30, 90))

    before = lib.orc_dec_range_errors()
    for _ in range(6000):
        a, b, k, e = money(), money(), small(), tiny()
        s_ = Decimal(rng.randrange(-99, 99))
        for op, x, y, want in (("+", a, b, a + b), ("-", a, b, a - b), ("*", k, a, k * a), ("/", a, k, a / k),
                               ("+", a, e, a + e), ("-", a, e, a - e), ("-", e, a, e - a), ("+", s_, e, s_ + e), ("-", s_, e, s_ - e)):
            lib.orc_dec_op(op.encode(), str(x).encode(), str(y).encode(), out, 96)
            assert Decimal(out.value.decode()) == want, (op, x, y, out.value, want)
        lib.orc_dec_op(b"c", str(a).encode(), str(b).encode(), out, 96)
        assert int(out.value) == (a > b) - (a < b)
    assert lib.orc_dec_range_errors() == before         # everything above fitted
    lib.orc_dec_op(b"*", b"1234567890123456789012345678", b"9876543210987654321098765432", out, 96)
    assert lib.orc_dec_range_errors() == before + 1     # 28 x 28 digits does not: reported, not mis-rounded


def _ledger_operands(rng, n):
    def money():
        v = Decimal(rng.randrange(-10**rng.randrange(1, 11), 10**rng.randrange(1, 11)))
        if rng.random() < 0.7:
            v = (v + Decimal(rng.randrange(-10**6, 10**6)).scaleb(-rng.randrange(16, 27))) * 1
        return v
    out = []
    for _ in range(n):
        a, b = money(), money()
        k = Decimal(rng.randrange(1, 10**rng.randrange(1, 7)))
        e = Decimal(rng.randrange(-10**28, 10**28)).scaleb(-rng.randrange(30, 90))
        out += [("+", a, b), ("-", a, b), ("*", k, a), ("/", a, k), ("+", a, e), ("-", a, e), ("-", e, a), ("c", a, b), ("c", a, a + e)]
    return out


def _want(op, x, y):
    if op == "c":
        return Decimal((x > y) - (x < y))
    return x + y if op == "+" else x - y if op == "-" else x * y if op == "*" else x / y


def test_product_header_host_side_matches_python_decimal():
    """csrc/cda_dec128.cuh (the header the device ledger will use), HOST compilation, through the C-ABI test entry."""
    from gym_continuousdoubleauction_b200 import _native
    L = _native.lib()
    out = ctypes.create_string_buffer(64)
    err = ctypes.c_int32(0)
    rng = random.Random(4242)
    for op, x, y in _ledger_operands(rng, 3000):
        assert L.cda_debug_dec_op(ord(op), str(x).encode(), str(y).encode(), out, 64, ctypes.byref(err)) == 0
        assert err.value == 0 and Decimal(out.value.decode()) == _want(op, x, y), (op, x, y, out.value)
    L.cda_debug_dec_op(ord("*"), b"1234567890123456789012345678", b"9876543210987654321098765432", out, 64, ctypes.byref(err))
    assert err.value == 1                                  # outside the 128-bit domain: reported


@pytest.mark.gpu
def test_product_header_device_side_matches_python_decimal():
    from gym_continuousdoubleauction_b200 import _native
    L = _native.lib()
    rng = random.Random(777)
    cases = _ledger_operands(rng, 500)
    for op in "+-*/c":
        sel = [(x, y) for o, x, y in cases if o == op]
        n = len(sel)
        A = (ctypes.c_char_p * n)(*[str(x).encode() for x, _ in sel])
        B = (ctypes.c_char_p * n)(*[str(y).encode() for _, y in sel])
        out = ctypes.create_string_buffer(64 * n)
        err = ctypes.c_int32(0)
        assert L.cda_debug_dec_op_device(ord(op), n, A, B, out, 64, ctypes.byref(err)) == 0 and err.value == 0
        for i, (x, y) in enumerate(sel):
            got = out.raw[64 * i:64 * (i + 1)].split(b"\0", 1)[0].decode()
            assert Decimal(got) == _want(op, x, y), (op, x, y, got)
