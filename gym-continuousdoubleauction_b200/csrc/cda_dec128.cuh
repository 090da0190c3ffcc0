// cda_dec128.cuh — Decimal(prec 28, ROUND_HALF_EVEN) add / sub / mul / div / compare on fixed-width integers, host + device.
//
// NOT USED BY THE STEP KERNEL YET.  The reference keeps money as Python `decimal.Decimal` (envs/account/account.py:124-231,
// calculate.py:5-55, cash_processor.py:15-97, agent/trader.py:108-151); its VWAP divisions leave ~1e-24 residues in cash /
// NAV, and a `cash >= order value` test at EXACT equality is decided by the residue's sign — the one place where the exact
// integer ledger of the kernel can disagree with the reference.  Reproducing that needs the reference's
// arithmetic itself: value = (-1)^sign * c * 10^exp with c < 10^28 in an unsigned __int128, every operation computed exactly
// in 128 bits and rounded ONCE.  This header is that arithmetic in the form the device ledger will use; it is the same
// algorithm as oracle/dec128.h, which reproduces the reference's Decimal fields exactly inside the oracle (700 fuzzed
// configurations).  The host side of THIS file is pinned against Python's decimal through cda_debug_dec_op
// (tests/test_dec28.py); the device side is compiled (cda_dec_selftest_kernel) and waits for its GPU test.
//
// Domain: additions are exact for all operands; a product needs digits(a) + digits(b) <= 38 and a quotient a divisor of
// at most 9 digits (in the ledger one factor is always a size or a price).  Outside it `*range_err` is incremented and the
// result is not to be used.
#pragma once
#include <stdint.h>

typedef unsigned __int128 cda_u128;
struct CdaDec { cda_u128 c; int exp; int sign; };   // c == 0 <=> zero (sign 0)

#define CDA_DEC_P 28
#if defined(__CUDACC__)
#define CDA_HD __host__ __device__ __forceinline__
#else
#define CDA_HD inline
#endif

CDA_HD cda_u128 cda_dec_pow10(int k) {               // 0 <= k <= 38; no table: shared by host and device, off the hot path
    cda_u128 r = 1, b = 10;
    for (; k; k >>= 1, b *= b) if (k & 1) r *= b;
    return r;
}
CDA_HD int cda_dec_ndigits(cda_u128 x) {             // 0 for x == 0
    int n = 0;
    cda_u128 p = 1;
    while (n < 38 && x >= p) { p *= 10; ++n; }       // p = 10^n
    if (n == 38 && x >= p) return 39;
    return n;
}
CDA_HD CdaDec cda_dec_zero() { CdaDec r; r.c = 0; r.exp = 0; r.sign = 0; return r; }

// x * 10^exp (+ sticky: something non-zero beyond x) -> 28 significant digits, half to even
CDA_HD CdaDec cda_dec_round(int sign, cda_u128 x, int exp, int sticky) {
    CdaDec r = cda_dec_zero();
    if (x == 0) return r;
    const int nd = cda_dec_ndigits(x);
    if (nd > CDA_DEC_P) {
        int k = nd - CDA_DEC_P;
        const cda_u128 p = cda_dec_pow10(k), half = p / 2;
        cda_u128 q = x / p;
        const cda_u128 rem = x % p;
        if (rem > half || (rem == half && (sticky || (q & 1)))) ++q;
        if (q == cda_dec_pow10(CDA_DEC_P)) { q = cda_dec_pow10(CDA_DEC_P - 1); ++k; }
        x = q; exp += k;
    }
    r.c = x; r.exp = exp; r.sign = sign;
    return r;
}
CDA_HD CdaDec cda_dec_from_i64(long long v) {
    const int sign = v < 0;
    const unsigned long long u = sign ? (unsigned long long)(-(v + 1)) + 1ULL : (unsigned long long)v;
    return cda_dec_round(sign, (cda_u128)u, 0, 0);
}
CDA_HD CdaDec cda_dec_neg(CdaDec a) { if (a.c) a.sign ^= 1; return a; }
CDA_HD CdaDec cda_dec_strip(CdaDec a) { while (a.c && a.c % 10 == 0) { a.c /= 10; ++a.exp; } return a; }

CDA_HD CdaDec cda_dec_add(CdaDec a, CdaDec b) {
    if (a.c == 0) return b;
    if (b.c == 0) return a;
    if (a.exp < b.exp) { const CdaDec t = a; a = b; b = t; }   // a has the larger exponent
    const int diff = a.exp - b.exp, na = cda_dec_ndigits(a.c);
    if (na + diff <= 38) {                                       // the aligned operands fit: exact sum, one rounding
        const cda_u128 x = a.c * cda_dec_pow10(diff), y = b.c;
        if (a.sign == b.sign) return cda_dec_round(a.sign, x + y, b.exp, 0);
        if (x == y) return cda_dec_zero();
        return x > y ? cda_dec_round(a.sign, x - y, b.exp, 0) : cda_dec_round(b.sign, y - x, b.exp, 0);
    }
    // b lies (partly) below the 28-digit window of a (then |b| < |a| * 1e-10): widen a to 30 digits, cut b there and keep
    // what was cut as a sticky bit.  With 0 < f < 1 cut off, a' + b' + f rounds like (a' + b', sticky) and
    // a' - b' - f like (a' - b' - 1, sticky).
    const int t = 30 - na, shift = diff - t;
    const cda_u128 x = a.c * cda_dec_pow10(t);
    cda_u128 y = 0; int sticky = 1;
    if (shift <= 38) { const cda_u128 p = cda_dec_pow10(shift); y = b.c / p; sticky = (b.c % p) != 0; }
    const cda_u128 r = a.sign == b.sign ? x + y : x - y - (sticky ? 1 : 0);
    return cda_dec_round(a.sign, r, a.exp - t, sticky);
}
CDA_HD CdaDec cda_dec_sub(CdaDec a, CdaDec b) { return cda_dec_add(a, cda_dec_neg(b)); }

CDA_HD CdaDec cda_dec_mul(CdaDec a, CdaDec b, int *range_err) {
    if (a.c == 0 || b.c == 0) return cda_dec_zero();
    if (cda_dec_ndigits(a.c) + cda_dec_ndigits(b.c) > 38) {
        a = cda_dec_strip(a); b = cda_dec_strip(b);
        if (cda_dec_ndigits(a.c) + cda_dec_ndigits(b.c) > 38) { ++*range_err; return cda_dec_zero(); }
    }
    return cda_dec_round(a.sign ^ b.sign, a.c * b.c, a.exp + b.exp, 0);
}
// a / b, b != 0 with at most 9 significant digits: dividend scaled to 38 digits, quotient >= 29 digits + sticky remainder
CDA_HD CdaDec cda_dec_div(CdaDec a, CdaDec b, int *range_err) {
    if (a.c == 0) return cda_dec_zero();
    b = cda_dec_strip(b);
    const int nb = cda_dec_ndigits(b.c), na = cda_dec_ndigits(a.c), k = 38 - na;
    if (na + k - nb < CDA_DEC_P + 1) { ++*range_err; return cda_dec_zero(); }
    const cda_u128 x = a.c * cda_dec_pow10(k);
    return cda_dec_round(a.sign ^ b.sign, x / b.c, a.exp - k - b.exp, (x % b.c) != 0);
}
CDA_HD int cda_dec_cmp(CdaDec a, CdaDec b) {          // -1, 0, +1 (the sign of a rounded difference is the sign of the exact one)
    const CdaDec d = cda_dec_sub(a, b);
    return d.c == 0 ? 0 : (d.sign ? -1 : 1);
}
