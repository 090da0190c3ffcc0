/* =====================================================================================
 * TEST INFRASTRUCTURE — the same Decimal(prec 28, ROUND_HALF_EVEN) arithmetic as dec28.h, on fixed-width integers
 * (coefficient < 10^28 in an unsigned __int128, decimal exponent, sign).  NOT PRODUCT CODE (yet): this is the form the
 * device ledger of DESIGN.md section 9 item 0 will take — nvcc supports unsigned __int128 in device code — validated
 * here first, on the CPU, inside the oracle (build with -DORC_DEC128) against the reference's own Decimal fields.
 *
 * Every operation rounds ONCE, from the exact result: additions are exact for every pair of operands (aligned in 128 bits
 * when they fit, otherwise on a 30-digit window with a sticky bit); a product needs digits(a) + digits(b) <= 38 and a
 * quotient a divisor of at most 9 digits — true for the ledger, where one factor is always a size or a price.  A
 * product or quotient that would not fit raises `dec_range_errors` instead of returning a wrong digit (the oracle turns
 * that into a status bit).
 * ===================================================================================== */
#ifndef CDA_DEC128_H
#define CDA_DEC128_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define DEC_P 28
typedef unsigned __int128 u128;
typedef struct { u128 c; int32_t exp; int32_t sign; } dec;   /* value = (-1)^sign * c * 10^exp, c < 10^28; c == 0 is zero (sign 0) */

static int dec_range_errors = 0;

static u128 dec_pow10(int k) {                 /* 0 <= k <= 38 */
    static u128 t[39]; static int init = 0;
    if (!init) { t[0] = 1; for (int i = 1; i < 39; ++i) t[i] = t[i - 1] * 10; init = 1; }
    return t[k];
}
static int dec_ndigits(u128 x) {               /* 0 for x == 0 */
    for (int n = 0; n <= 38; ++n) if (x < dec_pow10(n)) return n;   /* (a binary search over the table on the device) */
    return 39;
}
static dec dec_zero(void) { dec r; r.c = 0; r.exp = 0; r.sign = 0; return r; }

/* x * 10^exp (+ sticky beyond x) -> 28 significant digits, half to even */
static dec dec_round_u128(int sign, u128 x, int exp, int sticky) {
    dec r = dec_zero();
    if (x == 0) return r;
    int nd = dec_ndigits(x);
    if (nd > DEC_P) {
        int k = nd - DEC_P;
        u128 p = dec_pow10(k), q = x / p, rem = x % p, half = p / 2;
        int up = rem > half || (rem == half && (sticky || (q & 1)));
        if (up) ++q;
        if (q == dec_pow10(DEC_P)) { q = dec_pow10(DEC_P - 1); ++k; }
        x = q; exp += k;
    }
    r.c = x; r.exp = exp; r.sign = sign;
    return r;
}
static dec dec_from_i64(int64_t v) {
    int sign = v < 0;
    uint64_t u = sign ? (uint64_t)(-(v + 1)) + 1u : (uint64_t)v;
    return dec_round_u128(sign, (u128)u, 0, 0);
}
static dec dec_neg(dec a) { if (a.c) a.sign ^= 1; return a; }

/* strip trailing zeros (value-preserving; keeps coefficients short so that alignment has room) */
static dec dec_strip(dec a) {
    while (a.c && a.c % 10 == 0) { a.c /= 10; ++a.exp; }
    return a;
}

static dec dec_add(dec a, dec b) {
    if (a.c == 0) return b;
    if (b.c == 0) return a;
    if (a.exp < b.exp) { dec t = a; a = b; b = t; }            /* a has the larger exponent */
    int diff = a.exp - b.exp;
    int na = dec_ndigits(a.c);
    if (na + diff <= 38) {                                      /* the aligned operands fit: exact sum, one rounding */
        u128 x = a.c * dec_pow10(diff), y = b.c;
        if (a.sign == b.sign) return dec_round_u128(a.sign, x + y, b.exp, 0);
        if (x == y) return dec_zero();
        return x > y ? dec_round_u128(a.sign, x - y, b.exp, 0) : dec_round_u128(b.sign, y - x, b.exp, 0);
    }
    /* b lies (partly) below the 28-digit window of a (na + diff > 38 implies |b| < |a| * 1e-10): widen a to 30 digits (two
     * guard digits), cut b at that position and keep what was cut as a sticky bit.  With a fraction 0 < f < 1 cut off,
     * a' + b' + f rounds like (a' + b', sticky) and a' - b' - f like (a' - b' - 1, sticky). */
    int t = 30 - na, shift = diff - t;
    u128 x = a.c * dec_pow10(t), y = 0; int sticky;
    if (shift > 38) sticky = 1;
    else { u128 p = dec_pow10(shift); y = b.c / p; sticky = (b.c % p) != 0; }
    u128 r = a.sign == b.sign ? x + y : x - y - (sticky ? 1 : 0);
    return dec_round_u128(a.sign, r, a.exp - t, sticky);
}
static dec dec_sub(dec a, dec b) { return dec_add(a, dec_neg(b)); }

/* exact product of two coefficients can need 56 digits: split b so that every partial product fits, i.e. require
 * digits(a) + digits(b) <= 38 after stripping — true for the ledger (one factor is a size or a price) */
static dec dec_mul(dec a, dec b) {
    if (a.c == 0 || b.c == 0) return dec_zero();
    if (dec_ndigits(a.c) + dec_ndigits(b.c) > 38) {
        a = dec_strip(a); b = dec_strip(b);
        if (dec_ndigits(a.c) + dec_ndigits(b.c) > 38) { ++dec_range_errors; return dec_zero(); }
    }
    return dec_round_u128(a.sign ^ b.sign, a.c * b.c, a.exp + b.exp, 0);
}

/* a / b with a small divisor (digits(b) <= 10): dividend scaled to at most 38 digits, quotient >= 28 digits + sticky */
static dec dec_div(dec a, dec b) {
    if (a.c == 0) return dec_zero();
    b = dec_strip(b);
    int nb = dec_ndigits(b.c), na = dec_ndigits(a.c);
    int k = 38 - na;                                            /* as many zeros as fit */
    if (na + k - nb < DEC_P + 1) { ++dec_range_errors; return dec_zero(); }   /* divisor too long for one 128-bit step */
    u128 x = a.c * dec_pow10(k), q = x / b.c, rem = x % b.c;
    return dec_round_u128(a.sign ^ b.sign, q, a.exp - k - b.exp, rem != 0);
}

static int dec_cmp(dec a, dec b) {
    dec d = dec_sub(a, b);
    if (d.c == 0) return 0;
    return d.sign ? -1 : 1;
}
static int dec_sign(dec a) { return a.c == 0 ? 0 : (a.sign ? -1 : 1); }

static void dec_to_str(dec a, char *out, size_t cap) {
    if (a.c == 0) { snprintf(out, cap, "0"); return; }
    char t[48]; int n = 0; u128 x = a.c;
    while (x) { t[n++] = (char)('0' + (int)(x % 10)); x /= 10; }
    size_t p = 0;
    if (a.sign && p + 1 < cap) out[p++] = '-';
    while (n && p + 1 < cap) out[p++] = t[--n];
    snprintf(out + p, cap - p, "e%d", a.exp);
}
static double dec_to_double(dec a) {
    if (a.c == 0) return 0.0;
    char s[64]; dec_to_str(a, s, sizeof(s));
    return strtod(s, NULL);
}
static int64_t dec_to_i64_nearest(dec a) {
    if (a.c == 0) return 0;
    u128 v;
    if (a.exp >= 0) v = a.c * dec_pow10(a.exp > 18 ? 18 : a.exp);
    else if (-a.exp > 38) v = 0;
    else { u128 p = dec_pow10(-a.exp), q = a.c / p, rem = a.c % p, half = p / 2; if (rem > half || (rem == half && (q & 1))) ++q; v = q; }
    return a.sign ? -(int64_t)v : (int64_t)v;
}
static dec dec_from_str(const char *s) {
    int sign = 0; if (*s == '-') { sign = 1; ++s; } else if (*s == '+') ++s;
    u128 x = 0; int exp = 0, seen_pt = 0, nd = 0, sticky = 0;
    for (; *s && *s != 'e' && *s != 'E'; ++s) {
        if (*s == '.') { seen_pt = 1; continue; }
        if (nd < 38) { x = x * 10 + (unsigned)(*s - '0'); if (x) ++nd; if (seen_pt) --exp; }
        else { sticky |= *s != '0'; if (!seen_pt) ++exp; }
    }
    if (*s == 'e' || *s == 'E') exp += atoi(s + 1);
    return dec_round_u128(sign, x, exp, sticky);
}
#endif
