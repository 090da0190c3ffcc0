/* =====================================================================================
 * TEST INFRASTRUCTURE — decimal arithmetic of the reference's ledger.  NOT PRODUCT CODE.
 *
 * The reference keeps money as Python `decimal.Decimal` in the default context (precision 28 significant digits,
 * ROUND_HALF_EVEN; call sites: envs/account/account.py:124-231, calculate.py:5-55, cash_processor.py:15-97,
 * agent/trader.py:108-151).  CPython's decimal is libmpdec (a third-party dependency, not in /root/reference); its
 * published semantics for + - * / are: compute the exact result, round it ONCE to the context precision.  This header
 * restates exactly that on digit arrays (schoolbook, slow, obviously right), so that the oracle's optional
 * `decimal_ledger` mode reproduces the ~1e-24 residues the reference's VWAP divisions leave in cash / NAV — the only
 * place where an exact integer ledger can disagree with the reference (a `cash >= order value` test at exact equality).
 * Pinned against Python's own decimal by tests/test_dec28.py (random operands, all four operations, compare, to-float).
 * ===================================================================================== */
#ifndef CDA_DEC28_H
#define CDA_DEC28_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define DEC_P 28          /* context precision */
#define DEC_W 160         /* work-buffer digits */

typedef struct {
    int8_t sign;          /* 0 = +, 1 = - (zero is always +) */
    int16_t n;            /* number of coefficient digits, 0 = the value zero */
    int32_t exp;          /* value = coefficient * 10^exp */
    uint8_t d[DEC_P];     /* coefficient digits, most significant first, no leading zero, no trailing zero */
} dec;

static dec dec_zero(void) { dec r; memset(&r, 0, sizeof(r)); return r; }

/* digits dig[0..n) (most significant first, leading zeros allowed) * 10^exp, plus `sticky` = something non-zero beyond the
 * last digit -> rounded to DEC_P significant digits, half to even; trailing zeros stripped (value-preserving) */
static dec dec_round_into(int sign, const uint8_t *dig, int n, int exp, int sticky) {
    dec r = dec_zero();
    while (n > 0 && dig[0] == 0) { ++dig; --n; }
    if (n <= 0) return r;                       /* (a pure-sticky value cannot occur: callers keep >= DEC_P + 1 digits) */
    uint8_t buf[DEC_P + 1];
    int keep = n < DEC_P ? n : DEC_P;
    memcpy(buf + 1, dig, (size_t)keep); buf[0] = 0;
    if (n > DEC_P) {
        int guard = dig[DEC_P], rest = sticky;
        for (int i = DEC_P + 1; i < n && !rest; ++i) rest |= dig[i] != 0;
        exp += n - DEC_P;
        int up = guard > 5 || (guard == 5 && (rest || (buf[DEC_P] & 1)));
        if (up) {
            int i = DEC_P;
            while (++buf[i] == 10) { buf[i] = 0; --i; }
        }
    }
    const uint8_t *src = buf + 1; int len = keep;
    if (buf[0]) { src = buf; len = keep + 1; }   /* 99..9 rounded up to 100..0: one more digit, all trailing zeros */
    while (len > 0 && src[len - 1] == 0) { --len; ++exp; }
    if (len > DEC_P) { len = DEC_P; }            /* unreachable: the extra digit only appears with trailing zeros */
    r.sign = (int8_t)sign; r.n = (int16_t)len; r.exp = exp;
    memcpy(r.d, src, (size_t)len);
    return r;
}

static dec dec_from_i64(int64_t v) {
    uint8_t b[24]; int n = 0, sign = v < 0;
    uint64_t u = sign ? (uint64_t)(-(v + 1)) + 1u : (uint64_t)v;
    uint8_t t[24]; while (u) { t[n++] = (uint8_t)(u % 10); u /= 10; }
    for (int i = 0; i < n; ++i) b[i] = t[n - 1 - i];
    return dec_round_into(sign, b, n, 0, 0);
}

/* compare magnitudes of two digit strings of equal length */
static int dig_cmp(const uint8_t *a, const uint8_t *b, int n) {
    for (int i = 0; i < n; ++i) if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return 0;
}

static dec dec_neg(dec a) { if (a.n) a.sign ^= 1; return a; }

static dec dec_add(dec a, dec b) {
    if (a.n == 0) return b;
    if (b.n == 0) return a;
    /* far apart: the smaller one cannot reach the rounding digit of the larger (and both are already <= DEC_P digits) */
    int adja = a.exp + a.n, adjb = b.exp + b.n;
    if (adja - adjb > 2 * DEC_P + 4) return a;
    if (adjb - adja > 2 * DEC_P + 4) return b;
    int e = a.exp < b.exp ? a.exp : b.exp;
    int la = a.n + (a.exp - e), lb = b.n + (b.exp - e), L = (la > lb ? la : lb) + 1;
    uint8_t A[DEC_W], B[DEC_W], R[DEC_W];
    memset(A, 0, (size_t)L); memset(B, 0, (size_t)L);
    memcpy(A + (L - la), a.d, (size_t)a.n);      /* right-aligned at exponent e; the shifted-in low digits stay zero */
    memcpy(B + (L - lb), b.d, (size_t)b.n);
    int sign;
    if (a.sign == b.sign) {
        int c = 0;
        for (int i = L - 1; i >= 0; --i) { int s = A[i] + B[i] + c; R[i] = (uint8_t)(s % 10); c = s / 10; }
        sign = a.sign;
    } else {
        int cm = dig_cmp(A, B, L);
        if (cm == 0) return dec_zero();
        const uint8_t *X = cm > 0 ? A : B, *Y = cm > 0 ? B : A;
        int br = 0;
        for (int i = L - 1; i >= 0; --i) { int s = X[i] - Y[i] - br; br = s < 0; R[i] = (uint8_t)(s + 10 * br); }
        sign = cm > 0 ? a.sign : b.sign;
    }
    return dec_round_into(sign, R, L, e, 0);
}
static dec dec_sub(dec a, dec b) { return dec_add(a, dec_neg(b)); }

static dec dec_mul(dec a, dec b) {
    if (a.n == 0 || b.n == 0) return dec_zero();
    int L = a.n + b.n;
    int acc[2 * DEC_P + 2]; memset(acc, 0, sizeof(acc));
    for (int i = 0; i < a.n; ++i) for (int j = 0; j < b.n; ++j) acc[i + j + 1] += a.d[i] * b.d[j];
    uint8_t R[2 * DEC_P + 2];
    int c = 0;
    for (int i = L - 1; i >= 0; --i) { int s = acc[i] + c; R[i] = (uint8_t)(s % 10); c = s / 10; }
    return dec_round_into(a.sign ^ b.sign, R, L, a.exp + b.exp, 0);
}

/* a / b, b != 0: long division to DEC_P + 2 significant quotient digits + sticky remainder */
static dec dec_div(dec a, dec b) {
    if (a.n == 0) return dec_zero();
    int k = DEC_P + 2 + b.n;                     /* zeros appended to the dividend: quotient has >= DEC_P + 2 digits */
    int L = a.n + k;
    uint8_t Q[DEC_W], rem[DEC_P + 2], den[DEC_P + 2];
    int nd = b.n + 1;                            /* remainder width: one digit more than the divisor */
    memset(rem, 0, (size_t)nd); den[0] = 0; memcpy(den + 1, b.d, (size_t)b.n);
    for (int i = 0; i < L; ++i) {
        memmove(rem, rem + 1, (size_t)(nd - 1));  /* rem = rem * 10 + next digit (rem < den, so its top digit is 0) */
        rem[nd - 1] = i < a.n ? a.d[i] : 0;
        int q = 0;
        while (dig_cmp(rem, den, nd) >= 0) {
            int br = 0;
            for (int j = nd - 1; j >= 0; --j) { int s = rem[j] - den[j] - br; br = s < 0; rem[j] = (uint8_t)(s + 10 * br); }
            ++q;
        }
        Q[i] = (uint8_t)q;
    }
    int sticky = 0;
    for (int j = 0; j < nd; ++j) sticky |= rem[j] != 0;
    return dec_round_into(a.sign ^ b.sign, Q, L, a.exp - b.exp - k, sticky);
}

static int dec_cmp(dec a, dec b) {              /* -1, 0, +1 */
    dec d = dec_sub(a, b);                       /* the SIGN of a rounded difference is the sign of the exact one */
    if (d.n == 0) return 0;
    return d.sign ? -1 : 1;
}
static int dec_sign(dec a) { return a.n == 0 ? 0 : (a.sign ? -1 : 1); }

/* float(Decimal): correctly rounded decimal -> binary64 (glibc strtod is correctly rounded) */
static double dec_to_double(dec a) {
    if (a.n == 0) return 0.0;
    char s[DEC_P + 24]; int p = 0;
    if (a.sign) s[p++] = '-';
    for (int i = 0; i < a.n; ++i) s[p++] = (char)('0' + a.d[i]);
    p += snprintf(s + p, sizeof(s) - (size_t)p, "e%d", a.exp);
    return strtod(s, NULL);
}

/* nearest integer (half to even); the ledger's values are integers +- ~1e-20, so this recovers the exact-ledger value */
static int64_t dec_to_i64_nearest(dec a) {
    if (a.n == 0) return 0;
    uint8_t dig[DEC_W]; int n = a.n; memcpy(dig, a.d, (size_t)a.n);
    int64_t v = 0;
    if (a.exp >= 0) { for (int i = 0; i < n; ++i) v = v * 10 + dig[i]; for (int i = 0; i < a.exp; ++i) v *= 10; }
    else {
        int ip = n + a.exp;                       /* digits before the point */
        for (int i = 0; i < ip; ++i) v = v * 10 + dig[i];
        int guard = ip >= 0 && ip < n ? dig[ip] : 0, rest = 0;
        if (ip < 0) guard = 0, rest = 1;          /* |a| < 0.1 */
        for (int i = (ip < 0 ? 0 : ip + 1); i < n; ++i) rest |= dig[i] != 0;
        if (guard > 5 || (guard == 5 && (rest || (v & 1)))) ++v;
    }
    return a.sign ? -v : v;
}

/* "[-]ddd[.ddd][e[+-]x]" -> dec, rounded to the context (test entry) */
static dec dec_from_str(const char *s) {
    int sign = 0; if (*s == '-') { sign = 1; ++s; } else if (*s == '+') ++s;
    uint8_t dig[DEC_W]; int n = 0, exp = 0, seen_pt = 0;
    for (; *s && *s != 'e' && *s != 'E'; ++s) {
        if (*s == '.') { seen_pt = 1; continue; }
        if (n < DEC_W) { dig[n++] = (uint8_t)(*s - '0'); if (seen_pt) --exp; }
    }
    if (*s == 'e' || *s == 'E') exp += atoi(s + 1);
    return dec_round_into(sign, dig, n, exp, 0);
}
/* scientific string, e.g. "-3.96e+2" ("0" for zero): Decimal(str) gives the same value back */
static void dec_to_str(dec a, char *out, size_t cap) {
    if (a.n == 0) { snprintf(out, cap, "0"); return; }
    size_t p = 0;
    if (a.sign && p + 1 < cap) out[p++] = '-';
    for (int i = 0; i < a.n && p + 1 < cap; ++i) out[p++] = (char)('0' + a.d[i]);
    snprintf(out + p, cap - p, "e%d", a.exp);
}
#endif
