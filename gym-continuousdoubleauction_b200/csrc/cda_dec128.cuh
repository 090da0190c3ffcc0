// cda_dec128.cuh — Decimal(prec 28, ROUND_HALF_EVEN) add / sub / mul / div / compare on fixed-width integers, host + device.
//
// The reference keeps money as Python `decimal.Decimal` (envs/account/account.py:124-231, calculate.py:5-55,
// cash_processor.py:15-97, agent/trader.py:108-151); its VWAP divisions leave ~1e-24 residues in cash / NAV, and a
// `cash >= order value` test at EXACT equality is decided by the residue's sign — the one place where an exact integer ledger
// can disagree with the reference.  Reproducing that needs the reference's arithmetic itself:
// value = (-1)^sign * c * 10^exp with c < 10^28 in an unsigned __int128, every operation computed exactly in 128 bits and
// rounded ONCE.  This header is that arithmetic; the device twin ledger (cda_twin.cuh) is built on it.  It is the same algorithm
// as oracle/dec128.h, which reproduces the reference's Decimal fields exactly inside the oracle (700 fuzzed configurations).
// The host compilation of THIS file is pinned against Python's decimal through cda_debug_dec_op, the device compilation
// through cda_debug_dec_op_device (tests/test_dec28.py).
//
// Domain: additions are exact for all operands; a product needs digits(a) + digits(b) <= 38 and a quotient a divisor below
// 2^32 (in the ledger one factor is always a size or a price).  Outside it `*range_err` is incremented and the result is not to
// be used.
//
// Speed (device): no 128-bit division anywhere — powers of ten come from a table, digit counts from the bit length, and every
// quotient is a chain of (64-bit / 32-bit) steps done with one FP64 multiply and an exact integer fix-up.
#pragma once
#include <stdint.h>

typedef unsigned __int128 cda_u128;
struct CdaDec { cda_u128 c; int exp; int sign; };   // c == 0 <=> zero (sign 0)

#define CDA_DEC_P 28
#if defined(__CUDACC__)
#define CDA_HD __host__ __device__ __forceinline__
#define CDA_HDN __host__ __device__ __noinline__ inline   /* add / mul / div (each with its rounding inlined) are real calls: with everything
                                                              inlined one replay is 50 k instructions and every step kernel carries a copy */
#else
#define CDA_HD inline
#define CDA_HDN inline
#endif

#define CDA_POW10_ROWS \
    {0x0000000000000000ULL, 0x0000000000000001ULL}, \
    {0x0000000000000000ULL, 0x000000000000000aULL}, \
    {0x0000000000000000ULL, 0x0000000000000064ULL}, \
    {0x0000000000000000ULL, 0x00000000000003e8ULL}, \
    {0x0000000000000000ULL, 0x0000000000002710ULL}, \
    {0x0000000000000000ULL, 0x00000000000186a0ULL}, \
    {0x0000000000000000ULL, 0x00000000000f4240ULL}, \
    {0x0000000000000000ULL, 0x0000000000989680ULL}, \
    {0x0000000000000000ULL, 0x0000000005f5e100ULL}, \
    {0x0000000000000000ULL, 0x000000003b9aca00ULL}, \
    {0x0000000000000000ULL, 0x00000002540be400ULL}, \
    {0x0000000000000000ULL, 0x000000174876e800ULL}, \
    {0x0000000000000000ULL, 0x000000e8d4a51000ULL}, \
    {0x0000000000000000ULL, 0x000009184e72a000ULL}, \
    {0x0000000000000000ULL, 0x00005af3107a4000ULL}, \
    {0x0000000000000000ULL, 0x00038d7ea4c68000ULL}, \
    {0x0000000000000000ULL, 0x002386f26fc10000ULL}, \
    {0x0000000000000000ULL, 0x016345785d8a0000ULL}, \
    {0x0000000000000000ULL, 0x0de0b6b3a7640000ULL}, \
    {0x0000000000000000ULL, 0x8ac7230489e80000ULL}, \
    {0x0000000000000005ULL, 0x6bc75e2d63100000ULL}, \
    {0x0000000000000036ULL, 0x35c9adc5dea00000ULL}, \
    {0x000000000000021eULL, 0x19e0c9bab2400000ULL}, \
    {0x000000000000152dULL, 0x02c7e14af6800000ULL}, \
    {0x000000000000d3c2ULL, 0x1bcecceda1000000ULL}, \
    {0x0000000000084595ULL, 0x161401484a000000ULL}, \
    {0x000000000052b7d2ULL, 0xdcc80cd2e4000000ULL}, \
    {0x00000000033b2e3cULL, 0x9fd0803ce8000000ULL}, \
    {0x00000000204fce5eULL, 0x3e25026110000000ULL}, \
    {0x00000001431e0faeULL, 0x6d7217caa0000000ULL}, \
    {0x0000000c9f2c9cd0ULL, 0x4674edea40000000ULL}, \
    {0x0000007e37be2022ULL, 0xc0914b2680000000ULL}, \
    {0x000004ee2d6d415bULL, 0x85acef8100000000ULL}, \
    {0x0000314dc6448d93ULL, 0x38c15b0a00000000ULL}, \
    {0x0001ed09bead87c0ULL, 0x378d8e6400000000ULL}, \
    {0x0013426172c74d82ULL, 0x2b878fe800000000ULL}, \
    {0x00c097ce7bc90715ULL, 0xb34b9f1000000000ULL}, \
    {0x0785ee10d5da46d9ULL, 0x00f436a000000000ULL}, \
    {0x4b3b4ca85a86c47aULL, 0x098a224000000000ULL}
static const unsigned long long cda_pow10_host[39][2] = { CDA_POW10_ROWS };
#if defined(__CUDACC__)
__device__ const unsigned long long cda_pow10_dev[39][2] = { CDA_POW10_ROWS };
#endif

// 10^k for k <= 19, normalised (shifted left until its top bit is set), with the Moller-Granlund reciprocal floor((2^128 - 1) / d) - 2^64
// of the normalised divisor and the shift: a 128-bit / 10^k quotient is then two multiply-high steps instead of a long division
#define CDA_MG10_ROWS \
    {0x8000000000000000ULL, 0xffffffffffffffffULL, 63ULL}, \
    {0xa000000000000000ULL, 0x9999999999999999ULL, 60ULL}, \
    {0xc800000000000000ULL, 0x47ae147ae147ae14ULL, 57ULL}, \
    {0xfa00000000000000ULL, 0x0624dd2f1a9fbe76ULL, 54ULL}, \
    {0x9c40000000000000ULL, 0xa36e2eb1c432ca57ULL, 50ULL}, \
    {0xc350000000000000ULL, 0x4f8b588e368f0846ULL, 47ULL}, \
    {0xf424000000000000ULL, 0x0c6f7a0b5ed8d36bULL, 44ULL}, \
    {0x9896800000000000ULL, 0xad7f29abcaf48578ULL, 40ULL}, \
    {0xbebc200000000000ULL, 0x5798ee2308c39df9ULL, 37ULL}, \
    {0xee6b280000000000ULL, 0x12e0be826d694b2eULL, 34ULL}, \
    {0x9502f90000000000ULL, 0xb7cdfd9d7bdbab7dULL, 30ULL}, \
    {0xba43b74000000000ULL, 0x5fd7fe17964955fdULL, 27ULL}, \
    {0xe8d4a51000000000ULL, 0x19799812dea11197ULL, 24ULL}, \
    {0x9184e72a00000000ULL, 0xc25c268497681c26ULL, 20ULL}, \
    {0xb5e620f480000000ULL, 0x6849b86a12b9b01eULL, 17ULL}, \
    {0xe35fa931a0000000ULL, 0x203af9ee756159b2ULL, 14ULL}, \
    {0x8e1bc9bf04000000ULL, 0xcd2b297d889bc2b6ULL, 10ULL}, \
    {0xb1a2bc2ec5000000ULL, 0x70ef54646d496892ULL, 7ULL}, \
    {0xde0b6b3a76400000ULL, 0x2725dd1d243aba0eULL, 4ULL}, \
    {0x8ac7230489e80000ULL, 0xd83c94fb6d2ac34aULL, 0ULL}
static const unsigned long long cda_mg10_host[20][3] = { CDA_MG10_ROWS };
#if defined(__CUDACC__)
__device__ const unsigned long long cda_mg10_dev[20][3] = { CDA_MG10_ROWS };
#endif

CDA_HD cda_u128 cda_dec_pow10(int k) {               // 0 <= k <= 38
#if defined(__CUDA_ARCH__)
    return ((cda_u128)cda_pow10_dev[k][0] << 64) | cda_pow10_dev[k][1];
#else
    return ((cda_u128)cda_pow10_host[k][0] << 64) | cda_pow10_host[k][1];
#endif
}
CDA_HD int cda_dec_bitlen(cda_u128 x) {
    const unsigned long long hi = (unsigned long long)(x >> 64), lo = (unsigned long long)x;
#if defined(__CUDA_ARCH__)
    return hi ? 128 - __clzll((long long)hi) : (lo ? 64 - __clzll((long long)lo) : 0);
#else
    return hi ? 128 - __builtin_clzll(hi) : (lo ? 64 - __builtin_clzll(lo) : 0);
#endif
}
CDA_HD int cda_dec_ndigits(cda_u128 x) {             // 0 for x == 0; 39 for x >= 10^38
    const int t = (cda_dec_bitlen(x) * 1233) >> 12;  // floor(bits * log10 2), exact for bits <= 128 (checked exhaustively)
    if (t >= 38) return x >= cda_dec_pow10(38) ? 39 : 38;
    return t + (x >= cda_dec_pow10(t) ? 1 : 0);
}
// n / d and n % d for d < 2^32, n < d * 2^32 (one step of a long division): FP64 estimate (error <= 1), exact integer fix-up
CDA_HD unsigned cda_div64_32(unsigned long long n, unsigned d, unsigned *rem) {
#if defined(__CUDA_ARCH__)
    long long q = (long long)__double2ull_rz(__ull2double_rz(n) * __drcp_rn((double)d));
    long long r = (long long)(n - (unsigned long long)q * d);
    if (r < 0) { --q; r += d; }
    if (r < 0) { --q; r += d; }
    if (r >= (long long)d) { ++q; r -= d; }
    if (r >= (long long)d) { ++q; r -= d; }
    *rem = (unsigned)r;
    return (unsigned)q;
#else
    *rem = (unsigned)(n % d);
    return (unsigned)(n / d);
#endif
}
// x / d, remainder in *rem, for 0 < d < 2^32
CDA_HD cda_u128 cda_u128_divmod_small(cda_u128 x, unsigned d, unsigned *rem) {
    const unsigned long long hi = (unsigned long long)(x >> 64), lo = (unsigned long long)x;
    unsigned r = 0;
    const unsigned q3 = cda_div64_32(hi >> 32, d, &r);
    const unsigned q2 = cda_div64_32(((unsigned long long)r << 32) | (hi & 0xffffffffULL), d, &r);
    const unsigned q1 = cda_div64_32(((unsigned long long)r << 32) | (lo >> 32), d, &r);
    const unsigned q0 = cda_div64_32(((unsigned long long)r << 32) | (lo & 0xffffffffULL), d, &r);
    *rem = r;
    return ((cda_u128)(((unsigned long long)q3 << 32) | q2) << 64) | (((unsigned long long)q1 << 32) | q0);
}
CDA_HD unsigned long long cda_umulhi64(unsigned long long a, unsigned long long b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (unsigned long long)(((cda_u128)a * b) >> 64);
#endif
}
// (u1 * 2^64 + u0) / d for a normalised d (top bit set), u1 < d, with v = floor((2^128 - 1) / d) - 2^64 (Moller & Granlund 2011, alg. 4)
CDA_HD unsigned long long cda_div2by1(unsigned long long u1, unsigned long long u0, unsigned long long d, unsigned long long v, unsigned long long *rem) {
    unsigned long long q0 = v * u1, q1 = cda_umulhi64(v, u1);
    const unsigned long long s0 = q0 + u0;
    q1 += u1 + (s0 < q0 ? 1ULL : 0ULL) + 1ULL;
    q0 = s0;
    unsigned long long r = u0 - q1 * d;
    if (r > q0) { --q1; r += d; }
    if (r >= d) { ++q1; r -= d; }
    *rem = r;
    return q1;
}
// x / 10^k for 0 <= k <= 38; the remainder is x - q * 10^k (the caller multiplies back: exact and cheap)
CDA_HD cda_u128 cda_u128_div_pow10(cda_u128 x, int k) {
    while (k > 0) {
        const int j = k > 19 ? 19 : k;
        k -= j;
#if defined(__CUDA_ARCH__)
        const unsigned long long d = cda_mg10_dev[j][0], v = cda_mg10_dev[j][1]; const int s = (int)cda_mg10_dev[j][2];
#else
        const unsigned long long d = cda_mg10_host[j][0], v = cda_mg10_host[j][1]; const int s = (int)cda_mg10_host[j][2];
#endif
        const unsigned long long hi = (unsigned long long)(x >> 64), lo = (unsigned long long)x;
        const unsigned long long x2 = s ? hi >> (64 - s) : 0ULL, x1 = s ? (hi << s) | (lo >> (64 - s)) : hi, x0 = lo << s;
        unsigned long long r;
        const unsigned long long qh = cda_div2by1(x2, x1, d, v, &r);
        const unsigned long long ql = cda_div2by1(r, x0, d, v, &r);
        x = ((cda_u128)qh << 64) | ql;
    }
    return x;
}
CDA_HD CdaDec cda_dec_zero() { CdaDec r; r.c = 0; r.exp = 0; r.sign = 0; return r; }

// x * 10^exp (+ sticky: something non-zero beyond x) -> 28 significant digits, half to even
CDA_HD CdaDec cda_dec_round(int sign, cda_u128 x, int exp, int sticky) {
    CdaDec r = cda_dec_zero();
    if (x == 0) return r;
    const int nd = cda_dec_ndigits(x);
    if (nd > CDA_DEC_P) {
        int k = nd - CDA_DEC_P;
        const cda_u128 p = cda_dec_pow10(k), half = p >> 1;
        cda_u128 q = cda_u128_div_pow10(x, k);
        const cda_u128 rem = x - q * p;
        if (rem > half || (rem == half && (sticky || (q & 1)))) ++q;
        if (q == cda_dec_pow10(CDA_DEC_P)) { q = cda_dec_pow10(CDA_DEC_P - 1); ++k; }
        x = q; exp += k;
    }
    r.c = x; r.exp = exp; r.sign = sign;
    return r;
}
CDA_HD CdaDec cda_dec_from_i64(long long v) {        // |v| < 2^63 < 10^19: never rounded
    CdaDec r;
    r.sign = v < 0;
    r.c = (cda_u128)(r.sign ? (unsigned long long)(-(v + 1)) + 1ULL : (unsigned long long)v);
    r.exp = 0;
    if (r.c == 0) r.sign = 0;
    return r;
}
CDA_HD CdaDec cda_dec_neg(CdaDec a) { if (a.c) a.sign ^= 1; return a; }
CDA_HDN CdaDec cda_dec_strip(CdaDec a) {              // value-preserving: drop trailing zeros of the coefficient
    while (a.c && (a.c >> 64) == 0) {               // (the common case, a divisor that is a position size: plain 64-bit arithmetic)
        const unsigned long long c = (unsigned long long)a.c;
        if (c % 10ULL) return a;
        a.c = c / 10ULL; ++a.exp;
    }
    while (a.c) {
        unsigned r;
        const cda_u128 q = cda_u128_divmod_small(a.c, 10u, &r);
        if (r) break;
        a.c = q; ++a.exp;
    }
    return a;
}

CDA_HDN CdaDec cda_dec_add(CdaDec a, CdaDec b) {
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
    if (shift <= 38) { const cda_u128 p = cda_dec_pow10(shift); y = cda_u128_div_pow10(b.c, shift); sticky = (b.c - y * p) != 0; }
    const cda_u128 r = a.sign == b.sign ? x + y : x - y - (sticky ? 1 : 0);
    return cda_dec_round(a.sign, r, a.exp - t, sticky);
}
CDA_HD CdaDec cda_dec_sub(CdaDec a, CdaDec b) { return cda_dec_add(a, cda_dec_neg(b)); }

CDA_HDN CdaDec cda_dec_mul(CdaDec a, CdaDec b, int *range_err) {
    if (a.c == 0 || b.c == 0) return cda_dec_zero();
    if (cda_dec_ndigits(a.c) + cda_dec_ndigits(b.c) > 38) {
        a = cda_dec_strip(a); b = cda_dec_strip(b);
        if (cda_dec_ndigits(a.c) + cda_dec_ndigits(b.c) > 38) { ++*range_err; return cda_dec_zero(); }
    }
    return cda_dec_round(a.sign ^ b.sign, a.c * b.c, a.exp + b.exp, 0);
}
// a / b, b != 0 with a coefficient below 2^32 after stripping: dividend scaled to 38 digits, quotient >= 29 digits + sticky remainder
CDA_HDN CdaDec cda_dec_div(CdaDec a, CdaDec b, int *range_err) {
    if (a.c == 0) return cda_dec_zero();
    b = cda_dec_strip(b);
    const int nb = cda_dec_ndigits(b.c), na = cda_dec_ndigits(a.c), k = 38 - na;
    if (na + k - nb < CDA_DEC_P + 1 || (b.c >> 32) != 0) { ++*range_err; return cda_dec_zero(); }
    const cda_u128 x = a.c * cda_dec_pow10(k);
    unsigned rem;
    const cda_u128 q = cda_u128_divmod_small(x, (unsigned)b.c, &rem);
    return cda_dec_round(a.sign ^ b.sign, q, a.exp - k - b.exp, rem != 0);
}
CDA_HD int cda_dec_cmp(CdaDec a, CdaDec b) {          // -1, 0, +1 (the sign of a rounded difference is the sign of the exact one)
    const CdaDec d = cda_dec_sub(a, b);
    return d.c == 0 ? 0 : (d.sign ? -1 : 1);
}
