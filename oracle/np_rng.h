/* TEST INFRASTRUCTURE (CPU oracle) — not product code.
 *
 * Restatement of the part of numpy's `Generator(PCG64(SeedSequence(seed)))` stream that
 * the reference env consumes.  numpy is a third-party dependency of the reference
 * (requirements-lock.txt: numpy==2.5.2; this image: 2.3.5) and its source is absent from
 * /root/reference, so this follows numpy's published algorithms:
 *
 *   - SeedSequence entropy pool + generate_state   (numpy/random/bit_generator.pyx)
 *   - PCG64 = pcg_setseq_128_xsl_rr_64, 128-bit LCG, default multiplier (O'Neill 2014)
 *   - next_uint32 buffering of the high half        (numpy/random/src/pcg64/pcg64.h)
 *   - Lemire bounded 32-bit integers                 (distributions.c, buffered_bounded_lemire_uint32)
 *   - masked-rejection `random_interval` used by Generator.shuffle/permutation
 *   - 256-block ziggurat `random_standard_normal`    (distributions.c)
 *
 * anchored on the reference's own call sites:
 *   continuousDoubleAuction_env.py:221   self.np_random.integers(low, high + 1)
 *   action_helper.py:331-333             self.np_random.normal(mul * mean, sigma, 1)
 *   action_helper.py:198-199             rng.permutation(len(actions))
 * and pinned bit-for-bit against the installed numpy by tests/test_np_rng.py.
 */
#ifndef ORC_NP_RNG_H
#define ORC_NP_RNG_H
#include <math.h>
#include <stdint.h>
#include "zig_tables.h"

typedef unsigned __int128 orc_u128;

typedef struct {
    orc_u128 state;
    orc_u128 inc;
    uint32_t has_uint32;
    uint32_t uinteger;
} orc_rng;

#define ORC_PCG_MULT ((((orc_u128)2549297995355413924ULL) << 64) | (orc_u128)4865540595714422341ULL)

/* ---- SeedSequence (pool of 4 uint32) ---------------------------------------------------- */
static inline uint32_t orc_ss_hashmix(uint32_t value, uint32_t *hash_const) {
    value ^= *hash_const;
    *hash_const *= 0x931e8875u;
    value *= *hash_const;
    value ^= value >> 16;
    return value;
}
static inline uint32_t orc_ss_mix(uint32_t x, uint32_t y) {
    uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y;
    r ^= r >> 16;
    return r;
}
/* seed: non-negative integer < 2^64 (what `reset(seed=int)` passes).  out: 4 uint64 words =
 * SeedSequence(seed).generate_state(4, uint64). */
static inline void orc_seedseq_state(uint64_t seed, uint64_t out[4]) {
    uint32_t ent[2];
    int n_ent = 1;
    ent[0] = (uint32_t)(seed & 0xffffffffu);
    ent[1] = (uint32_t)(seed >> 32);
    if (ent[1] != 0) n_ent = 2;
    uint32_t pool[4];
    uint32_t hc = 0x43b0d7e5u;
    for (int i = 0; i < 4; ++i) pool[i] = orc_ss_hashmix(i < n_ent ? ent[i] : 0u, &hc);
    for (int s = 0; s < 4; ++s)
        for (int d = 0; d < 4; ++d)
            if (s != d) pool[d] = orc_ss_mix(pool[d], orc_ss_hashmix(pool[s], &hc));
    uint32_t hb = 0x8b51f9ddu;
    uint32_t w[8];
    for (int i = 0; i < 8; ++i) {
        uint32_t v = pool[i & 3];
        v ^= hb;
        hb *= 0x58f38dedu;
        v *= hb;
        v ^= v >> 16;
        w[i] = v;
    }
    for (int i = 0; i < 4; ++i) out[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
}

static inline void orc_rng_step(orc_rng *r) { r->state = r->state * ORC_PCG_MULT + r->inc; }

static inline void orc_rng_seed(orc_rng *r, uint64_t seed) {
    uint64_t s[4];
    orc_seedseq_state(seed, s);
    orc_u128 initstate = (((orc_u128)s[0]) << 64) | s[1];
    orc_u128 initseq = (((orc_u128)s[2]) << 64) | s[3];
    r->state = 0;
    r->inc = (initseq << 1) | 1u;
    orc_rng_step(r);
    r->state += initstate;
    orc_rng_step(r);
    r->has_uint32 = 0;
    r->uinteger = 0;
}

static inline uint64_t orc_next_u64(orc_rng *r) {
    orc_rng_step(r);
    uint64_t hi = (uint64_t)(r->state >> 64), lo = (uint64_t)r->state;
    uint64_t x = hi ^ lo;
    unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((-rot) & 63));
}
static inline uint32_t orc_next_u32(orc_rng *r) {
    if (r->has_uint32) {
        r->has_uint32 = 0;
        return r->uinteger;
    }
    uint64_t n = orc_next_u64(r);
    r->has_uint32 = 1;
    r->uinteger = (uint32_t)(n >> 32);
    return (uint32_t)(n & 0xffffffffu);
}
static inline double orc_next_double(orc_rng *r) {
    return (double)(orc_next_u64(r) >> 11) * (1.0 / 9007199254740992.0);
}

/* Generator.integers(lo, hi_exclusive) for a range that fits 32 bits (scalar int64 path). */
static inline int64_t orc_integers(orc_rng *r, int64_t lo, int64_t hi_excl) {
    uint64_t rng = (uint64_t)(hi_excl - 1 - lo);
    if (rng == 0) return lo;
    uint32_t rng32 = (uint32_t)rng;
    uint32_t rng_excl = rng32 + 1u;
    uint64_t m = (uint64_t)orc_next_u32(r) * rng_excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < rng_excl) {
        uint32_t threshold = (0xffffffffu - rng32) % rng_excl;
        while (leftover < threshold) {
            m = (uint64_t)orc_next_u32(r) * rng_excl;
            leftover = (uint32_t)m;
        }
    }
    return lo + (int64_t)(m >> 32);
}

/* random_interval(max): uniform on [0, max] by masked rejection over next_uint32. */
static inline uint32_t orc_interval(orc_rng *r, uint32_t max) {
    if (max == 0) return 0;
    uint32_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    while ((v = (orc_next_u32(r) & mask)) > max) {}
    return v;
}

/* Generator.permutation(n): arange(n) shuffled by Fisher-Yates from the top.  n<=1 draws nothing. */
static inline void orc_permutation(orc_rng *r, int n, int *out) {
    for (int i = 0; i < n; ++i) out[i] = i;
    for (int i = n - 1; i >= 1; --i) {
        uint32_t j = orc_interval(r, (uint32_t)i);
        int t = out[i]; out[i] = out[j]; out[j] = t;
    }
}

static inline double orc_standard_normal(orc_rng *g) {
    for (;;) {
        uint64_t r = orc_next_u64(g);
        int idx = (int)(r & 0xff);
        r >>= 8;
        int sign = (int)(r & 0x1);
        uint64_t rabs = (r >> 1) & 0x000fffffffffffffULL;
        double x = (double)rabs * orc_zig_wi[idx];
        if (sign) x = -x;
        if (rabs < orc_zig_ki[idx]) return x;
        if (idx == 0) {
            for (;;) {
                double xx = -ORC_ZIG_NOR_INV_R * log1p(-orc_next_double(g));
                double yy = -log1p(-orc_next_double(g));
                if (yy + yy > xx * xx)
                    return ((rabs >> 8) & 0x1) ? -(ORC_ZIG_NOR_R + xx) : ORC_ZIG_NOR_R + xx;
            }
        } else {
            if (((orc_zig_fi[idx - 1] - orc_zig_fi[idx]) * orc_next_double(g) + orc_zig_fi[idx]) <
                exp(-0.5 * x * x))
                return x;
        }
    }
}
#endif
