// cda_kernels.cuh — sm_100a device code of the vectorised continuous-double-auction env.
//
// One WARP steps one MARKET.  The market's resting orders live in a flat, dense, unsorted
// order pool per side (structure-of-arrays: price|trader, qty, order_id, timestamp, seq),
// staged global->shared with cp.async.bulk (TMA bulk copy, SASS UBLKCP) behind an mbarrier;
// the sorted-tree / FIFO-list / order-map machinery of the reference is replaced by
// warp-wide reductions over that pool (redux.sync min/max/add + ballot):
//     best price            = redux min/max over price
//     head of a price level = redux min over `seq` among orders at that price
//     a trader's order      = redux min over `seq` (limit/cancel) or `timestamp` (modify)
// `seq` is a per-market insertion counter: in the reference an order is appended to its level's
// FIFO list and to the side's order_map dict at the same moment (ordertree.py:44-55) and both
// positions are kept by an in-place quantity update (orderbook.py:245-248), so one key encodes
// both the time priority inside a level and the order_map iteration order that
// Trader._get_order_ID depends on (trader.py:254-287).
// The A accounts of the market live in the registers of lanes 0..A-1 for the whole step; the
// RNG (numpy's PCG64 + ziggurat, so reset(seed) reproduces the reference's stream) is
// evaluated redundantly by all lanes (warp-uniform, no broadcast needed).
//
// Reference call stack mirrored by market_step():  continuousDoubleAuction_env.py:265-309.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cda_b200.h"
#include "cda_zig_tables.cuh"
#include "cda_twin.cuh"

#define CDA_FULL 0xffffffffu
#define CDA_SCAN_PRAGMA _Pragma("unroll 1")
#ifndef CDA_BEST_CACHE
#define CDA_BEST_CACHE 1      /* 1: the best price of each side is cached (seeded from the header, maintained by append / remove): a scan only
                                 after an order AT the best price was removed.  (Defined HERE, above its first use: until r02e the switch sat
                                 below pool_best(), which therefore always scanned.) */
#endif
#ifndef CDA_SPEC_TILES
#define CDA_SPEC_TILES 0      /* > 0: the first CDA_SPEC_TILES tiles (32 orders each) of BOTH pool sides are fetched speculatively together with the
                                 account block, before the header has said how many orders are live: saves one dependent HBM round trip on a
                                 cold cache at the price of over-fetching up to 640 B * CDA_SPEC_TILES per side */
#endif
#ifndef CDA_WARP_ACT_TMA
#define CDA_WARP_ACT_TMA 0    /* 1: on the plain device step every WARP stages its own market's action rows behind its own mbarrier (five 4*A-byte bulk
                                 copies), so the kernel starts without a __syncthreads; the routed (host) bodies keep ONE set of copies per CTA:
                                 reads from host memory are bound by the number of requests */
#endif
#ifndef CDA_PREFETCH_TABLES
#define CDA_PREFETCH_TABLES 0 /* 1: every CTA asks L2 for the ziggurat / jump-ahead tables at kernel entry (they are indexed by data that arrives two
                                 dependent loads into the kernel; after an L2 flush the first touch per SM goes to HBM) */
#endif
#ifndef CDA_SCAN_UNROLL
#define CDA_SCAN_UNROLL 0     /* 1: the pool scans (best price, a trader's order, head of a level) read ALL tiles of a side first (CAP/32 predicated loads in
                                 flight at once) and compare afterwards, instead of one dependent load -> compare round per tile: the kernel is bound by
                                 dependent-issue latency, not by issue slots */
#endif
#ifndef CDA_TOPK_UNROLL
#define CDA_TOPK_UNROLL 0     /* 1: the same for the first sweep of the top-K snapshot (occupancy masks of both sides) */
#endif
#define CDA_HDR_BYTES 192
#define CDA_POOL_FIELDS 5 /* 0 pt (trader<<24|price), 1 qty, 2 order_id, 3 timestamp, 4 seq */
#define CDA_PRICE_MASK 0x00ffffffu
#define CDA_FLAG_TAPE 1u
#define CDA_TILE_WORDS (CDA_POOL_FIELDS * 32)   /* one tile: 5 fields x 32 orders */
// word offset of order i inside a side (field 0); field f is at + f*32
#define CDA_EOFF(i) ((((i) >> 5) * CDA_TILE_WORDS) + ((i) & 31))

// ------------------------------------------------------------------------------------------
// Per-market block in HBM (stride bytes, 128-B aligned), see DESIGN.md "Data layout":
//   [0,192)            header words (below)
//   [off_acct, ...)    accounts, SoA inside the market: cash[A] hold[A] cost[A] nav[A] prev_nav[A]
//                      max_nav[A] (i64), pos[A] (i32), num_trades[A] (u32), stepctr[A] (u32), twin flags[A] (u32): 64*A bytes
//   [off_twin, ...)    decimal_ledger only: CdaTwinStored[A] (64 B each), then the event journals u64[A][CDA_JRN_E] (cda_twin.cuh)
//   [off_hist, ...)    snapshot ring  f32[n_hist][42]
//   [off_pool, ...)    order pool     u32[2 sides][cap/32 tiles][5 fields][32]  ("blocked SoA": the live
//                      prefix of a side is ONE contiguous run of whole tiles => one bulk copy per side,
//                      and a lane-strided scan reads 32 consecutive words => no bank conflicts)
// Header words (u32 index):
//   0 time  1 next_order_id  2 seqctr  3 t_step  4 last_price  5 flags  6 done_mask  7 status
//   8 n_bid 9 n_ask 10 rng_has_uint32 11 rng_uinteger 12..19 rng state_hi,state_lo,inc_hi,inc_lo (u64)
//   20..39 pre-step raw top-K prices (bid[10], ask[10]; 0 = empty)  40 best_bid 41 best_ask 42..47 pad
// ------------------------------------------------------------------------------------------
struct CdaDevCfg {
    int A, n_hist, max_step, tick;
    long long init_cash;
    int min_size;
    float mkt_mul, lim_mul;   // action_helper.py:46-47 (cast to f32 like numpy's weak python float)
    int price_lo, price_hi;
    int cap, fill_cap;
    int fill_tape;            // 1: the fill log is a TAPE — a ring of the last fill_cap fills of the market across steps and launches (header word 42 counts
                              //    all fills since the reset; the reference's unbounded LOB.tape, orderbook.py:20,140), instead of one step's fills
    double c_order, c_trade, c_dd, c_passive, c_loss;
    unsigned off_acct, off_hist, off_pool, stride;
    int W;                    // n_hist * 42
    int dec;                  // 1: decimal_ledger — the deferred Decimal(28) twin of cda_twin.cuh decides exact-equality ties like the reference
    unsigned off_twin, off_jrn;
};

struct CdaStepParams {
    CdaDevCfg cfg;
    unsigned char *state;
    int M;
    const int *cat; const float *mean; const float *sigma; const int *pcode; const int *poff;
    int acct_tma;  // 1: the market's account block (60*A bytes) is staged global -> shared with cp.async.bulk (needs A % 4 == 0); 0: plain loads fill the same tile
    int act_tma;   // 1: every CTA stages its markets' five action rows global/pinned-host -> shared with cp.async.bulk (needs A % 4 == 0, 16-B aligned arrays)
    int act_mstride;   // words between consecutive markets' rows of an action array: A (five [M][A] arrays) or 5*A (ONE market-major block i32[M][5][A])
    int act_packed;    // 1: market-major block: the five rows of a CTA's markets are ONE contiguous run -> one bulk copy per CTA instead of five
    int rec_inline;    // 1 (with ring_out): the result record {f64 reward[A]; u8 terminated, truncated; pad} is stored right behind the newest snapshot
                       //    (the head of slot ring_slot + 1), in the same store instructions: no separate PCIe write transactions for it
    float *obs; double *reward; unsigned char *term; unsigned char *trunc;
    int *fills; int *fill_counts;
    int *act_log;      // optional i32[M][A][4]: the decoded actions of this step (type, side, size, price; side -1 = pass / absent) — the reference's LOB_actions
    // fused random-policy rollout (cda_rollout_random): num_steps > 0 => actions are generated
    int num_steps; unsigned long long policy_seed;
    unsigned long long *prof;   // CDA_PROFILE_PHASES builds: per-phase cycle sums [16]
    float *obs_hi; int obs_split;     // rows m >= obs_split go to obs_hi (device staging, DMA'd after the kernel) instead of obs
    int obs_stride;                   // floats between consecutive markets' obs rows (W when densely packed)
    int flag_pack;                    // 1: truncated lives in the byte after terminated (packed result records): both leave in one 16-bit store
    int reward_stride, flag_stride;   // doubles between markets' reward rows (A when dense); bytes between markets' flags (1 when dense)
    float *ring_out; int ring_slot;   // host ring / window: newest snapshot only, at slot ring_slot of row m (+ a mirror copy n_hist slots later)
    int ring_stride, ring_mirror;     // floats per market row of ring_out; 1 = also write the mirror copy
    int ring_pad;                     // > 0: the cell written at ring_slot is padded with zeros to this many words (dense plane output: a market's
                                      //      snapshot + record fill one aligned 256-B cell, so every store is a whole 128-B line)
    // fused all-gather epilogue: every output store (obs row / newest snapshot + record / reward / flags / completion flag) is REPLICATED to
    // rep_n destinations: the same address plus rep_delta[g] bytes (g = 0: the address itself).  The destinations are the ranks' gather
    // windows, peer-mapped over NVLink (CUDA IPC), all laid out alike, so one base pointer (already offset to this rank's rows) and one
    // byte delta per peer describe them all.  rep_n = 0 / 1: ordinary single-destination outputs.
    int rep_n;
    long long rep_delta[CDA_MAX_PEERS];
    // completion doorbell (host paths): every warp fences its output stores and counts itself on done_ctr (device memory); the last one
    // resets the counter and stores done_seq to done_flag (mapped pinned host word), which the host polls instead of synchronising the
    // stream: it sees the results a few microseconds before the driver sees the kernel retire
    unsigned *done_ctr; unsigned *done_flag; unsigned done_seq;
    unsigned *status_flag;            // mapped pinned host word: set to 1 by any market that ends the step with a non-zero sticky status
    // resident step server (cda_serve_*; the <ROLLOUT = 1, ROUTED = 1> body): the kernel stays on the SMs between steps with every market's
    // book and ledger in shared memory; the host rings a message word per step instead of launching (see "resident step server" below)
    const unsigned long long *srv_go_host;   // mapped pinned message word written by the host
    unsigned long long *srv_go_dev;          // CDA_SRV_COPIES copies of the newest message, 128 B apart (device memory), republished by the poller CTA
    unsigned *srv_done_dev;                  // device copy of the completion word (the poller's idle clock starts when a step has completed)
    unsigned *srv_err;                       // mapped pinned word: a worker's watchdog fired (the poller never answered)
    const unsigned char *srv_act_base;       // device alias of the pinned action area; a message carries the step's block as an offset from it
    unsigned srv_next;                       // number of the first step this launch serves
    unsigned long long srv_lease_ns, srv_watchdog_ns;
    int srv_act_mode;                        // 0: one cp.async.bulk per market and step from mapped host memory; 1: volatile loads
};

// ------------------------------------ numpy-exact RNG --------------------------------------
// PCG64 (pcg_setseq_128_xsl_rr_64) + numpy's buffered next_uint32 + ziggurat normal +
// masked-rejection interval + Lemire bounded ints.  Restated from the published algorithms
// (numpy is a third-party dependency of the reference); call sites in the reference:
// continuousDoubleAuction_env.py:221, action_helper.py:331-333, action_helper.py:198-199.
struct CdaRng {
    unsigned long long shi, slo, ihi, ilo;
    unsigned has32, u32;
};
__device__ __forceinline__ void rng_step(CdaRng &r) {
    const unsigned long long MH = 2549297995355413924ULL, ML = 4865540595714422341ULL;
    unsigned long long lo = r.slo * ML;
    unsigned long long hi = __umul64hi(r.slo, ML) + r.shi * ML + r.slo * MH;
    unsigned long long nlo = lo + r.ilo;
    unsigned long long carry = nlo < lo ? 1ULL : 0ULL;
    r.slo = nlo;
    r.shi = hi + r.ihi + carry;
}
__device__ __forceinline__ unsigned long long rng_u64(CdaRng &r) {
    rng_step(r);
    unsigned long long x = r.shi ^ r.slo;
    unsigned rot = (unsigned)(r.shi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}
__device__ __forceinline__ unsigned rng_u32(CdaRng &r) {
    if (r.has32) { r.has32 = 0; return r.u32; }
    unsigned long long n = rng_u64(r);
    r.has32 = 1; r.u32 = (unsigned)(n >> 32);
    return (unsigned)n;
}
__device__ __forceinline__ double rng_double(CdaRng &r) {
    return (double)(rng_u64(r) >> 11) * (1.0 / 9007199254740992.0);
}
// wedge / tail of the ziggurat (~1.2 % of draws); may consume further numbers.  Takes and returns the
// generator BY VALUE: a by-reference parameter of a non-inlined function would force the caller's
// generator state into local memory for the whole kernel.
struct CdaNormalRet { double x; unsigned long long shi, slo; };
__device__ __noinline__ CdaNormalRet rng_normal_slow(unsigned long long shi, unsigned long long slo, unsigned long long ihi,
                                                     unsigned long long ilo, int idx, unsigned long long rabs, double x) {
    CdaRng g; g.shi = shi; g.slo = slo; g.ihi = ihi; g.ilo = ilo; g.has32 = 0; g.u32 = 0;
    for (;;) {
        if (idx == 0) {
            for (;;) {
                double xx = -CDA_ZIG_NOR_INV_R * log1p(-rng_double(g));
                double yy = -log1p(-rng_double(g));
                if (yy + yy > xx * xx)
                    return CdaNormalRet{((rabs >> 8) & 1ULL) ? -(CDA_ZIG_NOR_R + xx) : CDA_ZIG_NOR_R + xx, g.shi, g.slo};
            }
        } else {
            double u = rng_double(g);
            if (((__ldg(&cda_zig_fi[idx - 1]) - __ldg(&cda_zig_fi[idx])) * u + __ldg(&cda_zig_fi[idx])) < exp(-0.5 * x * x)) return CdaNormalRet{x, g.shi, g.slo};
        }
        unsigned long long r = rng_u64(g);
        idx = (int)(r & 0xff);
        r >>= 8;
        int sign = (int)(r & 1ULL);
        rabs = (r >> 1) & 0x000fffffffffffffULL;
        x = (double)rabs * __ldg(&cda_zig_wi[idx]);
        if (sign) x = -x;
        if (rabs < __ldg(&cda_zig_ki[idx])) return CdaNormalRet{x, g.shi, g.slo};
    }
}
__device__ __forceinline__ double rng_normal(CdaRng &g) {
    unsigned long long r = rng_u64(g);
    int idx = (int)(r & 0xff);
    r >>= 8;
    int sign = (int)(r & 1ULL);
    unsigned long long rabs = (r >> 1) & 0x000fffffffffffffULL;
    double x = (double)rabs * __ldg(&cda_zig_wi[idx]);
    if (sign) x = -x;
    if (rabs < __ldg(&cda_zig_ki[idx])) return x;
    const CdaNormalRet rr = rng_normal_slow(g.shi, g.slo, g.ihi, g.ilo, idx, rabs, x);
    g.shi = rr.shi; g.slo = rr.slo;
    return rr.x;
}
__device__ __forceinline__ unsigned rng_interval(CdaRng &r, unsigned max) {
    if (max == 0) return 0;
    unsigned mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    unsigned v;
    while ((v = (rng_u32(r) & mask)) > max) {}
    return v;
}
__device__ __forceinline__ long long rng_integers(CdaRng &r, long long lo, long long hi_excl) {
    unsigned long long rng = (unsigned long long)(hi_excl - 1 - lo);
    if (rng == 0) return lo;
    unsigned rng32 = (unsigned)rng, rng_excl = rng32 + 1u;
    unsigned long long m = (unsigned long long)rng_u32(r) * rng_excl;
    unsigned leftover = (unsigned)m;
    if (leftover < rng_excl) {
        unsigned threshold = (0xffffffffu - rng32) % rng_excl;
        while (leftover < threshold) {
            m = (unsigned long long)rng_u32(r) * rng_excl;
            leftover = (unsigned)m;
        }
    }
    return lo + (long long)(m >> 32);
}
// LCG jump-ahead table: after r steps  state_r = A^r * state + G_r * inc  (mod 2^128) with
// G_r = 1 + A + ... + A^(r-1).  Row r = {A^r hi, A^r lo, G_r hi, G_r lo}; filled by cda_create.
// Lets lane a evaluate "its" draw of the sequential numpy stream without waiting for lanes < a.
__device__ unsigned long long cda_pcg_jump[CDA_MAX_AGENTS + 1][4];   // global (lane-indexed reads; see cda_zig_tables.cuh)
extern __shared__ __align__(128) unsigned smw[];
// jump_w: word index of the warp's copy of the table in smw, or -1: read the table in global memory
__device__ __forceinline__ void rng_jump(const CdaRng &g, int r, int jump_w, unsigned long long &shi, unsigned long long &slo) {
    ulonglong2 aa, gg;
    if (jump_w >= 0) { aa = *reinterpret_cast<const ulonglong2 *>(&smw[jump_w + 8 * r]); gg = *reinterpret_cast<const ulonglong2 *>(&smw[jump_w + 8 * r + 4]); }
    else { aa = __ldg(reinterpret_cast<const ulonglong2 *>(&cda_pcg_jump[r][0])); gg = __ldg(reinterpret_cast<const ulonglong2 *>(&cda_pcg_jump[r][2])); }
    const unsigned long long Ah = aa.x, Al = aa.y, Gh = gg.x, Gl = gg.y;
    const unsigned long long l1 = Al * g.slo, h1 = __umul64hi(Al, g.slo) + Ah * g.slo + Al * g.shi;
    const unsigned long long l2 = Gl * g.ilo, h2 = __umul64hi(Gl, g.ilo) + Gh * g.ilo + Gl * g.ihi;
    slo = l1 + l2;
    shi = h1 + h2 + (slo < l1 ? 1ULL : 0ULL);
}

// SeedSequence(seed).generate_state(4, uint64) -> PCG64 seeding (bit_generator.pyx, pcg64.c)
__host__ __device__ inline unsigned ss_hashmix(unsigned value, unsigned &hc) {
    value ^= hc; hc *= 0x931e8875u; value *= hc; value ^= value >> 16; return value;
}
__host__ __device__ inline unsigned ss_mix(unsigned x, unsigned y) {
    unsigned r = 0xca01f9ddu * x - 0x4973f715u * y; r ^= r >> 16; return r;
}
__host__ __device__ inline void seedseq_words(unsigned long long seed, unsigned long long out[4]) {
    unsigned ent[2] = {(unsigned)(seed & 0xffffffffu), (unsigned)(seed >> 32)};
    int n_ent = ent[1] ? 2 : 1;
    unsigned pool[4], hc = 0x43b0d7e5u;
    for (int i = 0; i < 4; ++i) pool[i] = ss_hashmix(i < n_ent ? ent[i] : 0u, hc);
    for (int s = 0; s < 4; ++s)
        for (int d = 0; d < 4; ++d)
            if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], hc));
    unsigned hb = 0x8b51f9ddu, w[8];
    for (int i = 0; i < 8; ++i) {
        unsigned v = pool[i & 3];
        v ^= hb; hb *= 0x58f38dedu; v *= hb; v ^= v >> 16;
        w[i] = v;
    }
    for (int i = 0; i < 4; ++i) out[i] = (unsigned long long)w[2 * i] | ((unsigned long long)w[2 * i + 1] << 32);
}
__device__ __forceinline__ void rng_seed(CdaRng &r, unsigned long long seed) {
    unsigned long long s[4];
    seedseq_words(seed, s);
    // inc = (initseq << 1) | 1 ; state = 0; step; state += initstate; step
    r.ihi = (s[2] << 1) | (s[3] >> 63);
    r.ilo = (s[3] << 1) | 1ULL;
    r.shi = 0; r.slo = 0;
    rng_step(r);
    unsigned long long nlo = r.slo + s[1];
    r.shi = r.shi + s[0] + (nlo < r.slo ? 1ULL : 0ULL);
    r.slo = nlo;
    rng_step(r);
    r.has32 = 0; r.u32 = 0;
}

__device__ __forceinline__ unsigned fresh_tid_x() { unsigned t; asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t)); return t; }   // not CSE'd with earlier reads
// ----------------------------------- async-copy helpers ------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, non-tensor form); bytes % 16 == 0, both addresses 16-B aligned
__device__ __forceinline__ void bulk_g2s(unsigned dst_smem, const void *src_gmem, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, unsigned src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// --------------------------------------- accounts ------------------------------------------
struct CdaAcct {
    long long cash, hold, cost, nav, pos;
    unsigned ntr;
    unsigned ctr;   // per-step counters, packed as stored: trades[0:12) passive[12:24) placed[24] rejected[25] is_pass[26]; bit 27: this agent's
                    // Decimal cash carries a residue ("tracked": every cash movement is journaled, cda_twin.cuh) — copied from the twin flags
};
// account.py:215-231 process_acc with the ledger restated on integers (cost = |pos|*VWAP):
//   open/increase: cost += q*p (account.py:124-133, :173-176); decrease: cost -= q*p (:151-157);
//   cover: cash += position_val - mkt_val, which is 0 for a long and 2*cost - 2*mkt for a short
//   (:135-149 with calculate.py:24-33); flip: cover |pos| then open (q-|pos|) at p (:163-171).
// party: 0 init_party, 1 counter_party; cash moves per cash_processor.py:31-53.
__device__ __forceinline__ void acct_fill(CdaAcct &a, int party, int side /*0 bid,1 ask*/, long long q, long long p) {
    a.ntr++;
    a.ctr += party ? 0x1001u : 0x1u;
    const long long tv = q * p;
    long long inc = 0, dec = 0;      // value moved by size_increase / size_decrease transfers
    const long long ap = a.pos < 0 ? -a.pos : a.pos;
    if (a.pos == 0) { a.cost = tv; inc = tv; }
    else if ((a.pos > 0) == (side == 0)) { a.cost += tv; inc = tv; }
    else if (ap > q) { a.cost -= tv; dec = tv; }
    else {
        const long long mkt = ap * p;
        if (a.pos < 0) a.cash += 2 * a.cost - 2 * mkt;   // size_zero_cash_transfer, short side
        a.cost = 0;
        if (ap == q) dec = tv;
        else { dec = mkt; inc = (q - ap) * p; a.cost = inc; }
    }
    if (party == 0) { a.cash += dec; a.cash -= inc; }
    else { a.cash += 2 * dec; a.hold -= dec; a.hold -= inc; }
    a.pos += side == 0 ? q : -q;
}

// ------------------------------------- market context --------------------------------------
// Warp-uniform book state, held in registers by every lane.  No member is an array that is
// indexed at run time (that would force the struct into local memory): the two sides are
// addressed arithmetically through SOFF(side).
// All shared-memory traffic goes through 32-bit word indices into ONE dynamic array: an LDS/STS then
// takes a 32-bit register + immediate, instead of re-deriving 64-bit generic addresses at every use.
#define SMW(i) smw[(i)]

// Per-warp shared-memory tile, in 32-bit words.
template <int CAP, bool DEC = false>
struct CdaSmemLayout {
    static constexpr int POOL = 0;                                   // u32[2 sides][CAP/32 tiles][5 fields][32]
    static constexpr int SNAP = 2 * CDA_POOL_FIELDS * CAP;           // f32[44] newest snapshot
    static constexpr int TOPK = SNAP + 44;                           // i32[20] frozen pre-step raw top-K prices
    static constexpr int VOL = TOPK + 2 * CDA_K_ROWS;                // u32[20] level volumes being accumulated
    static constexpr int LPX = VOL + 2 * CDA_K_ROWS;                 // u32[20] level prices of the snapshot being built
    static constexpr int ORDER = LPX + 2 * CDA_K_ROWS;               // u32[32] shuffled execution order
    static constexpr int ACT = ORDER + 32;                           // u32[32][3] decoded actions: type|side<<8, size, price
    static constexpr int PARK = ACT + 96;                            // 10 words: parked PCG64 state (+2 pad)
    static constexpr int TIE = PARK + 12;                            // u32[8] decimal_ledger: restart count | answers << 8, then up to 7 parked tie answers
    static constexpr int BAR = TIE + (DEC ? 8 : 0);                  // mbarrier (8-B aligned)
    static constexpr int WORDS = ((BAR + 4 + 3) / 4) * 4;            // (BAR + 2: a second mbarrier, the warp's own action copies) keep 16-B alignment of the next tile
    static constexpr int BYTES = WORDS * 4;
    static_assert(BAR % 2 == 0, "mbarrier must be 8-B aligned");
};

template <int CAP>
struct CdaMkt {
    int pool_w;          // word index of this warp's pool in smw
    int nb, na;          // live orders per side
    unsigned dirty;      // bit (side*8 + tile): pool tile modified this launch -> must be written back
    int bestb, besta;    // cached best bid / best ask: price, -1 = side empty, -2 = unknown (recomputed by a scan on demand)
    unsigned time, next_id, seqctr;
    int status_w;        // word index of the sticky status word in smw (rarely touched: kept out of the registers)
    __device__ __forceinline__ void raise(unsigned bits) const { smw[status_w] |= bits; }   // warp-uniform call: every lane writes the same value
    int tape_nonempty, tape_px;
    int lane;
    int *fills_base; int fill_cap, n_fills, mkt;   // fill log: row = fills_base + mkt * fill_cap * 8 (computed when a fill happens)
    int twf_w;           // decimal_ledger: word index (smw) of THIS lane's twin-flags word, -1 = ledger off / not an agent lane
    __device__ __forceinline__ int side_w(int side) const { return pool_w + side * (CDA_POOL_FIELDS * CAP); }
    __device__ __forceinline__ int count(int side) const { return side ? na : nb; }
    __device__ __forceinline__ void set_count(int side, int v) { if (side) na = v; else nb = v; }
    __device__ __forceinline__ void touch(int side, int idx) { dirty |= 1u << (side * 8 + (idx >> 5)); }
    __device__ __forceinline__ int cached_best(int side) const { return side ? besta : bestb; }
    __device__ __forceinline__ void set_best(int side, int v) { if (side) besta = v; else bestb = v; }
};

// ---- decimal_ledger: journal appends (cda_twin.cuh).  Executed by the lane that owns the account, only when k.twf_w >= 0.
#define CDA_TRACKED_BIT (1u << 27)
template <int CAP> __device__ __forceinline__ void twin_log(const CdaMkt<CAP> &k, const CdaStepParams &p, unsigned long long ev) {
    unsigned twf = SMW(k.twf_w);
    unsigned jn = twf & CDA_TWF_JN_MASK;
    unsigned char *blk = p.state + (size_t)k.mkt * p.cfg.stride;
    unsigned long long *jr = reinterpret_cast<unsigned long long *>(blk + p.cfg.off_jrn) + k.lane * CDA_JRN_E;
    // No call in here (this sits inside the matching loop).  The journal cannot fill up in normal operation: the host replays it every
    // CDA_TWIN_FLUSH_STEPS steps (~0.7 events per agent and step at the BASELINE fill rates, CDA_JRN_E fit); an agent that does
    // overflow it is flagged (CDA_ST_DEC_RANGE: its residues are no longer exact), never silently wrong.
    if (jn >= CDA_JRN_E) { SMW(k.twf_w) = twf | CDA_TWF_RANGE; return; }
    jr[jn] = ev;
    SMW(k.twf_w) = (twf & ~CDA_TWF_JN_MASK) | (jn + 1u);
}
// ---- ties that only the Decimal twin can decide: DETERMINISTIC RE-EXECUTION WITH PARKED ANSWERS -------------------------------------
// A call to the 128-bit arithmetic anywhere inside the step body costs the hot path its register allocation (measured: 0 -> 440
// bytes of spill), so the step body contains none.  Nothing a step does is committed to global memory before its epilogue, so when a
// warp meets a tie it (1) parks a request {what, who, operands, journal length so far}, (2) leaves the step body, (3) at the TOP LEVEL
// of the kernel — no live values — replays the agent's journal up to that point WITHOUT storing it (cda_twin_query) and parks the
// answer in a small per-warp table, and (4) runs the launch's steps for this market again from the state in global memory.  The
// re-execution is deterministic, reaches the same tie, finds the answer and carries on.  Keys: gate tie (step, action slot), NAV tie
// (step, agent).  Cost: one extra pass for that one market, about once per 10^5..10^6 agent-steps on low-cash configurations.
__device__ unsigned long long cda_debug_restarts = 0ULL;   // tie-resolution passes made by all step kernels so far (cda_debug_restart_count)
#define CDA_TIE_SLOTS 7
// words of a warp's account tile in shared memory: the whole 64*A-byte account block with the Decimal twin (its flags word is the 16th
// array), else the 60*A bytes of r1 rounded up to 16 B (32 B less per warp and agent quadruple: what lets 7 CTAs of 8-agent markets fit an SM)
#define CDA_ACCT_TILE_WORDS(DEC, A) ((DEC) ? 16 * (A) : ((15 * (A) + 3) & ~3))
#define CDA_TIE_KEY_GATE(it, q) (0x10000u | ((unsigned)(it) << 8) | (unsigned)(q))
#define CDA_TIE_KEY_NAV(it, a) (0x20000u | ((unsigned)(it) << 8) | (unsigned)(a))
// answer for `key` (2 bits) or -1; tie_w[0] = restarts | n << 8, tie_w[1..] = key << 2 | answer
__device__ __forceinline__ int tie_lookup(int tie_w, unsigned key) {
    const int n = (int)((SMW(tie_w) >> 8) & 0xffu);
    for (int i = 0; i < n && i < CDA_TIE_SLOTS; ++i) { const unsigned w = SMW(tie_w + 1 + i); if ((w >> 2) == key) return (int)(w & 3u); }
    return n >= CDA_TIE_SLOTS ? 2 : -1;   // table full (flagged CDA_ST_DEC_RANGE by the resolver): decide like the integers do, so that the passes terminate
}
// pure query on a COPY of the twin: replays jn journal entries, stores nothing.  mode 1: code of (cash - a0); mode 2: code of nav at price
// a2 with integer cash a0 and cash_on_hold a1.  Codes: 1 greater / positive, 2 equal / zero, 3 less / negative; bit 2 = range error.
__device__ __noinline__ unsigned cda_twin_query(const CdaTwinStored *st, const unsigned long long *jr, int jn, int mode, long long a0, long long a1, long long a2) {
    CdaTwin t;
    cda_twin_load(t, st);
    for (int i = 0; i < jn; ++i) cda_twin_apply(t, jr[i]);
    cda_twin_settle(t);
    int c;
    if (mode == 1) c = t.tracked ? cda_dec_cmp(t.cash, CDA_DI(a0)) : 0;
    else c = cda_twin_nav_sign(t, a0, a1, a2);
    return (c > 0 ? 1u : c == 0 ? 2u : 3u) | (t.err ? 4u : 0u);
}
// a fill on this lane's account: FILL event; a fill that COVERS the position (flat or flip) is where a residue can enter cash
// (account.py:135-149), so from there on this agent's cash is tracked: CASHSYNC carries the (still exact) integer cash
template <int CAP> __device__ __forceinline__ void twin_log_fill(const CdaMkt<CAP> &k, const CdaStepParams &p, CdaAcct &a, int party, int side, unsigned q, int px) {
    const long long ap = a.pos < 0 ? -a.pos : a.pos;
    if (a.pos != 0 && ((a.pos > 0) != (side == 0)) && ap <= (long long)q && !(a.ctr & CDA_TRACKED_BIT)) {
        twin_log(k, p, cda_ev_value(CDA_EV_CASHSYNC, a.cash));
        a.ctr |= CDA_TRACKED_BIT;
        SMW(k.twf_w) |= CDA_TWF_TRACKED;
    }
    twin_log(k, p, cda_ev_fill(party, side, q, (unsigned)px));
}
#define CDA_TWIN_ST(p, mkt, lane) (reinterpret_cast<CdaTwinStored *>((p).state + (size_t)(mkt) * (p).cfg.stride + (p).cfg.off_twin) + (lane))
#define CDA_TWIN_JR(p, mkt, lane) (reinterpret_cast<unsigned long long *>((p).state + (size_t)(mkt) * (p).cfg.stride + (p).cfg.off_jrn) + (lane) * CDA_JRN_E)
// nav > 0 as the reference's Decimal sees it (a zero integer NAV defers to the sign recorded at the last mark-to-market)
template <int CAP> __device__ __forceinline__ bool nav_positive(const CdaMkt<CAP> &k, long long nav) {
    if (nav != 0 || k.twf_w < 0) return nav > 0;
    return ((SMW(k.twf_w) >> CDA_TWF_NAVSIGN_SHIFT) & 3u) == 1u;   // recorded by the mark-to-market that produced the zero
}

template <int CAP> __device__ __forceinline__ int pool_best_scan(const CdaMkt<CAP> &k, int side) {
    int pt = k.side_w(side) + k.lane;
    const int n = k.count(side);
    if (n == 0) return -1;
    unsigned loc = side == 0 ? 0u : 0xffffffffu;
#if CDA_SCAN_UNROLL
    {
        unsigned v[CAP / 32];
#pragma unroll
        for (int t = 0; t < CAP / 32; ++t) v[t] = k.lane + 32 * t < n ? SMW(pt + t * CDA_TILE_WORDS) & CDA_PRICE_MASK : loc;
#pragma unroll
        for (int t = 0; t < CAP / 32; ++t) loc = side == 0 ? max(loc, v[t]) : min(loc, v[t]);
    }
#else
    CDA_SCAN_PRAGMA
    for (int i = k.lane; i < n; i += 32, pt += CDA_TILE_WORDS) {
        const unsigned p = SMW(pt) & CDA_PRICE_MASK;
        loc = side == 0 ? max(loc, p) : min(loc, p);
    }
#endif
    return (int)(side == 0 ? __reduce_max_sync(CDA_FULL, loc) : __reduce_min_sync(CDA_FULL, loc));
}
// best price of a side through the cache (a scan only after an order AT the best price was removed)
template <int CAP> __device__ __forceinline__ int pool_best(CdaMkt<CAP> &k, int side) {
#if CDA_BEST_CACHE
    int b = k.cached_best(side);
    if (b == -2) { b = pool_best_scan(k, side); k.set_best(side, b); }
    return b;
#else
    return pool_best_scan(k, side);
#endif
}
// index of the entry with the smallest key[field] among entries with (pt & mask) == want, or -1
template <int CAP> __device__ __forceinline__ int pool_argmin(const CdaMkt<CAP> &k, int side, unsigned mask, unsigned want, int field) {
    int pt = k.side_w(side) + k.lane;
    const int n = k.count(side);
    unsigned bk = 0xffffffffu; int bi = -1;
#if CDA_SCAN_UNROLL
    {
        unsigned hit = 0u;          // bit t: this lane's order in tile t matches
#pragma unroll
        for (int t = 0; t < CAP / 32; ++t) hit |= (k.lane + 32 * t < n && (SMW(pt + t * CDA_TILE_WORDS) & mask) == want) ? 1u << t : 0u;
#pragma unroll
        for (int t = 0; t < CAP / 32; ++t)
            if ((hit >> t) & 1u) { const unsigned kk = SMW(pt + t * CDA_TILE_WORDS + field * 32); if (kk < bk) { bk = kk; bi = k.lane + 32 * t; } }
    }
#else
    CDA_SCAN_PRAGMA
    for (int i = k.lane; i < n; i += 32, pt += CDA_TILE_WORDS) {
        if ((SMW(pt) & mask) == want) { const unsigned kk = SMW(pt + field * 32); if (kk < bk) { bk = kk; bi = i; } }
    }
#endif
    const unsigned mk = __reduce_min_sync(CDA_FULL, bk);
    if (mk == 0xffffffffu) return -1;
    const unsigned b = __ballot_sync(CDA_FULL, bk == mk);
    return __shfl_sync(CDA_FULL, bi, __ffs(b) - 1);
}
// ordertree.py:70-77 remove_order_by_id: dense pool => move the last entry into the hole
template <int CAP> __device__ __forceinline__ void pool_remove(CdaMkt<CAP> &k, int side, int idx) {
    const int last = k.count(side) - 1;
#if CDA_BEST_CACHE
    {   // best-price cache: an order leaving the best level may empty it -> unknown; an emptied side -> -1
        const int cb = k.cached_best(side);
        if (last == 0) k.set_best(side, -1);
        else if (cb >= 0 && (int)(SMW(k.side_w(side) + CDA_EOFF(idx)) & CDA_PRICE_MASK) == cb) k.set_best(side, -2);
    }
#endif
    __syncwarp();
    if (idx != last && k.lane < CDA_POOL_FIELDS) {
        const int f = k.side_w(side) + k.lane * 32;
        SMW(f + CDA_EOFF(idx)) = SMW(f + CDA_EOFF(last));
    }
    if (idx != last) k.touch(side, idx);
    k.set_count(side, last);
    __syncwarp();
}
// ordertree.py:44-55 insert_order: append with a fresh seq (tail of the level's FIFO and of order_map)
template <int CAP> __device__ __forceinline__ bool pool_append(CdaMkt<CAP> &k, int side, unsigned price, unsigned qty, int trader, unsigned oid, unsigned ts) {
    const int n = k.count(side);
    if (n >= CAP) { k.raise(CDA_ST_POOL_OVERFLOW); return false; }
    const unsigned seq = k.seqctr++;
#if CDA_BEST_CACHE
    {   // best-price cache: a better (or first) price becomes the best; unknown stays unknown
        const int cb = k.cached_best(side);
        if (cb == -1 || (cb >= 0 && (side == 0 ? (int)price > cb : (int)price < cb))) k.set_best(side, (int)price);
    }
#endif
    __syncwarp();
    if (k.lane < CDA_POOL_FIELDS) {
        const unsigned v = k.lane == 0 ? (((unsigned)trader << 24) | price) : k.lane == 1 ? qty : k.lane == 2 ? oid : k.lane == 3 ? ts : seq;
        SMW(k.side_w(side) + CDA_EOFF(n) + k.lane * 32) = v;
    }
    k.touch(side, n);
    k.set_count(side, n + 1);
    __syncwarp();
    return true;
}

// trader.py:49-106 place_order as ONE straight-line flow with a single matching loop (keeps the
// SASS small enough for the instruction cache).  All arguments are warp-uniform.
// type: 0 market, 1 limit, 2 modify, 3 cancel.   `ac` is this lane's account.
// Returns true when the gate hit an exact-equality tie that only the Decimal twin can decide and no answer is parked for it yet
// (decimal_ledger; nothing has been changed): the trader's lane has parked the gated value in req_w, the caller leaves the step body.
// tie_w: the warp's answer table; tie_key: this action's key; req_w: two scratch words for the value.
template <int CAP, bool DEC>
__device__ __forceinline__ bool place_order(CdaMkt<CAP> &k, const CdaStepParams &p, CdaAcct &ac, int t, int type, int side, long long size, int price,
                                            int tie_w, unsigned tie_key, int req_w) {
    const int opp = side ^ 1;
    const bool is_t = k.lane == t;
    // ---- trader.py:108-151 _order_approved (on the trader's lane; market orders need the best opposite quote)
    int best_opp = -1;
    if (type == 0) best_opp = pool_best(k, opp);
    int ok_l = 0;
    if (is_t && (DEC ? nav_positive(k, ac.nav) : ac.nav > 0)) {
        long long opening;
        if ((side == 0 && ac.pos >= 0) || (side == 1 && ac.pos <= 0)) opening = size;
        else { const long long ap = ac.pos < 0 ? -ac.pos : ac.pos; opening = size - ap; if (opening < 0) opening = 0; }
        if (opening <= 0) ok_l = 1;
        else {
            const long long est = type == 0 ? (best_opp > 0 ? best_opp : (k.tape_nonempty ? k.tape_px : 1)) : price;
            ok_l = ac.cash >= opening * est;
            if (DEC && (ac.ctr & CDA_TRACKED_BIT) && ac.cash == opening * est) {   // the Decimal residue decides (about 1 agent-step in 10^5 at low cash)
                const int ans = tie_lookup(tie_w, tie_key);
                if (ans >= 0) ok_l = ans != 3;             // cash >= value unless the Decimal cash is smaller
                else { const long long v = opening * est; SMW(req_w) = (unsigned)v; SMW(req_w + 1) = (unsigned)((unsigned long long)v >> 32); ok_l = 2; }
            }
        }
    }
    ok_l = __shfl_sync(CDA_FULL, ok_l, t);
    if (DEC && ok_l == 2) return true;
    if (!ok_l) { if (is_t) ac.ctr |= 1u << 25; return false; }
    if (type <= 1 && is_t) ac.ctr |= 1u << 24;                          // trader.py:75-76
    if (size <= 0 && type <= 1) { k.raise(CDA_ST_BAD_SIZE); return false; }  // reference: sys.exit in process_order

    // ---- trader.py:254-287 _get_order_ID: limit/cancel = first in order_map order at that price (min seq);
    //      modify = oldest timestamp at any price
    int idx = -1;
    if (type != 0) {
        const unsigned mask = type == 2 ? 0xff000000u : 0xffffffffu;
        const unsigned want = type == 2 ? ((unsigned)t << 24) : (((unsigned)t << 24) | (unsigned)price);
        idx = pool_argmin(k, side, mask, want, type == 2 ? 3 : 4);
    }
    if (type >= 2 && idx < 0) return false;                              // nothing to modify / cancel: no book op

    unsigned oid;
    if (idx >= 0) {
        // trader.py:219-235 / :237-252: release the old order's escrow (cash_processor.py:85-97), then touch the book
        const int pl = k.side_w(side) + CDA_EOFF(idx);
        const unsigned op = SMW(pl) & CDA_PRICE_MASK, oq = SMW(pl + 32);
        oid = SMW(pl + 64);
        if (is_t) {
            const long long ov = (long long)op * oq; ac.hold -= ov; ac.cash += ov;
            if (DEC && (ac.ctr & CDA_TRACKED_BIT)) twin_log(k, p, cda_ev_value(CDA_EV_ESCROW, ov));
        }
        k.time++;                                                        // orderbook.py:196-200, :212-215
        if (type == 3) { pool_remove(k, side, idx); return false; }
        if ((unsigned)price == op && (unsigned long long)size <= oq) {   // orderbook.py:245-248 in place
            __syncwarp();
            if (k.lane == 0) { SMW(pl + 32) = (unsigned)size; SMW(pl + 96) = k.time; }
            k.touch(side, idx);
            __syncwarp();
            if (is_t) {
                const long long v = (long long)price * size; ac.cash -= v; ac.hold += v;
                if (DEC && (ac.ctr & CDA_TRACKED_BIT)) twin_log(k, p, cda_ev_value(CDA_EV_ESCROW, -v));
            }
            return false;
        }
        pool_remove(k, side, idx);                                       // orderbook.py:250-266 re-process, same id
    } else {
        k.time++; k.next_id++;                                           // orderbook.py:33-44 process_order
        oid = k.next_id;
    }

    // ---- orderbook.py:61-194: sweep the opposite side by price-time priority; settle each fill
    //      (trader.py:303-328 settles after the sweep; the sweep never reads accounts, so the ledger
    //      sees the same sequence of updates)
    const int limit = type == 0 ? -1 : price;
    unsigned qty = (unsigned)size;
    while (qty > 0) {
        const int P = pool_best(k, opp);
        if (P < 0) break;
        if (limit >= 0 && (side == 0 ? limit < P : limit > P)) break;
        const int h = pool_argmin(k, opp, CDA_PRICE_MASK, (unsigned)P, 4);
        const int po = k.side_w(opp) + CDA_EOFF(h);
        const unsigned hq = SMW(po + 32);
        const int maker = (int)(SMW(po) >> 24);
        const unsigned moid = SMW(po + 64);
        unsigned traded; int left = -1;
        if (qty < hq) {                       // :73-85 partial: resting order shrinks in place, keeps its timestamp
            traded = qty; left = (int)(hq - qty);
            __syncwarp();
            if (k.lane == 0) SMW(po + 32) = hq - qty;
            k.touch(opp, h);
            __syncwarp();
            qty = 0;
        } else {                              // :86-100 resting order consumed
            traded = hq;
            pool_remove(k, opp, h);
            qty -= traded;
        }
        k.tape_nonempty = 1; k.tape_px = P;   // :140 tape.append (trade price = resting price)
        if (k.fills_base) {
            if (k.n_fills < k.fill_cap || p.cfg.fill_tape) {
                if (k.lane < CDA_FILL_WORDS) {
                    int *frow = k.fills_base + (size_t)k.mkt * k.fill_cap * CDA_FILL_WORDS;
                    const int v = k.lane == 0 ? (int)k.time : k.lane == 1 ? P : k.lane == 2 ? (int)traded : k.lane == 3 ? maker
                                : k.lane == 4 ? (int)moid : k.lane == 5 ? left : k.lane == 6 ? t : side;
                    frow[(p.cfg.fill_tape ? k.n_fills % k.fill_cap : k.n_fills) * CDA_FILL_WORDS + k.lane] = v;
                }
            } else k.raise(CDA_ST_FILL_OVERFLOW);
        }
        k.n_fills++;
        if (maker != t) {                     // trader.py:311-322: counter party, then initiator (disjoint lanes)
            if (k.lane == maker || is_t) {
                if (DEC) twin_log_fill(k, p, ac, is_t ? 0 : 1, is_t ? side : opp, traded, P);
                acct_fill(ac, is_t ? 0 : 1, is_t ? side : opp, traded, P);
            }
        } else if (is_t) {                    // cash_processor.py:55-62 self-trade: escrow back to cash
            const long long tv = (long long)traded * P;
            ac.hold -= tv; ac.cash += tv;
            if (DEC && (ac.ctr & CDA_TRACKED_BIT)) twin_log(k, p, cda_ev_value(CDA_EV_ESCROW, tv));
        }
    }
    // ---- residue rests (orderbook.py:174-191) and is escrowed (cash_processor.py:15-29); market remainder dropped
    if (type != 0 && qty > 0 && pool_append(k, side, (unsigned)price, qty, t, oid, k.time)) {
        if (is_t) {
            const long long v = (long long)price * qty; ac.cash -= v; ac.hold += v;
            if (DEC && (ac.ctr & CDA_TRACKED_BIT)) twin_log(k, p, cda_ev_value(CDA_EV_ESCROW, -v));
        }
    }
    return false;
}

// counter-based generator for the fused random-policy rollout (NOT the env stream)
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// ------------------------------------------------------------------------------------------
// Resident step server (the <ROLLOUT = 1, ROUTED = 1> body; cda_serve_step in include/cda_b200.h).
// The end-to-end host path costs a launch, a completion hand-shake and a state load / store per step on top of the step itself.  In this
// mode the kernel is launched ONCE and stays resident: every warp keeps its market's order pool and accounts in shared memory between
// steps (like the fused rollout), and a step is triggered by a 64-bit MESSAGE the host writes to a mapped pinned word:
//     bits [0,24)  step number, as (seq mod (2^24 - 1)) + 1 (never 0: a cleared word matches nothing)
//     bits [24,32) slot of the plane ring that receives this step's outputs; 0xff = STOP (store the state and exit)
//     bits [32,64) signed offset of the step's pinned action block from srv_act_base, in 4-byte words
// CTA 0 is the POLLER: thread 0 reads the host word over PCIe (no other thread of the grid does), and the CTA republishes a new message into
// CDA_SRV_COPIES device words in different 128-B lines, which the worker warps poll in L2.  A worker that sees its next step number
// stages the market's action record (one bulk copy from host memory per warp), runs the step, stores the newest snapshot + result record
// into the plane, fences, and counts itself; the last one rings the completion word the host spins on.
// Nothing can wait forever: the poller gives the SMs back (publishes STOP itself) when no message has arrived for srv_lease_ns after the
// last step completed — the host notices that the kernel has retired and launches it again with the next step — and a worker that has
// not heard from the poller for srv_watchdog_ns (which cannot happen while the poller is resident) raises srv_err and exits.
// ------------------------------------------------------------------------------------------
#ifndef CDA_SRV_FENCE_ACQREL
#define CDA_SRV_FENCE_ACQREL 0   /* 1: the last warp's system-scope fence is fence.acq_rel.sys instead of the sequentially consistent __threadfence_system() */
#endif
#if CDA_SRV_FENCE_ACQREL
#define CDA_FINAL_SYS_FENCE() asm volatile("fence.acq_rel.sys;" ::: "memory")
#define CDA_WARP_GPU_FENCE() asm volatile("fence.acq_rel.gpu;" ::: "memory")
#else
#define CDA_FINAL_SYS_FENCE() __threadfence_system()
#define CDA_WARP_GPU_FENCE() __threadfence()
#endif
#define CDA_SRV_COPIES 128
/* measurement only (cda_debug_serve_timeline, tools/serve_timeline.py): lane 0 of every warp leaves the time of the step's milestones */
#define CDA_SRV_STAMP(i) do { if (p.prof && lane == 0) p.prof[(size_t)m * 16 + (i)] = globaltimer_ns(); } while (0)
#define CDA_SRV_STOP 0xffu
__host__ __device__ __forceinline__ unsigned cda_srv_seq24(unsigned seq) { return seq % 0xffffffu + 1u; }
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    unsigned long long v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) {
    unsigned v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// the poller CTA (all its threads; smw[0..1] is the mailbox between thread 0 and the others)
// (arguments by value: taking the address of the kernel's parameter block would copy it to every thread's stack)
__device__ __noinline__ void cda_serve_poller(const unsigned long long *go_host, unsigned long long *go_dev, const unsigned *done_dev, unsigned next, unsigned long long lease_ns,
                                              unsigned long long *prof_row) {
    bool idle = true;                      // no step in flight
    unsigned long long t_idle = globaltimer_ns();
    for (;;) {
        if (threadIdx.x == 0) {
            unsigned long long v;
            for (;;) {
                v = ld_volatile_u64(go_host);
                if (((unsigned)v & 0xffffffu) == cda_srv_seq24(next)) break;
                const unsigned long long now = globaltimer_ns();
                if (!idle) { if (ld_volatile_u32(done_dev) == next - 1u) { idle = true; t_idle = now; } }
                else if (now - t_idle > lease_ns) { v = (unsigned long long)cda_srv_seq24(next) | ((unsigned long long)CDA_SRV_STOP << 24); break; }
            }
            *reinterpret_cast<volatile unsigned long long *>(smw) = v;
            if (prof_row && (((unsigned)v >> 24) & 0xffu) != CDA_SRV_STOP) prof_row[1] = globaltimer_ns();   // (timeline tool) step message read from host memory
        }
        __syncthreads();
        const unsigned long long v = *reinterpret_cast<volatile unsigned long long *>(smw);
        if (threadIdx.x < CDA_SRV_COPIES) *reinterpret_cast<volatile unsigned long long *>(go_dev + threadIdx.x * 16) = v;
        if ((((unsigned)v >> 24) & 0xffu) == CDA_SRV_STOP) return;
        __syncthreads();
        next++; idle = false;
    }
}
// a worker warp's lane 0 waits for the message of step `seq` (or STOP)
__device__ __forceinline__ unsigned long long cda_serve_wait(const CdaStepParams &p, unsigned seq, int cta) {   // (inlined: p stays in the constant bank)
    const unsigned long long *w = p.srv_go_dev + (cta & (CDA_SRV_COPIES - 1)) * 16;
    const unsigned want = cda_srv_seq24(seq);
    unsigned long long v = ld_volatile_u64(w);
    if (((unsigned)v & 0xffffffu) == want) return v;
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        __nanosleep(100);
        v = ld_volatile_u64(w);
        if (((unsigned)v & 0xffffffu) == want) return v;
        if (globaltimer_ns() - t0 > p.srv_watchdog_ns) { *reinterpret_cast<volatile unsigned *>(p.srv_err) = 1u; return (unsigned long long)want | ((unsigned long long)CDA_SRV_STOP << 24); }
    }
}

// ------------------------------------------------------------------------------------------
// The fused step kernel: load -> decode -> shuffle -> match/settle -> MTM -> snapshot -> reward
// -> store, one warp per market.  WARPS warps per CTA share nothing but the CTA's shared memory
// carve-up, so there is no __syncthreads anywhere.
// ------------------------------------------------------------------------------------------
#ifndef CDA_MIN_CTAS
#define CDA_MIN_CTAS 7   /* 7 CTAs x 4 warps = 28 warps/SM -> 4144 resident markets on 148 SMs (>= 4096 in one wave) */
#endif
#ifdef CDA_PROFILE_PHASES
#define CDA_TICK(i) do { const long long t__ = clock64(); if (lane == 0 && p.prof) p.prof[(size_t)m * 16 + (i)] += (unsigned long long)(t__ - tprev); tprev = t__; } while (0)
#else
#define CDA_TICK(i) do {} while (0)
#endif
/* registers per lane for the old-snapshot prefetch (covers n_hist <= 4): five 128-B chunks when the output row is cut at the 128-B
   boundaries of its destination (routed outputs: PCIe / NVLink write transactions), four when chunk 0 starts at the row start */
#define CDA_HIST_PREFETCH (ROUTED ? 5 : 4)

template <int CAP, int WARPS, bool ROLLOUT, bool ROUTED, bool DEC>
__global__ void __launch_bounds__(WARPS * 32, CDA_MIN_CTAS) cda_step_kernel(const CdaStepParams p) {
    using L = CdaSmemLayout<CAP, DEC>;
    constexpr bool SERVE = ROLLOUT && ROUTED;   // resident step server: CTA 0 polls the host, CTA b > 0 steps markets 4(b-1) .. 4(b-1)+3 whenever a message arrives
    static_assert(!(SERVE && DEC), "the resident step server runs the integer ledger only");
    if (SERVE && blockIdx.x == 0) { cda_serve_poller(p.srv_go_host, p.srv_go_dev, p.srv_done_dev, p.srv_next, p.srv_lease_ns, p.prof ? p.prof + (size_t)p.M * 16 : nullptr); return; }
    {   // ---- CTA prologue (its values die here: nothing defined above `restart` may be live across the resolve block's call) ----
        // action tile of this CTA: five bulk copies (one per field, the CTA's markets are adjacent rows of every [M][A] array) behind one
        // CTA mbarrier.  When the arrays live in pinned host memory (end-to-end path) this turns 20 sector-sized PCIe reads per CTA into
        // 5 requests (ONE for a market-major block) issued at the very start of the kernel.
#if CDA_PREFETCH_TABLES
        if (threadIdx.x < 48) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(threadIdx.x < 16 ? (const void *)cda_zig_ki : threadIdx.x < 32 ? (const void *)cda_zig_wi : (const void *)cda_zig_fi) + (threadIdx.x & 15) * 128));
        else if (threadIdx.x < 56) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(&cda_pcg_jump[0][0]) + (threadIdx.x - 48) * 128));
#endif
        const int A0 = p.cfg.A, actb0 = WARPS * L::WORDS, cbar_w0 = actb0 + 5 * WARPS * A0;
        constexpr bool WARP_ACT0 = CDA_WARP_ACT_TMA && !ROUTED;
        if (!ROLLOUT && p.act_tma && WARP_ACT0) {
            const int w = threadIdx.x >> 5, m0 = blockIdx.x * WARPS + w;
            if ((threadIdx.x & 31) == 0 && m0 < p.M) {
                const unsigned abar = smem_u32(smw) + (unsigned)(w * L::WORDS + L::BAR + 2) * 4u, fb = (unsigned)A0 * 4u;
                mbar_init(abar, 1);
                mbar_expect_tx(abar, 5u * fb);
                const size_t so = (size_t)m0 * p.act_mstride;
                if (p.act_packed) bulk_g2s(smem_u32(smw) + (unsigned)(actb0 + w * 5 * A0) * 4u, p.cat + so, 5u * fb, abar);
                else {
                    const unsigned dst = smem_u32(smw) + (unsigned)(actb0 + w * A0) * 4u, fs = (unsigned)(WARPS * A0) * 4u;
                    bulk_g2s(dst, p.cat + so, fb, abar);
                    bulk_g2s(dst + fs, p.mean + so, fb, abar);
                    bulk_g2s(dst + 2u * fs, p.sigma + so, fb, abar);
                    bulk_g2s(dst + 3u * fs, p.pcode + so, fb, abar);
                    bulk_g2s(dst + 4u * fs, p.poff + so, fb, abar);
                }
            }
        }
        if (!ROLLOUT && p.act_tma && !WARP_ACT0) {
            if (threadIdx.x == 0) {
                const unsigned cbar = smem_u32(smw) + (unsigned)cbar_w0 * 4u;
                const int m0 = blockIdx.x * WARPS, nm = min(WARPS, p.M - m0);
                const unsigned fb = (unsigned)(nm * A0) * 4u;
                mbar_init(cbar, 1);
                mbar_expect_tx(cbar, 5u * fb);
                const unsigned dst = smem_u32(smw) + (unsigned)actb0 * 4u, fs = (unsigned)(WARPS * A0) * 4u;
                const size_t so = (size_t)m0 * p.act_mstride;
                if (p.act_packed) bulk_g2s(dst, p.cat + so, 5u * fb, cbar);   // tile = u32[markets][5][A]
                else {                                                        // tile = u32[5][WARPS][A]
                    bulk_g2s(dst, p.cat + so, fb, cbar);
                    bulk_g2s(dst + fs, p.mean + so, fb, cbar);
                    bulk_g2s(dst + 2u * fs, p.sigma + so, fb, cbar);
                    bulk_g2s(dst + 3u * fs, p.pcode + so, fb, cbar);
                    bulk_g2s(dst + 4u * fs, p.poff + so, fb, cbar);
                }
            }
        }
        if (!ROLLOUT && p.act_tma && !WARP_ACT0) __syncthreads();   // the CTA's mbarrier is initialised before any warp goes on
        if (SERVE) {   // the CTA's action mbarrier (one bulk copy per CTA and step, issued by whichever warp sees the message first) and its claim word
            if (threadIdx.x == 0) { mbar_init(smem_u32(smw) + (unsigned)cbar_w0 * 4u, 1); smw[cbar_w0 + 2] = 0u; }
            __syncthreads();
        }
        if ((int)((SERVE ? blockIdx.x - 1u : blockIdx.x) * WARPS + (threadIdx.x >> 5)) >= p.M) return;
        if (DEC) {
            if ((threadIdx.x & 31) == 0) smw[(threadIdx.x >> 5) * L::WORDS + L::TIE] = 0u;  // decimal_ledger: no restart yet, no parked answers
            __syncwarp();
        }
    }
    // decimal_ledger: a tie only the Decimal twin can decide makes the warp leave the step body (`goto resolve`, nothing committed),
    // answer it at the bottom of the kernel and come back HERE to run this launch's steps for its market again (see tie_lookup)
restart:;
  {
    // (decimal_ledger: fresh reads, not the prologue's copies kept alive across `resolve`)
    const int warp = DEC ? (int)(fresh_tid_x() >> 5) : (int)(threadIdx.x >> 5), lane = DEC ? (int)(fresh_tid_x() & 31u) : (int)(threadIdx.x & 31);
    const int m = (SERVE ? blockIdx.x - 1u : blockIdx.x) * WARPS + warp;
    const CdaDevCfg &cfg = p.cfg;
    const int A = cfg.A;
    // Output routing.  ROUTED = false is the plain device step (dense obs / reward / flag arrays): everything only the host window /
    // ring, packed-record, split-row and fused all-gather paths execute is compiled out (720 of 4600 SASS instructions; the
    // kernel is several times larger than the 32 KB L1.5 instruction cache, and the smaller body measures 2-3 % faster once
    // the grid runs in more than one wave, when resident warps are spread over all phases of the step).
    const int o_rep_n = ROUTED && p.rep_n > 1 ? p.rep_n : 1, o_rec_inline = ROUTED ? p.rec_inline : 0, o_flag_pack = ROUTED ? p.flag_pack : 0;
    const int o_obs_split = ROUTED ? p.obs_split : p.M, o_ring_mirror = ROUTED ? p.ring_mirror : 0;
    float *const o_ring_out = ROUTED ? p.ring_out : nullptr;
    const int o_obs_stride = ROUTED ? p.obs_stride : cfg.W, o_reward_stride = ROUTED ? p.reward_stride : A, o_flag_stride = ROUTED ? p.flag_stride : 1;
    const int actb = WARPS * L::WORDS;                 // word index of the CTA's action tile: u32[5][WARPS][A], then the mbarrier
    const int cbar_w = actb + 5 * WARPS * A;           // (16-B aligned: 20*A words)
    // this warp's copy of the PCG jump-ahead rows 0..A: loaded NOW, stored to shared memory just before the normal draws (by then the
    // load has landed: nothing waits for it).  It lives in the tail of the decoded-action tile (u32[32][3], of which 3A words are
    // used): no extra shared memory, so 7 CTAs per SM still fit; with more than 8 agents it does not fit and the draws read
    // the table in global memory.
    const bool jump_sm = ((3 * A + 3) & ~3) + 8 * (A + 1) <= 96;
    ulonglong2 jrow = make_ulonglong2(0ULL, 0ULL);   // lane j < 2(A+1): 16-byte piece j of the table (row j/2: A^r for even j, G_r for odd j)
    if (jump_sm && lane < 2 * (A + 1)) jrow = __ldg(reinterpret_cast<const ulonglong2 *>(&cda_pcg_jump[0][0]) + lane);
#ifdef CDA_PROFILE_PHASES
    long long tprev = clock64();
    { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); if (lane == 0 && p.prof) { p.prof[(size_t)m * 16 + 12] = gt; unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); p.prof[(size_t)m * 16 + 14] = sm; } }   // warp start (ns), SM id
#endif
    const int wb = warp * L::WORDS;                       // this warp's tile in smw
    const unsigned sa = smem_u32(smw) + (unsigned)wb * 4u;   // shared-window byte address of this warp's tile
    const unsigned bar = sa + L::BAR * 4u;
    unsigned char *blk = p.state + (size_t)m * cfg.stride;
    unsigned *hdr = reinterpret_cast<unsigned *>(blk);

    // ---- loads are issued where their values are first needed (anything loaded far ahead of its use is
    //      spilled by the register allocator, and the spill store then waits for the load)
    // account arrays of this market (pointers are re-derived where needed instead of living in registers)
#define CDA_ACCT_PTRS \
    long long *g_cash = reinterpret_cast<long long *>(blk + cfg.off_acct); \
    long long *g_hold = g_cash + A, *g_cost = g_cash + 2 * A, *g_nav = g_cash + 3 * A, *g_prev = g_cash + 4 * A, *g_max = g_cash + 5 * A; \
    int *g_pos = reinterpret_cast<int *>(g_cash + 6 * A); \
    unsigned *g_ntr = reinterpret_cast<unsigned *>(g_pos + A), *g_ctr = g_ntr + A, *g_twf = g_ctr + A; \
    (void)g_hold; (void)g_cost; (void)g_nav; (void)g_prev; (void)g_max; (void)g_pos; (void)g_ntr; (void)g_ctr; (void)g_twf;
    CdaAcct ac = CdaAcct{0, 0, 0, 0, 0, 0, 0};
    // ---- account tile: the market's whole account block (cash hold cost nav prev_nav max_nav i64[A], pos ntr ctr twf
    //      u32[A]: 64*A contiguous bytes) goes global -> shared with ONE bulk copy issued before anything else; the
    //      lanes pick their fields out of shared memory when do_actions / mark-to-market need them, so no register
    //      holds an account value across the decode / RNG phases and no global-load latency is exposed later.
    //      The warp's mbarrier counts two arrivals: this copy and the order-pool copy issued once the header is here.
    const int acct_w = cbar_w + 4 + warp * CDA_ACCT_TILE_WORDS(DEC, A);   // word index of this warp's account tile (16-B aligned)
    const unsigned acct_b = DEC ? 64u * (unsigned)A : 60u * (unsigned)A;   // bytes staged: the twin-flags array only with the Decimal twin
    if (lane == 0) {
        if (!DEC || (SMW(wb + L::TIE) & 0xffu) == 0u) mbar_init(bar, 2);      // (a restarted pass uses the barrier's next phase)
        constexpr unsigned spec_b = (CDA_SPEC_TILES * 32 > CAP ? CAP / 32 : CDA_SPEC_TILES) * (CDA_TILE_WORDS * 4u);
        if (p.acct_tma) {
            mbar_expect_tx(bar, acct_b + 2u * spec_b);
            bulk_g2s(smem_u32(smw) + (unsigned)acct_w * 4u, blk + cfg.off_acct, acct_b, bar);
            if (spec_b) {
                const unsigned *gp = reinterpret_cast<const unsigned *>(blk + cfg.off_pool);
                bulk_g2s(sa + L::POOL * 4u, gp, spec_b, bar);
                bulk_g2s(sa + (L::POOL + CDA_POOL_FIELDS * CAP) * 4u, gp + CDA_POOL_FIELDS * CAP, spec_b, bar);
            }
        }
    }
    if (!p.acct_tma) {   // (debug switch) same tile, filled by plain loads
        const unsigned *ga = reinterpret_cast<const unsigned *>(blk + cfg.off_acct);
        for (int i = lane; i < (int)(acct_b >> 2); i += 32) SMW(acct_w + i) = ga[i];
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
    }
    const unsigned tk0 = lane < 2 * CDA_K_ROWS ? hdr[20 + lane] : 0u;   // consumed below, after the header loads are in flight

    // ---- header (warp-uniform 128-bit loads: one request each, value in every lane)
    const uint4 h0 = *reinterpret_cast<const uint4 *>(hdr + 0);
    const uint4 h1 = *reinterpret_cast<const uint4 *>(hdr + 4);
    const uint4 h2 = *reinterpret_cast<const uint4 *>(hdr + 8);

    if (lane < 2 * CDA_K_ROWS) SMW(wb + L::TOPK + lane) = tk0;
    __syncwarp();
    CdaMkt<CAP> k;
    k.lane = lane;
    if (h0.x == 0xffffffffu) return;  // (keeps the header loads ahead of the first tick)
    CDA_TICK(0);   // header arrived
    k.pool_w = wb + L::POOL;
    k.time = h0.x; k.next_id = h0.y; k.seqctr = h0.z;
    unsigned t_step = h0.w;
    // last_price (exchg_helper.py:62-63: the latest tape price, or the reset anchor while the tape is empty) and the tape price the
    // matching loop maintains are the same number at every point where either is read: ONE register, k.tape_px
    k.tape_nonempty = (h1.y & CDA_FLAG_TAPE) ? 1 : 0;
    // done_mask and the sticky status are not needed until the very end: parked in shared memory (held in registers they get
    // spilled, and the reload at the end of the kernel misses the small L1: measured 2.8 % of the kernel in one LDL)
    k.status_w = wb + L::PARK + 11;
    if (lane == 0) { SMW(wb + L::PARK + 10) = h1.z; SMW(wb + L::PARK + 11) = h1.w; }
    k.nb = (int)h2.x; k.na = (int)h2.y;
    CdaRng rng;
    rng.has32 = h2.z; rng.u32 = h2.w;
    k.tape_px = (int)h1.x;
    k.fills_base = p.fills; k.mkt = m;
    k.fill_cap = cfg.fill_cap; k.n_fills = (cfg.fill_tape && p.fills) ? (int)hdr[42] : 0; k.dirty = 0;
    k.bestb = -2; k.besta = -2;
    k.twf_w = (DEC && lane < A) ? acct_w + 15 * A + lane : -1;

    // ---- order pool: ONE TMA bulk copy per side of the live tiles (640 B per 32 orders)
    unsigned *gpool = reinterpret_cast<unsigned *>(blk + cfg.off_pool);
    unsigned bytes_b = (((unsigned)k.nb + 31u) >> 5) * (CDA_TILE_WORDS * 4u), bytes_a = (((unsigned)k.na + 31u) >> 5) * (CDA_TILE_WORDS * 4u);
    if (lane == 0) {
        // (tiles already on their way speculatively are skipped; only the plain-load debug path fetches everything here)
        const unsigned skip = p.acct_tma ? (CDA_SPEC_TILES * 32 > CAP ? CAP / 32 : CDA_SPEC_TILES) * (CDA_TILE_WORDS * 4u) : 0u;
        bytes_b = bytes_b > skip ? bytes_b - skip : 0u; bytes_a = bytes_a > skip ? bytes_a - skip : 0u;
        if (bytes_b | bytes_a) mbar_expect_tx(bar, bytes_b + bytes_a); else mbar_arrive(bar);
        if (bytes_b) bulk_g2s(sa + L::POOL * 4u + skip, reinterpret_cast<unsigned char *>(gpool) + skip, bytes_b, bar);
        if (bytes_a) bulk_g2s(sa + (L::POOL + CDA_POOL_FIELDS * CAP) * 4u + skip, reinterpret_cast<unsigned char *>(gpool + CDA_POOL_FIELDS * CAP) + skip, bytes_a, bar);
    }

    float *g_hist = reinterpret_cast<float *>(blk + cfg.off_hist);
    const int W_old = cfg.W - CDA_SNAPSHOT_DIM;       // obs elements that come from older snapshots

    bool waited = false;
    long long nav_max_carry = 0, nav_prev_carry = 0;   // multi-step rollout: carry max_nav / prev_nav between steps
    const int n_iter = ROLLOUT ? p.num_steps : 1;
    int it = 0;
    for (; SERVE || it < n_iter; ++it) {
        const bool last_it = SERVE || !ROLLOUT || it == n_iter - 1;   // this step's outputs leave the SM (resident server: every step's do)
        float hv[CDA_HIST_PREFETCH];
        int slot_new = 0;
        if (SERVE) {   // wait for the host's message for step srv_next + it; STOP (or the poller's idle lease running out) ends the launch
            unsigned long long msg = 0ULL;
            if (lane == 0) msg = cda_serve_wait(p, p.srv_next + (unsigned)it, (int)blockIdx.x - 1);
            msg = __shfl_sync(CDA_FULL, msg, 0);
            const unsigned pslot = ((unsigned)msg >> 24) & 0xffu;
            if (pslot == CDA_SRV_STOP) break;
            CDA_SRV_STAMP(0);   // message seen
            // the step's action records i32[markets][5][A] of this CTA's markets are ONE contiguous run of the caller's block: one bulk copy per
            // CTA (reads from host memory are bound by the number of requests: 4096 copies of 80 B took 17 us, 1024 of 320 B take 7), issued by
            // the first of the CTA's warps to see the message
            const unsigned char *ab = p.srv_act_base + (long long)(int)(msg >> 32) * 4LL;
            if (lane == 0) SMW(wb + L::SNAP + 42) = pslot;                         // (parked: needed again when the outputs are stored)
            if (p.srv_act_mode == 0) {
                if (lane == 0 && atomicMax(&smw[cbar_w + 2], (unsigned)it + 1u) <= (unsigned)it) {
                    const int m0 = m - warp, nm = min(WARPS, p.M - m0);
                    const unsigned cbar = smem_u32(smw) + (unsigned)cbar_w * 4u, nb = (unsigned)(nm * 20 * A);
                    mbar_expect_tx(cbar, nb);
                    bulk_g2s(smem_u32(smw) + (unsigned)actb * 4u, ab + (size_t)m0 * (size_t)(20 * A), nb, cbar);
                }
            } else {
                const unsigned *aw = reinterpret_cast<const unsigned *>(ab + (size_t)m * (size_t)(20 * A));
                for (int i = lane; i < 5 * A; i += 32) SMW(actb + warp * 5 * A + i) = ld_volatile_u32(aw + i);
            }
            __syncwarp();
        }
        // ================= set_actions: action_helper.py:145-172, :241-397 =================
        int a_cat = -1, a_pcode = 0, a_poff = 1; float a_mean = 0.f, a_sigma = 0.f;
        if (lane < A) {
            if (ROLLOUT && !SERVE) {   // fused uniform random policy (model_handler.py:38-78)
                const unsigned long long h = splitmix64(p.policy_seed ^ splitmix64(((unsigned long long)m << 32) ^ ((unsigned long long)(t_step) * 64ULL + lane)));
                a_cat = (int)(((h & 0xffffu) * 9u) >> 16);
                a_pcode = (int)((((h >> 16) & 0xffffu) * 10u) >> 16);
                a_poff = (int)((((h >> 32) & 0xffffu) * 3u) >> 16);
                const unsigned long long h2 = splitmix64(h);
                a_mean = (float)((double)(h2 & 0xffffffu) * (2.0 / 16777216.0) - 1.0);
                a_sigma = (float)((double)((h2 >> 24) & 0xffffffu) * (1.0 / 16777216.0));
            } else if (!SERVE && !p.act_tma) {
                const size_t o = (size_t)m * p.act_mstride + lane;
                a_cat = p.cat[o]; a_mean = p.mean[o]; a_sigma = p.sigma[o]; a_pcode = p.pcode[o]; a_poff = p.poff[o];
            }
        }
        if (SERVE) {   // this market's action record i32[5][A], staged above
            if (p.srv_act_mode == 0) mbar_wait(smem_u32(smw) + (unsigned)cbar_w * 4u, (unsigned)it & 1u);
            CDA_SRV_STAMP(1);   // actions here
            if (lane < A) {
                const int o = actb + warp * 5 * A + lane;
                a_cat = (int)SMW(o); a_mean = __uint_as_float(SMW(o + A)); a_sigma = __uint_as_float(SMW(o + 2 * A));
                a_pcode = (int)SMW(o + 3 * A); a_poff = (int)SMW(o + 4 * A);
            }
        }
        if (!ROLLOUT && p.act_tma) {
            mbar_wait((CDA_WARP_ACT_TMA && !ROUTED) ? smem_u32(smw) + (unsigned)(wb + L::BAR + 2) * 4u : smem_u32(smw) + (unsigned)cbar_w * 4u, 0);
            if (lane < A) {
                const int o = actb + (p.act_packed ? warp * 5 * A : warp * A) + lane, fs = p.act_packed ? A : WARPS * A;
                a_cat = (int)SMW(o); a_mean = __uint_as_float(SMW(o + fs)); a_sigma = __uint_as_float(SMW(o + 2 * fs));
                a_pcode = (int)SMW(o + 3 * fs); a_poff = (int)SMW(o + 4 * fs);
            }
        }
        if (!ROLLOUT || it == 0) {   // the generator is needed from here on (not earlier)
            const ulonglong2 r0 = *reinterpret_cast<const ulonglong2 *>(hdr + 12);
            const ulonglong2 r1 = *reinterpret_cast<const ulonglong2 *>(hdr + 16);
            rng.shi = r0.x; rng.slo = r0.y; rng.ihi = r1.x; rng.ilo = r1.y;
        }
        if (!cfg.fill_tape) k.n_fills = 0;
        ac.ctr = 0;
        const bool bad = lane < A && (a_cat > 8 || (a_cat > 0 && ((a_cat - 1) & 3) != 0 && (a_pcode < 0 || a_pcode >= CDA_K_ROWS || a_poff < 0 || a_poff > 2)));
        if (__any_sync(CDA_FULL, bad)) k.raise(CDA_ST_BAD_ACTION);
        if (a_cat > 8) a_cat = 0;
        if (a_pcode < 0 || a_pcode >= CDA_K_ROWS) a_pcode = 0;
        if (a_poff < 0 || a_poff > 2) a_poff = 1;
        const unsigned present = __ballot_sync(CDA_FULL, lane < A && a_cat >= 0);
        SMW(wb + L::ORDER + lane) = (unsigned)a_cat;   // parked across the draws (the register-pressure peak); the order tile is free until the shuffle
        CDA_TICK(10);  // actions arrived
        if (!ROLLOUT || it == 0) {
            if (jump_sm && lane < 2 * (A + 1)) *reinterpret_cast<ulonglong2 *>(&smw[wb + L::ACT + ((3 * A + 3) & ~3) + 4 * lane]) = jrow;
            __syncwarp();
        }
        // one standard-normal draw per PRESENT agent, in agent order, pass agents included (:311-339).
        // Lane a jumps the LCG ahead by (its rank + 1) steps and evaluates its own draw; this is the sequential
        // stream as long as every draw returns from the first ziggurat test (98.8 % each).
        // A lane's finished draw is parked in the (not yet used) decoded-action tile instead of a register pair: the slow path is a
        // real call, and a double held across it is spilled to local memory and reloaded through the small L1.
        const int zw = wb + L::ACT + 2 * lane;
#pragma unroll 1
        for (unsigned todo = present; todo;) {   // agents whose draw is still to be made, in agent order
            const bool mine = (todo >> lane) & 1u;
            const int rnk = __popc(todo & ((1u << lane) - 1u));
            unsigned long long jh, jl;
            rng_jump(rng, mine ? rnk + 1 : 0, jump_sm ? wb + L::ACT + ((3 * A + 3) & ~3) : -1, jh, jl);
            const unsigned long long xr = jh ^ jl;
            const unsigned rot = (unsigned)(jh >> 58);
            unsigned long long r = (xr >> rot) | (xr << ((64u - rot) & 63u));
            const int idx = (int)(r & 0xff);
            r >>= 8;
            const int sign = (int)(r & 1ULL);
            const unsigned long long rabs = (r >> 1) & 0x000fffffffffffffULL;
            double zx = (double)rabs * __ldg(&cda_zig_wi[idx]);
            if (sign) zx = -zx;
            const unsigned fail = __ballot_sync(CDA_FULL, mine && !(rabs < __ldg(&cda_zig_ki[idx])));
            // the draws AHEAD of the first wedge/tail draw are exactly the sequential stream: keep them, move the
            // generator past them, make that one draw the slow way (it consumes extra numbers), and go round again
            // for the agents behind it (one round in 95 % of the steps)
            const unsigned acc = fail ? (todo & ((1u << (__ffs(fail) - 1)) - 1u)) : todo;
            if ((acc >> lane) & 1u) { const unsigned long long zb = (unsigned long long)__double_as_longlong(zx); SMW(zw) = (unsigned)zb; SMW(zw + 1) = (unsigned)(zb >> 32); }
            if (acc) {
                const int src = 31 - __clz(acc);               // last accepted lane holds the state after its draw
                rng.shi = __shfl_sync(CDA_FULL, jh, src);
                rng.slo = __shfl_sync(CDA_FULL, jl, src);
            }
            todo &= ~acc;
            if (fail) {
                const int a = __ffs(fail) - 1;
                const double za = rng_normal(rng);
                if (lane == a) { const unsigned long long zb = (unsigned long long)__double_as_longlong(za); SMW(zw) = (unsigned)zb; SMW(zw + 1) = (unsigned)(zb >> 32); }
                todo &= ~(1u << a);
            }
        }
        double z = 0.0;
        if ((present >> lane) & 1u) z = __longlong_as_double((long long)(((unsigned long long)SMW(zw + 1) << 32) | SMW(zw)));
        __syncwarp();   // every lane has its draw back before the tile receives the decoded actions
        CDA_TICK(11);  // draws done
        a_cat = (int)SMW(wb + L::ORDER + lane);
        const int a_side = a_cat <= 0 ? -1 : (a_cat <= 4 ? 0 : 1);
        const int a_type = a_cat <= 0 ? 0 : ((a_cat - 1) & 3);
        long long a_size = 0; int a_price = -1;
        if (lane < A && a_cat >= 0) {
            const float loc = __fmul_rn(a_type == 0 ? cfg.mkt_mul : cfg.lim_mul, a_mean);  // f32 product (NEP 50)
            const double x = (double)loc + (double)a_sigma * z;                              // numpy: loc + scale*z
            a_size = __double2ll_rn(fabs(x)) + cfg.min_size;                                 // rint half-even, :339, :276
            if (a_cat == 0) ac.ctr |= 1u << 26;
            if (a_side >= 0 && a_type != 0) {            // _set_price :341-397 on the frozen pre-step top-K
                const int raw = (int)SMW(wb + L::TOPK + a_side * CDA_K_ROWS + a_pcode);
                const int off = a_poff - 1;
                int base, pr;
                if (a_side == 0) { base = raw == 0 ? k.tape_px - (a_pcode + 1) * cfg.tick : raw; pr = base + off * cfg.tick; }
                else             { base = raw == 0 ? k.tape_px + (a_pcode + 1) * cfg.tick : raw; pr = base - off * cfg.tick; }
                if (pr < cfg.tick) pr = cfg.tick;
                a_price = pr;
            }
        }
        if (__any_sync(CDA_FULL, a_price >= (int)CDA_PRICE_MASK)) { k.raise(CDA_ST_PRICE_RANGE); if (a_price >= (int)CDA_PRICE_MASK) a_price = CDA_PRICE_MASK - 1; }

        CDA_TICK(1);   // accounts + actions arrived, draws + decode done
        // park the decoded actions in shared memory: the matching phase reads them with uniform loads, and the
        // dozen registers they occupied are free while the book is being worked on
        if ((!ROLLOUT || SERVE) && p.act_log && lane < A)
            *reinterpret_cast<int4 *>(p.act_log + ((size_t)m * A + lane) * 4) = a_side >= 0 ? make_int4(a_type, a_side, (int)a_size, a_price) : make_int4(-1, -1, 0, -1);
        if (lane < A && a_side >= 0) {
            SMW(wb + L::ACT + 3 * lane) = (unsigned)a_type | ((unsigned)a_side << 8);
            SMW(wb + L::ACT + 3 * lane + 1) = (unsigned)a_size;
            SMW(wb + L::ACT + 3 * lane + 2) = (unsigned)a_price;
        }
        // ================= rand_exec_seq: action_helper.py:174-199 ==========================
        const unsigned active = __ballot_sync(CDA_FULL, lane < A && a_side >= 0);
        const int n_act = __popc(active);
        // the execution order lives in a register: lane q holds the q-th action's agent (scatter by rank, read back by lane)
        if ((active >> lane) & 1u) SMW(wb + L::ORDER + __popc(active & ((1u << lane) - 1u))) = (unsigned)lane;
        __syncwarp();
        int ord = (int)SMW(wb + L::ORDER + lane);
        for (int i = n_act - 1; i >= 1; --i) {           // Generator.permutation: Fisher-Yates from the top
            const int j = (int)rng_interval(rng, (unsigned)i);
            const int vi = __shfl_sync(CDA_FULL, ord, i), vj = __shfl_sync(CDA_FULL, ord, j);
            ord = lane == i ? vj : (lane == j ? vi : ord);
        }
        if (lane == 0) {   // park the generator: it is not needed again until the next step / the final store
            unsigned long long *pk = reinterpret_cast<unsigned long long *>(&smw[wb + L::PARK]);
            pk[0] = rng.shi; pk[1] = rng.slo; pk[2] = rng.ihi; pk[3] = rng.ilo;
            SMW(wb + L::PARK + 8) = rng.has32; SMW(wb + L::PARK + 9) = rng.u32;
        }
        __syncwarp();

        CDA_TICK(2);   // shuffle done
        if (!waited) { mbar_wait(bar, DEC ? (SMW(wb + L::TIE) & 1u) : 0u); waited = true; }
        {   // this lane's account, out of the account tile.  In a multi-step rollout the tile is also where the
                                             // accounts live BETWEEN steps (written back at the end of every step): nothing account-related is
                                             // carried in registers across the decode / RNG phases of the next step.  Unconditional loads (lanes
                                             // >= A read agent 0's words and never use them) so that the old values are dead at the loop head.
            const int al = lane < A ? lane : 0;
            const long long *sq = reinterpret_cast<const long long *>(&smw[acct_w]);
            ac.cash = sq[al]; ac.hold = sq[A + al]; ac.cost = sq[2 * A + al]; ac.nav = sq[3 * A + al];
            ac.pos = (int)SMW(acct_w + 12 * A + al); ac.ntr = SMW(acct_w + 13 * A + al);
            if (DEC && k.twf_w >= 0 && (SMW(k.twf_w) & CDA_TWF_TRACKED)) ac.ctr |= CDA_TRACKED_BIT;
        }
        CDA_TICK(3);   // pool + account tiles landed

        // ================= do_actions: action_helper.py:201-239 =============================
        {   // best-price cache: seeded from level 0 of the frozen pre-step top-K (0 = that side was empty), here rather than at kernel
            // entry / across rollout steps so that it occupies registers only while the book is being worked on
            const unsigned b0 = SMW(wb + L::TOPK), a0 = SMW(wb + L::TOPK + CDA_K_ROWS);
            k.bestb = k.nb ? (b0 ? (int)b0 : -2) : -1; k.besta = k.na ? (a0 ? (int)a0 : -2) : -1;
        }
        for (int q = 0; q < n_act; ++q) {
            const int t = __shfl_sync(CDA_FULL, ord, q);
            const unsigned ts_ = SMW(wb + L::ACT + 3 * t);
            const long long size = (long long)SMW(wb + L::ACT + 3 * t + 1);
            const int price = (int)SMW(wb + L::ACT + 3 * t + 2);
            if (place_order<CAP, DEC>(k, p, ac, t, (int)(ts_ & 0xffu), (int)(ts_ >> 8), size, price, wb + L::TIE, CDA_TIE_KEY_GATE(it, q), wb + L::SNAP) && DEC) {
                // gate tie without an answer (decimal_ledger, rare): request = {mode 1, value, -, -, key} in this lane's slot of the pool tile
                // (the pool is reloaded by the next pass), then leave
                const int rq = wb + L::POOL + 8 * lane;
                SMW(rq) = lane == t ? 1u : 0u;
                if (lane == t) { SMW(rq + 1) = SMW(wb + L::SNAP); SMW(rq + 2) = SMW(wb + L::SNAP + 1); SMW(rq + 6) = CDA_TIE_KEY_GATE(it, q); }
                goto resolve;
            }
        }

        CDA_TICK(4);   // do_actions done
        // mark-to-market needs max_nav / prev_nav from the state block: issue those loads now, do the top-K sweep
        // (which does not depend on the accounts), then mark to market
        const int last_price = k.tape_px;              // exchg_helper.py:62-63 (the snapshot's midpoint fallback reads it)
        // ================= mark_to_mkt: exchg_helper.py:56-66, calculate.py:35-55 ===========
        // (the new NAV right here — few values are live, which matters for the rare tie below; prev_nav / max_nav follow after the sweep)
        if (k.tape_nonempty) {
            if (lane < A) {
                const long long ap = ac.pos < 0 ? -ac.pos : ac.pos;
                const long long pv = ac.pos >= 0 ? ap * last_price : 2 * ac.cost - ap * last_price;
                ac.nav = ac.cash + ac.hold + pv;
            }
            // decimal_ledger: an integer NAV of exactly 0 — the Decimal NAV's sign decides bankruptcy and the next gate (done_helper.py,
            // trader.py:112).  Answer parked by an earlier pass -> record it in the twin flags; none yet -> park the request and leave.
            if (DEC && __any_sync(CDA_FULL, lane < A && ac.nav == 0)) {
                int code = -2;                                   // -2: no tie on this lane
                if (lane < A && ac.nav == 0) code = tie_lookup(wb + L::TIE, CDA_TIE_KEY_NAV(it, lane));
                if (__any_sync(CDA_FULL, code == -1)) {
                    const int rq = wb + L::POOL + 8 * lane;
                    SMW(rq) = code == -1 ? 2u : 0u;
                    if (code == -1) {
                        SMW(rq + 1) = (unsigned)ac.cash; SMW(rq + 2) = (unsigned)((unsigned long long)ac.cash >> 32);
                        SMW(rq + 3) = (unsigned)ac.hold; SMW(rq + 4) = (unsigned)((unsigned long long)ac.hold >> 32);
                        SMW(rq + 5) = (unsigned)last_price; SMW(rq + 6) = CDA_TIE_KEY_NAV(it, lane);
                    }
                    goto resolve;
                }
                if (code >= 0) SMW(k.twf_w) = (SMW(k.twf_w) & ~(3u << CDA_TWF_NAVSIGN_SHIFT)) | ((unsigned)code << CDA_TWF_NAVSIGN_SHIFT);
            }
        }

        // ================= set_agg_LOB: state_helper.py:113-214 =============================
        // Top-K levels per side in ONE sweep: the distinct prices within 64 ticks of the best form
        // a 64-bit occupancy mask (redux.or); a level's rank is the popcount below its bit; level
        // volumes accumulate with shared-memory atomics.  Levels beyond the window (rare) are
        // finished by the generic next-best search.
        int myP = 0; unsigned myV = 0;     // lane l<10: bid level l; 10<=l<20: ask level l-10
        __syncwarp();
        if (lane < 2 * CDA_K_ROWS) SMW(wb + L::VOL + lane) = 0;
        __syncwarp();
        // Both sides go through ONE loop body per pass: the bid chain and the ask chain are independent, so their
        // shared-memory loads, the four redux.or and the rank arithmetic overlap instead of running back to back
        // (the kernel is bound by dependent-issue latency at 7 warps per scheduler, not by issue slots).
        unsigned far_sides = 0;            // bit s: side s has levels beyond the 64-tick window AND fewer than K inside it
        unsigned bestB = 0, bestA = 0; int nlevB = 0, nlevA = 0;
        {
            const int ptB = k.side_w(0) + lane, ptA = k.side_w(1) + lane;
            const int ntB = (k.nb - lane + 31) >> 5, ntA = (k.na - lane + 31) >> 5;   // tiles in which this lane owns a live order (0: none)
            const int ntm = (max(k.nb, k.na) + 31) >> 5;
            bestB = k.nb ? (unsigned)pool_best(k, 0) : 0u;      // usually cached by the matching phase
            bestA = k.na ? (unsigned)pool_best(k, 1) : 0u;
            unsigned bl = 0, bh = 0, al = 0, ah = 0; bool fB = false, fA = false;
#if CDA_TOPK_UNROLL
            {   // all tiles' prices in flight at once (predicated), masks afterwards
                unsigned vb[CAP / 32], va[CAP / 32];
#pragma unroll
                for (int t = 0; t < CAP / 32; ++t) {
                    vb[t] = t < ntB ? bestB - (SMW(ptB + t * CDA_TILE_WORDS) & CDA_PRICE_MASK) : 0xffffffffu;
                    va[t] = t < ntA ? (SMW(ptA + t * CDA_TILE_WORDS) & CDA_PRICE_MASK) - bestA : 0xffffffffu;
                }
#pragma unroll
                for (int t = 0; t < CAP / 32; ++t) {
                    const unsigned db = vb[t], da = va[t];
                    bl |= db < 32u ? 1u << db : 0u; bh |= (db - 32u) < 32u ? 1u << (db - 32u) : 0u; fB |= (db - 64u) < 0xffffffbfu;
                    al |= da < 32u ? 1u << da : 0u; ah |= (da - 32u) < 32u ? 1u << (da - 32u) : 0u; fA |= (da - 64u) < 0xffffffbfu;
                }
            }
#else
            CDA_SCAN_PRAGMA
            for (int it = 0; it < ntm; ++it) {
                const unsigned db = it < ntB ? bestB - (SMW(ptB + it * CDA_TILE_WORDS) & CDA_PRICE_MASK) : 0xffffffffu;
                const unsigned da = it < ntA ? (SMW(ptA + it * CDA_TILE_WORDS) & CDA_PRICE_MASK) - bestA : 0xffffffffu;
                bl |= db < 32u ? 1u << db : 0u; bh |= (db - 32u) < 32u ? 1u << (db - 32u) : 0u; fB |= (db - 64u) < 0xffffffbfu;   // 64 <= db < 2^32-1
                al |= da < 32u ? 1u << da : 0u; ah |= (da - 32u) < 32u ? 1u << (da - 32u) : 0u; fA |= (da - 64u) < 0xffffffbfu;
            }
#endif
            bl = __reduce_or_sync(CDA_FULL, bl); bh = __reduce_or_sync(CDA_FULL, bh);
            al = __reduce_or_sync(CDA_FULL, al); ah = __reduce_or_sync(CDA_FULL, ah);
            const unsigned farb = __ballot_sync(CDA_FULL, fB), fara = __ballot_sync(CDA_FULL, fA);
            const int nbl = __popc(bl), nal = __popc(al);
            nlevB = nbl + __popc(bh); nlevA = nal + __popc(ah);
            CDA_SCAN_PRAGMA
            for (int it = 0; it < ntm; ++it) {
                if (it < ntB) {
                    const unsigned pp = SMW(ptB + it * CDA_TILE_WORDS) & CDA_PRICE_MASK, d = bestB - pp;
                    if (d < 64u) {
                        const int rank = d < 32u ? __popc(bl & ((1u << d) - 1u)) : nbl + __popc(bh & ((1u << (d - 32u)) - 1u));
                        if (rank < CDA_K_ROWS) {      // every order of a level writes the same price: benign same-value race
                            atomicAdd(&smw[wb + L::VOL + rank], SMW(ptB + it * CDA_TILE_WORDS + 32));
                            SMW(wb + L::LPX + rank) = pp;
                        }
                    }
                }
                if (it < ntA) {
                    const unsigned pp = SMW(ptA + it * CDA_TILE_WORDS) & CDA_PRICE_MASK, d = pp - bestA;
                    if (d < 64u) {
                        const int rank = d < 32u ? __popc(al & ((1u << d) - 1u)) : nal + __popc(ah & ((1u << (d - 32u)) - 1u));
                        if (rank < CDA_K_ROWS) {
                            atomicAdd(&smw[wb + L::VOL + CDA_K_ROWS + rank], SMW(ptA + it * CDA_TILE_WORDS + 32));
                            SMW(wb + L::LPX + CDA_K_ROWS + rank) = pp;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane < CDA_K_ROWS ? lane < nlevB : (lane < 2 * CDA_K_ROWS && lane - CDA_K_ROWS < nlevA)) { myP = (int)SMW(wb + L::LPX + lane); myV = SMW(wb + L::VOL + lane); }
            far_sides = (farb && nlevB < CDA_K_ROWS ? 1u : 0u) | (fara && nlevA < CDA_K_ROWS ? 2u : 0u);
        }
#pragma unroll 1
        for (int side = 0; side < 2 && far_sides; ++side) {        // levels further than 64 ticks from the best (rare): generic next-best search
            if (!((far_sides >> side) & 1u)) continue;
            const int pt = k.side_w(side) + lane;
            const int nt = (k.count(side) - lane + 31) >> 5;
            const unsigned B = side == 0 ? bestB : bestA;
            unsigned prev = side == 0 ? B - 63u : B + 63u;
            for (int lv = side == 0 ? nlevB : nlevA; lv < CDA_K_ROWS; ++lv) {
                unsigned l2 = side == 0 ? 0u : 0xffffffffu;
                _Pragma("unroll 1")
                for (int it = 0; it < nt; ++it) {
                    const unsigned pp = SMW(pt + it * CDA_TILE_WORDS) & CDA_PRICE_MASK;
                    if (side == 0 ? pp < prev : pp > prev) l2 = side == 0 ? max(l2, pp) : min(l2, pp);
                }
                const unsigned P = side == 0 ? __reduce_max_sync(CDA_FULL, l2) : __reduce_min_sync(CDA_FULL, l2);
                if (P == (side == 0 ? 0u : 0xffffffffu)) break;
                unsigned sv = 0;
                _Pragma("unroll 1")
                for (int it = 0; it < nt; ++it) if ((SMW(pt + it * CDA_TILE_WORDS) & CDA_PRICE_MASK) == P) sv += SMW(pt + it * CDA_TILE_WORDS + 32);
                const unsigned V = __reduce_add_sync(CDA_FULL, sv);
                if (lane == side * CDA_K_ROWS + lv) { myP = (int)P; myV = V; }
                prev = P;
            }
        }
        // ---- fetch the older snapshots of the stacked observation (state_helper.py:88-90); latency hides behind
        //      mark-to-market and the observation math.  Lane mapping: the output row of market m starts 32-B
        //      aligned (672-B rows), so element e is handled by lane (e + mis) & 31 of chunk (e + mis) >> 5, where
        //      mis = floats between the previous 128-B boundary and the row start: every warp store then covers
        //      ONE aligned 128-B line (matters for DRAM sectors and doubles the PCIe/NVLink write efficiency when
        //      the row lives in pinned host or peer memory).
        slot_new = (int)(t_step % (unsigned)cfg.n_hist);
        // the warp's tile index again, re-derived from %tid: the copy made at kernel entry would otherwise be spilled for the whole
        // matching phase and reloaded here through a cold L1
        const int wbL = (int)(fresh_tid_x() >> 5) * L::WORDS;
        float *orow = nullptr; int mis = 0;
        if (p.obs && last_it) {
            orow = (m < o_obs_split ? p.obs : p.obs_hi) + (size_t)m * o_obs_stride;
            if (ROUTED) mis = (int)((reinterpret_cast<size_t>(orow) >> 2) & 31);
            // the ring holds exactly n_hist snapshots, so the stacked old part (oldest first) is ONE circular run of the ring
            // starting at the slot after the newest: element e lives at ring position (first + e) mod W — no division by 42
            const int first = (slot_new + 1) * CDA_SNAPSHOT_DIM;
#pragma unroll
            for (int q = 0; q < CDA_HIST_PREFETCH; ++q) {
                const int e = lane + 32 * q - mis;
                hv[q] = 0.f;
                if (e >= 0 && e < W_old) { int ri = first + e; if (ri >= cfg.W) ri -= cfg.W; hv[q] = g_hist[ri]; }
            }
        }
        // ---- mark_to_mkt, second half (calculate.py:49-51): prev_nav = the NAV before this step's mark-to-market, max_nav = high-water mark.
        //      The old values still sit in the account tile (the new NAV was computed right after do_actions, see above).
        long long nav_prev = 0, nav_max = 0;
        if (lane < A) {
            const long long *sq = reinterpret_cast<const long long *>(&smw[acct_w]);
            nav_max = sq[5 * A + lane];
            nav_prev = k.tape_nonempty ? sq[3 * A + lane] : sq[4 * A + lane];   // never marked yet: keep the stored value
            if (k.tape_nonempty && ac.nav > nav_max) nav_max = ac.nav;
        }
        nav_max_carry = nav_max; nav_prev_carry = nav_prev;
        CDA_TICK(5);   // top-K levels + mtm done
        const int best_bid = __shfl_sync(CDA_FULL, myP, 0), best_ask = __shfl_sync(CDA_FULL, myP, CDA_K_ROWS);
        double Mid;
        if (best_bid > 0 && best_ask > 0) Mid = ((double)best_bid + (double)best_ask) / 2.0;
        else if (best_bid > 0) Mid = (double)best_bid;
        else if (best_ask > 0) Mid = (double)best_ask;
        else { Mid = (double)last_price; if (Mid <= 0) Mid = 100.0; }
        // The reference evaluates these in f64 and casts once to f32 (state_helper.py:180-212).  For the
        // price/size rows the operands are small exact integers/half-integers, so a correctly rounded
        // f32 divide / sqrt returns the same bits as f64-then-round (double rounding is innocuous for
        // / and sqrt when 53 >= 2*24+2).  log_mid and log1p_spread go through ONE f64 log() call:
        // lane 20 takes log(M), lane 21 log(1 + spread_ticks) (1 + st is exact).
        if (lane < 2 * CDA_K_ROWS) {
            const float Mf = (float)Mid, pz = (float)myP;       // exact: prices < 2^24, Mid is a half-integer
            float pn = 0.f, sn = 0.f;
            if (myP > 0) pn = lane < CDA_K_ROWS ? __fdiv_rn(Mf - pz, Mf) : -__fdiv_rn(pz - Mf, Mf);
            if (myV > 0) { const float sq = __fsqrt_rn((float)myV); sn = lane < CDA_K_ROWS ? sq : -sq; }
            const int l = lane < CDA_K_ROWS ? lane : lane - CDA_K_ROWS;
            const int b = lane < CDA_K_ROWS ? 0 : 2 * CDA_K_ROWS;
            SMW(wbL + L::SNAP + b + l) = __float_as_uint(pn);
            SMW(wbL + L::SNAP + b + CDA_K_ROWS + l) = __float_as_uint(sn);
            SMW(wbL + L::TOPK + lane) = (unsigned)myP;                          // frozen raw top-K for the next step's _set_price
            if (!ROLLOUT || last_it) hdr[20 + (fresh_tid_x() & 31u)] = (unsigned)myP;   // (fresh lane index: reusing the entry-time &hdr[20 + lane] would keep that pointer spilled across the whole step)
        } else if (lane < 22) {
            double x = Mid; bool live = true;
            if (lane == 21) {
                live = best_bid > 0 && best_ask > 0;
                const double st = ((double)best_ask - (double)best_bid) / (double)cfg.tick;
                x = 1.0 + (st > 0.0 ? st : 0.0);
            }
            const double lg = log(x);
            SMW(wbL + L::SNAP + 20 + lane) = __float_as_uint(live ? (float)lg : 0.0f);
        }
        __syncwarp();

        CDA_TICK(6);   // obs math done
        // ================= prep_next_state: state_helper.py:80-92 (ring + stacked obs) ======
        if (p.obs && last_it) {
            // one destination normally; with the fused all-gather, row (row0 + m) of EVERY peer's buffer
            // (plain stores to peer-mapped addresses: they travel over NVLink while other warps still match)
            const int nd = o_rep_n;
#pragma unroll 1
            for (int g = 0; g < nd; ++g) {
                float *o = ROUTED ? reinterpret_cast<float *>(reinterpret_cast<char *>(orow) + p.rep_delta[g]) : orow;
#pragma unroll
                for (int q = 0; q < CDA_HIST_PREFETCH; ++q) {
                    const int e = lane + 32 * q - mis;
                    if (e >= 0 && e < cfg.W) o[e] = e < W_old ? hv[q] : __uint_as_float(SMW(wbL + L::SNAP + e - W_old));
                }
                for (int e = lane + 32 * CDA_HIST_PREFETCH - mis; e < cfg.W; e += 32) {   // beyond the prefetched chunks
                    float v;
                    if (e < W_old) { int ri = (slot_new + 1) * CDA_SNAPSHOT_DIM + e; if (ri >= cfg.W) ri -= cfg.W; v = g_hist[ri]; }
                    else v = __uint_as_float(SMW(wbL + L::SNAP + e - W_old));
                    o[e] = v;
                }
            }
        }
        __syncwarp();
        for (int cc = lane; cc < CDA_SNAPSHOT_DIM; cc += 32) g_hist[slot_new * CDA_SNAPSHOT_DIM + cc] = __uint_as_float(SMW(wbL + L::SNAP + cc));

        CDA_TICK(7);   // obs + ring written
        // ================= set_reward / set_done: reward_helper.py:35-103, done_helper.py ===
        bool broke = false;
        if (lane < A) {
            const double nav_change = (double)(ac.nav - nav_prev);
            const double nav_term = nav_change * (nav_change < 0 ? cfg.c_loss : 1.0);
            long long ddi = nav_max - ac.nav; if (ddi < 0) ddi = 0;
            double r = 0.0;
            r = r + nav_term;
            r = r + -(cfg.c_order * (double)((ac.ctr >> 24) & 1u));
            r = r + -(cfg.c_trade * (double)(ac.ctr & 0xfffu));
            r = r + -(cfg.c_dd * (double)ddi);
            r = r + cfg.c_passive * (double)((ac.ctr >> 12) & 0xfffu);
            if (p.reward && last_it) {
                double *rp = p.reward + (size_t)m * o_reward_stride + lane;
                *rp = r;
                for (int g = 1; g < o_rep_n; ++g) *reinterpret_cast<double *>(reinterpret_cast<char *>(rp) + p.rep_delta[g]) = r;
            }
            broke = DEC ? !nav_positive(k, ac.nav) : ac.nav <= 0;
            if (o_rec_inline) { const unsigned long long rb = (unsigned long long)__double_as_longlong(r); SMW(wbL + L::ACT + 2 * lane) = (unsigned)rb; SMW(wbL + L::ACT + 2 * lane + 1) = (unsigned)(rb >> 32); }
        }
        const unsigned done_mask = SMW(wbL + L::PARK + 10) | __ballot_sync(CDA_FULL, broke);
        __syncwarp();
        if (lane == 0) SMW(wbL + L::PARK + 10) = done_mask;
        const unsigned all = A >= 32 ? 0xffffffffu : ((1u << A) - 1u);
        if (SERVE) CDA_SRV_STAMP(2);   // step computed, outputs about to leave
        if (o_ring_out && last_it) {   // host ring / sliding window: only the newest 42 floats leave the GPU (128-B aligned chunks, like the stack)
            int nw = CDA_SNAPSHOT_DIM;
            if (o_rec_inline) {        // ... followed by the result record, which then shares the snapshot's last write transaction (the
                                       // decoded-action words are dead by now: their tile carries the record)
                if (lane == 0) { SMW(wbL + L::ACT + 2 * A) = ((done_mask & all) == all ? 1u : 0u) | (t_step + 1 >= (unsigned)cfg.max_step ? 0x100u : 0u); SMW(wbL + L::ACT + 2 * A + 1) = 0u; }
                nw += 2 * A + 2;
                __syncwarp();
            }
            const int nw_data = nw;
            if (p.ring_pad > nw) nw = p.ring_pad;
            // (resident server: cell m of the plane the message named; planes are M * ring_stride floats apart)
            float *rg0 = SERVE ? o_ring_out + ((size_t)SMW(wbL + L::SNAP + 42) * (size_t)p.M + (size_t)m) * (size_t)p.ring_stride
                               : o_ring_out + (size_t)m * p.ring_stride + p.ring_slot * CDA_SNAPSHOT_DIM;
#pragma unroll 1
            for (int g = 0; g < o_rep_n; ++g) {   // (fused all-gather: the same cell of every rank's window, over NVLink)
                float *rg = reinterpret_cast<float *>(reinterpret_cast<char *>(rg0) + (g ? p.rep_delta[g] : 0LL));
                for (int cc = lane - (int)((reinterpret_cast<size_t>(rg) >> 2) & 31); cc < nw; cc += 32) {
                    if (cc < 0) continue;
                    const float v = cc >= nw_data ? 0.f : __uint_as_float(cc < CDA_SNAPSHOT_DIM ? SMW(wbL + L::SNAP + cc) : SMW(wbL + L::ACT + cc - CDA_SNAPSHOT_DIM));
                    rg[cc] = v;
                    if (o_ring_mirror && cc < CDA_SNAPSHOT_DIM) rg[cfg.n_hist * CDA_SNAPSHOT_DIM + cc] = v;
                }
            }
        }
        if (lane == 0 && last_it) {
            const unsigned char f_term = (done_mask & all) == all, f_trunc = (t_step + 1 >= (unsigned)cfg.max_step);
            for (int g = 0; g < o_rep_n; ++g) {
                const long long dl = g ? p.rep_delta[g] : 0LL;
                if (o_flag_pack) *reinterpret_cast<unsigned short *>(p.term + (size_t)m * o_flag_stride + dl) = (unsigned short)(f_term | (f_trunc << 8));   // adjacent bytes: one store
                else {
                    if (p.term) p.term[(size_t)m * o_flag_stride + dl] = f_term;
                    if (p.trunc) p.trunc[(size_t)m * o_flag_stride + dl] = f_trunc;
                }
            }
            if (p.fill_counts) p.fill_counts[m] = k.n_fills;
            hdr[40] = (unsigned)best_bid; hdr[41] = (unsigned)best_ask;   // (inside `last_it`: a restarted rollout pass must find the header untouched)
        }
        t_step++;
        __syncwarp();
        if (SERVE) {   // this step's completion: outputs fenced, every warp counts itself, the last one rings the host (and tells the poller)
            if (lane == 0 && SMW(wbL + L::PARK + 11)) *p.status_flag = 1u;
            // every warp orders its stores before its count at GPU scope (cheap); the LAST warp's system-scope fence, made after it has
            // observed all the counts, is cumulative: everything the others stored is visible to the host before the completion word.
            // (4096 system-scope fences in flight at once cost every warp 4 - 14 us, profiles/r04a_serve_timeline.txt.)
            CDA_WARP_GPU_FENCE();
            __syncwarp();
            CDA_SRV_STAMP(3);   // outputs issued and ordered
            if (lane == 0 && atomicAdd(p.done_ctr, 1u) == (unsigned)p.M - 1u) {
                *p.done_ctr = 0u;          // (every other warp has counted itself, and none starts the next step before the host has seen this one)
                CDA_FINAL_SYS_FENCE();
                const unsigned sq = p.srv_next + (unsigned)it;
                *reinterpret_cast<volatile unsigned *>(p.done_flag) = sq;
                *reinterpret_cast<volatile unsigned *>(p.srv_done_dev) = sq;
                if (p.prof) p.prof[(size_t)p.M * 16 + 0] = globaltimer_ns();   // (timeline tool) completion rung
            }
        }
        if (ROLLOUT && (SERVE || !last_it)) {    // multi-step rollout: the accounts go back to their tile, the generator comes back for the next step's draws
            if (lane < A) {
                long long *sq = reinterpret_cast<long long *>(&smw[acct_w]);
                sq[lane] = ac.cash; sq[A + lane] = ac.hold; sq[2 * A + lane] = ac.cost; sq[3 * A + lane] = ac.nav;
                sq[4 * A + lane] = nav_prev_carry; sq[5 * A + lane] = nav_max_carry;
                SMW(acct_w + 12 * A + lane) = (unsigned)(int)ac.pos; SMW(acct_w + 13 * A + lane) = ac.ntr;
            }
            const unsigned long long *pk = reinterpret_cast<const unsigned long long *>(&smw[wbL + L::PARK]);
            rng.shi = pk[0]; rng.slo = pk[1]; rng.ihi = pk[2]; rng.ilo = pk[3];
            rng.has32 = SMW(wbL + L::PARK + 8); rng.u32 = SMW(wbL + L::PARK + 9);
        }
        if (SERVE) {
            // The CTA's warps meet before the next step: the next step's action copy (issued by whichever warp sees the message first)
            // overwrites the tile the others read THIS step's actions from.  That is already ordered through the completion count, the
            // host and the next message; the barrier makes it explicit inside the CTA (and visible to racecheck) at no cost — a warp
            // that is done has nothing to do before every warp of the grid is done.  (Named barrier with the live warps' thread count:
            // warps beyond the last market have left the kernel.)
            const int live = min(WARPS, p.M - (int)(blockIdx.x - 1u) * WARPS);
            asm volatile("bar.sync 1, %0;" ::"r"(live * 32) : "memory");
        }
    }

    if (SERVE && it == 0) {   // stopped before the first step: nothing has changed (the staged copies must have landed before the CTA may go)
        if (!waited) mbar_wait(bar, 0u);
        return;
    }
    CDA_TICK(8);   // reward/done
    const int wbL = (int)(fresh_tid_x() >> 5) * L::WORDS;   // (as inside the loop: not the entry-time copy)
    // ---- store: header, accounts, pool prefix
    if (DEC) {   // a Decimal operation left the 128-bit domain (sizes / prices far beyond the reference's ranges): sticky status
        const bool re = k.twf_w >= 0 && (SMW(k.twf_w) & CDA_TWF_RANGE);
        if (__any_sync(CDA_FULL, re)) k.raise(CDA_ST_DEC_RANGE);
        __syncwarp();
    }
    if (lane == 0) {
        *reinterpret_cast<uint4 *>(hdr + 0) = make_uint4(k.time, k.next_id, k.seqctr, t_step);
        if (cfg.fill_tape && p.fills) hdr[42] = (unsigned)k.n_fills;
        const unsigned stv = SMW(wbL + L::PARK + 11);
        if (stv) *p.status_flag = 1u;   // (rare) lets the host notice a sticky status without a gather
        *reinterpret_cast<uint4 *>(hdr + 4) = make_uint4((unsigned)k.tape_px, k.tape_nonempty ? CDA_FLAG_TAPE : 0u, SMW(wbL + L::PARK + 10), stv);
        const unsigned long long *pk = reinterpret_cast<const unsigned long long *>(&smw[wbL + L::PARK]);
        *reinterpret_cast<uint4 *>(hdr + 8) = make_uint4((unsigned)k.nb, (unsigned)k.na, SMW(wbL + L::PARK + 8), SMW(wbL + L::PARK + 9));
        *reinterpret_cast<ulonglong2 *>(hdr + 12) = make_ulonglong2(pk[0], pk[1]);
        *reinterpret_cast<ulonglong2 *>(hdr + 16) = make_ulonglong2(pk[2], pk[3]);
    }
    if (lane < A) {
        CDA_ACCT_PTRS
        g_cash[lane] = ac.cash; g_hold[lane] = ac.hold; g_cost[lane] = ac.cost; g_nav[lane] = ac.nav;
        g_prev[lane] = nav_prev_carry; g_max[lane] = nav_max_carry; g_pos[lane] = (int)ac.pos; g_ntr[lane] = ac.ntr;
        g_ctr[lane] = ac.ctr;
        if (DEC && k.twf_w >= 0) {
            const unsigned twf = SMW(k.twf_w);
            g_twf[lane] = twf & ~CDA_TWF_RANGE;
        }
    }
    fence_proxy_async();
    __syncwarp();
    {
        const unsigned ob = (((unsigned)k.nb + 31u) >> 5) * (CDA_TILE_WORDS * 4u), oa = (((unsigned)k.na + 31u) >> 5) * (CDA_TILE_WORDS * 4u);
        if (lane == 0 && (ob | oa)) {
            if (ob) bulk_s2g(gpool, sa + L::POOL * 4u, ob);
            if (oa) bulk_s2g(gpool + CDA_POOL_FIELDS * CAP, sa + (L::POOL + CDA_POOL_FIELDS * CAP) * 4u, oa);
            bulk_commit();
            bulk_wait_read0();
        }
    }
    if (ROUTED && !SERVE && p.done_flag) {
        // this warp's stores are ordered before its count; host memory only: at GPU scope — the last warp's system-scope fence below is
        // cumulative over everything it has observed (thousands of system-scope fences in flight at once cost each warp microseconds,
        // profiles/r04a_serve_timeline.txt); peer windows over NVLink keep the system-scope fence per warp
        if (p.rep_n > 1) __threadfence_system(); else __threadfence();
        __syncwarp();
        if (lane == 0 && atomicAdd(p.done_ctr, 1u) == (unsigned)p.M - 1u) {
            *p.done_ctr = 0u;          // (every other warp has counted itself: the next launch starts from zero)
            __threadfence_system();
            *reinterpret_cast<volatile unsigned *>(p.done_flag) = p.done_seq;
            for (int g = 1; g < o_rep_n; ++g)    // fused all-gather: this rank's "step done" word in every peer's flag array
                *reinterpret_cast<volatile unsigned *>(reinterpret_cast<char *>(p.done_flag) + p.rep_delta[g]) = p.done_seq;
        }
    }
    CDA_TICK(9);   // state stored
#ifdef CDA_PROFILE_PHASES
    { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); if (lane == 0 && p.prof) p.prof[(size_t)m * 16 + 13] = gt; }   // warp end (ns)
#endif
    return;
  }
resolve:
    if (!DEC) return;        // (unreachable without the Decimal twin: nothing jumps here)
    // ---- decimal_ledger, rare: answer the parked tie requests with the Decimal twin (the ONLY place the step kernel calls the 128-bit
    //      arithmetic: top level, nothing live), park the answers, run the launch's steps for this market again
    {
        const int wbR = (int)(fresh_tid_x() >> 5) * L::WORDS, lnR = (int)(fresh_tid_x() & 31u);
        const int mR = blockIdx.x * WARPS + (int)(fresh_tid_x() >> 5);
        const int rq = wbR + L::POOL + 8 * lnR;
        __syncwarp();
        const unsigned mode = SMW(rq);
        if (mode) {
            const long long a0 = (long long)(((unsigned long long)SMW(rq + 2) << 32) | SMW(rq + 1)), a1 = (long long)(((unsigned long long)SMW(rq + 4) << 32) | SMW(rq + 3));
            const int twfR = WARPS * L::WORDS + 5 * WARPS * p.cfg.A + 4 + (int)(fresh_tid_x() >> 5) * (16 * p.cfg.A) + 15 * p.cfg.A + lnR;   // this lane's twin flags (account tile)
            const unsigned r = cda_twin_query(CDA_TWIN_ST(p, mR, lnR), CDA_TWIN_JR(p, mR, lnR), (int)(SMW(twfR) & CDA_TWF_JN_MASK), (int)mode, a0, a1, (long long)(int)SMW(rq + 5));
            const unsigned idx = (atomicAdd(&smw[wbR + L::TIE], 0x100u) >> 8) & 0xffu;
            unsigned *hdrR = reinterpret_cast<unsigned *>(p.state + (size_t)mR * p.cfg.stride);
            if (idx < CDA_TIE_SLOTS) SMW(wbR + L::TIE + 1 + idx) = (SMW(rq + 6) << 2) | (r & 3u);
            if (idx >= CDA_TIE_SLOTS || (r & 4u)) { atomicOr(hdrR + 7, CDA_ST_DEC_RANGE); *p.status_flag = 1u; }   // table full / out of the 128-bit domain: flagged
        }
        __syncwarp();
        if (lnR == 0) { const unsigned w = SMW(wbR + L::TIE); SMW(wbR + L::TIE) = (w & ~0xffu) | ((w + 1u) & 0xffu); atomicAdd(&cda_debug_restarts, 1ULL); }
        __syncwarp();
    }
    goto restart;
}

// ------------------------------------------------------------------------------------------
// reset: continuousDoubleAuction_env.py:175-231 — one thread per market (cold path).
// ------------------------------------------------------------------------------------------
__global__ void cda_reset_kernel(CdaDevCfg cfg, unsigned char *state, int M, const unsigned long long *seeds,
                                 const unsigned char *mask, float *obs, int *fill_counts) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    if (mask && !mask[m]) return;
    if (fill_counts) fill_counts[m] = 0;      // the fill log / tape of a reset market is empty
    unsigned char *blk = state + (size_t)m * cfg.stride;
    unsigned *hdr = reinterpret_cast<unsigned *>(blk);
    CdaRng rng;
    if (seeds) rng_seed(rng, seeds[m]);
    else {
        rng.has32 = hdr[10]; rng.u32 = hdr[11];
        const unsigned long long *r = reinterpret_cast<const unsigned long long *>(hdr + 12);
        rng.shi = r[0]; rng.slo = r[1]; rng.ihi = r[2]; rng.ilo = r[3];
    }
    const int anchor = (int)rng_integers(rng, cfg.price_lo, (long long)cfg.price_hi + 1);   // :219-221
    for (int i = 0; i < CDA_HDR_BYTES / 4; ++i) hdr[i] = 0;
    hdr[4] = (unsigned)anchor;
    hdr[10] = rng.has32; hdr[11] = rng.u32;
    unsigned long long *r = reinterpret_cast<unsigned long long *>(hdr + 12);
    r[0] = rng.shi; r[1] = rng.slo; r[2] = rng.ihi; r[3] = rng.ilo;
    const int A = cfg.A;
    long long *g_cash = reinterpret_cast<long long *>(blk + cfg.off_acct);
    for (int a = 0; a < A; ++a) {                                                          // account.py:55-82
        g_cash[a] = cfg.init_cash; g_cash[A + a] = 0; g_cash[2 * A + a] = 0;
        g_cash[3 * A + a] = cfg.init_cash; g_cash[4 * A + a] = cfg.init_cash; g_cash[5 * A + a] = cfg.init_cash;
    }
    int *g_pos = reinterpret_cast<int *>(g_cash + 6 * A);
    for (int a = 0; a < 4 * A; ++a) g_pos[a] = 0;                                         // position, num_trades, step counters, twin flags
    if (cfg.dec) {                                                                         // Decimal twin: VWAP 0, cash untracked, empty journal
        unsigned long long *tw = reinterpret_cast<unsigned long long *>(blk + cfg.off_twin);
        for (int a = 0; a < A * (CDA_TWIN_BYTES / 8); ++a) tw[a] = 0ULL;
    }
    // empty-book snapshot (state_helper.py:66-78, :163-175): zeros, log(anchor), 0
    float *g_hist = reinterpret_cast<float *>(blk + cfg.off_hist);
    double Mid = (double)anchor; if (Mid <= 0) Mid = 100.0;
    const float lm = (float)log(Mid);
    for (int h = 0; h < cfg.n_hist; ++h)
        for (int cc = 0; cc < CDA_SNAPSHOT_DIM; ++cc) {
            const float v = cc == 40 ? lm : 0.0f;
            g_hist[h * CDA_SNAPSHOT_DIM + cc] = v;
            if (obs) obs[(size_t)m * cfg.W + h * CDA_SNAPSHOT_DIM + cc] = v;
        }
}

// stacked observation of every market rebuilt from its snapshot ring (state_helper.py:88-90): row m of `dst`
// (row stride in floats) receives n_hist x 42 floats, oldest snapshot first.  Cold path (reset of the host window).
__global__ void cda_emit_obs_kernel(CdaDevCfg cfg, const unsigned char *state, int M, float *dst, int stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * cfg.W) return;
    const int m = i / cfg.W, e = i - m * cfg.W, j = e / CDA_SNAPSHOT_DIM, cc = e - j * CDA_SNAPSHOT_DIM;
    const unsigned char *blk = state + (size_t)m * cfg.stride;
    const unsigned t_step = reinterpret_cast<const unsigned *>(blk)[3];
    const float *g_hist = reinterpret_cast<const float *>(blk + cfg.off_hist);
    dst[(size_t)m * stride + e] = g_hist[((t_step + (unsigned)j) % (unsigned)cfg.n_hist) * CDA_SNAPSHOT_DIM + cc];
}

// dense plane ring (cda_step_planes): the n_hist most recent snapshots of every market go to the planes ending at slot `pos`
// (plane (pos - n_hist + 1 + j) mod slots receives the j-th oldest), cell m of each plane.  Cold path (reset / attach).
__global__ void cda_emit_planes_kernel(CdaDevCfg cfg, const unsigned char *state, int M, float *planes, int slots, int cell, int pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * cfg.W) return;
    const int m = i / cfg.W, e = i - m * cfg.W, j = e / CDA_SNAPSHOT_DIM, cc = e - j * CDA_SNAPSHOT_DIM;
    const unsigned char *blk = state + (size_t)m * cfg.stride;
    const unsigned t_step = reinterpret_cast<const unsigned *>(blk)[3];
    const float *g_hist = reinterpret_cast<const float *>(blk + cfg.off_hist);
    const int slot = ((pos - cfg.n_hist + 1 + j) % slots + slots) % slots;
    planes[((size_t)slot * M + m) * cell + cc] = g_hist[((t_step + (unsigned)j) % (unsigned)cfg.n_hist) * CDA_SNAPSHOT_DIM + cc];
}

// fill every slot of the mirrored host ring of the selected markets with their current newest snapshot
// (after a reset all device ring slots hold the initial snapshot)
__global__ void cda_ring_fill_kernel(CdaDevCfg cfg, const unsigned char *state, int M, const unsigned char *mask, float *ring) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = 2 * cfg.n_hist * CDA_SNAPSHOT_DIM;
    if (i >= M * per) return;
    const int m = i / per, e = i - m * per;
    if (mask && !mask[m]) return;
    const float *g_hist = reinterpret_cast<const float *>(state + (size_t)m * cfg.stride + cfg.off_hist);
    ring[i] = g_hist[e % CDA_SNAPSHOT_DIM];   // slot 0 (all slots are equal right after a reset)
}

// fused all-gather: wait (on the consumer's stream) until every rank has published step `seq` into this rank's flag array.  A rank may be
// one step ahead (flag == seq + 1): compare as signed distance.  Gives up after ~2 s (a dead peer must not hang the GPU): *err = 1.
__global__ void cda_gather_wait_kernel(const volatile unsigned *flags, int world, unsigned seq, unsigned *err) {
    if ((int)threadIdx.x < world) {
        const long long t0 = clock64();
        while ((int)(flags[threadIdx.x] - seq) < 0) {
            __nanosleep(200);
            if (clock64() - t0 > 4000000000LL) { *err = 1u; break; }
        }
    }
    __threadfence_system();
}

// ------------------------------------------------------------------------------------------
// decimal_ledger: periodic journal replay, one THREAD per (market, agent) (cda_twin.cuh); runs between steps on the env's stream
// ------------------------------------------------------------------------------------------
__global__ void cda_twin_flush_kernel(CdaDevCfg cfg, unsigned char *state, int M, unsigned *status_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * cfg.A) return;
    const int m = i / cfg.A, a = i - m * cfg.A;
    unsigned char *blk = state + (size_t)m * cfg.stride;
    unsigned *twfp = reinterpret_cast<unsigned *>(blk + cfg.off_acct) + 15 * cfg.A + a;
    const unsigned twf = *twfp;
    const int jn = (int)(twf & CDA_TWF_JN_MASK);
    if (jn == 0) return;
    const unsigned r = cda_twin_replay(reinterpret_cast<CdaTwinStored *>(blk + cfg.off_twin) + a,
                                       reinterpret_cast<const unsigned long long *>(blk + cfg.off_jrn) + a * CDA_JRN_E, jn, 0, 0, 0, 0);
    *twfp = (twf & ~(CDA_TWF_JN_MASK | CDA_TWF_TRACKED)) | ((r & 1u) ? CDA_TWF_TRACKED : 0u);
    if (r & 2u) { atomicOr(reinterpret_cast<unsigned *>(blk) + 7, CDA_ST_DEC_RANGE); *status_flag = 1u; }
}
// the twins (after a flush) for inspection: out i64[M][A][8] = vwap lo, hi, exp, sign, cash lo, hi, exp | sign << 32, with word 7's
// bit 40 = cash tracked and the position in word 3's upper half (see VecCDAEnv.decimal_fields)
__global__ void cda_twin_dump_kernel(CdaDevCfg cfg, const unsigned char *state, int M, long long *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * cfg.A) return;
    const int m = i / cfg.A, a = i - m * cfg.A;
    const unsigned char *blk = state + (size_t)m * cfg.stride;
    const CdaTwinStored *s = reinterpret_cast<const CdaTwinStored *>(blk + cfg.off_twin) + a;
    long long *o = out + (size_t)i * 8;
    o[0] = (long long)s->vwap_lo; o[1] = (long long)s->vwap_hi; o[2] = s->vwap_exp; o[3] = (long long)(unsigned)s->vwap_sign | ((long long)s->pos << 32);
    o[4] = (long long)s->cash_lo; o[5] = (long long)s->cash_hi; o[6] = s->cash_exp; o[7] = (long long)(unsigned)s->cash_sign | ((long long)(s->flags & 1u) << 40);
}

// ------------------------------------------------------------------------------------------
// lazy info gather (info_helper.py:30-116): one thread per (market, agent)
// ------------------------------------------------------------------------------------------
__global__ void cda_info_kernel(CdaDevCfg cfg, const unsigned char *state, int M, int field, long long *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int A = cfg.A;
    if (field < 0) {   // all fields: [CDA_INFO_MARKET][M][A] then the market block [M][8]
        const int per = M * A;
        if (i >= CDA_INFO_MARKET * per + M) return;
        if (i < CDA_INFO_MARKET * per) { field = i / per; i -= field * per; out += (size_t)field * per; }
        else { i -= CDA_INFO_MARKET * per; field = CDA_INFO_MARKET; out += (size_t)CDA_INFO_MARKET * per; }
    }
    if (field == CDA_INFO_MARKET) {
        if (i >= M) return;
        const unsigned *hdr = reinterpret_cast<const unsigned *>(state + (size_t)i * cfg.stride);
        long long *o = out + (size_t)i * 8;
        o[0] = (int)hdr[4]; o[1] = (int)hdr[40]; o[2] = (int)hdr[41]; o[3] = hdr[0]; o[4] = hdr[1]; o[5] = hdr[3]; o[6] = hdr[6]; o[7] = hdr[7];
        return;
    }
    if (i >= M * A) return;
    const int m = i / A, a = i - m * A;
    const unsigned char *blk = state + (size_t)m * cfg.stride;
    const unsigned *hdr = reinterpret_cast<const unsigned *>(blk);
    const long long *g = reinterpret_cast<const long long *>(blk + cfg.off_acct);
    const int *g_pos = reinterpret_cast<const int *>(g + 6 * A);
    const unsigned *g_ntr = reinterpret_cast<const unsigned *>(g_pos + A), *g_ctr = g_ntr + A;
    long long v = 0;
    const unsigned ctr = g_ctr[a];
    switch (field) {
        case CDA_INFO_CASH: v = g[a]; break;
        case CDA_INFO_CASH_ON_HOLD: v = g[A + a]; break;
        case CDA_INFO_COST_BASIS: v = g[2 * A + a]; break;
        case CDA_INFO_NAV: v = g[3 * A + a]; break;
        case CDA_INFO_PREV_NAV: v = g[4 * A + a]; break;
        case CDA_INFO_MAX_NAV: v = g[5 * A + a]; break;
        case CDA_INFO_NET_POSITION: v = g_pos[a]; break;
        case CDA_INFO_POSITION_VAL: {   // calculate.py:35-55 at the last mark-to-market
            const long long pos = g_pos[a], ap = pos < 0 ? -pos : pos, lp = (int)hdr[4];
            v = (hdr[5] & CDA_FLAG_TAPE) ? (pos >= 0 ? ap * lp : 2 * g[2 * A + a] - ap * lp) : 0;
            break;
        }
        case CDA_INFO_NUM_TRADES: v = g_ntr[a]; break;
        case CDA_INFO_NUM_TRADES_STEP: v = ctr & 0xfffu; break;
        case CDA_INFO_NUM_PASSIVE_FILLS_STEP: v = (ctr >> 12) & 0xfffu; break;
        case CDA_INFO_ORDER_STEP_PLACED: v = (ctr >> 24) & 1u; break;
        case CDA_INFO_NUM_REJECTED_STEP: v = (ctr >> 25) & 1u; break;
        case CDA_INFO_IS_PASS_ACTION: v = (ctr >> 26) & 1u; break;
        default: break;
    }
    out[i] = v;
}
