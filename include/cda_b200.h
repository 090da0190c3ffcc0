/* cda_b200.h — C-ABI of the B200-native vectorised continuous-double-auction environment.
 *
 * This is the drop-in boundary for the reference's per-step env hot path.  The reference
 * (ChuaCheowHuan/gym-continuousDoubleAuction) is pure Python and has no FFI of its own; the
 * boundary it exposes is the `continuousDoubleAuctionEnv` class surface
 * (gym_continuousDoubleAuction/envs/continuousDoubleAuction_env.py:21-359).  Each entry point
 * below names the reference call it replaces.  All pointers are plain device or host pointers,
 * sizes are plain integers, `stream` is a cudaStream_t passed as void* (NULL = default
 * stream).  No torch types appear here; PyTorch only supplies the buffers and the stream.
 *
 * Shapes (M = num_markets, A = num_agents, W = n_hist * 42):
 *   actions   category i32[M][A]  (0..8, reference _CATEGORY_MAP action_helper.py:12-22;
 *                                  -1 = agent absent from the action dict: no RNG draw)
 *             size_mean f32[M][A], size_sigma f32[M][A], price i32[M][A] (0..9),
 *             price_offset i32[M][A] (0..2)            (action space: action_helper.py:126-138)
 *   obs       f32[M][W]   stacked observation, oldest snapshot first (state_helper.py:80-92)
 *   reward    f64[M][A]   (reward_helper.py:35-103; f64 because |r| reaches 1e4 and parity is 1e-6)
 *   terminated/truncated u8[M]  == terminateds["__all__"] / truncateds["__all__"]
 *                                  (done_helper.py:20-55)
 *
 * Every function returns 0 on success or a negative CDA_E* code; cda_strerror() names it.
 * Calls on one handle are stream-ordered and not re-entrant; use one handle per GPU/process.
 */
#ifndef CDA_B200_H
#define CDA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDA_K_ROWS 10   /* observation_layout.k_rows  (config/tunable_constants.json) */
#define CDA_SNAPSHOT_DIM 42
#define CDA_MAX_AGENTS 32
#define CDA_MAX_HIST 16
#define CDA_FILL_WORDS 8 /* time, price, qty, maker, maker_order_id, maker_left(-1=None), taker, taker_side(0 bid/1 ask) */

/* error codes */
#define CDA_OK 0
#define CDA_EINVAL (-1)   /* bad argument / config (reference raises ValueError at construction) */
#define CDA_ECUDA (-2)    /* a CUDA runtime call failed; see cda_last_cuda_error() */
#define CDA_ENOMEM (-3)
#define CDA_ESTATE (-4)   /* handle used before reset, or after a sticky device error */

/* per-market sticky status bits (replace the reference's sys.exit() paths, orderbook.py:42,58,159) */
#define CDA_ST_POOL_OVERFLOW 1u   /* more resting orders on one side than order_capacity: order dropped */
#define CDA_ST_FILL_OVERFLOW 2u   /* more fills in one step than the fill log holds (log truncated only) */
#define CDA_ST_BAD_ACTION 4u      /* category/price/price_offset outside the action space */
#define CDA_ST_PRICE_RANGE 8u     /* a price left [1, 2^24): outside the exactly-representable range */
#define CDA_ST_BAD_SIZE 16u       /* order size <= 0 (reference: sys.exit in process_order) */
#define CDA_ST_DEC_RANGE 32u      /* decimal_ledger: a Decimal(28) operation left the 128-bit domain (position >= 1e10 or a price beyond 2^24) */

/* Mirrors the 17 env config keys (continuousDoubleAuction_env.py:35-53) that touch the hot path. */
typedef struct CdaConfig {
    int32_t num_agents;          /* num_of_agents, 1..32 */
    int32_t n_hist;              /* 1..16 */
    int32_t max_step;
    int32_t tick_size;           /* integral ticks only (>=1) */
    int64_t init_cash;           /* > 0 */
    int32_t min_size, mkt_max_size, limit_size_multiple;
    int32_t initial_price_min, initial_price_max;   /* inclusive anchor range */
    int32_t order_capacity;      /* resting orders per side per market: 64, 128, 160, 192 or 256; 0 = auto (160 up to 8 agents, else 256) */
    int32_t fill_capacity;       /* fills logged per market per step (0 = no fill log) */
    double order_penalty, trade_penalty, drawdown_penalty, passive_bonus, loss_multiplier;
    int32_t decimal_ledger;      /* 1: keep the reference's Decimal(prec 28) residues (deferred twin, csrc/cda_twin.cuh) so that a cash gate or a
                                    bankruptcy test at EXACT integer equality is decided like the reference's Decimal compare
                                    (agent/trader.py:108-151); 0: exact int64 ledger only (identical except at such ties) */
    int32_t fill_tape;           /* 1: the fill log is a tape — a ring holding the last fill_capacity fills of each market ACROSS steps and launches
                                    (the reference's LOB.tape, orderbook.py:20,140, bounded), row of fill number n at n % fill_capacity, and the count
                                    cda_get_fills returns is the number of fills since the reset; 0: one step's fills, overwritten by the next step */
} CdaConfig;

typedef struct CdaEnv CdaEnv; /* opaque handle */

/* info fields for cda_get_info (info_helper.py:30-116) */
enum CdaInfoField {
    CDA_INFO_CASH = 0,        /* i64[M][A] */
    CDA_INFO_CASH_ON_HOLD,    /* i64[M][A] */
    CDA_INFO_COST_BASIS,      /* i64[M][A]  |net_position| * VWAP  (VWAP = this / |pos|) */
    CDA_INFO_NAV,             /* i64[M][A] */
    CDA_INFO_PREV_NAV,        /* i64[M][A] */
    CDA_INFO_MAX_NAV,         /* i64[M][A] */
    CDA_INFO_NET_POSITION,    /* i64[M][A] */
    CDA_INFO_POSITION_VAL,    /* i64[M][A] */
    CDA_INFO_NUM_TRADES,      /* i64[M][A] */
    CDA_INFO_NUM_TRADES_STEP, /* i64[M][A] */
    CDA_INFO_NUM_PASSIVE_FILLS_STEP, /* i64[M][A] */
    CDA_INFO_ORDER_STEP_PLACED,      /* i64[M][A] */
    CDA_INFO_NUM_REJECTED_STEP,      /* i64[M][A] */
    CDA_INFO_IS_PASS_ACTION,         /* i64[M][A] */
    CDA_INFO_MARKET,          /* i64[M][8]: last_price, best_bid(0=None), best_ask(0=None), time,
                                 next_order_id, t_step, done_mask, status */
    CDA_INFO__COUNT
};

/* continuousDoubleAuctionEnv.__init__ (continuousDoubleAuction_env.py:27-119), for M markets.
 * One handle takes at most INT_MAX / max(2*n_hist*42, 15*num_agents) markets (6.39 M with the defaults, 48 GB of
 * state): CDA_EINVAL beyond that — shard the markets over several handles / GPUs.
 * Host buffers handed to the cda_step_host* / window / ring calls are looked up ONCE per distinct pointer (pinned +
 * mapped?) and the answer is cached in the handle: keep a buffer pinned for as long as the handle may see its address. */
int cda_create(const CdaConfig *cfg, int32_t num_markets, int32_t device, CdaEnv **out);
int cda_destroy(CdaEnv *env);

/* env.reset(seed=...) (continuousDoubleAuction_env.py:175-231).
 *   d_seeds: u64[M] device pointer, or NULL to keep each market's stream (reset(seed=None)).
 *            Market m is seeded exactly like gymnasium: Generator(PCG64(SeedSequence(seed))).
 *   d_mask:  u8[M] device pointer selecting markets to reset, or NULL for all.
 *   d_obs:   f32[M][W] output (rows of unselected markets untouched), may be NULL. */
int cda_reset(CdaEnv *env, const uint64_t *d_seeds, const uint8_t *d_mask, float *d_obs, void *stream);

/* env.step(action_dict) (continuousDoubleAuction_env.py:265-309) for all M markets: one fused
 * kernel launch; no host synchronisation.  All pointers are DEVICE pointers. */
int cda_step(CdaEnv *env, const int32_t *d_category, const float *d_size_mean, const float *d_size_sigma,
             const int32_t *d_price, const int32_t *d_price_offset, float *d_obs, double *d_reward,
             uint8_t *d_terminated, uint8_t *d_truncated, void *stream);

/* Same call with HOST buffers (pinned for full speed): copies the five action arrays to the
 * device, steps, copies obs/reward/flags back, all on `stream`; the caller synchronises the
 * stream before reading.  This is the end-to-end path a host-side policy uses. */
int cda_step_host(CdaEnv *env, const int32_t *h_category, const float *h_size_mean, const float *h_size_sigma,
                  const int32_t *h_price, const int32_t *h_price_offset, float *h_obs, double *h_reward,
                  uint8_t *h_terminated, uint8_t *h_truncated, void *stream);

/* Host path with a MIRRORED OBSERVATION RING (halves-and-more the PCIe traffic of cda_step_host):
 * the stacked observation is n_hist snapshots of 42 floats of which only the newest is new each step, so
 * the host keeps, per market, a ring of 2*n_hist snapshot slots  h_ring f32[M][2*n_hist][42]  (pinned +
 * mapped).  Each step the kernel stores ONLY the newest snapshot, at slots `pos` and `pos + n_hist`
 * (pos = ring_pos mod n_hist); the stacked observation of every market is then the contiguous window of
 * n_hist slots starting at slot pos + 1 — a zero-copy strided view, oldest snapshot first, identical to
 * cda_step_host's obs.  cda_reset_host_ring resets markets and fills all their slots with the initial
 * snapshot.  The caller advances ring_pos by one per step (any start value).
 * MEASURED (B200, PCIe Gen5): no faster than cda_step_host — posted PCIe writes from the SMs are bound by the
 * number of write transactions, and 2 x 168 B per market needs as many as one 672-B row.  Kept as an option for
 * hosts where the byte count matters (e.g. a remote/virtualised PCIe path). */
int cda_step_host_ring(CdaEnv *env, const int32_t *h_category, const float *h_size_mean, const float *h_size_sigma,
                       const int32_t *h_price, const int32_t *h_price_offset, float *h_ring, double *h_reward,
                       uint8_t *h_terminated, uint8_t *h_truncated, int64_t ring_pos, void *stream);
int cda_reset_host_ring(CdaEnv *env, const uint64_t *d_seeds, const uint8_t *d_mask, float *h_ring, void *stream);

/* Host path with a SLIDING OBSERVATION WINDOW — the lowest-traffic end-to-end path.
 * Of the n_hist snapshots in a stacked observation only the newest is new each step (state_helper.py:80-92:
 * the deque drops the oldest and appends one), so the host keeps, per market, a row of `slots` snapshot slots
 *     h_window f32[M][slots][42]   (pinned; slots >= n_hist, e.g. 64)
 * and each step only the newest snapshot crosses PCIe, into slot `pos` of every row (ONE strided copy-engine
 * transfer of M x 168 B instead of M x 672 B).  The stacked observation of market m is then the contiguous run of
 * n_hist slots ending at `pos` — h_window[m][pos-n_hist+1 .. pos] — a zero-copy view with the same 168 floats, oldest
 * first, as cda_step_host's obs row.  The caller advances pos by one per step; when it would reach `slots` it passes
 * pos = n_hist-1 instead, which re-sends the whole stack into slots 0..n_hist-1 (amortised over slots-n_hist+1 steps).
 * cda_reset_host_window resets the selected markets and sends every market's stack to slots 0..n_hist-1 (the next
 * step uses pos = n_hist).  Actions as in cda_step_host (read in place when the five arrays are one pinned block).
 * The other per-step results arrive as ONE packed record per market,
 *     h_records: M records of cda_record_bytes() bytes = { double reward[A]; uint8_t terminated, truncated; pad }
 * (8*(A+1) rounded up to a multiple of 64, so a record never straddles a 64-byte line: with a pinned + mapped, 64-B
 * aligned array the kernel writes it with two stores per market; otherwise it is staged and copied contiguously).
 * sync != 0: the call returns after the stream has been synchronised (outputs readable).  Use this path consistently
 * between resets: it keeps the device copy of the full stack in the handle's own staging buffer. */
int cda_step_host_window(CdaEnv *env, const int32_t *h_category, const float *h_size_mean, const float *h_size_sigma,
                         const int32_t *h_price, const int32_t *h_price_offset, float *h_window, int32_t slots, int32_t pos,
                         void *h_records, int32_t sync, void *stream);
int cda_reset_host_window(CdaEnv *env, const uint64_t *d_seeds, const uint8_t *d_mask, float *h_window, int32_t slots, void *stream);
/* Bound form of the same call for tight host loops: register the window, the record array and the stream once;
 * each step then passes only the pinned action block (category, size_mean bits, size_sigma bits, price,
 * price_offset as 4-byte words), the slot position and `flags`:
 *   CDA_WIN_SYNC           return after the stream has been synchronised
 *   CDA_WIN_MARKET_MAJOR   the block is i32[M][5][A] (one 20*A-byte action record per market) instead of i32[5][M][A]:
 *                          the markets of a CTA are then ONE contiguous run, fetched over PCIe by one bulk copy per CTA
 *                          instead of five (reads from host memory are bound by the number of requests)
 *   CDA_WIN_INLINE_RECORD  this step's result record { double reward[A]; uint8_t terminated, truncated; pad to 8 } is
 *                          placed right BEHIND the newest snapshot — at the head of slot pos+1 of every row, i.e. at
 *                          h_window[m][pos+1][0..] — instead of in h_records: the kernel then emits it in the same store
 *                          instructions as the snapshot's tail, which saves two PCIe write transactions per market and
 *                          step (posted writes from the SMs are bound by the number of transactions: measured
 *                          tools/pcie_store_bench.cu).  Slot pos+1 is overwritten by the next step's snapshot, so the
 *                          record is valid until the next call, like the observation view.  Needs pos+1 < slots (the
 *                          caller wraps one slot earlier) and 2*A+2 <= 42. */
#define CDA_WIN_SYNC 1
#define CDA_WIN_MARKET_MAJOR 2
#define CDA_WIN_INLINE_RECORD 4
int cda_window_bind(CdaEnv *env, float *h_window, int32_t slots, void *h_records, void *stream);
int cda_step_window(CdaEnv *env, const int32_t *h_action_block, int32_t pos, int32_t flags);

/* Host path with a DENSE PLANE RING — the end-to-end path that scales on a multi-GPU node.
 * The sliding window above keeps every market's stack contiguous, at the price of scattering each step's output over M rows
 * (one 168 + 8(A+1)-byte piece per 5 KB row: every market touches its own page and 3-4 partial cache lines of host memory).  On a
 * node where 8 GPUs store into one host that is what saturates first (profiles/r03b_diag8_split.txt: the output leg grows from
 * 12 to 57 us per step between 1 and 8 active GPUs while the input leg and the kernels do not change).  Here the host keeps
 *     h_planes f32[slots][M][cell_words]   (pinned + mapped; cell_words even, >= 42 + 2A + 2, e.g. 64 = one aligned 256-B cell)
 * and step t stores, for every market m, the newest snapshot followed by the result record { double reward[A]; uint8_t terminated,
 * truncated; pad } into cell m of plane (t mod slots), zero-padded to the cell size: ONE contiguous, line-aligned M*cell region per
 * step (two whole 128-B stores per market), no per-market pages, nothing is ever re-sent.  The stacked observation of market m is
 * the n_hist cells  h_planes[(pos-n_hist+1 .. pos) mod slots][m][0..41]  (oldest first) — a strided [M][n_hist][42] view, not one
 * contiguous row: a consumer that needs f32[M][168] contiguous makes that one copy itself (uploads can take the planes as they are).
 * `h_plane` in cda_step_planes is the address of plane `pos` (h_planes + pos*M*cell_words); flags: CDA_WIN_SYNC, CDA_WIN_MARKET_MAJOR.
 * cda_reset_planes resets the selected markets and (re)writes the n_hist planes ending at `pos` for every market. */
int cda_step_planes(CdaEnv *env, const int32_t *h_action_block, float *h_plane, int32_t cell_words, int32_t flags, void *stream);
int cda_reset_planes(CdaEnv *env, const uint64_t *d_seeds, const uint8_t *d_mask, float *h_planes, int32_t slots, int32_t cell_words, int32_t pos, void *stream);

/* ---- RESIDENT STEP SERVER: the plane path without a launch, a stream hand-shake or a state load / store per step ----
 * For a policy that lives on the HOST (the reference's: RLlib workers call env.step from Python, train/train.py:509-518) the per-step
 * cost of cda_step_planes is a kernel launch, the completion hand-shake and the market state's round trip through HBM — more than the
 * matching itself.  cda_serve_step removes all three: the step kernel is launched once and STAYS RESIDENT (one warp per market; books
 * and ledgers kept in shared memory between steps, as in the fused rollout), and a step is one 8-byte message the host writes into a
 * mapped pinned word: a poller CTA reads it over PCIe and republishes it in L2, every warp fetches its market's 20*A-byte action record
 * from the caller's pinned block, steps, stores snapshot + result record into cell m of plane `slot` (same layout and bytes as
 * cda_step_planes) and counts itself; the last warp rings a pinned completion word the call spins on.  Results are identical to
 * cda_step_planes' (tests/test_gpu_serve.py compares them step by step, and the state afterwards).
 *   cda_serve_bind   registers the plane ring (pinned + mapped, as above).  Returns CDA_EUNSUPPORTED when the mode cannot be used and
 *                    the caller should stay on cda_step_planes: decimal_ledger handles, more markets than one resident wave holds
 *                    (7 CTAs x 4 markets per SM: 4140 markets of <= 4 agents on a B200), buffers that are not mapped.
 *   cda_serve_step   h_action_block: pinned i32[M][5][A] (market-major) for THIS step — any block within +-4 GB of the first one
 *                    seen; slot: plane that receives the outputs (the caller advances it modulo slots).  Returns when the outputs
 *                    are in host memory.  `stream`: the caller's stream; a (re)launch is ordered behind the work already queued there.
 *   cda_serve_stop   stores the state back and retires the kernel.  Every other entry point that touches the handle's state
 *                    (reset, step*, rollout, info, fills, dump, save / load) does this implicitly, so the modes can be mixed freely.
 * The kernel also retires by itself when no step has been requested for the lease (2 ms; $CDA_SERVE_LEASE_US) — a host that went away
 * cannot pin the GPU — and the next cda_serve_step launches it again: an idle server costs one launch, not a hang.  While it is resident
 * it occupies every SM: other kernels on the device wait for the lease to run out, so this mode is for host-side policies (GPU-side
 * policies use cda_step / cda_step_gather, which never leave the device).  A server that had to be relaunched for 48 of the last 64 steps
 * (the host needs longer than the lease between steps, or other work keeps claiming the GPU) switches itself off: cda_serve_step returns
 * CDA_EUNSUPPORTED without having stepped, and the caller continues with cda_step_planes (cda_serve_bind switches it on again). */
#define CDA_EUNSUPPORTED (-5)   /* the requested mode cannot serve this handle; use the fallback the header names */
int cda_serve_bind(CdaEnv *env, float *h_planes, int32_t slots, int32_t cell_words);
int cda_serve_step(CdaEnv *env, const int32_t *h_action_block, int32_t slot, void *stream);
int cda_serve_stop(CdaEnv *env);
int64_t cda_serve_launches(const CdaEnv *env);   /* how many times the resident kernel has been (re)launched so far */

/* Fused T-step rollout with the on-device uniform random policy (the RandomRLModule /
 * CDA_rand.py workload: category U{0..8}, price U{0..9}, offset U{0..2}, mean U(-1,1), sigma U(0,1),
 * gym_continuousDoubleAuction/train/model/model_handler.py:38-78).  Policy draws come from a
 * separate counter-based generator keyed by (policy_seed, market, step, agent) so env streams
 * stay numpy-exact.  Outputs hold the LAST step.  d_obs etc. may be NULL. */
int cda_rollout_random(CdaEnv *env, int32_t num_steps, uint64_t policy_seed, float *d_obs, double *d_reward,
                       uint8_t *d_terminated, uint8_t *d_truncated, void *stream);

/* ---- fused step + all-gather over NVLink peer memory (one policy batch spanning G GPUs) ----
 * Each rank owns M markets (global rows [rank*M, (rank+1)*M)).  cda_gather_create allocates this rank's GATHER WINDOW: G*M rows of
 *     f32 snapshot[CDA_GATHER_SLOTS][42], then two result records { double reward[A]; uint8_t terminated, truncated; pad to 8 }
 * (row = cda_gather_row_words() 4-byte words), followed by u32 flags[64]; it returns the window's 64-byte CUDA IPC handle.  After the
 * ranks exchange handles (any transport), cda_gather_connect maps every peer's window and cda_gather_publish sends every local
 * market's current stack to slots 0..n_hist-1 of its row in EVERY rank's window (call it after a reset, on all ranks).
 * cda_step_gather is then cda_step whose epilogue stores, for every local market, the newest 42-float snapshot into the next slot of
 * that market's row, and the result record into record slot (step parity), in ALL G windows — plain stores to peer-mapped addresses
 * that travel over NVLink / NVSwitch while other warps are still matching.  Of the 168-float observation only those 42 floats are new
 * each step, so this moves 3.5x fewer bytes than an all-gather of the stacked observations (what NCCL would be given), and the
 * stacked observation of global row r stays ONE contiguous run of its window row:
 *     obs(r)    = row(r).snapshot[pos-n_hist+1 .. pos][0..41]     (pos = cda_gather_pos(); a strided [G*M, 168] view, oldest first)
 *     record(r) = row(r).record[cda_gather_record_parity()]
 * When the row is full the whole stack is re-sent into slots 0..n_hist-1 (once per CDA_GATHER_SLOTS - n_hist + 1 steps).
 * Ordering is fused too: every warp fences its peer stores and counts itself; the last one writes this rank's step number into every
 * rank's flag array.  cda_gather_wait enqueues a one-warp kernel on `stream` that returns once all G ranks have published the current
 * step — no NCCL call anywhere on the path.  A rank may run at most one step ahead of its peers' consumers: the step it writes lands
 * in a different snapshot slot and the other record slot than the ones being read, and it cannot start a second step before every
 * peer has published the first (the flag wait): no second, consumer-done barrier.  Needs CDA_GATHER_SLOTS >= 2*n_hist + 2. */
#define CDA_MAX_PEERS 8
#define CDA_GATHER_SLOTS 32
int cda_gather_create(CdaEnv *env, int32_t world, int32_t rank, void *ipc_handle_out64, void **d_local_buf, uint64_t *bytes);
int cda_gather_connect(CdaEnv *env, const void *all_ipc_handles /* [world][64] */);
int cda_gather_publish(CdaEnv *env, void *stream);
int cda_step_gather(CdaEnv *env, const int32_t *d_category, const float *d_size_mean, const float *d_size_sigma,
                    const int32_t *d_price, const int32_t *d_price_offset, void *stream);
int cda_gather_wait(CdaEnv *env, void *stream);
int32_t cda_gather_pos(const CdaEnv *env);
int32_t cda_gather_row_words(const CdaEnv *env);
int32_t cda_gather_record_parity(const CdaEnv *env);

/* Lazy info (info_helper.py:30-116): gathers one field for all markets into d_out. */
int cda_get_info(CdaEnv *env, int32_t field, int64_t *d_out, void *stream);

/* All per-agent fields at once: d_out i64[CDA_INFO_MARKET][M][A] followed by the market block
 * i64[M][8] (one kernel launch; what the dict adapter uses to build the reference's info dict). */
int cda_get_info_all(CdaEnv *env, int64_t *d_out, void *stream);

/* Fill log of the last step (the reference's per-step `seq_trades`, action_helper.py:201-239):
 * d_fills i32[M][fill_capacity][8], d_counts i32[M].  Requires fill_capacity > 0.  With CdaConfig.fill_tape the same buffer is a tape:
 * the last fill_capacity fills of every market across steps (and across the steps of a fused rollout), d_counts = fills since the reset. */
int cda_get_fills(CdaEnv *env, int32_t *d_fills, int32_t *d_counts, void *stream);

/* Decoded-action log (the reference's `LOB_actions`, continuousDoubleAuction_env.py:285: what Action_Helper.set_actions made of the
 * model's actions).  d_log i32[M][A][4] (device; NULL switches the log off): every later step stores, per agent, order type
 * (0 market, 1 limit, 2 modify, 3 cancel), side (0 bid, 1 ask; -1 = pass or absent, other fields then meaningless), size and price. */
int cda_set_action_log(CdaEnv *env, int32_t *d_log);

/* Canonical dump of one market for bit-exact checks (host buffers; synchronises).
 *   book rows (priority order: bids price desc / asks price asc, FIFO inside a level):
 *     price, qty, trader, order_id, timestamp           -> h_bids/h_asks i64[max_rows][5]
 *   map rows: order ids in the reference's order_map iteration order -> i64[max_rows]
 *   h_counts[2] receive the number of orders per side. */
int cda_dump_market(CdaEnv *env, int32_t market, int64_t *h_bids, int64_t *h_asks, int64_t *h_bids_map,
                    int64_t *h_asks_map, int32_t max_rows, int32_t *h_counts, uint64_t *h_rng6);

/* decimal_ledger: bring every agent's Decimal twin up to date (replays the event journals; the library does this by itself every few
 * steps and on the spot whenever a decision needs it) and, optionally, copy the twins out: d_out i64[M][A][8] =
 * { vwap coefficient lo, hi, exponent, sign | position << 32, cash coefficient lo, hi, exponent, sign | tracked << 40 } where a value is
 * (-1)^sign * (hi * 2^64 + lo) * 10^exponent and `cash` is only meaningful while tracked (otherwise Decimal cash == integer cash). */
int cda_twin_sync(CdaEnv *env, int64_t *d_out /* may be NULL */, void *stream);

/* Device-state checkpoint (the reference has none). */
size_t cda_state_bytes(const CdaEnv *env);
int cda_save_state(CdaEnv *env, void *h_dst, void *stream);
/* Layout of one market's block inside the checkpoint (bytes): out = { stride, off_accounts, off_snapshot_ring, off_order_pool,
 * order_capacity, num_agents, n_hist, header_bytes, decimal_ledger, off_decimal_twins, off_event_journals, journal_entries }.  Market m starts at m * stride.  Header words (u32): 0 time, 1 next_order_id,
 * 2 insertion counter, 3 t_step, 4 last_price, 5 flags (bit 0: tape non-empty), 6 done_mask, 7 status, 8 n_bid, 9 n_ask,
 * 10 rng.has_uint32, 11 rng.uinteger, 12..19 PCG64 state_hi, state_lo, inc_hi, inc_lo (u64 each), 20..39 raw top-K prices, 40 best_bid,
 * 41 best_ask.  Accounts: cash, hold, cost_basis, nav, prev_nav, max_nav (i64[A] each), position (i32[A]), num_trades (u32[A]),
 * step counters (u32[A]: trades[0:12) passive[12:24) placed[24] rejected[25] is_pass[26]), Decimal-twin flags (u32[A]).  Order pool: u32[2 sides][cap/32][5][32],
 * fields trader<<24|price, qty, order_id, timestamp, insertion seq; live orders are entries 0..n-1 of their side, unsorted. */
int cda_state_layout(const CdaEnv *env, int32_t out[12]);
int cda_load_state(CdaEnv *env, const void *h_src, void *stream);

/* Introspection */
int32_t cda_num_markets(const CdaEnv *env);
int32_t cda_record_bytes(const CdaEnv *env);   /* size of one packed result record of cda_step_host_window */
int32_t cda_obs_dim(const CdaEnv *env);
int32_t cda_order_capacity(const CdaEnv *env);
int64_t cda_kernel_launches(const CdaEnv *env);   /* kernels launched by this handle so far */

/* Sticky-status early warning (the reference calls sys.exit() where this library sets a per-market status bit): a pinned host
 * word that every step kernel sets to 1 when ANY market ends the step with a non-zero status.  Reading it costs nothing (no
 * launch, no synchronisation; it reflects the steps that have completed so far), so a host loop can poll it after every step
 * and only then pay for the precise per-market gather (cda_get_info(CDA_INFO_MARKET), column 7).  cda_status_flag_clear re-arms it. */
const volatile uint32_t *cda_status_flag(const CdaEnv *env);
int cda_status_flag_clear(CdaEnv *env);

const char *cda_strerror(int code);
const char *cda_last_cuda_error(void);
const char *cda_build_info(void);

/* Host-side helper: the PCG64 state numpy gives for `seed` (state_hi, state_lo, inc_hi, inc_lo). */
void cda_seed_to_pcg64(uint64_t seed, uint64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* CDA_B200_H */
