// cda_b200.cu — C-ABI (include/cda_b200.h) over the sm_100a kernels in cda_kernels.cuh.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -shared ...
// (-fmad=false: the reference computes loc + scale*z, the reward sum and the ziggurat wedge
//  test with separately rounded multiplies and adds; contracting them into FMAs would change
//  results in the last bit.)
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include <time.h>

#include "cda_kernels.cuh"
#include "cda_b200_testing.h"
#include "cda_dec128.cuh"

#ifndef CDA_WARPS_PER_CTA
#define CDA_WARPS_PER_CTA 4
#endif

static thread_local char g_cuda_err[256] = "";

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            snprintf(g_cuda_err, sizeof(g_cuda_err), "%s failed: %s", #expr, cudaGetErrorString(e__)); \
            return CDA_ECUDA;                                                                   \
        }                                                                                       \
    } while (0)

// Every entry point that launches or copies runs on the handle's device and leaves the caller's current device as it was.
struct DevGuard {
    int prev = -1; bool switched = false;
    explicit DevGuard(int dev) { if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess; }
    ~DevGuard() { if (switched) cudaSetDevice(prev); }
};

struct CdaEnv {
    CdaConfig cfg;
    CdaDevCfg dev;
    int M, device;
    unsigned char *state;      // M * stride bytes
    size_t state_bytes;
    int *fills; int *fill_counts;
    int *act_log;              // cda_set_action_log: decoded-action log of every step (caller's device buffer), or NULL
    // device staging for cda_step_host
    int *s_cat; float *s_mean; float *s_sigma; int *s_pcode; int *s_poff;
    float *s_obs; double *s_reward; unsigned char *s_term; unsigned char *s_trunc;
    unsigned char *s_rec;      // packed result records [M][A*8+8] (window host path)
    float *s_plane;            // device staging of one output plane (cda_step_planes fallback when the host plane is not mapped)
    bool was_reset;
    // fused all-gather
    int g_world, g_rank; unsigned char *g_local; size_t g_bytes; unsigned char *g_peer[CDA_MAX_PEERS]; bool g_connected;
    int zerocopy;              // cda_step_host: let the kernel store outputs straight into mapped pinned host memory
    const void *zc_host; void *zc_dev;   // last host obs pointer checked and its device alias (NULL = not mapped)
    const void *zr_host; void *zr_dev;   // same for the window path's record array
    float *w_window; int w_slots; void *w_records; void *w_stream;   // cda_window_bind
    int zerocopy_in; const void *zi_host; void *zi_dev;
    int act_tma;               // CDA_ACT_TMA (default 1): stage the action rows with cp.async.bulk
    double zc_fraction;        // share of the obs rows the kernel writes straight to host memory (the rest is DMA'd)   // same for the action block (kernel reads pinned host memory)
    long long launches;
    unsigned *done_ctr; unsigned done_seq; int doorbell;   // completion doorbell (status_host[8] is the host word the kernel rings): CDA_DOORBELL=0 disables
    unsigned *status_host, *status_dev;   // one mapped pinned word: any step kernel that ends with a non-zero sticky market status stores 1 here
    int twin_steps, twin_every;            // decimal_ledger: steps since the last journal flush; flush cadence (CDA_TWIN_FLUSH_STEPS, $CDA_TWIN_FLUSH overrides: measurements)
    int g_pos; unsigned g_seq;             // fused all-gather: slot of the newest snapshot in the gather windows; steps published so far
    size_t smem_bytes;
    // resident step server (cda_serve_*): status_host word 16/17 = the message word the host rings, word 32 = watchdog flag
    cudaStream_t srv_stream; cudaEvent_t srv_event;
    unsigned long long *srv_go_dev; unsigned *srv_done_dev;
    float *srv_planes_host, *srv_planes_dev; int srv_slots, srv_cell;
    const char *srv_act_host0; const unsigned char *srv_act_dev0;     // first action block seen: the base the messages' offsets count from
    const void *srv_ptr_cache_h[64]; const unsigned char *srv_ptr_cache_d[64];   // host pointer -> device alias of recent action blocks
    bool srv_bound, srv_running; unsigned srv_seq; long long srv_launches;
    int srv_win_steps, srv_win_launches;   // relaunches over the current window of 64 served steps (a server that is relaunched for most steps is no server)
    unsigned long long srv_lease_ns, srv_last_ns;   // lease; host clock (CLOCK_MONOTONIC) at the end of the last served step
    int host_ctas, dev_ctas;   // resident CTAs per SM for the host paths / the device path (0 = as many as fit)
};

static unsigned align_up(unsigned v, unsigned a) { return (v + a - 1) / a * a; }

// ctas_per_sm > 0 caps the CTAs resident on one SM by padding the dynamic shared memory request (228 KB per SM, 1 KB
// reserved per CTA).  Used by the host paths: with fewer resident warps the grid runs as a stream of short CTAs instead
// of ONE wave in which every market finishes at the same moment, so the outputs of early markets cross PCIe while later
// markets are still being matched (and the action reads of later CTAs overlap the matching of earlier ones).
template <int CAP, bool DEC>
static cudaError_t launch_step(const CdaEnv *e, const CdaStepParams &p, cudaStream_t st, int ctas_per_sm) {
    const int grid = (e->M + CDA_WARPS_PER_CTA - 1) / CDA_WARPS_PER_CTA;
    // per-warp tiles, then the CTA's action tile u32[5][WARPS][A] and its mbarrier
    //     then one 16-B aligned account tile per warp
    size_t smem = (size_t)CdaSmemLayout<CAP, DEC>::BYTES * CDA_WARPS_PER_CTA + (size_t)5 * CDA_WARPS_PER_CTA * e->dev.A * 4 + 16 +
                  (size_t)CDA_WARPS_PER_CTA * (CDA_ACCT_TILE_WORDS(DEC, e->dev.A) * 4);
    if (ctas_per_sm > 0) {
        const size_t per_sm = 233472, pad = per_sm / (size_t)(ctas_per_sm + 1) - 1024 + 256;   // ctas_per_sm + 1 CTAs no longer fit
        if (pad > smem && pad <= 232448 - 1024) smem = pad;
    }
    static size_t attr_set[16] = {0};
    if (attr_set[e->device & 15] < smem) {
        auto set = [&](const void *f) {
            cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        };
        set((const void *)cda_step_kernel<CAP, CDA_WARPS_PER_CTA, false, false, DEC>);
        set((const void *)cda_step_kernel<CAP, CDA_WARPS_PER_CTA, false, true, DEC>);
        set((const void *)cda_step_kernel<CAP, CDA_WARPS_PER_CTA, true, false, DEC>);
        attr_set[e->device & 15] = smem;
    }
    // routed outputs (host window / ring / planes, packed or strided records, split rows, fused all-gather, completion doorbell) take the full
    // kernel; the plain device step and the fused rollout take the body without that code
    const bool routed = p.rep_n > 1 || p.done_flag || p.ring_out || p.rec_inline || p.flag_pack || p.obs_hi || p.obs_split != e->M ||
                        p.obs_stride != e->dev.W || p.reward_stride != e->dev.A || p.flag_stride != 1;
    if (p.num_steps > 0) {
        if (routed) return cudaErrorInvalidValue;   // the rollout writes dense device arrays only
        cda_step_kernel<CAP, CDA_WARPS_PER_CTA, true, false, DEC><<<grid, CDA_WARPS_PER_CTA * 32, smem, st>>>(p);
    } else if (routed) cda_step_kernel<CAP, CDA_WARPS_PER_CTA, false, true, DEC><<<grid, CDA_WARPS_PER_CTA * 32, smem, st>>>(p);
    else cda_step_kernel<CAP, CDA_WARPS_PER_CTA, false, false, DEC><<<grid, CDA_WARPS_PER_CTA * 32, smem, st>>>(p);
    return cudaGetLastError();
}
template <bool DEC>
static cudaError_t launch_step_cap(const CdaEnv *e, const CdaStepParams &p, cudaStream_t st, int ctas_per_sm) {
    switch (e->dev.cap) {
        case 64: return launch_step<64, DEC>(e, p, st, ctas_per_sm);
        case 128: return launch_step<128, DEC>(e, p, st, ctas_per_sm);
        case 160: return launch_step<160, DEC>(e, p, st, ctas_per_sm);
        case 192: return launch_step<192, DEC>(e, p, st, ctas_per_sm);
        default: return launch_step<256, DEC>(e, p, st, ctas_per_sm);
    }
}
static cudaError_t launch_step_any(const CdaEnv *e, const CdaStepParams &p, cudaStream_t st, int ctas_per_sm) {
    return e->dev.dec ? launch_step_cap<true>(e, p, st, ctas_per_sm) : launch_step_cap<false>(e, p, st, ctas_per_sm);
}


// ---- resident step server: host side (device side: "Resident step server" in cda_kernels.cuh; contract: include/cda_b200.h) ----------
#define CDA_SRV_GO_WORD 16      /* status_host word index of the 64-bit message the host rings (its own 64-B line) */
#define CDA_SRV_DONE_WORD 10    /* completion word the kernel's last warp writes (NOT word 8: the launch paths' doorbell counts differently) */
#define CDA_SRV_ERR_WORD 32     /* a worker's watchdog fired */
static unsigned long long *g_srv_prof = nullptr;   // cda_debug_serve_timeline: u64[(M + 1)][16] milestone times of the last served step
static int g_srv_dbg = getenv("CDA_SERVE_DEBUG") ? atoi(getenv("CDA_SERVE_DEBUG")) : 0;
static unsigned long long host_now_ns() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (unsigned long long)ts.tv_sec * 1000000000ULL + (unsigned long long)ts.tv_nsec; }
static size_t srv_smem_bytes(const CdaEnv *e, int cap_words_bytes) {
    return (size_t)cap_words_bytes * CDA_WARPS_PER_CTA + (size_t)5 * CDA_WARPS_PER_CTA * e->dev.A * 4 + 16 +
           (size_t)CDA_WARPS_PER_CTA * (CDA_ACCT_TILE_WORDS(false, e->dev.A) * 4);
}
// check_only: can every CTA of the grid (one poller + one per four markets) be resident at once?  cudaErrorLaunchOutOfResources if not.
template <int CAP>
static cudaError_t launch_serve(CdaEnv *e, const CdaStepParams &p, bool check_only) {
    auto kern = cda_step_kernel<CAP, CDA_WARPS_PER_CTA, true, true, false>;
    const int grid = (e->M + CDA_WARPS_PER_CTA - 1) / CDA_WARPS_PER_CTA + 1;
    const size_t smem = srv_smem_bytes(e, CdaSmemLayout<CAP, false>::BYTES);
    cudaError_t err = cudaSuccess;
    static size_t attr_set[16] = {0};
    if (attr_set[e->device & 15] < smem) {
        err = cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err == cudaSuccess) err = cudaFuncSetAttribute((const void *)kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
        attr_set[e->device & 15] = smem;
    }
    if (check_only) {
        int per_sm = 0, sms = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)kern, CDA_WARPS_PER_CTA * 32, smem);
        if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
        if (err != cudaSuccess) return err;
        return (long long)per_sm * sms >= grid ? cudaSuccess : cudaErrorLaunchOutOfResources;
    }
    kern<<<grid, CDA_WARPS_PER_CTA * 32, smem, e->srv_stream>>>(p);
    return cudaGetLastError();
}
static cudaError_t launch_serve_cap(CdaEnv *e, const CdaStepParams &p, bool check_only) {
    switch (e->dev.cap) {
        case 64: return launch_serve<64>(e, p, check_only);
        case 128: return launch_serve<128>(e, p, check_only);
        case 160: return launch_serve<160>(e, p, check_only);
        case 192: return launch_serve<192>(e, p, check_only);
        default: return launch_serve<256>(e, p, check_only);
    }
}
// (re)launch the resident kernel for step srv_seq + 1, behind whatever the caller has queued on `st`
static int srv_launch(CdaEnv *e, cudaStream_t st) {
    CdaStepParams p;
    memset(&p, 0, sizeof(p));
    p.cfg = e->dev; p.state = e->state; p.M = e->M; p.status_flag = e->status_dev;
    p.obs_split = e->M; p.obs_stride = e->dev.W; p.reward_stride = e->dev.A; p.flag_stride = 1;
    p.act_mstride = 5 * e->dev.A; p.act_packed = 1; p.num_steps = 1;
    p.ring_out = e->srv_planes_dev; p.ring_stride = e->srv_cell; p.ring_pad = e->srv_cell; p.rec_inline = 1;
    p.fills = e->fills; p.fill_counts = e->fill_counts; p.act_log = e->act_log;
    p.acct_tma = (e->dev.A % 4) == 0;
    p.done_ctr = e->done_ctr; p.done_flag = e->status_dev + CDA_SRV_DONE_WORD;
    p.srv_go_host = reinterpret_cast<const unsigned long long *>(e->status_dev + CDA_SRV_GO_WORD);
    p.srv_go_dev = e->srv_go_dev; p.srv_done_dev = e->srv_done_dev; p.srv_err = e->status_dev + CDA_SRV_ERR_WORD;
    p.srv_act_base = e->srv_act_dev0; p.srv_next = e->srv_seq + 1u;
    p.srv_lease_ns = e->srv_lease_ns; p.srv_watchdog_ns = 1000000000ULL + 4ULL * e->srv_lease_ns;
    static const int dbg_mode = getenv("CDA_SERVE_ACT_MODE") ? atoi(getenv("CDA_SERVE_ACT_MODE")) : 0;
    p.srv_act_mode = ((e->dev.A % 4) == 0 && dbg_mode == 0) ? 0 : 1;
    p.prof = g_srv_prof;
    if (g_srv_dbg & 1) {   // timing experiments (tools/serve_timeline.py): actions read from a device copy of the first block (no input transfer)
        CUDA_TRY(cudaMemcpyAsync(e->s_cat, e->srv_act_dev0, (size_t)e->M * e->dev.A * 20, cudaMemcpyDefault, e->srv_stream));
        p.srv_act_base = reinterpret_cast<const unsigned char *>(e->s_cat);
    }
    if (g_srv_dbg & 2) {   // ... outputs kept on the device (no output transfer)
        static float *dbg_planes = nullptr;   // (one env per process in the tool)
        if (!dbg_planes) CUDA_TRY(cudaMalloc(&dbg_planes, (size_t)e->srv_slots * e->M * e->srv_cell * 4));
        p.ring_out = dbg_planes;
    }   // bulk copies move multiples of 16 B from 16-B aligned addresses
    CUDA_TRY(cudaEventRecord(e->srv_event, st));
    CUDA_TRY(cudaStreamWaitEvent(e->srv_stream, e->srv_event, 0));
    CUDA_TRY(cudaMemsetAsync(e->srv_go_dev, 0, (size_t)CDA_SRV_COPIES * 128, e->srv_stream));   // (a message left by an earlier launch must not match)
    CUDA_TRY(launch_serve_cap(e, p, false));
    e->launches++; e->srv_launches++;
    e->srv_running = true;
    return CDA_OK;
}
// retire the resident kernel (state back in HBM) before anything else touches the handle's state
static int srv_quiesce(CdaEnv *e) {
    if (!e->srv_running) return CDA_OK;
    __atomic_store_n(reinterpret_cast<unsigned long long *>(e->status_host + CDA_SRV_GO_WORD),
                     (unsigned long long)cda_srv_seq24(e->srv_seq + 1u) | ((unsigned long long)CDA_SRV_STOP << 24), __ATOMIC_RELEASE);
    e->srv_running = false;
    CUDA_TRY(cudaStreamSynchronize(e->srv_stream));
    if (e->status_host[CDA_SRV_ERR_WORD]) { snprintf(g_cuda_err, sizeof(g_cuda_err), "resident step server: worker watchdog fired"); return CDA_ECUDA; }
    return CDA_OK;
}
#define SRV_QUIESCE(e) do { if ((e)->srv_running) { DevGuard g__((e)->device); int rc__ = srv_quiesce(e); if (rc__) return rc__; } } while (0)

extern "C" {

const char *cda_strerror(int code) {
    switch (code) {
        case CDA_OK: return "ok";
        case CDA_EINVAL: return "invalid argument or configuration";
        case CDA_ECUDA: return "CUDA runtime error (see cda_last_cuda_error)";
        case CDA_ENOMEM: return "out of memory";
        case CDA_ESTATE: return "environment used before reset";
        case CDA_EUNSUPPORTED: return "mode not available for this handle (see the header for the fallback)";
        default: return "unknown error";
    }
}
const char *cda_last_cuda_error(void) { return g_cuda_err; }
const char *cda_build_info(void) {
    return "cda_b200 sm_100a; warp-per-market fused step; cp.async.bulk staged order pool; nvcc " __DATE__;
}
void cda_seed_to_pcg64(uint64_t seed, uint64_t out[4]) {
    unsigned long long s[4];
    seedseq_words(seed, s);
    // same arithmetic as rng_seed, on the host with 128-bit integers
    unsigned __int128 mult = (((unsigned __int128)2549297995355413924ULL) << 64) | 4865540595714422341ULL;
    unsigned __int128 initstate = (((unsigned __int128)s[0]) << 64) | s[1];
    unsigned __int128 initseq = (((unsigned __int128)s[2]) << 64) | s[3];
    unsigned __int128 inc = (initseq << 1) | 1u, state = 0;
    state = state * mult + inc;
    state += initstate;
    state = state * mult + inc;
    out[0] = (uint64_t)(state >> 64); out[1] = (uint64_t)state; out[2] = (uint64_t)(inc >> 64); out[3] = (uint64_t)inc;
}

int cda_create(const CdaConfig *cfg, int32_t num_markets, int32_t device, CdaEnv **out) {
    if (!cfg || !out || num_markets < 1) return CDA_EINVAL;
    if (cfg->num_agents < 1 || cfg->num_agents > CDA_MAX_AGENTS) return CDA_EINVAL;
    if (cfg->n_hist < 1 || cfg->n_hist > CDA_MAX_HIST) return CDA_EINVAL;
    if (cfg->tick_size < 1 || cfg->init_cash <= 0 || cfg->max_step < 1) return CDA_EINVAL;
    if (cfg->initial_price_min < 1 || cfg->initial_price_max < cfg->initial_price_min) return CDA_EINVAL;
    if (cfg->min_size < 0 || cfg->mkt_max_size < 1 || cfg->limit_size_multiple < 1) return CDA_EINVAL;
    int cap = cfg->order_capacity;
    if (cap == 0) cap = cfg->num_agents <= 8 ? 160 : 256;
    if (cap != 64 && cap != 128 && cap != 160 && cap != 192 && cap != 256) return CDA_EINVAL;
    if (cfg->fill_capacity < 0 || cfg->fill_capacity > 1024) return CDA_EINVAL;
    {   // the cold-path kernels (reset window / ring fill, info gather) index with 32-bit ints: one handle takes at most
        // INT_MAX / max(2 * n_hist * 42, 15 * A) markets (6.39 M with the defaults, 48 GB of state); shard beyond that
        const long long per = std::max<long long>(2LL * cfg->n_hist * CDA_SNAPSHOT_DIM, 15LL * cfg->num_agents);
        if ((long long)num_markets * per > 2147483647LL) return CDA_EINVAL;
    }
    DevGuard guard(device);
    { int cur = -1; if (cudaGetDevice(&cur) != cudaSuccess || cur != device) { snprintf(g_cuda_err, sizeof(g_cuda_err), "cudaSetDevice(%d) failed", device); cudaGetLastError(); return CDA_ECUDA; } }
    CdaEnv *e = new (std::nothrow) CdaEnv();
    if (!e) return CDA_ENOMEM;
    memset(e, 0, sizeof(*e));
    e->cfg = *cfg; e->cfg.order_capacity = cap;
    e->M = num_markets; e->device = device;
    CdaDevCfg &d = e->dev;
    d.A = cfg->num_agents; d.n_hist = cfg->n_hist; d.max_step = cfg->max_step; d.tick = cfg->tick_size;
    d.init_cash = cfg->init_cash; d.min_size = cfg->min_size;
    // action_helper.py:46-47 — python floats, then weak-scalar cast to f32 when multiplied with the f32 mean
    d.mkt_mul = (float)((cfg->mkt_max_size - cfg->min_size) / 2.0);
    d.lim_mul = (float)(((double)cfg->mkt_max_size * cfg->limit_size_multiple - cfg->min_size) / 2.0);
    d.price_lo = cfg->initial_price_min; d.price_hi = cfg->initial_price_max;
    d.cap = cap; d.fill_cap = cfg->fill_capacity; d.fill_tape = cfg->fill_tape && cfg->fill_capacity > 0 ? 1 : 0;
    d.c_order = cfg->order_penalty; d.c_trade = cfg->trade_penalty; d.c_dd = cfg->drawdown_penalty;
    d.c_passive = cfg->passive_bonus; d.c_loss = cfg->loss_multiplier;
    d.W = cfg->n_hist * CDA_SNAPSHOT_DIM;
    d.off_acct = CDA_HDR_BYTES;
    d.off_hist = align_up(d.off_acct + (unsigned)d.A * 64u, 16);
    d.off_pool = align_up(d.off_hist + (unsigned)d.W * 4u, 16);
    unsigned end = d.off_pool + 2u * CDA_POOL_FIELDS * (unsigned)cap * 4u;
    d.dec = cfg->decimal_ledger ? 1 : 0;
    if (d.dec) {   // Decimal twins + event journals behind the pool (cda_twin.cuh): 64 + 8 * CDA_JRN_E bytes per agent
        d.off_twin = align_up(end, 64);
        d.off_jrn = d.off_twin + (unsigned)d.A * CDA_TWIN_BYTES;
        end = d.off_jrn + (unsigned)d.A * CDA_JRN_E * 8u;
    }
    d.stride = align_up(end, 128);
    e->state_bytes = (size_t)d.stride * (size_t)num_markets;
    const size_t MA = (size_t)num_markets * d.A;
    cudaError_t err = cudaMalloc(&e->state, e->state_bytes);
    if (err == cudaSuccess && d.fill_cap > 0) {
        err = cudaMalloc(&e->fills, (size_t)num_markets * d.fill_cap * CDA_FILL_WORDS * sizeof(int));
        if (err == cudaSuccess) err = cudaMalloc(&e->fill_counts, (size_t)num_markets * sizeof(int));
    }
    if (err == cudaSuccess) err = cudaMalloc(&e->s_cat, MA * 4 * 5);
    // one contiguous output staging block: obs | reward | terminated | truncated  (single D2H when the
    // caller's host buffers are laid out the same way)
    if (err == cudaSuccess) err = cudaMalloc(&e->s_obs, (size_t)num_markets * d.W * 4 + MA * 8 + (size_t)num_markets * 2);
    if (err == cudaSuccess) err = cudaMalloc(&e->s_rec, (size_t)num_markets * (((size_t)d.A * 8 + 8 + 63) / 64 * 64));
    if (err == cudaSuccess) err = cudaHostAlloc(reinterpret_cast<void **>(&e->status_host), 256, cudaHostAllocMapped | cudaHostAllocPortable);
    if (err == cudaSuccess) { memset(e->status_host, 0, 256); err = cudaHostGetDevicePointer(reinterpret_cast<void **>(&e->status_dev), e->status_host, 0); }
    if (err == cudaSuccess) err = cudaMalloc(&e->done_ctr, 64);
    if (err == cudaSuccess) err = cudaMemset(e->done_ctr, 0, 64);
    if (err != cudaSuccess) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "cudaMalloc failed: %s", cudaGetErrorString(err));
        cda_destroy(e);
        return err == cudaErrorMemoryAllocation ? CDA_ENOMEM : CDA_ECUDA;
    }
    e->s_mean = reinterpret_cast<float *>(e->s_cat + MA);
    e->s_sigma = reinterpret_cast<float *>(e->s_cat + 2 * MA);
    e->s_pcode = e->s_cat + 3 * MA;
    e->s_poff = e->s_cat + 4 * MA;
    e->s_reward = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(e->s_obs) + (size_t)num_markets * d.W * 4);
    e->s_term = reinterpret_cast<unsigned char *>(e->s_reward) + MA * 8;
    e->s_trunc = e->s_term + num_markets;
    CUDA_TRY(cudaMemset(e->state, 0, e->state_bytes));
    {   // PCG64 jump-ahead table (see rng_jump)
        unsigned long long tab[CDA_MAX_AGENTS + 1][4];
        const unsigned __int128 mult = (((unsigned __int128)2549297995355413924ULL) << 64) | 4865540595714422341ULL;
        unsigned __int128 a = 1, g = 0;
        for (int r = 0; r <= CDA_MAX_AGENTS; ++r) {
            tab[r][0] = (unsigned long long)(a >> 64); tab[r][1] = (unsigned long long)a;
            tab[r][2] = (unsigned long long)(g >> 64); tab[r][3] = (unsigned long long)g;
            g = g + a;       // G_{r+1} = G_r + A^r
            a = a * mult;
        }
        CUDA_TRY(cudaMemcpyToSymbol(cda_pcg_jump, tab, sizeof(tab)));
    }
    {
        const char *zc = getenv("CDA_ZEROCOPY");
        e->zerocopy = zc ? atoi(zc) : 1;
        const char *zi = getenv("CDA_ZEROCOPY_IN");
        e->zerocopy_in = zi ? atoi(zi) : 1;
        const char *at = getenv("CDA_ACT_TMA");
        e->act_tma = at ? atoi(at) : 1;
        const char *hcs = getenv("CDA_HOST_CTAS"), *dcs = getenv("CDA_DEV_CTAS");
        e->host_ctas = hcs ? atoi(hcs) : 0;
        e->dev_ctas = dcs ? atoi(dcs) : 0;
        const char *db = getenv("CDA_DOORBELL");
        e->doorbell = db ? atoi(db) : 1;
        const char *tf = getenv("CDA_TWIN_FLUSH");
        e->twin_every = tf && atoi(tf) > 0 ? atoi(tf) : CDA_TWIN_FLUSH_STEPS;
        const char *zf = getenv("CDA_ZC_FRACTION");
        e->zc_fraction = zf ? atof(zf) : 0.25;   // SM stores to host reach ~25 GB/s but overlap the kernel; the copy engine does ~53 GB/s after it   // measured: kernel reading the pinned action block beats a separate H2D copy by ~10 us
    }
    *out = e;
    return CDA_OK;
}

int cda_destroy(CdaEnv *e) {
    if (!e) return CDA_OK;
    DevGuard guard(e->device);
    if (e->srv_running) srv_quiesce(e);
    if (e->srv_stream) { cudaStreamDestroy(e->srv_stream); cudaEventDestroy(e->srv_event); cudaFree(e->srv_go_dev); }
    if (e->status_host) cudaFreeHost(e->status_host);
    if (e->g_connected) for (int g = 0; g < e->g_world; ++g) if (g != e->g_rank && e->g_peer[g]) cudaIpcCloseMemHandle(e->g_peer[g]);
    cudaFree(e->g_local);
    cudaFree(e->state); cudaFree(e->fills); cudaFree(e->fill_counts);
    cudaFree(e->s_cat); cudaFree(e->s_obs); cudaFree(e->s_rec); cudaFree(e->s_plane); cudaFree(e->done_ctr);
    delete e;
    return CDA_OK;
}

int cda_reset(CdaEnv *e, const uint64_t *d_seeds, const uint8_t *d_mask, float *d_obs, void *stream) {
    if (!e) return CDA_EINVAL;
    if (!d_seeds && !e->was_reset) return CDA_ESTATE;   // reset(seed=None) needs an existing stream
    DevGuard guard(e->device);
    SRV_QUIESCE(e);
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 128, grid = (e->M + threads - 1) / threads;
    cda_reset_kernel<<<grid, threads, 0, st>>>(e->dev, e->state, e->M, (const unsigned long long *)d_seeds, d_mask, d_obs, e->fill_counts);
    CUDA_TRY(cudaGetLastError());
    e->launches++;
    if (!d_mask) e->was_reset = true;
    else e->was_reset = true;   // partial first reset: unselected markets keep zeroed (inert) state
    return CDA_OK;
}

// Host side of the completion doorbell: spin on the pinned word the kernel's last warp writes (results are fenced before it); if it
// does not ring within ~2 ms (a debugger, a preempted GPU) fall back to the stream synchronisation, which is always correct.
static int doorbell_wait(CdaEnv *e, unsigned seq, cudaStream_t st) {
    volatile unsigned *flag = e->status_host + 8;
    for (int spin = 0; spin < 400000; ++spin) {
        if (*flag == seq) return CDA_OK;
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return CDA_OK;
}
static int twin_flush(CdaEnv *e, cudaStream_t st) {
    const int n = e->M * e->dev.A, threads = 128;
    cda_twin_flush_kernel<<<(n + threads - 1) / threads, threads, 0, st>>>(e->dev, e->state, e->M, e->status_dev);
    CUDA_TRY(cudaGetLastError());
    e->launches++;
    e->twin_steps = 0;
    return CDA_OK;
}
static unsigned long long *g_prof = nullptr;
static int step_common(CdaEnv *e, CdaStepParams &p, cudaStream_t st, bool host_path = false) {
    SRV_QUIESCE(e);   // (a resident step server holds the state in shared memory: retire it first)
    p.cfg = e->dev; p.state = e->state; p.M = e->M;
    p.status_flag = e->status_dev;
    if (!p.obs_hi) p.obs_split = e->M;   // no split: every row goes to p.obs
    if (!p.obs_stride) p.obs_stride = e->dev.W;
    if (!p.reward_stride) p.reward_stride = e->dev.A;
    if (!p.flag_stride) p.flag_stride = 1;
    if (!p.act_mstride) p.act_mstride = e->dev.A;
    if (!p.ring_stride) { p.ring_stride = 2 * e->dev.n_hist * CDA_SNAPSHOT_DIM; p.ring_mirror = 1; }
    p.prof = g_prof;
    p.fills = e->fills; p.fill_counts = e->fill_counts; p.act_log = e->act_log;
    // TMA staging of the action rows: rows of A 4-byte words must be multiples of 16 B and the arrays 16-B aligned
    static const int dbg_acct_tma = getenv("CDA_ACCT_TMA") ? atoi(getenv("CDA_ACCT_TMA")) : 1;
    p.acct_tma = dbg_acct_tma && (e->dev.dec || (e->dev.A % 4) == 0);   // bulk copies move multiples of 16 B: 64*A with the twin flags, else 60*A
    p.act_tma = e->act_tma && p.num_steps == 0 && (e->dev.A % 4) == 0 &&
                (((uintptr_t)p.cat | (uintptr_t)p.mean | (uintptr_t)p.sigma | (uintptr_t)p.pcode | (uintptr_t)p.poff) & 15) == 0;
    CUDA_TRY(launch_step_any(e, p, st, host_path ? e->host_ctas : e->dev_ctas));
    e->launches++;
    if (e->dev.dec) {   // deferred Decimal twin: replay the event journals every CDA_TWIN_FLUSH_STEPS steps (one thread per agent, same stream)
        e->twin_steps += p.num_steps > 0 ? p.num_steps : 1;
        if (e->twin_steps >= e->twin_every) {
            int rc = twin_flush(e, st);
            if (rc) return rc;
        }
    }
    return CDA_OK;
}

int cda_step(CdaEnv *e, const int32_t *d_category, const float *d_size_mean, const float *d_size_sigma,
             const int32_t *d_price, const int32_t *d_price_offset, float *d_obs, double *d_reward,
             uint8_t *d_terminated, uint8_t *d_truncated, void *stream) {
    if (!e || !d_category || !d_size_mean || !d_size_sigma || !d_price || !d_price_offset) return CDA_EINVAL;
    DevGuard guard(e->device);
    if (!e->was_reset) return CDA_ESTATE;
    CdaStepParams p;
    memset(&p, 0, sizeof(p));
    p.cat = d_category; p.mean = d_size_mean; p.sigma = d_size_sigma; p.pcode = d_price; p.poff = d_price_offset;
    p.obs = d_obs; p.reward = d_reward; p.term = d_terminated; p.trunc = d_truncated;
    return step_common(e, p, (cudaStream_t)stream);
}

int cda_step_host(CdaEnv *e, const int32_t *h_category, const float *h_size_mean, const float *h_size_sigma,
                  const int32_t *h_price, const int32_t *h_price_offset, float *h_obs, double *h_reward,
                  uint8_t *h_terminated, uint8_t *h_truncated, void *stream) {
    if (!e || !h_category || !h_size_mean || !h_size_sigma || !h_price || !h_price_offset) return CDA_EINVAL;
    DevGuard guard(e->device);
    if (!e->was_reset) return CDA_ESTATE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t MA = (size_t)e->M * e->dev.A;
    // actions: one H2D when the five host arrays are one contiguous [5][M][A] block, else five
    const char *hc = reinterpret_cast<const char *>(h_category);
    const bool in_contig = reinterpret_cast<const char *>(h_size_mean) == hc + MA * 4 && reinterpret_cast<const char *>(h_size_sigma) == hc + 2 * MA * 4 &&
                           reinterpret_cast<const char *>(h_price) == hc + 3 * MA * 4 && reinterpret_cast<const char *>(h_price_offset) == hc + 4 * MA * 4;
    char *zi = nullptr;
    if (e->zerocopy_in && in_contig) {
        if (e->zi_host != h_category) {
            cudaPointerAttributes at;
            e->zi_host = h_category; e->zi_dev = nullptr;
            if (cudaPointerGetAttributes(&at, h_category) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) e->zi_dev = at.devicePointer;
            else cudaGetLastError();
        }
        zi = reinterpret_cast<char *>(e->zi_dev);
    }
    if (zi) {
        // the kernel reads this step's actions straight out of the caller's pinned block
    } else if (in_contig) {
        CUDA_TRY(cudaMemcpyAsync(e->s_cat, h_category, MA * 4 * 5, cudaMemcpyHostToDevice, st));
    } else {
        CUDA_TRY(cudaMemcpyAsync(e->s_cat, h_category, MA * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->s_mean, h_size_mean, MA * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->s_sigma, h_size_sigma, MA * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->s_pcode, h_price, MA * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->s_poff, h_price_offset, MA * 4, cudaMemcpyHostToDevice, st));
    }
    const size_t obs_bytes = (size_t)e->M * e->dev.W * 4;
    char *ho = reinterpret_cast<char *>(h_obs);
    const bool out_contig = h_obs && h_reward && h_terminated && h_truncated && reinterpret_cast<char *>(h_reward) == ho + obs_bytes &&
                            reinterpret_cast<char *>(h_terminated) == ho + obs_bytes + MA * 8 && reinterpret_cast<char *>(h_truncated) == ho + obs_bytes + MA * 8 + e->M;
    // Zero-copy outputs: when the caller's output block is pinned + mapped (UVA: every cudaHostAlloc /
    // torch pin_memory buffer is), the kernel stores obs/reward/flags straight into host memory, so the
    // PCIe transfer overlaps the step instead of following it.
    char *zc = nullptr;
    if (e->zerocopy && out_contig) {
        if (e->zc_host != h_obs) {
            cudaPointerAttributes at;
            e->zc_host = h_obs; e->zc_dev = nullptr;
            if (cudaPointerGetAttributes(&at, h_obs) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) e->zc_dev = at.devicePointer;
            else cudaGetLastError();
        }
        zc = reinterpret_cast<char *>(e->zc_dev);
    }
    CdaStepParams p;
    memset(&p, 0, sizeof(p));
    p.cat = e->s_cat; p.mean = e->s_mean; p.sigma = e->s_sigma; p.pcode = e->s_pcode; p.poff = e->s_poff;
    if (zi) {
        p.cat = reinterpret_cast<const int *>(zi); p.mean = reinterpret_cast<const float *>(zi + MA * 4);
        p.sigma = reinterpret_cast<const float *>(zi + 2 * MA * 4); p.pcode = reinterpret_cast<const int *>(zi + 3 * MA * 4);
        p.poff = reinterpret_cast<const int *>(zi + 4 * MA * 4);
    }
    if (zc) {
        // hybrid output: the first `split` rows are stored by the kernel straight into the pinned block (overlapping the
        // step), the remaining rows go to device staging and follow with ONE copy-engine transfer; reward/flags zero-copy
        int split = (int)(e->zc_fraction * e->M + 0.5);
        if (split < 0) split = 0;
        if (split > e->M) split = e->M;
        p.obs = reinterpret_cast<float *>(zc); p.reward = reinterpret_cast<double *>(zc + obs_bytes);
        p.term = reinterpret_cast<unsigned char *>(zc + obs_bytes + MA * 8); p.trunc = p.term + e->M;
        if (split < e->M) { p.obs_hi = e->s_obs; p.obs_split = split; }
        int rc2 = step_common(e, p, st, true);
        if (rc2) return rc2;
        if (split < e->M) {
            const size_t off = (size_t)split * e->dev.W * 4;
            CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<char *>(h_obs) + off, reinterpret_cast<char *>(e->s_obs) + off, obs_bytes - off, cudaMemcpyDeviceToHost, st));
        }
        return CDA_OK;
    }
    p.obs = e->s_obs; p.reward = e->s_reward; p.term = e->s_term; p.trunc = e->s_trunc;
    int rc = step_common(e, p, st, true);
    if (rc) return rc;
    if (out_contig) {
        CUDA_TRY(cudaMemcpyAsync(h_obs, e->s_obs, obs_bytes + MA * 8 + 2 * (size_t)e->M, cudaMemcpyDeviceToHost, st));
    } else {
        if (h_obs) CUDA_TRY(cudaMemcpyAsync(h_obs, e->s_obs, obs_bytes, cudaMemcpyDeviceToHost, st));
        if (h_reward) CUDA_TRY(cudaMemcpyAsync(h_reward, e->s_reward, MA * 8, cudaMemcpyDeviceToHost, st));
        if (h_terminated) CUDA_TRY(cudaMemcpyAsync(h_terminated, e->s_term, e->M, cudaMemcpyDeviceToHost, st));
        if (h_truncated) CUDA_TRY(cudaMemcpyAsync(h_truncated, e->s_trunc, e->M, cudaMemcpyDeviceToHost, st));
    }
    return CDA_OK;
}

static int g_dbg_window = getenv("CDA_DEBUG_WINDOW") ? atoi(getenv("CDA_DEBUG_WINDOW")) : 0;   // timing experiments (tools/): see step_window_impl
static void *mapped_alias(const void *h) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) return at.devicePointer;
    cudaGetLastError();
    return nullptr;
}

int cda_step_host_ring(CdaEnv *e, const int32_t *h_category, const float *h_size_mean, const float *h_size_sigma,
                       const int32_t *h_price, const int32_t *h_price_offset, float *h_ring, double *h_reward,
                       uint8_t *h_terminated, uint8_t *h_truncated, int64_t ring_pos, void *stream) {
    if (!e || !h_category || !h_size_mean || !h_size_sigma || !h_price || !h_price_offset || !h_ring || !h_reward || !h_terminated || !h_truncated) return CDA_EINVAL;
    DevGuard guard(e->device);
    if (!e->was_reset) return CDA_ESTATE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t MA = (size_t)e->M * e->dev.A;
    // all host buffers must be pinned + mapped: the kernel reads the actions and writes the outputs in place
    static thread_local const void *c_key[6] = {nullptr}; static thread_local void *c_val[6] = {nullptr};
    const void *hp[6] = {h_category, h_ring, h_reward, h_terminated, h_truncated, nullptr};
    void *dp[5];
    for (int i = 0; i < 5; ++i) {
        if (c_key[i] != hp[i]) { c_key[i] = hp[i]; c_val[i] = mapped_alias(hp[i]); }
        dp[i] = c_val[i];
        if (!dp[i]) return CDA_EINVAL;
    }
    const char *hc = reinterpret_cast<const char *>(h_category);
    if (!(reinterpret_cast<const char *>(h_size_mean) == hc + MA * 4 && reinterpret_cast<const char *>(h_size_sigma) == hc + 2 * MA * 4 &&
          reinterpret_cast<const char *>(h_price) == hc + 3 * MA * 4 && reinterpret_cast<const char *>(h_price_offset) == hc + 4 * MA * 4)) return CDA_EINVAL;
    CdaStepParams p;
    memset(&p, 0, sizeof(p));
    char *zi = reinterpret_cast<char *>(dp[0]);
    p.cat = reinterpret_cast<const int *>(zi); p.mean = reinterpret_cast<const float *>(zi + MA * 4);
    p.sigma = reinterpret_cast<const float *>(zi + 2 * MA * 4); p.pcode = reinterpret_cast<const int *>(zi + 3 * MA * 4);
    p.poff = reinterpret_cast<const int *>(zi + 4 * MA * 4);
    p.ring_out = reinterpret_cast<float *>(dp[1]);
    p.ring_slot = (int)(((ring_pos % e->dev.n_hist) + e->dev.n_hist) % e->dev.n_hist);
    p.reward = reinterpret_cast<double *>(dp[2]); p.term = reinterpret_cast<unsigned char *>(dp[3]); p.trunc = reinterpret_cast<unsigned char *>(dp[4]);
    return step_common(e, p, st, true);
}

int cda_reset_host_ring(CdaEnv *e, const uint64_t *d_seeds, const uint8_t *d_mask, float *h_ring, void *stream) {
    if (!e || !h_ring) return CDA_EINVAL;
    DevGuard guard(e->device);
    void *dr = mapped_alias(h_ring);
    if (!dr) return CDA_EINVAL;
    int rc = cda_reset(e, d_seeds, d_mask, nullptr, stream);
    if (rc) return rc;
    const int n = e->M * 2 * e->dev.n_hist * CDA_SNAPSHOT_DIM, threads = 256;
    cda_ring_fill_kernel<<<(n + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(e->dev, e->state, e->M, d_mask, reinterpret_cast<float *>(dr));
    CUDA_TRY(cudaGetLastError());
    e->launches++;
    return CDA_OK;
}

// ---- sliding observation window (see include/cda_b200.h) ------------------------------------------------
// market_major: the five pointers address ONE block i32[M][5][A] (h_category = its start).  inline_rec: the record of this step is
// stored behind the newest snapshot, i.e. at the head of slot pos + 1 of every row (needs pos + 1 < slots and 2A + 2 <= 42 words).
static int step_window_impl(CdaEnv *e, const int32_t *h_category, const float *h_size_mean, const float *h_size_sigma,
                            const int32_t *h_price, const int32_t *h_price_offset, bool market_major, float *h_window, int32_t slots, int32_t pos,
                            void *h_records, bool inline_rec, int32_t sync, void *stream) {
    if (!e || !h_category || !h_size_mean || !h_size_sigma || !h_price || !h_price_offset || !h_window || !h_records) return CDA_EINVAL;
    DevGuard guard(e->device);
    const int H = e->dev.n_hist, A = e->dev.A;
    if (slots < H || pos < H - 1 || pos >= slots) return CDA_EINVAL;
    if (inline_rec && (pos + 1 >= slots || 2 * A + 2 > CDA_SNAPSHOT_DIM)) return CDA_EINVAL;
    if (!e->was_reset) return CDA_ESTATE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t MA = (size_t)e->M * e->dev.A;
    const char *hc = reinterpret_cast<const char *>(h_category);
    const bool in_contig = market_major ||
                           (reinterpret_cast<const char *>(h_size_mean) == hc + MA * 4 && reinterpret_cast<const char *>(h_size_sigma) == hc + 2 * MA * 4 &&
                            reinterpret_cast<const char *>(h_price) == hc + 3 * MA * 4 && reinterpret_cast<const char *>(h_price_offset) == hc + 4 * MA * 4);
    char *zi = nullptr;
    if (e->zerocopy_in && in_contig) {
        if (e->zi_host != h_category) { e->zi_host = h_category; e->zi_dev = mapped_alias(h_category); }
        zi = reinterpret_cast<char *>(e->zi_dev);
    }
    // timing experiments only (tools/e2e_timeline.py): 1 = reuse the actions already staged on the device (no input
    // transfer after the first call), 2 = keep the outputs on the device (no output transfer), 3 = both
    const int dbg = g_dbg_window;
    static int dbg_calls = 0;
    if (!(dbg & 1)) dbg_calls = 0;
    const bool dbg_skip_in = (dbg & 1) && dbg_calls++ > 0, dbg_dev_out = (dbg & 2) != 0;
    if (dbg & 1) zi = nullptr;
    CdaStepParams p;
    memset(&p, 0, sizeof(p));
    p.cat = e->s_cat; p.mean = e->s_mean; p.sigma = e->s_sigma; p.pcode = e->s_pcode; p.poff = e->s_poff;
    const size_t fstep = market_major ? (size_t)A * 4 : MA * 4;   // bytes from one field's array to the next inside the block
    if (market_major) {   // (also the layout of the staged copy)
        const char *b = reinterpret_cast<const char *>(e->s_cat);
        p.mean = reinterpret_cast<const float *>(b + fstep); p.sigma = reinterpret_cast<const float *>(b + 2 * fstep);
        p.pcode = reinterpret_cast<const int *>(b + 3 * fstep); p.poff = reinterpret_cast<const int *>(b + 4 * fstep);
        p.act_mstride = 5 * A; p.act_packed = 1;
    }
    if (dbg_skip_in) {
    } else if (zi) {
        p.cat = reinterpret_cast<const int *>(zi); p.mean = reinterpret_cast<const float *>(zi + fstep);
        p.sigma = reinterpret_cast<const float *>(zi + 2 * fstep); p.pcode = reinterpret_cast<const int *>(zi + 3 * fstep);
        p.poff = reinterpret_cast<const int *>(zi + 4 * fstep);
    } else if (in_contig) {
        CUDA_TRY(cudaMemcpyAsync(e->s_cat, h_category, MA * 4 * 5, cudaMemcpyHostToDevice, st));
    } else {
        CUDA_TRY(cudaMemcpyAsync(e->s_cat, h_category, MA * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->s_mean, h_size_mean, MA * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->s_sigma, h_size_sigma, MA * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->s_pcode, h_price, MA * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->s_poff, h_price_offset, MA * 4, cudaMemcpyHostToDevice, st));
    }
    // Everything the host does not hold yet is stored by the kernel STRAIGHT into pinned host memory (posted PCIe
    // writes that overlap the step; no copy-engine hand-off): the newest snapshot into slot `pos` of the market's window
    // row (the whole stack when the window restarts at slot 0) and the packed result record.  When the buffers are not
    // mapped (or CDA_ZEROCOPY=0) the same bytes are staged in HBM and follow with one strided + one contiguous copy.
    const size_t rec_bytes = cda_record_bytes(e);
    float *zw = nullptr; unsigned char *zr = nullptr;
    if (e->zerocopy) {
        if (e->zc_host != h_window) { e->zc_host = h_window; e->zc_dev = mapped_alias(h_window); }
        if (e->zr_host != h_records) { e->zr_host = h_records; e->zr_dev = mapped_alias(h_records); }
        zw = reinterpret_cast<float *>(e->zc_dev); zr = reinterpret_cast<unsigned char *>(e->zr_dev);
        if (!zw || !zr || dbg_dev_out) { zw = nullptr; zr = nullptr; }
    }
    const int wstride = slots * CDA_SNAPSHOT_DIM;
    unsigned char *rec = zr ? zr : e->s_rec;
    // inline_rec: the record's place is behind the newest snapshot (head of slot pos + 1).  On the zero-copy newest-snapshot
    // steps it rides in the snapshot's own store instructions; on a window restart the kernel stores it there separately; on
    // the staged fallback it is staged in HBM and copied there.
    const bool rides = inline_rec && zw && pos != H - 1;
    if (rides) p.rec_inline = 1;
    else if (inline_rec && zw) {
        unsigned char *ir = reinterpret_cast<unsigned char *>(zw + (size_t)(pos + 1) * CDA_SNAPSHOT_DIM);
        p.reward = reinterpret_cast<double *>(ir); p.reward_stride = wstride / 2;
        p.term = ir + (size_t)A * 8; p.trunc = p.term + 1; p.flag_stride = wstride * 4; p.flag_pack = 1;
    } else {
        if (inline_rec) rec = e->s_rec;
        p.reward = reinterpret_cast<double *>(rec); p.reward_stride = (int)(rec_bytes / 8);
        p.term = rec + (size_t)e->dev.A * 8; p.trunc = p.term + 1; p.flag_stride = (int)rec_bytes; p.flag_pack = 1;
    }
    if (zw) {
        if (pos == H - 1) { p.obs = zw; p.obs_stride = wstride; }
        else { p.ring_out = zw; p.ring_stride = wstride; p.ring_slot = pos; p.ring_mirror = 0; }
    } else p.obs = e->s_obs;
    const bool bell = zw && sync && e->doorbell && !(e->dev.dec && e->twin_steps + 1 >= e->twin_every);
    if (bell) { p.done_ctr = e->done_ctr; p.done_flag = e->status_dev + 8; p.done_seq = ++e->done_seq; }
    int rc = step_common(e, p, st, true);
    if (rc) return rc;
    if (bell) return doorbell_wait(e, p.done_seq, st);
    if (!zw && !dbg_dev_out) {
        const size_t dpitch = (size_t)wstride * 4, spitch = (size_t)e->dev.W * 4;
        if (pos == H - 1) CUDA_TRY(cudaMemcpy2DAsync(h_window, dpitch, e->s_obs, spitch, spitch, e->M, cudaMemcpyDeviceToHost, st));
        else CUDA_TRY(cudaMemcpy2DAsync(h_window + (size_t)pos * CDA_SNAPSHOT_DIM, dpitch, e->s_obs + (size_t)(H - 1) * CDA_SNAPSHOT_DIM, spitch,
                                        CDA_SNAPSHOT_DIM * 4, e->M, cudaMemcpyDeviceToHost, st));
        if (inline_rec) CUDA_TRY(cudaMemcpy2DAsync(h_window + (size_t)(pos + 1) * CDA_SNAPSHOT_DIM, dpitch, e->s_rec, rec_bytes, (size_t)A * 8 + 8, e->M, cudaMemcpyDeviceToHost, st));
        else CUDA_TRY(cudaMemcpyAsync(h_records, e->s_rec, (size_t)e->M * rec_bytes, cudaMemcpyDeviceToHost, st));
    }
    if (sync) CUDA_TRY(cudaStreamSynchronize(st));
    return CDA_OK;
}

int cda_step_host_window(CdaEnv *e, const int32_t *h_category, const float *h_size_mean, const float *h_size_sigma,
                         const int32_t *h_price, const int32_t *h_price_offset, float *h_window, int32_t slots, int32_t pos,
                         void *h_records, int32_t sync, void *stream) {
    return step_window_impl(e, h_category, h_size_mean, h_size_sigma, h_price, h_price_offset, false, h_window, slots, pos, h_records, false, sync, stream);
}

// ---- bound form: the buffers are registered once, the per-step call carries three arguments -------------------
int cda_window_bind(CdaEnv *e, float *h_window, int32_t slots, void *h_records, void *stream) {
    if (!e || !h_window || !h_records || slots < e->dev.n_hist) return CDA_EINVAL;
    e->w_window = h_window; e->w_slots = slots; e->w_records = h_records; e->w_stream = stream;
    return CDA_OK;
}
int cda_step_window(CdaEnv *e, const int32_t *h_action_block, int32_t pos, int32_t flags) {
    if (!e || !e->w_window || !h_action_block) return CDA_EINVAL;
    const bool mm = (flags & CDA_WIN_MARKET_MAJOR) != 0;
    const size_t fs = mm ? (size_t)e->dev.A : (size_t)e->M * e->dev.A;   // words between the five field arrays
    const int32_t *b = h_action_block;
    return step_window_impl(e, b, reinterpret_cast<const float *>(b + fs), reinterpret_cast<const float *>(b + 2 * fs), b + 3 * fs, b + 4 * fs, mm,
                            e->w_window, e->w_slots, pos, e->w_records, (flags & CDA_WIN_INLINE_RECORD) != 0, flags & CDA_WIN_SYNC, e->w_stream);
}

// ---- dense plane output (see include/cda_b200.h): one aligned cell per market and step, consecutive markets contiguous -------------
int cda_step_planes(CdaEnv *e, const int32_t *h_action_block, float *h_plane, int32_t cell_words, int32_t flags, void *stream) {
    if (!e || !h_action_block || !h_plane) return CDA_EINVAL;
    DevGuard guard(e->device);
    const int A = e->dev.A;
    if (cell_words < CDA_SNAPSHOT_DIM + 2 * A + 2 || (cell_words & 1)) return CDA_EINVAL;
    if (!e->was_reset) return CDA_ESTATE;
    cudaStream_t st = (cudaStream_t)stream;
    const bool mm = (flags & CDA_WIN_MARKET_MAJOR) != 0;
    const size_t MA = (size_t)e->M * A, fstep = mm ? (size_t)A * 4 : MA * 4;
    if (e->zi_host != h_action_block) { e->zi_host = h_action_block; e->zi_dev = e->zerocopy_in ? mapped_alias(h_action_block) : nullptr; }
    if (e->zc_host != h_plane) { e->zc_host = h_plane; e->zc_dev = e->zerocopy ? mapped_alias(h_plane) : nullptr; }
    const char *zi = reinterpret_cast<const char *>(e->zi_dev);
    float *zp = reinterpret_cast<float *>(e->zc_dev);
    if (!zi) { CUDA_TRY(cudaMemcpyAsync(e->s_cat, h_action_block, MA * 20, cudaMemcpyHostToDevice, st)); zi = reinterpret_cast<const char *>(e->s_cat); }
    if (!zp && !e->s_plane) CUDA_TRY(cudaMalloc(&e->s_plane, (size_t)e->M * 64 * 4 > (size_t)e->M * cell_words * 4 ? (size_t)e->M * 64 * 4 : (size_t)e->M * cell_words * 4));
    CdaStepParams p;
    memset(&p, 0, sizeof(p));
    p.cat = reinterpret_cast<const int *>(zi); p.mean = reinterpret_cast<const float *>(zi + fstep); p.sigma = reinterpret_cast<const float *>(zi + 2 * fstep);
    p.pcode = reinterpret_cast<const int *>(zi + 3 * fstep); p.poff = reinterpret_cast<const int *>(zi + 4 * fstep);
    if (mm) { p.act_mstride = 5 * A; p.act_packed = 1; }
    p.ring_out = zp ? zp : e->s_plane; p.ring_stride = cell_words; p.ring_slot = 0; p.ring_mirror = 0; p.rec_inline = 1; p.ring_pad = cell_words;
    const bool bell = zp && (flags & CDA_WIN_SYNC) && e->doorbell && !(e->dev.dec && e->twin_steps + 1 >= e->twin_every);   // (a replay kernel follows this step: synchronise)
    if (bell) { p.done_ctr = e->done_ctr; p.done_flag = e->status_dev + 8; p.done_seq = ++e->done_seq; }
    int rc = step_common(e, p, st, true);
    if (rc) return rc;
    if (!zp) CUDA_TRY(cudaMemcpyAsync(h_plane, e->s_plane, (size_t)e->M * cell_words * 4, cudaMemcpyDeviceToHost, st));   // ONE contiguous copy
    if (bell) return doorbell_wait(e, p.done_seq, st);
    if (flags & CDA_WIN_SYNC) CUDA_TRY(cudaStreamSynchronize(st));
    return CDA_OK;
}

int cda_reset_planes(CdaEnv *e, const uint64_t *d_seeds, const uint8_t *d_mask, float *h_planes, int32_t slots, int32_t cell_words, int32_t pos, void *stream) {
    if (!e || !h_planes || slots < e->dev.n_hist || cell_words < CDA_SNAPSHOT_DIM || pos < 0 || pos >= slots) return CDA_EINVAL;
    DevGuard guard(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = cda_reset(e, d_seeds, d_mask, nullptr, stream);
    if (rc) return rc;
    float *zp = reinterpret_cast<float *>(mapped_alias(h_planes));
    if (!zp) return CDA_EINVAL;                 // cold path: the plane ring must be pinned + mapped
    const int n = e->M * e->dev.W, threads = 256;
    cda_emit_planes_kernel<<<(n + threads - 1) / threads, threads, 0, st>>>(e->dev, e->state, e->M, zp, slots, cell_words, pos);
    CUDA_TRY(cudaGetLastError());
    e->launches++;
    CUDA_TRY(cudaStreamSynchronize(st));
    return CDA_OK;
}


int cda_serve_bind(CdaEnv *e, float *h_planes, int32_t slots, int32_t cell_words) {
    if (!e || !h_planes || slots < e->dev.n_hist + 1 || slots >= (int)CDA_SRV_STOP) return CDA_EINVAL;
    if (cell_words < CDA_SNAPSHOT_DIM + 2 * e->dev.A + 2 || (cell_words & 1)) return CDA_EINVAL;
    DevGuard guard(e->device);
    SRV_QUIESCE(e);
    e->srv_bound = false;
    static const int enabled = getenv("CDA_SERVE") ? atoi(getenv("CDA_SERVE")) : 1;
    if (!enabled || e->dev.dec || !e->zerocopy || !e->zerocopy_in) return CDA_EUNSUPPORTED;
    float *zp = reinterpret_cast<float *>(mapped_alias(h_planes));
    if (!zp) return CDA_EUNSUPPORTED;
    if (!e->srv_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&e->srv_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&e->srv_event, cudaEventDisableTiming));
        CUDA_TRY(cudaMalloc(&e->srv_go_dev, (size_t)CDA_SRV_COPIES * 128 + 128));
        CUDA_TRY(cudaMemset(e->srv_go_dev, 0, (size_t)CDA_SRV_COPIES * 128 + 128));
        e->srv_done_dev = reinterpret_cast<unsigned *>(e->srv_go_dev + CDA_SRV_COPIES * 16);
        const char *ls = getenv("CDA_SERVE_LEASE_US");
        e->srv_lease_ns = (unsigned long long)(ls && atoi(ls) > 0 ? atoi(ls) : 2000) * 1000ULL;
    }
    CdaStepParams none;
    memset(&none, 0, sizeof(none));
    const cudaError_t fit = launch_serve_cap(e, none, true);
    if (fit == cudaErrorLaunchOutOfResources) return CDA_EUNSUPPORTED;   // more markets than one resident wave holds
    CUDA_TRY(fit);
    e->srv_planes_host = h_planes; e->srv_planes_dev = zp; e->srv_slots = slots; e->srv_cell = cell_words;
    e->srv_win_steps = e->srv_win_launches = 0;
    e->srv_bound = true;
    return CDA_OK;
}

int cda_serve_step(CdaEnv *e, const int32_t *h_action_block, int32_t slot, void *stream) {
    if (!e || !e->srv_bound || !h_action_block || slot < 0 || slot >= e->srv_slots) return CDA_EINVAL;
    if (!e->was_reset) return CDA_ESTATE;
    // device alias of this step's action block (recent blocks cached: cudaPointerGetAttributes costs about a microsecond)
    const unsigned ci = (unsigned)((reinterpret_cast<uintptr_t>(h_action_block) >> 6) * 0x9E3779B1u >> 26) & 63u;
    if (e->srv_ptr_cache_h[ci] != h_action_block) {
        DevGuard guard(e->device);
        const unsigned char *d = reinterpret_cast<const unsigned char *>(mapped_alias(h_action_block));
        if (!d) return CDA_EINVAL;            // the action block must be pinned + mapped
        e->srv_ptr_cache_h[ci] = h_action_block; e->srv_ptr_cache_d[ci] = d;
    }
    const unsigned char *dblk = e->srv_ptr_cache_d[ci];
    if (!e->srv_act_dev0) {
        if (e->srv_running) SRV_QUIESCE(e);
        e->srv_act_dev0 = dblk;
    }
    // (bulk copies need 16-B aligned sources: blocks of 4k-agent markets from cudaHostAlloc / torch pin_memory are; the volatile-load path takes any int32 block)
    if ((reinterpret_cast<uintptr_t>(dblk) & 3) || ((e->dev.A % 4) == 0 && (reinterpret_cast<uintptr_t>(dblk) & 15))) return CDA_EINVAL;
    long long off = dblk - e->srv_act_dev0;
    if (off < -(1LL << 32) || off >= (1LL << 32)) {   // out of a message's reach: move the base (retires the kernel once)
        SRV_QUIESCE(e);
        e->srv_act_dev0 = dblk; off = 0;
    }
    if (g_srv_dbg & 1) off = 0;
    // A server that has to be relaunched for most steps — the host takes longer than the lease between steps, or other work keeps claiming the
    // GPU — only adds its (larger) launch to every step: decline from here on, the caller goes back to cda_step_planes.
    if (e->srv_win_steps >= 64) {
        const bool thrash = e->srv_win_launches >= 48;
        e->srv_win_steps = e->srv_win_launches = 0;
        if (thrash) { SRV_QUIESCE(e); e->srv_bound = false; return CDA_EUNSUPPORTED; }
    }
    const long long launches0 = e->srv_launches;
    const unsigned seq = e->srv_seq + 1u;
    const unsigned long long msg = (unsigned long long)cda_srv_seq24(seq) | ((unsigned long long)(unsigned)slot << 24) |
                                   ((unsigned long long)(unsigned)(int)(off >> 2) << 32);
    volatile unsigned *done = e->status_host + CDA_SRV_DONE_WORD;
    __atomic_store_n(reinterpret_cast<unsigned long long *>(e->status_host + CDA_SRV_GO_WORD), msg, __ATOMIC_RELEASE);   // (the caller's action writes come first: TSO + release)
    if (e->srv_running && host_now_ns() - e->srv_last_ns > e->srv_lease_ns / 2) {   // idle for a while: the kernel may have given the SMs back
        DevGuard guard(e->device);
        if (cudaStreamQuery(e->srv_stream) == cudaSuccess) e->srv_running = false; else cudaGetLastError();
    }
    if (!e->srv_running) {
        DevGuard guard(e->device);
        int rc = srv_launch(e, (cudaStream_t)stream);
        if (rc) return rc;
    }
    const unsigned long long t_start = host_now_ns();
    for (unsigned spins = 1;; ++spins) {
        if (*done == seq) break;
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((spins & 0x3ffu) == 0) {          // every ~50 us: did the kernel retire (lease ran out just before the message) without serving this step?
            DevGuard guard(e->device);
            const cudaError_t q = cudaStreamQuery(e->srv_stream);
            if (q == cudaSuccess) {
                if (*done == seq) break;
                e->srv_running = false;
                if (e->status_host[CDA_SRV_ERR_WORD]) { snprintf(g_cuda_err, sizeof(g_cuda_err), "resident step server: worker watchdog fired"); return CDA_ECUDA; }
                int rc = srv_launch(e, (cudaStream_t)stream);
                if (rc) return rc;
            } else if (q != cudaErrorNotReady) {
                snprintf(g_cuda_err, sizeof(g_cuda_err), "resident step server: %s", cudaGetErrorString(q));
                e->srv_running = false;
                return CDA_ECUDA;
            } else cudaGetLastError();
            if (host_now_ns() - t_start > 3000000000ULL) {   // 3 s without a completion (a step takes tens of microseconds): give up, retire the kernel
                snprintf(g_cuda_err, sizeof(g_cuda_err), "resident step server: step %u not completed within 3 s", seq);
                srv_quiesce(e);
                return CDA_ECUDA;
            }
        }
    }
    e->srv_seq = seq;
    e->srv_last_ns = host_now_ns();
    e->srv_win_steps++; e->srv_win_launches += (int)(e->srv_launches - launches0);
    return CDA_OK;
}

int cda_serve_stop(CdaEnv *e) {
    if (!e) return CDA_EINVAL;
    SRV_QUIESCE(e);
    return CDA_OK;
}
int64_t cda_serve_launches(const CdaEnv *e) { return e ? e->srv_launches : 0; }

int cda_reset_host_window(CdaEnv *e, const uint64_t *d_seeds, const uint8_t *d_mask, float *h_window, int32_t slots, void *stream) {
    if (!e || !h_window || slots < e->dev.n_hist) return CDA_EINVAL;
    DevGuard guard(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = cda_reset(e, d_seeds, d_mask, nullptr, stream);
    if (rc) return rc;
    // every market's stack (reset or not) goes to slots 0..n_hist-1 of its row, rebuilt from the snapshot ring in the state
    const int n = e->M * e->dev.W, threads = 256, wstride = slots * CDA_SNAPSHOT_DIM;
    float *zw = e->zerocopy ? reinterpret_cast<float *>(mapped_alias(h_window)) : nullptr;
    if (zw) cda_emit_obs_kernel<<<(n + threads - 1) / threads, threads, 0, st>>>(e->dev, e->state, e->M, zw, wstride);
    else {
        cda_emit_obs_kernel<<<(n + threads - 1) / threads, threads, 0, st>>>(e->dev, e->state, e->M, e->s_obs, e->dev.W);
        CUDA_TRY(cudaMemcpy2DAsync(h_window, (size_t)wstride * 4, e->s_obs, (size_t)e->dev.W * 4, (size_t)e->dev.W * 4, e->M, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaGetLastError());
    e->launches++;
    CUDA_TRY(cudaStreamSynchronize(st));
    return CDA_OK;
}

int cda_rollout_random(CdaEnv *e, int32_t num_steps, uint64_t policy_seed, float *d_obs, double *d_reward,
                       uint8_t *d_terminated, uint8_t *d_truncated, void *stream) {
    if (!e || num_steps < 1) return CDA_EINVAL;
    DevGuard guard(e->device);
    if (!e->was_reset) return CDA_ESTATE;
    CdaStepParams p;
    memset(&p, 0, sizeof(p));
    p.obs = d_obs; p.reward = d_reward; p.term = d_terminated; p.trunc = d_truncated;
    p.policy_seed = policy_seed;
    // decimal_ledger: the journals are flushed between launches, so a long rollout runs as chunks of at most CDA_TWIN_FLUSH_STEPS steps
    const int chunk = e->dev.dec ? std::min(e->twin_every, CDA_TWIN_FLUSH_STEPS) : num_steps;
    for (int done = 0; done < num_steps;) {
        CdaStepParams q = p;
        q.num_steps = std::min(chunk, num_steps - done);
        int rc = step_common(e, q, (cudaStream_t)stream);
        if (rc) return rc;
        done += q.num_steps;
    }
    return CDA_OK;
}

// ---- fused step + all-gather over NVLink peer memory (include/cda_b200.h) ---------------------------------------------------------
// Every rank owns a GATHER WINDOW  f32[world*M][CDA_GATHER_SLOTS][42]  (the sliding-window layout of the host path: a row's n_hist most
// recent slots are its stacked observation, the result record rides behind the newest snapshot) followed by a flag array u32[64].
// Row = CDA_GATHER_SLOTS snapshot slots, then TWO result records (step parity): a rank that is one step ahead writes the other record
// and a different slot than the ones its peers' consumers are still reading.
static int gather_row_words(const CdaEnv *e) { return (CDA_GATHER_SLOTS * CDA_SNAPSHOT_DIM + 2 * (2 * e->dev.A + 2) + 3) / 4 * 4; }
static size_t gather_window_bytes(const CdaEnv *e, int world) { return (size_t)world * e->M * gather_row_words(e) * 4; }

int cda_gather_create(CdaEnv *e, int32_t world, int32_t rank, void *ipc_handle_out64, void **d_local_buf, uint64_t *bytes) {
    if (!e || world < 1 || world > CDA_MAX_PEERS || rank < 0 || rank >= world || !ipc_handle_out64) return CDA_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (CDA_GATHER_SLOTS < 2 * e->dev.n_hist + 2) return CDA_EINVAL;
    DevGuard guard(e->device);
    if (e->g_local) return CDA_EINVAL;
    e->g_bytes = gather_window_bytes(e, world) + 256;
    CUDA_TRY(cudaMalloc(&e->g_local, e->g_bytes));
    CUDA_TRY(cudaMemset(e->g_local, 0, e->g_bytes));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, e->g_local));
    memcpy(ipc_handle_out64, &h, 64);
    e->g_world = world; e->g_rank = rank; e->g_connected = false; e->g_seq = 0; e->g_pos = -1;
    if (d_local_buf) *d_local_buf = e->g_local;
    if (bytes) *bytes = e->g_bytes;
    return CDA_OK;
}

int cda_gather_connect(CdaEnv *e, const void *all_handles) {
    if (!e || !e->g_local || !all_handles) return CDA_EINVAL;
    DevGuard guard(e->device);
    for (int g = 0; g < e->g_world; ++g) {
        if (g == e->g_rank) { e->g_peer[g] = e->g_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, reinterpret_cast<const char *>(all_handles) + (size_t)g * 64, 64);
        void *ptr = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        e->g_peer[g] = reinterpret_cast<unsigned char *>(ptr);
    }
    e->g_connected = true;
    return CDA_OK;
}

// destinations of this rank's rows: its own window first (delta 0), then the peers'
static void gather_targets(const CdaEnv *e, CdaStepParams &p) {
    p.rep_n = e->g_world;
    int n = 1;
    p.rep_delta[0] = 0;
    for (int g = 0; g < e->g_world; ++g) if (g != e->g_rank) p.rep_delta[n++] = (long long)(e->g_peer[g] - e->g_local);
}

int cda_gather_publish(CdaEnv *e, void *stream) {
    if (!e) return CDA_EINVAL;
    if (!e->was_reset || !e->g_connected) return CDA_ESTATE;
    DevGuard guard(e->device);
    SRV_QUIESCE(e);
    cudaStream_t st = (cudaStream_t)stream;
    const int n = e->M * e->dev.W, threads = 256, wstride = gather_row_words(e);
    for (int g = 0; g < e->g_world; ++g) {   // every market's current stack into slots 0..n_hist-1 of its row of every rank's window (cold path)
        float *dst = reinterpret_cast<float *>(e->g_peer[g]) + (size_t)e->g_rank * e->M * wstride;
        cda_emit_obs_kernel<<<(n + threads - 1) / threads, threads, 0, st>>>(e->dev, e->state, e->M, dst, wstride);
    }
    CUDA_TRY(cudaGetLastError());
    e->launches += e->g_world;
    e->g_pos = e->dev.n_hist - 1;
    ++e->g_seq;
    CUDA_TRY(cudaStreamSynchronize(st));     // the copies have landed in the peers' memory ...
    const unsigned seq = e->g_seq;           // ... then this rank's flag in every window
    for (int g = 0; g < e->g_world; ++g)
        CUDA_TRY(cudaMemcpyAsync(e->g_peer[g] + gather_window_bytes(e, e->g_world) + 4 * e->g_rank, &seq, 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return CDA_OK;
}

int cda_step_gather(CdaEnv *e, const int32_t *d_category, const float *d_size_mean, const float *d_size_sigma,
                    const int32_t *d_price, const int32_t *d_price_offset, void *stream) {
    if (!e || !d_category || !d_size_mean || !d_size_sigma || !d_price || !d_price_offset) return CDA_EINVAL;
    DevGuard guard(e->device);
    if (!e->was_reset || !e->g_connected || e->g_pos < 0) return CDA_ESTATE;
    const int H = e->dev.n_hist, A = e->dev.A, wstride = gather_row_words(e);
    int pos = e->g_pos + 1;
    if (pos >= CDA_GATHER_SLOTS) pos = H - 1;              // the window restarts: the whole stack is re-sent into slots 0..n_hist-1
    CdaStepParams p;
    memset(&p, 0, sizeof(p));
    p.cat = d_category; p.mean = d_size_mean; p.sigma = d_size_sigma; p.pcode = d_price; p.poff = d_price_offset;
    float *row0 = reinterpret_cast<float *>(e->g_local) + (size_t)e->g_rank * e->M * wstride;   // this rank's rows in its own window
    gather_targets(e, p);
    if (pos == H - 1) { p.obs = row0; p.obs_stride = wstride; }                                              // the whole stack
    else { p.ring_out = row0; p.ring_stride = wstride; p.ring_slot = pos; p.ring_mirror = 0; }              // the newest snapshot only
    {   // the result record of this step: record slot (seq & 1) behind the snapshot slots of the row
        unsigned char *ir = reinterpret_cast<unsigned char *>(row0 + (size_t)CDA_GATHER_SLOTS * CDA_SNAPSHOT_DIM + (size_t)((e->g_seq + 1) & 1u) * (2 * A + 2));
        p.reward = reinterpret_cast<double *>(ir); p.reward_stride = wstride / 2;
        p.term = ir + (size_t)A * 8; p.trunc = p.term + 1; p.flag_stride = wstride * 4; p.flag_pack = 1;
    }
    // completion: the last warp publishes "step seq of rank r is in your window" to every rank's flag array
    p.done_ctr = e->done_ctr; p.done_seq = ++e->g_seq;
    p.done_flag = reinterpret_cast<unsigned *>(e->g_local + gather_window_bytes(e, e->g_world)) + e->g_rank;
    int rc = step_common(e, p, (cudaStream_t)stream);
    if (rc) return rc;
    e->g_pos = pos;
    return CDA_OK;
}

int cda_gather_wait(CdaEnv *e, void *stream) {
    if (!e || !e->g_connected) return CDA_EINVAL;
    DevGuard guard(e->device);
    unsigned *flags = reinterpret_cast<unsigned *>(e->g_local + gather_window_bytes(e, e->g_world));
    cda_gather_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, e->g_world, e->g_seq, flags + 32);
    CUDA_TRY(cudaGetLastError());
    e->launches++;
    return CDA_OK;
}
int32_t cda_gather_pos(const CdaEnv *e) { return e ? e->g_pos : -1; }
int32_t cda_gather_row_words(const CdaEnv *e) { return e ? gather_row_words(e) : 0; }
int32_t cda_gather_record_parity(const CdaEnv *e) { return e ? (int32_t)(e->g_seq & 1u) : 0; }

int cda_get_info(CdaEnv *e, int32_t field, int64_t *d_out, void *stream) {
    if (!e || !d_out || field < 0 || field >= CDA_INFO__COUNT) return CDA_EINVAL;
    DevGuard guard(e->device);
    SRV_QUIESCE(e);
    const int n = field == CDA_INFO_MARKET ? e->M : e->M * e->dev.A;
    const int threads = 256, grid = (n + threads - 1) / threads;
    cda_info_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(e->dev, e->state, e->M, field, (long long *)d_out);
    CUDA_TRY(cudaGetLastError());
    e->launches++;
    return CDA_OK;
}

int cda_get_info_all(CdaEnv *e, int64_t *d_out, void *stream) {
    if (!e || !d_out) return CDA_EINVAL;
    DevGuard guard(e->device);
    SRV_QUIESCE(e);
    const int n = CDA_INFO_MARKET * e->M * e->dev.A + e->M;
    const int threads = 256, grid = (n + threads - 1) / threads;
    cda_info_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(e->dev, e->state, e->M, -1, (long long *)d_out);
    CUDA_TRY(cudaGetLastError());
    e->launches++;
    return CDA_OK;
}

int cda_set_action_log(CdaEnv *e, int32_t *d_log) {
    if (!e) return CDA_EINVAL;
    SRV_QUIESCE(e);
    e->act_log = d_log;
    return CDA_OK;
}
int cda_get_fills(CdaEnv *e, int32_t *d_fills, int32_t *d_counts, void *stream) {
    if (!e || !e->fills) return CDA_EINVAL;
    DevGuard guard(e->device);
    SRV_QUIESCE(e);
    cudaStream_t st = (cudaStream_t)stream;
    if (d_fills) CUDA_TRY(cudaMemcpyAsync(d_fills, e->fills, (size_t)e->M * e->dev.fill_cap * CDA_FILL_WORDS * 4, cudaMemcpyDeviceToDevice, st));
    if (d_counts) CUDA_TRY(cudaMemcpyAsync(d_counts, e->fill_counts, (size_t)e->M * 4, cudaMemcpyDeviceToDevice, st));
    return CDA_OK;
}

int cda_dump_market(CdaEnv *e, int32_t market, int64_t *h_bids, int64_t *h_asks, int64_t *h_bids_map,
                    int64_t *h_asks_map, int32_t max_rows, int32_t *h_counts, uint64_t *h_rng6) {
    if (!e || market < 0 || market >= e->M || !h_counts) return CDA_EINVAL;
    DevGuard guard(e->device);
    SRV_QUIESCE(e);
    std::vector<unsigned char> blk(e->dev.stride);
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(blk.data(), e->state + (size_t)market * e->dev.stride, e->dev.stride, cudaMemcpyDeviceToHost));
    const unsigned *hdr = reinterpret_cast<const unsigned *>(blk.data());
    const unsigned *pool = reinterpret_cast<const unsigned *>(blk.data() + e->dev.off_pool);
    const int cap = e->dev.cap;
    for (int side = 0; side < 2; ++side) {
        const int n = (int)hdr[8 + side];
        const unsigned *sb = pool + (size_t)side * CDA_POOL_FIELDS * cap;
        auto fld = [&](int i, int f) { return sb[CDA_EOFF(i) + f * 32]; };
        std::vector<int> idx(n);
        for (int i = 0; i < n; ++i) idx[i] = i;
        std::sort(idx.begin(), idx.end(), [&](int a, int b) {
            unsigned pa = fld(a, 0) & CDA_PRICE_MASK, pb = fld(b, 0) & CDA_PRICE_MASK;
            if (pa != pb) return side == 0 ? pa > pb : pa < pb;
            return fld(a, 4) < fld(b, 4);
        });
        int64_t *rows = side == 0 ? h_bids : h_asks;
        if (rows)
            for (int q = 0; q < n && q < max_rows; ++q) {
                int i = idx[q];
                rows[q * 5 + 0] = fld(i, 0) & CDA_PRICE_MASK; rows[q * 5 + 1] = fld(i, 1); rows[q * 5 + 2] = fld(i, 0) >> 24;
                rows[q * 5 + 3] = fld(i, 2); rows[q * 5 + 4] = fld(i, 3);
            }
        std::sort(idx.begin(), idx.end(), [&](int a, int b) { return fld(a, 4) < fld(b, 4); });
        int64_t *mp = side == 0 ? h_bids_map : h_asks_map;
        if (mp) for (int q = 0; q < n && q < max_rows; ++q) mp[q] = fld(idx[q], 2);
        h_counts[side] = n;
    }
    if (h_rng6) {
        const unsigned long long *r = reinterpret_cast<const unsigned long long *>(hdr + 12);
        h_rng6[0] = r[0]; h_rng6[1] = r[1]; h_rng6[2] = r[2]; h_rng6[3] = r[3]; h_rng6[4] = hdr[10]; h_rng6[5] = hdr[11];
    }
    return CDA_OK;
}

void cda_debug_set_window_mode(int32_t mode) { g_dbg_window = mode; }
int64_t cda_debug_restart_count(void) {
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, cda_debug_restarts, sizeof(v)) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (int64_t)v;
}
// debug builds (-DCDA_PROFILE_PHASES): returns the device buffer of 16 per-phase cycle sums (allocated on first call)
unsigned long long *cda_debug_phase_buffer(void) {
#ifdef CDA_PROFILE_PHASES
    if (!g_prof) { cudaMalloc(&g_prof, (size_t)(1 << 20) * 16 * sizeof(unsigned long long)); cudaMemset(g_prof, 0, (size_t)(1 << 20) * 16 * sizeof(unsigned long long)); }  // per-market rows, M <= 2^20
#endif
    return g_prof;
}
int cda_twin_sync(CdaEnv *e, int64_t *d_out, void *stream) {
    if (!e) return CDA_EINVAL;
    if (!e->dev.dec) return d_out ? CDA_EINVAL : CDA_OK;
    DevGuard guard(e->device);
    SRV_QUIESCE(e);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = twin_flush(e, st);
    if (rc || !d_out) return rc;
    const int n = e->M * e->dev.A, threads = 128;
    cda_twin_dump_kernel<<<(n + threads - 1) / threads, threads, 0, st>>>(e->dev, e->state, e->M, (long long *)d_out);
    CUDA_TRY(cudaGetLastError());
    e->launches++;
    return CDA_OK;
}
size_t cda_state_bytes(const CdaEnv *e) { return e ? e->state_bytes : 0; }
int cda_state_layout(const CdaEnv *e, int32_t out[12]) {
    if (!e || !out) return CDA_EINVAL;
    out[0] = (int32_t)e->dev.stride; out[1] = (int32_t)e->dev.off_acct; out[2] = (int32_t)e->dev.off_hist; out[3] = (int32_t)e->dev.off_pool;
    out[4] = e->dev.cap; out[5] = e->dev.A; out[6] = e->dev.n_hist; out[7] = CDA_HDR_BYTES;
    out[8] = e->dev.dec; out[9] = (int32_t)e->dev.off_twin; out[10] = (int32_t)e->dev.off_jrn; out[11] = CDA_JRN_E;
    return CDA_OK;
}
int cda_save_state(CdaEnv *e, void *h_dst, void *stream) {
    if (!e || !h_dst) return CDA_EINVAL;
    DevGuard guard(e->device);
    SRV_QUIESCE(e);
    CUDA_TRY(cudaMemcpyAsync(h_dst, e->state, e->state_bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return CDA_OK;
}
int cda_load_state(CdaEnv *e, const void *h_src, void *stream) {
    if (!e || !h_src) return CDA_EINVAL;
    DevGuard guard(e->device);
    SRV_QUIESCE(e);
    CUDA_TRY(cudaMemcpyAsync(e->state, h_src, e->state_bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    e->was_reset = true;
    return CDA_OK;
}
const volatile uint32_t *cda_status_flag(const CdaEnv *e) { return e ? e->status_host : nullptr; }
int cda_status_flag_clear(CdaEnv *e) { if (!e) return CDA_EINVAL; *e->status_host = 0; return CDA_OK; }
int32_t cda_record_bytes(const CdaEnv *e) { return e ? (int32_t)(((size_t)e->dev.A * 8 + 8 + 63) / 64 * 64) : 0; }
int32_t cda_num_markets(const CdaEnv *e) { return e ? e->M : 0; }
int32_t cda_obs_dim(const CdaEnv *e) { return e ? e->dev.W : 0; }
int32_t cda_order_capacity(const CdaEnv *e) { return e ? e->dev.cap : 0; }
int64_t cda_kernel_launches(const CdaEnv *e) { return e ? e->launches : 0; }

// ---- cda_dec128.cuh test entries (the Decimal(28) arithmetic of the future device ledger; not on any product path) --------
static CdaDec dec_parse(const char *s) {
    int sign = 0; if (*s == '-') { sign = 1; ++s; } else if (*s == '+') ++s;
    cda_u128 x = 0; int exp = 0, seen_pt = 0, nd = 0, sticky = 0;
    for (; *s && *s != 'e' && *s != 'E'; ++s) {
        if (*s == '.') { seen_pt = 1; continue; }
        if (nd < 38) { x = x * 10 + (unsigned)(*s - '0'); if (x) ++nd; if (seen_pt) --exp; }
        else { sticky |= *s != '0'; if (!seen_pt) ++exp; }
    }
    if (*s == 'e' || *s == 'E') exp += atoi(s + 1);
    return cda_dec_round(sign, x, exp, sticky);
}
static void dec_format(CdaDec a, char *out, int cap) {
    if (a.c == 0) { snprintf(out, (size_t)cap, "0"); return; }
    char t[48]; int n = 0; cda_u128 x = a.c;
    while (x) { t[n++] = (char)('0' + (int)(x % 10)); x /= 10; }
    int p = 0;
    if (a.sign && p + 1 < cap) out[p++] = '-';
    while (n && p + 1 < cap) out[p++] = t[--n];
    snprintf(out + p, (size_t)(cap - p), "e%d", a.exp);
}
__host__ __device__ static inline CdaDec dec_apply(int op, CdaDec a, CdaDec b, int *err) {
    switch (op) {
        case '+': return cda_dec_add(a, b);
        case '-': return cda_dec_sub(a, b);
        case '*': return cda_dec_mul(a, b, err);
        case '/': return cda_dec_div(a, b, err);
        default: { CdaDec r = cda_dec_zero(); const int c = cda_dec_cmp(a, b); r.c = c ? 1 : 0; r.sign = c < 0; return r; }   // 'c': -1 / 0 / +1
    }
}
__global__ void cda_dec_selftest_kernel(int op, int n, const CdaDec *a, const CdaDec *b, CdaDec *out, int *err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int e = 0;
    out[i] = dec_apply(op, a[i], b[i], &e);
    if (e) atomicAdd(err, e);
}
int cda_debug_dec_op(int32_t op, const char *a, const char *b, char *out, int32_t cap, int32_t *range_err) {
    if (!a || !b || !out || cap < 48) return CDA_EINVAL;
    int e = 0;
    dec_format(dec_apply(op, dec_parse(a), dec_parse(b), &e), out, cap);
    if (range_err) *range_err = e;
    return CDA_OK;
}
int cda_debug_dec_op_device(int32_t op, int32_t n, const char *const *a, const char *const *b, char *out, int32_t cap, int32_t *range_err) {
    if (!a || !b || !out || n < 1 || cap < 48) return CDA_EINVAL;
    std::vector<CdaDec> ha((size_t)n), hb((size_t)n), ho((size_t)n);
    for (int i = 0; i < n; ++i) { ha[i] = dec_parse(a[i]); hb[i] = dec_parse(b[i]); }
    CdaDec *da = nullptr, *db = nullptr, *dout = nullptr; int *derr = nullptr, herr = 0;
    const size_t bytes = (size_t)n * sizeof(CdaDec);
    CUDA_TRY(cudaMalloc(&da, bytes)); CUDA_TRY(cudaMalloc(&db, bytes)); CUDA_TRY(cudaMalloc(&dout, bytes)); CUDA_TRY(cudaMalloc(&derr, sizeof(int)));
    CUDA_TRY(cudaMemcpy(da, ha.data(), bytes, cudaMemcpyHostToDevice)); CUDA_TRY(cudaMemcpy(db, hb.data(), bytes, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(derr, 0, sizeof(int)));
    cda_dec_selftest_kernel<<<(n + 127) / 128, 128>>>(op, n, da, db, dout, derr);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(ho.data(), dout, bytes, cudaMemcpyDeviceToHost)); CUDA_TRY(cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(derr);
    for (int i = 0; i < n; ++i) dec_format(ho[i], out + (size_t)i * cap, cap);
    if (range_err) *range_err = herr;
    return CDA_OK;
}

// measurement only: the resident step server leaves per-market milestone times (globaltimer ns) of the LAST served step in rows of 16 u64
// (0 message seen, 1 actions here, 2 step computed, 3 outputs fenced); row M: 0 completion rung, 1 poller read the message.  NULL = off.
unsigned long long *cda_debug_serve_timeline(int32_t markets) {
    if (markets <= 0) { g_srv_prof = nullptr; return nullptr; }
    unsigned long long *b = nullptr;
    if (cudaMalloc(&b, (size_t)(markets + 1) * 16 * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
    cudaMemset(b, 0, (size_t)(markets + 1) * 16 * sizeof(unsigned long long));
    g_srv_prof = b;
    return b;
}
}  // extern "C"
