/* cda_b200_testing.h — TEST / MEASUREMENT entry points of libcda_b200.so.  NOT part of the product ABI (include/cda_b200.h):
 * nothing a user of the environment calls.  tests/ and tools/ bind them; they are exported from the same shared library because
 * only one library ships with the package. */
#ifndef CDA_B200_TESTING_H
#define CDA_B200_TESTING_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Test entries for csrc/cda_dec128.cuh — Decimal(prec 28, ROUND_HALF_EVEN) arithmetic on fixed-width integers, the form the
 * reference's Decimal ledger (envs/account/account.py:124-231, calculate.py:5-55) will take on the device.
 * op: '+', '-', '*', '/' or 'c' (compare: "-1e0" / "0" / "1e0"); operands and results are decimal strings ("-396e0",
 * "3.5e-24"); result slots are `cap` (>= 48) bytes each.  *range_err counts products / quotients outside the 128-bit domain.
 * cda_debug_dec_op runs the HOST compilation of the header, cda_debug_dec_op_device the DEVICE one (n operand pairs, one
 * thread each).  Neither is on a product path. */
int cda_debug_dec_op(int32_t op, const char *a, const char *b, char *out, int32_t cap, int32_t *range_err);
int cda_debug_dec_op_device(int32_t op, int32_t n, const char *const *a, const char *const *b, char *out, int32_t cap, int32_t *range_err);

/* Debug builds only (-DCDA_PROFILE_PHASES): device buffer of 16 per-phase cycle sums, else NULL. */
unsigned long long *cda_debug_phase_buffer(void);

/* Timing experiments on the host-window path (tools/e2e_timeline.py, tools/e2e_scale_diag.py): bit 0 = reuse the actions already
 * staged on the device (no input transfer after the first call), bit 1 = keep the outputs on the device (no output transfer).
 * Process-wide; initialised from $CDA_DEBUG_WINDOW. */
void cda_debug_set_window_mode(int32_t mode);

/* decimal_ledger: how many tie-resolution passes (a market's steps re-executed with a parked answer, csrc/cda_kernels.cuh `resolve`) the step
 * kernels of this process have made so far on the current device; -1 on error.  Synchronises. */
int64_t cda_debug_restart_count(void);

/* Resident step server (cda_serve_*): device buffer u64[markets + 1][16] in which every warp leaves the globaltimer (ns) of the milestones
 * of the LAST served step — 0 message seen, 1 actions here, 2 step computed, 3 outputs fenced; row `markets`: 0 completion rung, 1 poller read
 * the host's message.  Applies to launches made after the call; markets <= 0 switches it off.  tools/serve_timeline.py. */
unsigned long long *cda_debug_serve_timeline(int32_t markets);

#ifdef __cplusplus
}
#endif
#endif /* CDA_B200_TESTING_H */
