"""VecCDAEnv — tensor API over M independent continuous-double-auction markets on one B200.

One `step` == one fused CUDA kernel launch stepping every market once (all A agents act):
the whole `continuousDoubleAuctionEnv.step` of the reference
(gym_continuousDoubleAuction/envs/continuousDoubleAuction_env.py:265-309) per market.
PyTorch only supplies device buffers and the stream; the work is in csrc/ behind the C-ABI.
"""
import ctypes

import numpy as np
import torch

from . import _native
from .config import SNAPSHOT_DIM, resolve

STATUS_BITS = {1: "order pool overflow", 2: "fill log overflow", 4: "bad action",
               8: "price out of range", 16: "bad order size"}
FATAL_STATUS = 1 | 4 | 8 | 16      # the reference calls sys.exit() / corrupts nothing here; bit 2 only truncates the per-step fill LOG


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class StackedPlanes:
    """Stacked observation of every market as n_hist zero-copy [M, 42] views (oldest snapshot first) of the pinned plane ring.
    Indexing / np.asarray() give the reference's layout: obs[m] is the f32[n_hist*42] vector of market m (a copy)."""
    __slots__ = ("planes",)

    def __init__(self, planes):
        self.planes = planes

    @property
    def shape(self):
        return (self.planes[0].shape[0], len(self.planes) * SNAPSHOT_DIM)

    def stacked(self, out=None):
        M, W = self.shape
        out = np.empty((M, W), np.float32) if out is None else out
        for j, p in enumerate(self.planes):
            out[:, j * SNAPSHOT_DIM:(j + 1) * SNAPSHOT_DIM] = p
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.stacked()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __getitem__(self, idx):
        if isinstance(idx, tuple) and len(idx) == 2 and isinstance(idx[0], (int, np.integer)) and isinstance(idx[1], (int, np.integer)):
            j, c = divmod(int(idx[1]) % self.shape[1], SNAPSHOT_DIM)
            return self.planes[j][idx[0], c]
        if isinstance(idx, (int, np.integer)):
            return np.concatenate([p[idx] for p in self.planes])
        return self.stacked()[idx]


class VecCDAEnv:
    def __init__(self, config=None, num_markets=1, device=0, order_capacity=0, fill_capacity=0, status_policy="raise", decimal_ledger=False, fill_tape=False):
        """fill_tape: with fill_capacity > 0 the fill log becomes a tape — the last fill_capacity fills of every market across steps and
        fused-rollout launches (the reference's LOB.tape, bounded) instead of the last step's fills; see tape().
        decimal_ledger: besides the exact int64 ledger the env carries the reference's Decimal(prec 28) residues (VWAP and cash,
        csrc/cda_twin.cuh: event journal + deferred replay, off the step's critical path), so that a cash gate or bankruptcy test that
        lands on EXACT integer equality is decided like the reference's Decimal compare (agent/trader.py:108-151) — results are then
        identical to the reference's unconditionally.  It costs a journal replay kernel every few steps (measured in DESIGN.md §4.3),
        so the tensor API leaves it off by default: integers only, identical trajectories except at such ties (about one agent-step
        in 10^5..10^6 when cash is of the order of single order values; none observed at the default cash of 10^6).  The dict
        adapters (continuousDoubleAuctionEnv, VectorCDAEnv) — the drop-in surface — switch it on by default.
        status_policy: what step*/reset* do when a market carries a sticky status bit (see STATUS_BITS; noticed through a
        pinned flag word the step kernel sets, i.e. at the first call after the offending step has completed — no launch and
        no synchronisation while all markets are clean): "raise" RuntimeError on pool overflow / bad action / price range /
        bad size (book and ledger of that market no longer follow the reference), "warn" once per bit, or "ignore".
        A fill-log overflow (bit 2) only truncates the per-step log — book and ledger stay exact — and never raises."""
        if status_policy not in ("raise", "warn", "ignore"):
            raise ValueError("status_policy must be 'raise', 'warn' or 'ignore'")
        self.status_policy = status_policy
        self._status_warned = 0
        if not torch.cuda.is_available():
            raise RuntimeError("VecCDAEnv needs a CUDA device (sm_100a); there is no CPU fallback")
        cfg = resolve(config)
        self.config = cfg
        self.num_markets = self.M = int(num_markets)
        self.num_of_agents = self.A = int(cfg["num_of_agents"])
        self.n_hist = int(cfg["n_hist"])
        self.max_step = int(cfg["max_step"])
        self.init_cash = cfg["init_cash"]
        self.obs_dim = self.W = self.n_hist * SNAPSHOT_DIM
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self._L = _native.lib()
        c = _native.CdaConfig(
            self.A, self.n_hist, self.max_step, int(cfg["tick_size"]), int(cfg["init_cash"]),
            int(cfg["min_size"]), int(cfg["mkt_max_size"]), int(cfg["limit_size_multiple"]),
            int(cfg["initial_price_min"]), int(cfg["initial_price_max"]), int(order_capacity),
            int(fill_capacity), float(cfg["order_penalty"]), float(cfg["trade_penalty"]),
            float(cfg["drawdown_penalty"]), float(cfg["passive_bonus"]), float(cfg["loss_multiplier"]), 1 if decimal_ledger else 0, 1 if fill_tape else 0)
        self.fill_tape = bool(fill_tape) and int(fill_capacity) > 0
        self.decimal_ledger = bool(decimal_ledger)
        h = ctypes.c_void_p()
        _native.check(self._L.cda_create(ctypes.byref(c), self.M, self.device.index, ctypes.byref(h)))
        self._h = h
        self._status_flag = self._L.cda_status_flag(h)          # POINTER(c_uint32) into pinned host memory
        self.fill_capacity = int(fill_capacity)
        self.order_capacity = self._L.cda_order_capacity(h)
        with torch.cuda.device(self.device):
            self.obs = torch.zeros((self.M, self.W), dtype=torch.float32, device=self.device)
            self.reward = torch.zeros((self.M, self.A), dtype=torch.float64, device=self.device)
            self.terminated = torch.zeros(self.M, dtype=torch.uint8, device=self.device)
            self.truncated = torch.zeros(self.M, dtype=torch.uint8, device=self.device)
        self._pinned = None

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._L.cda_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ sticky status
    def _poll_status(self):
        """Called by every step / reset entry: free while all markets are clean (one read of a pinned host word)."""
        if self._status_flag[0] and self.status_policy != "ignore":
            self._on_status()

    def _on_status(self):
        self._L.cda_status_flag_clear(self._h)
        bits = 0
        for v in torch.unique(self.status()).cpu().tolist():
            bits |= int(v)
        names = ", ".join(v for b, v in STATUS_BITS.items() if bits & b)
        if (bits & FATAL_STATUS) and self.status_policy == "raise":
            raise RuntimeError("cda_b200 market status: " + names + " (the affected markets no longer follow the reference; "
                               "reset them, raise order_capacity, or construct the env with status_policy='warn')")
        new = bits & ~self._status_warned
        if new:
            import warnings
            warnings.warn("cda_b200 market status: " + ", ".join(v for b, v in STATUS_BITS.items() if new & b))
            self._status_warned |= new

    # ------------------------------------------------------------------ reset
    def _seed_mask_tensors(self, seed, mask):
        """seed: None | int (market m gets seed + m) | one non-negative integer per market; mask: None | bool/uint8 [M].
        Returns device tensors (u64 bits as int64 [M] or None, uint8 [M] or None) — the ONE place the reset paths convert
        and validate their arguments (cda_reset_kernel indexes seeds[m] / mask[m] for every market)."""
        seeds_t = None
        if seed is not None:
            if isinstance(seed, (int, np.integer)):
                if int(seed) < 0:
                    raise ValueError("seed must be non-negative")
                seeds = np.arange(self.M, dtype=np.uint64) + np.uint64(seed)
            else:
                raw = np.asarray(seed.cpu() if isinstance(seed, torch.Tensor) else seed)
                if raw.shape != (self.M,) or raw.dtype.kind not in "iu" or (raw.dtype.kind == "i" and (raw < 0).any()):
                    raise ValueError("seed must be None, a non-negative int, or one non-negative integer seed per market")
                seeds = raw.astype(np.uint64)
            seeds_t = torch.from_numpy(seeds.view(np.int64)).to(self.device)
        mask_t = None
        if mask is not None:
            mask_t = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
            if mask_t.shape != (self.M,):
                raise ValueError("mask must have one entry per market")
        return seeds_t, mask_t

    def reset(self, seed=None, mask=None):
        """seed: None (keep every market's stream, reference `reset(seed=None)`), an int (market m
        is seeded with seed + m), or a length-M sequence/tensor of per-market integer seeds.
        mask: optional bool/uint8 [M] selecting the markets to reset.  Returns obs [M, W]."""
        seeds_t, mask_t = self._seed_mask_tensors(seed, mask)
        _native.check(self._L.cda_reset(self._h, _ptr(seeds_t), _ptr(mask_t), _ptr(self.obs), self._stream()))
        return self.obs

    # ------------------------------------------------------------------ step (device tensors)
    def step(self, category, size_mean, size_sigma, price, price_offset, out=None):
        """All inputs are CUDA tensors shaped [M, A] (int32 / float32).  Returns
        (obs f32[M,W], reward f64[M,A], terminated u8[M], truncated u8[M]) — tensors owned by the
        env and overwritten by the next step unless `out` (a 4-tuple) is given."""
        for t, dt in ((category, torch.int32), (size_mean, torch.float32), (size_sigma, torch.float32),
                      (price, torch.int32), (price_offset, torch.int32)):
            if t.dtype != dt or not t.is_cuda or not t.is_contiguous() or t.numel() != self.M * self.A:
                raise ValueError("actions must be contiguous CUDA tensors [M, A] of int32/float32")
        obs, rew, term, trunc = out if out is not None else (self.obs, self.reward, self.terminated, self.truncated)
        if self._status_flag[0]:
            self._poll_status()
        _native.check(self._L.cda_step(self._h, _ptr(category), _ptr(size_mean), _ptr(size_sigma), _ptr(price),
                                       _ptr(price_offset), _ptr(obs), _ptr(rew), _ptr(term), _ptr(trunc),
                                       self._stream()))
        return obs, rew, term, trunc

    # ------------------------------------------------------------------ step (host buffers, e2e)
    def _ensure_pinned(self):
        """Pinned host buffers laid out so that cda_step_host needs ONE H2D and ONE D2H copy:
        actions = one [5, M, A] 4-byte block; outputs = obs | reward | terminated | truncated."""
        if self._pinned is None:
            M, A, W = self.M, self.A, self.W
            act = torch.empty((5, M, A), dtype=torch.int32, pin_memory=True)
            nb_obs, nb_rew = M * W * 4, M * A * 8
            out = torch.empty(nb_obs + nb_rew + 2 * M, dtype=torch.uint8, pin_memory=True)
            self._pinned = dict(
                act=act, out=out,
                cat=act[0], mean=act[1].view(torch.float32), sigma=act[2].view(torch.float32), price=act[3], off=act[4],
                obs=out[:nb_obs].view(torch.float32).view(M, W),
                reward=out[nb_obs:nb_obs + nb_rew].view(torch.float64).view(M, A),
                term=out[nb_obs + nb_rew:nb_obs + nb_rew + M], trunc=out[nb_obs + nb_rew + M:])
            pp = self._pinned
            self._out_ptrs = tuple(_ptr(pp[k]) for k in ("obs", "reward", "term", "trunc"))
            self._out_np = (pp["obs"].numpy(), pp["reward"].numpy(), pp["term"].numpy(), pp["trunc"].numpy())
        return self._pinned

    def step_host(self, category, size_mean, size_sigma, price, price_offset, sync=True):
        """Host arrays in, host arrays out (numpy views of pinned buffers): the H2D copy of the
        actions, the kernel and the D2H copy of obs/reward/flags are all inside this call."""
        p = self._ensure_pinned()
        srcs = (category, size_mean, size_sigma, price, price_offset)
        dts = (torch.int32, torch.float32, torch.float32, torch.int32, torch.int32)
        if all(isinstance(s, torch.Tensor) and s.is_pinned() and s.is_contiguous() and s.dtype == d
               and s.numel() == self.M * self.A for s, d in zip(srcs, dts)):
            # zero-copy: the caller's pinned tensors are the H2D source
            _native.check(self._L.cda_step_host(self._h, *[_ptr(s) for s in srcs], _ptr(p["obs"]), _ptr(p["reward"]),
                                                _ptr(p["term"]), _ptr(p["trunc"]), self._stream()))
            if sync:
                torch.cuda.current_stream(self.device).synchronize()
            return p["obs"].numpy(), p["reward"].numpy(), p["term"].numpy(), p["trunc"].numpy()
        for key, src in zip(("cat", "mean", "sigma", "price", "off"), srcs):
            if isinstance(src, torch.Tensor):
                p[key].copy_(src.reshape(self.M, self.A))
            else:
                np.copyto(p[key].numpy(), np.asarray(src).reshape(self.M, self.A), casting="same_kind")
        return self.step_pinned(sync=sync)

    def step_host_block(self, action_block, sync=True):
        """Lowest-overhead host path: `action_block` is ONE pinned int32 tensor [5, M, A] holding
        category, size_mean (float32 bits), size_sigma (float32 bits), price, price_offset.  The kernel
        reads it in place (mapped pinned memory) and writes obs/reward/flags into this env's pinned
        output block; returns numpy views of that block."""
        p = self._ensure_pinned()
        if self._status_flag[0]:
            self._poll_status()
        base = action_block.data_ptr()
        n = self.M * self.A * 4
        vp = ctypes.c_void_p
        _native.check(self._L.cda_step_host(self._h, vp(base), vp(base + n), vp(base + 2 * n), vp(base + 3 * n), vp(base + 4 * n),
                                            self._out_ptrs[0], self._out_ptrs[1], self._out_ptrs[2], self._out_ptrs[3], self._stream()))
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        return self._out_np

    # ---- mirrored host ring: the kernel ships only the newest snapshot; the stacked obs is a strided view
    def _ensure_ring(self):
        if getattr(self, "_ring", None) is None:
            M, A, H = self.M, self.A, self.n_hist
            self._ring = torch.empty((M, 2 * H * SNAPSHOT_DIM), dtype=torch.float32, pin_memory=True)
            tail = torch.empty(M * A * 8 + 2 * M, dtype=torch.uint8, pin_memory=True)
            self._ring_rew = tail[:M * A * 8].view(torch.float64).view(M, A)
            self._ring_term, self._ring_trunc = tail[M * A * 8:M * A * 8 + M], tail[M * A * 8 + M:]
            self._ring_np = self._ring.numpy()
            self._ring_out_np = (self._ring_rew.numpy(), self._ring_term.numpy(), self._ring_trunc.numpy())
            self._ring_ptrs = tuple(_ptr(t) for t in (self._ring, self._ring_rew, self._ring_term, self._ring_trunc))
            self._ring_pos = -1
        return self._ring

    def _ring_view(self):
        s0 = ((self._ring_pos % self.n_hist) + 1) * SNAPSHOT_DIM
        return self._ring_np[:, s0:s0 + self.W]          # [M, n_hist*42], oldest snapshot first, zero-copy

    def reset_host_ring(self, seed=None, mask=None):
        """reset() for the ring host path: returns the stacked observation as a view of the pinned ring."""
        self._ensure_ring()
        seeds_t, mask_t = self._seed_mask_tensors(seed, mask)
        _native.check(self._L.cda_reset_host_ring(self._h, _ptr(seeds_t), _ptr(mask_t), self._ring_ptrs[0], self._stream()))
        torch.cuda.current_stream(self.device).synchronize()
        return self._ring_view()

    def step_host_ring(self, action_block, sync=True):
        """Like step_host_block, but only the newest 42-float snapshot of every market crosses PCIe (twice,
        mirrored); the returned obs is a strided numpy view [M, n_hist*42] of the pinned ring (row stride
        2*n_hist*42 floats), bit-identical to step_host_block's obs.  Valid until the next call."""
        self._ensure_ring()
        self._ring_pos += 1
        base = action_block.data_ptr()
        n = self.M * self.A * 4
        vp = ctypes.c_void_p
        _native.check(self._L.cda_step_host_ring(self._h, vp(base), vp(base + n), vp(base + 2 * n), vp(base + 3 * n), vp(base + 4 * n),
                                                 self._ring_ptrs[0], self._ring_ptrs[1], self._ring_ptrs[2], self._ring_ptrs[3],
                                                 ctypes.c_int64(self._ring_pos), self._stream()))
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        return (self._ring_view(),) + self._ring_out_np

    # ---- sliding host window: only the newest snapshot of every market crosses PCIe (one strided DMA)
    WINDOW_SLOTS = 32

    def _ensure_window(self):
        if getattr(self, "_win", None) is None:
            M, A = self.M, self.A
            self._win = torch.empty((M, self.WINDOW_SLOTS * SNAPSHOT_DIM), dtype=torch.float32, pin_memory=True)
            rs = self._L.cda_record_bytes(self._h)             # packed record: reward f64[A] | terminated u8 | truncated u8 | pad (multiple of 64 B)
            self._win_rec = torch.zeros(M * rs, dtype=torch.uint8, pin_memory=True)
            rec = self._win_rec.numpy()
            self._win_np = self._win.numpy()
            self._win_out_np = (np.ndarray((M, A), np.float64, rec, 0, (rs, 8)), np.ndarray((M,), np.uint8, rec, 8 * A, (rs,)),
                                np.ndarray((M,), np.uint8, rec, 8 * A + 1, (rs,)))
            self._win_ptrs = (_ptr(self._win), _ptr(self._win_rec))
            self._win_pos = None
            H, W = self.n_hist, self.W
            # The result record rides behind the newest snapshot (CDA_WIN_INLINE_RECORD: head of slot pos+1, same PCIe write
            # transactions as the snapshot's tail) whenever it fits into a slot; the last slot is then only ever a record carrier.
            self._win_inline = 2 * A + 2 <= SNAPSHOT_DIM and self.WINDOW_SLOTS > H
            self._win_last = self.WINDOW_SLOTS - (2 if self._win_inline else 1)     # last slot position a snapshot may take
            # one precomputed result tuple per slot position: (obs view [M, W], reward [M, A], terminated [M], truncated [M])
            self._win_views = [None] * self.WINDOW_SLOTS
            row = self.WINDOW_SLOTS * SNAPSHOT_DIM * 4                               # bytes per market row
            for pos in range(H - 1, self._win_last + 1):
                s0 = (pos - H + 1) * SNAPSHOT_DIM
                if self._win_inline:
                    o = (pos + 1) * SNAPSHOT_DIM * 4
                    outs = (np.ndarray((M, A), np.float64, self._win_np, o, (row, 8)), np.ndarray((M,), np.uint8, self._win_np, o + 8 * A, (row,)),
                            np.ndarray((M,), np.uint8, self._win_np, o + 8 * A + 1, (row,)))
                else:
                    outs = self._win_out_np
                self._win_views[pos] = (self._win_np[:, s0:s0 + W],) + outs
        return self._win

    def _window_view(self):
        s0 = (self._win_pos - self.n_hist + 1) * SNAPSHOT_DIM
        return self._win_np[:, s0:s0 + self.W]          # [M, n_hist*42] oldest snapshot first; rows are contiguous

    def reset_host_window(self, seed=None, mask=None):
        """reset() for the window host path: returns the stacked observation as a view of the pinned window."""
        self._ensure_window()
        seeds_t, mask_t = self._seed_mask_tensors(seed, mask)
        stream = self._stream()
        _native.check(self._L.cda_reset_host_window(self._h, _ptr(seeds_t), _ptr(mask_t), self._win_ptrs[0], self.WINDOW_SLOTS, stream))
        # tight-loop form: buffers + the CURRENT stream are bound once; step_host_window then makes a 4-argument call
        _native.check(self._L.cda_window_bind(self._h, self._win_ptrs[0], self.WINDOW_SLOTS, self._win_ptrs[1], stream))
        self._win_step = self._L.cda_step_window
        self._win_pos = self.n_hist - 1
        return self._window_view()

    def attach_host_window(self):
        """Start (or re-synchronise) the host window from the device state WITHOUT resetting any market — e.g. after
        stepping through another path.  Returns the stacked observation view."""
        return self.reset_host_window(seed=None, mask=np.zeros(self.M, dtype=np.uint8))

    def step_host_window(self, action_block, sync=True, market_major=False):
        """Lowest-traffic host path.  `action_block` as in step_host_block (ONE pinned int32 tensor [5, M, A], read in
        place by the kernel) or, with market_major=True, [M, 5, A] (one 20*A-byte action record per market: a CTA's
        markets are then fetched over PCIe by one bulk copy instead of five).  The reward / terminated / truncated
        views are valid until the next call, like the observation (they live behind the newest snapshot in the window).
        Per step only the newest 42-float snapshot of every market crosses PCIe, into the
        next slot of that market's row of a pinned [M, 32, 42] window; the returned obs is the numpy view
        [M, n_hist*42] of the n_hist most recent slots (row stride 32*42 floats, each row contiguous), bit-identical
        to step_host_block's obs.  Views are valid until the next call.  Launch and stream synchronisation happen
        inside ONE C call, on the stream that was current when reset_host_window()/attach_host_window() was called."""
        pos = getattr(self, "_win_pos", None)
        if pos is None:
            raise RuntimeError("call reset_host_window() before step_host_window()")
        pos += 1
        if pos > self._win_last:
            pos = self.n_hist - 1        # window restarts: the whole stack is re-sent into slots 0..n_hist-1
        rc = self._win_step(self._h, action_block.data_ptr(), pos, (1 if sync else 0) | (2 if market_major else 0) | (4 if self._win_inline else 0))
        if rc:
            _native.check(rc)
        self._win_pos = pos
        if self._status_flag[0]:
            self._poll_status()
        return self._win_views[pos]

    # ---- dense plane ring: each step's output is ONE contiguous, line-aligned region (the host path that scales to 8 GPUs / node)
    PLANE_SLOTS = 8
    PLANE_CELL_WORDS = 0           # words per market and step: 42-float snapshot | reward f64[A] | terminated | truncated; 0 = exactly that (52 words for 4
                                   # agents: bytes are what an 8-GPU node runs out of, profiles/r03f_e2e_scale_diag_8gpu_layout.txt), or a larger even size

    def _ensure_planes(self):
        if getattr(self, "_planes", None) is None:
            M, A, S = self.M, self.A, self.PLANE_SLOTS
            need = SNAPSHOT_DIM + 2 * A + 2
            cell = self.PLANE_CELL_WORDS if self.PLANE_CELL_WORDS >= need and self.PLANE_CELL_WORDS % 2 == 0 else need + (need & 1)
            self._plane_cell = cell
            self._planes = torch.zeros((S, M, cell), dtype=torch.float32, pin_memory=True)
            pn = self._planes.numpy()
            self._planes_np = pn
            self._plane_ptrs = [self._planes[s].data_ptr() for s in range(S)]
            cb = cell * 4
            self._plane_snap = [pn[s, :, :SNAPSHOT_DIM] for s in range(S)]                                     # [M, 42] view per slot
            self._plane_out = [(np.ndarray((M, A), np.float64, pn, (s * M * cell + SNAPSHOT_DIM) * 4, (cb, 8)),
                                np.ndarray((M,), np.uint8, pn, (s * M * cell + SNAPSHOT_DIM) * 4 + 8 * A, (cb,)),
                                np.ndarray((M,), np.uint8, pn, (s * M * cell + SNAPSHOT_DIM) * 4 + 8 * A + 1, (cb,))) for s in range(S)]
            H = self.n_hist
            # one precomputed result tuple per ring position (the per-step call then only indexes a list)
            self._plane_results = [(StackedPlanes([self._plane_snap[(s - H + 1 + j) % S] for j in range(H)]),) + self._plane_out[s] for s in range(S)]
            self._plane_step = self._L.cda_step_planes
            self._plane_pos = None
        return self._planes

    def _plane_stack(self):
        """The stacked observation as n_hist [M, 42] views, oldest snapshot first."""
        return self._plane_results[self._plane_pos][0]

    def reset_host_planes(self, seed=None, mask=None):
        """reset() for the dense-plane host path (see include/cda_b200.h cda_step_planes): returns a StackedPlanes."""
        self._ensure_planes()
        if self.PLANE_SLOTS < self.n_hist + 1:
            raise ValueError("PLANE_SLOTS must exceed n_hist")
        seeds_t, mask_t = self._seed_mask_tensors(seed, mask)
        pos = self._plane_pos if self._plane_pos is not None else self.n_hist - 1
        _native.check(self._L.cda_reset_planes(self._h, _ptr(seeds_t), _ptr(mask_t), ctypes.c_void_p(self._plane_ptrs[0]), self.PLANE_SLOTS,
                                               self._plane_cell, pos, self._stream()))
        self._plane_pos = pos
        self._plane_stream = self._stream()
        return self._plane_stack()

    def attach_host_planes(self):
        """Start (or re-synchronise) the plane ring from the device state without resetting any market."""
        return self.reset_host_planes(seed=None, mask=np.zeros(self.M, dtype=np.uint8))

    def step_host_planes(self, action_block, sync=True, market_major=True):
        """Dense-plane host path.  `action_block`: ONE pinned int32 tensor, [M, 5, A] (market_major) or [5, M, A], read in place by
        the kernel.  Returns (StackedPlanes obs, reward f64[M, A], terminated u8[M], truncated u8[M]) — views of the pinned plane
        ring, valid until PLANE_SLOTS - n_hist further steps have been made.  `np.asarray(obs)` / `obs.stacked()` materialises the
        contiguous f32[M, n_hist*42] array (one host copy); `obs.planes` are the n_hist zero-copy [M, 42] views, oldest first.
        With `host_resident = True` (see serve()) the market-major synchronous form goes through the resident step server."""
        pos = self._plane_pos
        if pos is None:
            raise RuntimeError("call reset_host_planes() before step_host_planes()")
        pos = (pos + 1) % self.PLANE_SLOTS
        rc = -5
        if self._serve_on and sync and market_major:
            rc = self._serve_step(self._h, action_block.data_ptr(), pos, self._plane_stream)
            if rc == -5:       # CDA_EUNSUPPORTED: the server was being relaunched for most steps (slow policy / shared GPU) and has switched itself off
                self._serve_on = False
        if rc == -5:
            rc = self._plane_step(self._h, action_block.data_ptr(), self._plane_ptrs[pos], self._plane_cell,
                                  (1 if sync else 0) | (2 if market_major else 0), self._plane_stream)
        if rc:
            _native.check(rc)
        self._plane_pos = pos
        if self._status_flag[0]:
            self._poll_status()
        return self._plane_results[pos]

    # ---- resident step server (include/cda_b200.h, cda_serve_*): the plane path without a launch per step
    _serve_on = False

    def serve(self, on=True):
        """Switch the RESIDENT STEP SERVER on / off for step_host_planes(market_major=True): the step kernel is launched once and stays
        on the SMs with every market's book and ledger in shared memory; a step is then one doorbell write by the host (no launch, no
        stream hand-shake, no state round trip through HBM) and the same outputs in the same planes.  For HOST-side policies: while
        the server is resident (until 2 ms after the last step, or any other call on this env) it occupies the whole GPU.
        Returns True when the mode is active, False when this env cannot use it (decimal_ledger, more markets than one resident
        wave holds, unmapped buffers) and step_host_planes keeps launching per step."""
        if not on:
            if self._serve_on:
                _native.check(self._L.cda_serve_stop(self._h))
            self._serve_on = False
            return False
        self._ensure_planes()
        rc = self._L.cda_serve_bind(self._h, ctypes.c_void_p(self._plane_ptrs[0]), self.PLANE_SLOTS, self._plane_cell)
        if rc == -5:     # CDA_EUNSUPPORTED
            self._serve_on = False
            return False
        _native.check(rc)
        self._serve_step = self._L.cda_serve_step
        self._serve_on = True
        return True

    host_resident = property(lambda self: self._serve_on, lambda self, v: self.serve(bool(v)))

    def serve_stop(self):
        """Retire the resident kernel now (state back in HBM, SMs free); the next step_host_planes launches it again."""
        _native.check(self._L.cda_serve_stop(self._h))

    @property
    def serve_launches(self):
        return int(self._L.cda_serve_launches(self._h))

    def step_pinned(self, sync=True):
        """Like step_host but the caller has already written the actions into `pinned_buffers()`."""
        p = self._ensure_pinned()
        if self._status_flag[0]:
            self._poll_status()
        _native.check(self._L.cda_step_host(self._h, _ptr(p["cat"]), _ptr(p["mean"]), _ptr(p["sigma"]), _ptr(p["price"]),
                                            _ptr(p["off"]), _ptr(p["obs"]), _ptr(p["reward"]), _ptr(p["term"]),
                                            _ptr(p["trunc"]), self._stream()))
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        return p["obs"].numpy(), p["reward"].numpy(), p["term"].numpy(), p["trunc"].numpy()

    def pinned_buffers(self):
        return self._ensure_pinned()

    # ------------------------------------------------------------------ fused step + all-gather (multi-GPU)
    GATHER_SLOTS = 32

    def enable_peer_gather(self, group=None):
        """Set up the NVLink peer-memory gather (one process per GPU, torch.distributed initialised; call after reset(), on every rank).
        Every rank gets a gather WINDOW (G*M rows of 32 snapshot slots + two result records) that all ranks' step kernels fill directly.  Returns the stacked
        observations of all G*M markets (see step_gather)."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        handle = (ctypes.c_ubyte * 64)()
        ptr, nbytes = ctypes.c_void_p(), ctypes.c_uint64()
        with torch.cuda.device(self.device):
            _native.check(self._L.cda_gather_create(self._h, world, rank, handle, ctypes.byref(ptr), ctypes.byref(nbytes)))
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle), group=group)
            blob = (ctypes.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
            _native.check(self._L.cda_gather_connect(self._h, blob))
        rows = world * self.M

        class _Raw:   # expose the cudaMalloc'ed buffer to torch without copying
            def __init__(s, p, n):
                s.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (p, False), "version": 2}
        raw = torch.as_tensor(_Raw(ptr.value, nbytes.value), device=self.device)
        self._gather_row = self._L.cda_gather_row_words(self._h)        # 32 snapshot slots + two result records, in 4-byte words
        wbytes = rows * self._gather_row * 4
        self._gather_raw = raw
        self._gather_f32 = raw[:wbytes].view(torch.float32)
        self._gather_f64 = raw[:wbytes].view(torch.float64)
        self._gather_u8 = raw[:wbytes]
        self._gather_rows = rows
        dist.barrier(group=group)               # every window exists and is mapped everywhere
        return self.publish_gather()

    def publish_gather(self):
        """Send every local market's current stack to all ranks' windows (after reset() / a masked reset; every rank calls it)."""
        _native.check(self._L.cda_gather_publish(self._h, self._stream()))
        _native.check(self._L.cda_gather_wait(self._h, self._stream()))
        return self._gather_views()[0]

    def _gather_views(self):
        pos, S, H, A, rows = self._L.cda_gather_pos(self._h), self.GATHER_SLOTS, self.n_hist, self.A, self._gather_rows
        row_f = self._gather_row
        obs = torch.as_strided(self._gather_f32, (rows, self.W), (row_f, 1), (pos - H + 1) * SNAPSHOT_DIM)
        rec_b = (S * SNAPSHOT_DIM + self._L.cda_gather_record_parity(self._h) * (2 * A + 2)) * 4     # byte offset of this step's record inside a row
        rew = torch.as_strided(self._gather_f64, (rows, A), (row_f // 2, 1), rec_b // 8)
        term = torch.as_strided(self._gather_u8, (rows,), (row_f * 4,), rec_b + 8 * A)
        trunc = torch.as_strided(self._gather_u8, (rows,), (row_f * 4,), rec_b + 8 * A + 1)
        return obs, rew, term, trunc

    def step_gather(self, category, size_mean, size_sigma, price, price_offset, wait=True):
        """cda_step whose epilogue writes this rank's newest snapshots + result records into EVERY rank's gather window (P2P stores
        over NVLink) and publishes a completion flag to every rank; with wait=True a one-warp kernel on the current stream then waits
        for all ranks' flags (no NCCL call).  Returns (obs f32[G*M, W], reward f64[G*M, A], terminated u8[G*M], truncated u8[G*M]):
        strided CUDA views of THIS rank's window (rows contiguous), valid until the next step_gather."""
        with torch.cuda.device(self.device):
            _native.check(self._L.cda_step_gather(self._h, _ptr(category), _ptr(size_mean), _ptr(size_sigma), _ptr(price),
                                                  _ptr(price_offset), self._stream()))
            if wait:
                _native.check(self._L.cda_gather_wait(self._h, self._stream()))
        return self._gather_views()

    # ------------------------------------------------------------------ fused random rollout
    def rollout_random(self, num_steps, policy_seed=0):
        _native.check(self._L.cda_rollout_random(self._h, int(num_steps), ctypes.c_uint64(policy_seed), _ptr(self.obs),
                                                 _ptr(self.reward), _ptr(self.terminated), _ptr(self.truncated),
                                                 self._stream()))
        return self.obs, self.reward, self.terminated, self.truncated

    # ------------------------------------------------------------------ lazy info / fills
    def info(self, field):
        """One info field for all markets as an int64 CUDA tensor ([M, A], or [M, 8] for 'market')."""
        idx = _native.INFO_FIELDS.index(field)
        shape = (self.M, 8) if field == "market" else (self.M, self.A)
        out = torch.empty(shape, dtype=torch.int64, device=self.device)
        _native.check(self._L.cda_get_info(self._h, idx, _ptr(out), self._stream()))
        return out

    def info_all(self):
        """Every per-agent info field in one launch: dict name -> int64 CUDA tensor [M, A], plus
        'market' [M, 8] (columns: _native.INFO_MARKET_COLS)."""
        n = len(_native.INFO_FIELDS) - 1
        buf = torch.empty(n * self.M * self.A + self.M * 8, dtype=torch.int64, device=self.device)
        _native.check(self._L.cda_get_info_all(self._h, _ptr(buf), self._stream()))
        out = {name: buf[i * self.M * self.A:(i + 1) * self.M * self.A].view(self.M, self.A)
               for i, name in enumerate(_native.INFO_FIELDS[:-1])}
        out["market"] = buf[n * self.M * self.A:].view(self.M, 8)
        return out

    def status(self):
        return self.info("market")[:, 7]

    def check_status(self, include_fill_log=False):
        """Raise if any market carries a FATAL sticky status bit (pool overflow, bad action, price range, bad size: the
        reference would have sys.exit()ed or kept an order this library dropped).  A fill-log overflow (bit 2) only truncates
        the per-step fill log — book and ledger stay exact — and is reported only when include_fill_log=True."""
        bits = 0
        for v in torch.unique(self.status()).cpu().tolist():
            bits |= int(v)
        if not include_fill_log:
            bits &= FATAL_STATUS
        if bits:
            raise RuntimeError("cda_b200 market status: " + ", ".join(v for b, v in STATUS_BITS.items() if bits & b))
        return 0

    def decimal_fields(self, markets=None):
        """decimal_ledger: the reference's Decimal money fields of the given markets (default all), residues included, as
        {market: {"cash": [Decimal]*A, "VWAP": [...], "cash_on_hold": [...], "position_val": [...], "nav": [...]}}.  cash and VWAP are
        the device twin's values (after bringing the twins up to date); cash_on_hold is always an integer; position_val and nav are
        derived with Python's decimal exactly as the reference's mark-to-market does (calculate.py:35-55) — the reference never
        accumulates them, it recomputes them from VWAP at every fill and every mark-to-market."""
        from decimal import Decimal, localcontext
        if not self.decimal_ledger:
            raise ValueError("construct the env with decimal_ledger=True")
        buf = torch.empty((self.M, self.A, 8), dtype=torch.int64, device=self.device)
        _native.check(self._L.cda_twin_sync(self._h, _ptr(buf), self._stream()))
        tw = buf.cpu().numpy().view(np.uint64)
        info = {k: v.cpu().numpy() for k, v in self.info_all().items()}

        def dec(lo, hi, exp, sign):
            c = (int(hi) << 64) | int(lo)
            return Decimal((int(sign) & 1, tuple(int(ch) for ch in str(c)), int(np.int64(exp)))) if c else Decimal(0)
        out = {}
        with localcontext() as ctx:
            ctx.prec = 28
            for m in (range(self.M) if markets is None else markets):
                tape = int(info["market"][m, 0]), bool(info["position_val"][m].any() or info["num_trades"][m].any())
                d = {"cash": [], "VWAP": [], "cash_on_hold": [], "position_val": [], "nav": []}
                for a in range(self.A):
                    w = tw[m, a]
                    vwap = dec(w[0], w[1], w[2], w[3] & np.uint64(0xffffffff))
                    tracked = (int(w[7]) >> 40) & 1
                    cash = dec(w[4], w[5], w[6], w[7] & np.uint64(0xffffffff)) if tracked else Decimal(int(info["cash"][m, a]))
                    hold, pos, p = Decimal(int(info["cash_on_hold"][m, a])), int(info["net_position"][m, a]), Decimal(tape[0])
                    ap = Decimal(abs(pos))
                    if tape[1]:                                            # the tape is non-empty: marked to market every step
                        diff = (p - vwap) if pos >= 0 else (vwap - p)
                        pv = ap * vwap + ap * diff
                    else:
                        pv = Decimal(0)
                    d["cash"].append(cash); d["VWAP"].append(vwap); d["cash_on_hold"].append(hold); d["position_val"].append(pv)
                    d["nav"].append((cash + hold) + pv)
                out[m] = d
        return out

    def tape(self, m=0):
        """fill_tape mode: (rows, total) — the most recent min(total, fill_capacity) fills of market m in execution order, oldest first
        (rows of time, price, qty, maker, maker_order_id, maker_left, taker, taker_side), and the number of fills since the reset."""
        if not self.fill_tape:
            raise ValueError("construct the env with fill_capacity > 0 and fill_tape=True")
        f, n = self.fills()
        total, cap = int(n[m].item()), self.fill_capacity
        rows = f[m].cpu().numpy()
        if total <= cap:
            return rows[:total], total
        k = total % cap
        return np.concatenate([rows[k:], rows[:k]]), total

    def enable_action_log(self, on=True):
        """Keep the decoded actions of every step (the reference's LOB_actions): last_actions() then returns int32 [M, A, 4] =
        order type (0 market, 1 limit, 2 modify, 3 cancel), side (0 bid, 1 ask, -1 = pass / absent), size, price."""
        self._act_log = torch.full((self.M, self.A, 4), -1, dtype=torch.int32, device=self.device) if on else None
        _native.check(self._L.cda_set_action_log(self._h, _ptr(self._act_log)))

    def last_actions(self):
        if getattr(self, "_act_log", None) is None:
            raise ValueError("call enable_action_log() first")
        return self._act_log

    def fills(self):
        if not self.fill_capacity:
            raise ValueError("construct the env with fill_capacity > 0 to log fills")
        f = torch.empty((self.M, self.fill_capacity, 8), dtype=torch.int32, device=self.device)
        n = torch.empty(self.M, dtype=torch.int32, device=self.device)
        _native.check(self._L.cda_get_fills(self._h, _ptr(f), _ptr(n), self._stream()))
        return f, n

    # ------------------------------------------------------------------ canonical dump (tests)
    def dump(self, m=0, max_rows=1024):
        """Same schema as the oracle's dump: used for bit-exact parity checks."""
        bids = np.zeros((max_rows, 5), np.int64)
        asks = np.zeros((max_rows, 5), np.int64)
        bmap = np.zeros(max_rows, np.int64)
        amap = np.zeros(max_rows, np.int64)
        counts = np.zeros(2, np.int32)
        rng = np.zeros(6, np.uint64)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _native.check(self._L.cda_dump_market(self._h, int(m), vp(bids), vp(asks), vp(bmap), vp(amap), max_rows,
                                              vp(counts), vp(rng)))
        out = {"bids": bids[:counts[0]].copy(), "asks": asks[:counts[1]].copy(),
               "bids_map": bmap[:counts[0]].copy(), "asks_map": amap[:counts[1]].copy(), "rng": rng}
        mk = self.info("market")[m].cpu().numpy()
        out.update(last_price=int(mk[0]), best_bid=int(mk[1]), best_ask=int(mk[2]), time=int(mk[3]),
                   next_order_id=int(mk[4]), t_step=int(mk[5]), done_mask=int(mk[6]), status=int(mk[7]))
        cols = ("cash", "cash_on_hold", "position_val", "cost_basis", "nav", "prev_nav", "max_nav", "net_position",
                "num_trades", "num_trades_step", "num_passive_fills_step", "order_step_placed",
                "num_rejected_step", "is_pass_action")
        out["accounts"] = np.stack([self.info(c)[m].cpu().numpy() for c in cols], axis=1)
        if self.fill_capacity:
            f, n = self.fills()
            nf = int(n[m].item())
            out["n_fills"] = nf
            out["fills"] = f[m, :min(nf, self.fill_capacity)].cpu().numpy()
        return out

    def dump_all(self, markets=None):
        """Canonical dumps (same schema as dump()) of many markets from ONE checkpoint copy, parsed on the host with the layout
        cda_state_layout() reports — what the full-size parity tests use (dump() costs ~20 launches per market)."""
        lay = (ctypes.c_int32 * 12)()
        _native.check(self._L.cda_state_layout(self._h, lay))
        stride, off_acct, off_hist, off_pool, cap, A = (int(lay[i]) for i in range(6))
        raw = self.state_dict()["state"].numpy().reshape(self.M, stride)
        hdr = raw[:, :192].view(np.uint32)
        acc = np.ascontiguousarray(raw[:, off_acct:off_acct + 60 * A])
        i64 = acc[:, :48 * A].view(np.int64).reshape(self.M, 6, A)                  # cash hold cost nav prev max
        pos = acc[:, 48 * A:52 * A].view(np.int32).astype(np.int64)
        ntr = acc[:, 52 * A:56 * A].view(np.uint32).astype(np.int64)
        ctr = acc[:, 56 * A:60 * A].view(np.uint32).astype(np.int64)
        pool = np.ascontiguousarray(raw[:, off_pool:off_pool + 2 * 5 * cap * 4]).view(np.uint32).reshape(self.M, 2, cap // 32, 5, 32)
        fills = counts = None
        if self.fill_capacity:
            f, n = self.fills()
            fills, counts = f.cpu().numpy(), n.cpu().numpy()
        out = {}
        for m in (range(self.M) if markets is None else markets):
            h = hdr[m]
            d = {"time": int(h[0]), "next_order_id": int(h[1]), "t_step": int(h[3]), "last_price": int(np.int32(h[4])), "done_mask": int(h[6]),
                 "status": int(h[7]), "best_bid": int(np.int32(h[40])), "best_ask": int(np.int32(h[41])),
                 "rng": np.array([(int(h[13]) << 32) | int(h[12]), (int(h[15]) << 32) | int(h[14]), (int(h[17]) << 32) | int(h[16]),
                                  (int(h[19]) << 32) | int(h[18]), int(h[10]), int(h[11])], dtype=np.uint64)}
            for side, name in ((0, "bids"), (1, "asks")):
                n = int(h[8 + side])
                fl = pool[m, side].transpose(1, 0, 2).reshape(5, cap)[:, :n].astype(np.int64)   # field-major, live prefix
                price, trader = fl[0] & 0xffffff, fl[0] >> 24
                order = np.lexsort((fl[4], -price if side == 0 else price))                    # priority: price, then insertion seq
                d[name] = np.stack([price, fl[1], trader, fl[2], fl[3]], axis=1)[order]
                d[name + "_map"] = fl[2][np.argsort(fl[4], kind="stable")]
            tape = bool(h[5] & 1)
            lp = d["last_price"]
            ap = np.abs(pos[m])
            pv = np.where(pos[m] >= 0, ap * lp, 2 * i64[m, 2] - ap * lp) if tape else np.zeros(A, np.int64)
            c = ctr[m]
            d["accounts"] = np.stack([i64[m, 0], i64[m, 1], pv, i64[m, 2], i64[m, 3], i64[m, 4], i64[m, 5], pos[m], ntr[m], c & 0xfff,
                                      (c >> 12) & 0xfff, (c >> 24) & 1, (c >> 25) & 1, (c >> 26) & 1], axis=1)
            if fills is not None:
                d["n_fills"] = int(counts[m])
                d["fills"] = fills[m, :min(int(counts[m]), self.fill_capacity)]
            out[m] = d
        return out

    # ------------------------------------------------------------------ checkpoint
    def state_dict(self):
        nbytes = self._L.cda_state_bytes(self._h)
        buf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        _native.check(self._L.cda_save_state(self._h, _ptr(buf), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()
        return {"state": buf.clone(), "num_markets": self.M, "config": {k: v for k, v in self.config.items() if not k.startswith("_")},
                "order_capacity": self.order_capacity}

    def load_state_dict(self, sd):
        if sd["num_markets"] != self.M or sd["order_capacity"] != self.order_capacity or sd["state"].numel() != self._L.cda_state_bytes(self._h):
            raise ValueError("checkpoint does not match this env's shape")
        buf = sd["state"].contiguous().pin_memory()
        _native.check(self._L.cda_load_state(self._h, _ptr(buf), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    @property
    def kernel_launches(self):
        return int(self._L.cda_kernel_launches(self._h))
