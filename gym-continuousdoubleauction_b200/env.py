"""continuousDoubleAuctionEnv — the reference's MultiAgentEnv surface over the CUDA env.

Same constructor (`config` dict with the reference's keys, JSON fallbacks), `reset(*, seed,
options)`, `step(action_dict)`, `observation_spaces` / `action_spaces`, `agents`,
`possible_agents`, `get_*_space`, `close`, and the same `info` keys the league callback and the
episode recorder read (reference: gym_continuousDoubleAuction/envs/continuousDoubleAuction_env.py:21-359,
exchg/info_helper.py:30-116).  One object == one market, like the reference; for throughput use
VecCDAEnv (thousands of markets per launch) — this class exists so `train/`, `CDA_rand` and the
reference's tests can run unchanged on the GPU path.

Differences, all documented in INTEGRATION.md:
  * the RNG lives on the device (numpy-exact PCG64 stream), there is no `np_random` attribute;
  * normal draws are consumed in agent-index order; the reference consumes them in the action
    dict's iteration order, so pass dicts in agent order (RLlib does) for seed-level parity;
  * money is an exact int64 ledger: `info["NAV"]` is an integer string (the reference prints a
    Decimal with a ~1e-21 residue from its VWAP division; `Decimal(info["NAV"])` agrees to 1e-20).
    The residues themselves are carried by the Decimal twin (`decimal_ledger`, on by default here):
    they decide the rare exact-equality ties like the reference does; `decimal_fields()` shows them.
"""
import os
import warnings

import numpy as np

from . import config as _config

try:  # the real thing when available (not installed in the build container)
    import gymnasium as _gym
    _spaces = _gym.spaces
except Exception:  # pragma: no cover - exercised in this image
    _gym = None

    class _Space:
        def __init__(self):
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

    class _Box(_Space):
        def __init__(self, low, high, shape, dtype):
            super().__init__()
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

        def sample(self):
            if np.isinf(self.low):
                return self._rng.normal(size=self.shape).astype(self.dtype)
            return self._rng.uniform(self.low, self.high, self.shape).astype(self.dtype)

        def contains(self, x):
            return np.shape(x) == self.shape

    class _Discrete(_Space):
        def __init__(self, n):
            super().__init__()
            self.n = n

        def sample(self):
            return np.int64(self._rng.integers(0, self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

    class _Dict(_Space):
        def __init__(self, d):
            super().__init__()
            self.spaces = dict(d)

        def __getitem__(self, k):
            return self.spaces[k]

        def seed(self, seed=None):
            for sp, q in zip(self.spaces.values(), np.random.SeedSequence(seed).spawn(len(self.spaces))):
                sp._rng = np.random.default_rng(q)

        def sample(self):
            return {k: v.sample() for k, v in self.spaces.items()}

    class _spaces:  # noqa: N801
        Box, Discrete, Dict = _Box, _Discrete, _Dict

try:
    from ray.rllib.env.multi_agent_env import MultiAgentEnv as _Base
except Exception:  # pragma: no cover
    _Base = object

_INFO_INT_FIELDS = ("num_trades_step", "num_passive_fills_step", "order_step_placed", "num_rejected_step")


def _plain(v):
    if isinstance(v, np.ndarray):
        return [_plain(x) for x in v.tolist()]
    if isinstance(v, np.generic):
        return v.item()
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    return v


def build_infos(info, m, agents, rewards, actions, coeffs):
    """The reference's per-agent info dicts (exchg/info_helper.py:30-116, reward terms of reward_helper.py:75-81) of market
    `m`, from one `info_all` gather (`info`: numpy arrays).  `coeffs` carries the five reward coefficients.
    Returns (infos, pass_agents, done_agents)."""
    mk = info["market"][m]
    last_price = float(mk[0])
    best_bid = float(mk[1]) if mk[1] > 0 else None
    best_ask = float(mk[2]) if mk[2] > 0 else None
    spread = (best_ask - best_bid) if (best_bid is not None and best_ask is not None) else None
    infos, pass_agents, done = {}, set(), set()
    for i, a in enumerate(agents):
        nav, prev_nav, max_nav = int(info["nav"][m, i]), int(info["prev_nav"][m, i]), int(info["max_nav"][m, i])
        if nav <= 0:
            done.add(a)
        if info["is_pass_action"][m, i]:
            pass_agents.add(a)
        pos = int(info["net_position"][m, i])
        cost = int(info["cost_basis"][m, i])
        placed, trades_step = int(info["order_step_placed"][m, i]), int(info["num_trades_step"][m, i])
        passive = int(info["num_passive_fills_step"][m, i])
        nav_change = float(nav - prev_nav)
        dd = float(max(0, max_nav - nav))
        terms = {                                             # reward_helper.py:75-81, same IEEE ops as the kernel
            "nav_term": nav_change * (coeffs.loss_multiplier if nav_change < 0 else 1.0),
            "order_penalty": -(coeffs.order_penalty * placed),
            "trade_penalty": -(coeffs.trade_penalty * trades_step),
            "drawdown_penalty": -(coeffs.drawdown_penalty * dd),
            "passive_bonus": coeffs.passive_bonus * passive,
        }
        d = {
            "reward": rewards[a], "NAV": str(nav), "num_trades": int(info["num_trades"][m, i]),
            "net_position": pos, "VWAP": (cost / abs(pos)) if pos else 0.0,
            "cash": float(info["cash"][m, i]), "cash_on_hold": float(info["cash_on_hold"][m, i]),
            "position_val": float(info["position_val"][m, i]), "drawdown": dd, "max_nav": float(max_nav),
            "num_trades_step": trades_step, "num_passive_fills_step": passive, "order_step_placed": placed,
            "num_rejected_step": int(info["num_rejected_step"][m, i]),
            "is_pass_action": a in pass_agents, "reward_terms": terms,
            "last_price": last_price, "best_bid": best_bid, "best_ask": best_ask, "spread": spread,
        }
        if actions is not None and a in actions:
            d["model_action"] = _plain(actions[a])
        infos[a] = _plain(d)
    return infos, pass_agents, done


def pack_actions(actions, A, cat, mean, sigma, price, off, m=0):
    """One reference action dict ({agent_i: {category, size_mean, size_sigma, price, price_offset}}) -> row m of the five
    [M, A] arrays (absent agent: category -1).  Returns the agent indices in dict order."""
    cat[m, :] = -1
    idxs = []
    for key, val in actions.items():
        i = int(key.split("_")[1])
        idxs.append(i)
        cat[m, i] = int(val["category"])
        mean[m, i] = np.float32(np.asarray(val["size_mean"], dtype=np.float32).reshape(-1)[0])
        sigma[m, i] = np.float32(np.asarray(val["size_sigma"], dtype=np.float32).reshape(-1)[0])
        price[m, i] = int(val.get("price", 0))
        off[m, i] = int(val.get("price_offset", _config.PRICE_OFFSET_N // 2))
    return idxs


_SIDES, _TYPES = ("bid", "ask"), ("market", "limit", "modify", "cancel")


class _AccountView:
    """Read-only view of one agent's account on the device (the reference's `env.traders[i].acc`): the fields the callback, the
    recorder and the reference's tests read.  Money is the exact integer ledger (`str(acc.nav)` == `info["NAV"]`)."""
    _MAP = {"cash": "cash", "cash_on_hold": "cash_on_hold", "nav": "nav", "prev_nav": "prev_nav", "max_nav": "max_nav", "net_position": "net_position",
            "position_val": "position_val", "num_trades": "num_trades", "num_trades_step": "num_trades_step",
            "num_passive_fills_step": "num_passive_fills_step", "order_step_placed": "order_step_placed", "num_rejected_step": "num_rejected_step"}

    def __init__(self, env, i):
        self._env, self._i = env, i

    def __getattr__(self, name):
        if name == "VWAP":
            pos, cost = int(self._env._vec.info("net_position")[0, self._i].item()), int(self._env._vec.info("cost_basis")[0, self._i].item())
            return cost / abs(pos) if pos else 0.0
        if name in self._MAP:
            return int(self._env._vec.info(self._MAP[name])[0, self._i].item())
        raise AttributeError(name)


class _TraderView:
    def __init__(self, env, i):
        self.ID, self.acc = i, _AccountView(env, i)


class continuousDoubleAuctionEnv(_Base):
    metadata = {"render.modes": ["human"]}

    def __init__(self, config=None):
        if _Base is not object:
            try:
                super().__init__()
            except Exception:
                pass
        self.config = config or {}
        cfg = _config.resolve(self.config)
        from .vec_env import VecCDAEnv  # imports torch; needs a GPU
        self.num_of_agents = int(cfg["num_of_agents"])
        self.init_cash = cfg["init_cash"]
        self.max_step = int(cfg["max_step"])
        self.n_hist = int(cfg["n_hist"])
        self.is_render = cfg["is_render"]
        self.tick_size = cfg["tick_size"]
        self.min_tick = cfg["tick_size"]                      # action_helper.py:55-58
        self.k_rows, self.book_rows, self.extra_dim = _config.K_ROWS, _config.BOOK_ROWS, _config.EXTRA_DIM
        self.book_dim = self.k_rows * self.book_rows
        self.snapshot_dim = self.book_dim + self.extra_dim
        self.order_penalty, self.trade_penalty = float(cfg["order_penalty"]), float(cfg["trade_penalty"])
        self.drawdown_penalty, self.passive_bonus = float(cfg["drawdown_penalty"]), float(cfg["passive_bonus"])
        self.loss_multiplier = float(cfg["loss_multiplier"])
        self._vec = VecCDAEnv(cfg, num_markets=1, device=int(self.config.get("device", 0)),
                              order_capacity=int(self.config.get("order_capacity", 0)), fill_capacity=int(self.config.get("fill_capacity", 64)),
                              decimal_ledger=bool(self.config.get("decimal_ledger", True)))
        agent_ids = [f"agent_{i}" for i in range(self.num_of_agents)]
        self._agent_ids = set(agent_ids)
        self.agents = list(agent_ids)
        self.possible_agents = list(agent_ids)
        self.observation_spaces = {
            a: _spaces.Box(low=-np.inf, high=np.inf, shape=(self.n_hist * self.snapshot_dim,), dtype=np.float32)
            for a in agent_ids}
        self.action_spaces = self.act_space(self.num_of_agents)
        self._vec.enable_action_log()                         # LOB_actions
        self.traders = [_TraderView(self, i) for i in range(self.num_of_agents)]
        self.LOB_actions = None
        self.t_step = 0
        self.last_price = None
        self.best_bid = self.best_ask = self.spread = None
        self.pass_agents = set()
        self.done_set = set()
        self.model_actions = None
        self._warned_order = False
        self._ever_reset = False

    # ---- spaces (action_helper.py:103-143) -------------------------------------------------
    def act_space(self, num_agents):
        agent_space = _spaces.Dict({
            "category": _spaces.Discrete(_config.CATEGORY_N),
            "size_mean": _spaces.Box(low=-1.0, high=1.0, shape=(1,), dtype=np.float32),
            "size_sigma": _spaces.Box(low=0.0, high=1.0, shape=(1,), dtype=np.float32),
            "price": _spaces.Discrete(self.k_rows),
            "price_offset": _spaces.Discrete(_config.PRICE_OFFSET_N),
        })
        return {f"agent_{i}": agent_space for i in range(num_agents)}

    def get_action_space(self, agent_id):
        return self.action_spaces[agent_id]

    def get_observation_space(self, agent_id):
        return self.observation_spaces[agent_id]

    # ---- reset (continuousDoubleAuction_env.py:175-231) -------------------------------------
    def reset(self, *, seed=None, options=None):
        if seed is None and not self._ever_reset:
            seed = int.from_bytes(os.urandom(8), "little")   # gymnasium seeds from OS entropy
        obs = self._vec.reset(seed=None if seed is None else [int(seed)])
        self._ever_reset = True
        self.t_step = 0
        self.done_set = set()
        self.pass_agents = set()
        self.LOB_actions = None
        o = obs[0].cpu().numpy()
        self.last_price = float(self._vec.info("market")[0, 0].item())
        return {a: o for a in self.agents}, {a: {} for a in self._agent_ids}

    # ---- step (continuousDoubleAuction_env.py:265-309) --------------------------------------
    def step(self, actions):
        self.model_actions = actions
        A = self.num_of_agents
        cat = np.full((1, A), -1, np.int32)
        mean = np.zeros((1, A), np.float32); sigma = np.zeros((1, A), np.float32)
        price = np.zeros((1, A), np.int32); off = np.ones((1, A), np.int32)
        idxs = pack_actions(actions, A, cat, mean, sigma, price, off)
        if idxs != sorted(idxs) and not self._warned_order:
            warnings.warn("action dict is not in agent order: the reference draws order sizes in dict order, "
                          "cda_b200 draws them in agent order (results stay valid, seed-level parity is lost)")
            self._warned_order = True
        obs, rew, term, trunc = self._vec.step_host(cat, mean, sigma, price, off)
        o = obs[0].copy()
        info = {k: v.cpu().numpy() for k, v in self._vec.info_all().items()}
        mk = info["market"][0]
        self.last_price = float(mk[0])
        self.best_bid = float(mk[1]) if mk[1] > 0 else None
        self.best_ask = float(mk[2]) if mk[2] > 0 else None
        self.spread = (self.best_ask - self.best_bid) if (self.best_bid is not None and self.best_ask is not None) else None
        la = self._vec.last_actions()[0].cpu().numpy()            # continuousDoubleAuction_env.py:285: the decoded, non-pass actions in dict order
        self.LOB_actions = [{"ID": f"agent_{i}", "side": _SIDES[int(la[i, 1])], "type": _TYPES[int(la[i, 0])], "size": int(la[i, 2]),
                             "price": -1.0 if int(la[i, 0]) == 0 else float(la[i, 3])} for i in idxs if la[i, 1] >= 0]
        self._vec_status = int(mk[7])
        if self._vec_status & 29:                                  # fatal bits only: a fill-LOG overflow (bit 2) leaves book and ledger exact
            self._vec.check_status()
        next_states = {a: o for a in self.agents}                 # one shared array, like the reference
        rewards = {a: float(rew[0, i]) for i, a in enumerate(self.agents)}
        infos, self.pass_agents, newly_done = build_infos(info, 0, self.agents, rewards, actions, self)
        self.done_set |= newly_done
        terminateds = {a: False for a in self.agents}
        truncateds = {a: False for a in self.agents}
        terminateds["__all__"] = bool(term[0])
        truncateds["__all__"] = bool(trunc[0])
        self.t_step += 1
        return next_states, rewards, terminateds, truncateds, infos

    def decimal_fields(self):
        """The reference's Decimal money fields of this market, residues included (VecCDAEnv.decimal_fields)."""
        return self._vec.decimal_fields([0])[0]

    @property
    def np_random(self):
        """A numpy Generator positioned exactly where this market's device-side stream is (numpy's PCG64, state copied out of the
        device): what the reference's `env.np_random` would draw next.  A SNAPSHOT — drawing from it does not advance the env."""
        st = np.asarray(self._vec.dump(0)["rng"], dtype=np.uint64)
        bg = np.random.PCG64()
        bg.state = {"bit_generator": "PCG64", "state": {"state": (int(st[0]) << 64) | int(st[1]), "inc": (int(st[2]) << 64) | int(st[3])},
                    "has_uint32": int(st[4]), "uinteger": int(st[5])}
        return np.random.Generator(bg)

    def fills(self):
        """Trades of the last step (the reference's seq_trades), rows of
        (time, price, qty, maker, maker_order_id, maker_left, taker, taker_side)."""
        f, n = self._vec.fills()
        return f[0, :int(n[0].item())].cpu().numpy()

    def render(self):
        return None

    def close(self):
        self._vec.close()
