"""VectorCDAEnv — `num_envs` reference-style multi-agent envs behind ONE kernel launch per step.

Sub-env m is market m of a VecCDAEnv.  The surface is the reference's dict surface, vectorised the way RLlib's
env runners consume it (`train/train.py:495-514`: `num_envs_per_env_runner` envs stepped one after another by a
`MultiAgentEnvRunner`; here they are stepped together):

    reset(*, seed=None, options=None)      -> ([obs_dict] * num_envs, [info_dict] * num_envs)
    step([action_dict] * num_envs)         -> (obs, rewards, terminateds, truncateds, infos), each a list of the
                                              dicts `continuousDoubleAuctionEnv.step` returns for that market
    reset_at(index, seed=None)             -> (obs_dict, info_dict)        (episode boundary of one sub-env)

Observations are zero-copy views of the pinned sliding window (`VecCDAEnv.step_host_window`): valid until the next
call, one array shared by all agents of a market (like the reference, state_helper.py:76,109).  `infos` are LAZY:
the per-agent dicts of `exchg/info_helper.py:30-116` are only built, from ONE device gather for the whole batch, when
somebody reads them (the league callback reads a few keys at episode end; the hot loop usually reads none).

RLlib itself is not installed in the build image, so this class is checked against
`continuousDoubleAuctionEnv` (one market per object) and the reference's golden trajectories, not against Ray.
"""
from collections.abc import Mapping

import numpy as np

from . import config as _config
from .env import build_infos, pack_actions


class _LazyInfos(Mapping):
    """{agent: info dict} of one market; materialised on first access from the step's shared gather."""

    def __init__(self, owner, step_id, m, rewards, actions):
        self._owner, self._step_id, self._m, self._rewards, self._actions = owner, step_id, m, rewards, actions
        self._d = None

    def _get(self):
        if self._d is None:
            info = self._owner._gather(self._step_id)
            self._d, _, _ = build_infos(info, self._m, self._owner.agents, self._rewards, self._actions, self._owner)
        return self._d

    def __getitem__(self, k):
        return self._get()[k]

    def __iter__(self):
        return iter(self._owner.agents)

    def __len__(self):
        return len(self._owner.agents)


class VectorCDAEnv:
    def __init__(self, config=None, num_envs=1, device=0, order_capacity=0):
        import torch
        from .vec_env import VecCDAEnv
        self.config = config or {}
        cfg = _config.resolve(self.config)
        self.num_envs = int(num_envs)
        self.num_of_agents = int(cfg["num_of_agents"])
        self.init_cash, self.max_step, self.n_hist = cfg["init_cash"], int(cfg["max_step"]), int(cfg["n_hist"])
        self.order_penalty, self.trade_penalty = float(cfg["order_penalty"]), float(cfg["trade_penalty"])
        self.drawdown_penalty, self.passive_bonus = float(cfg["drawdown_penalty"]), float(cfg["passive_bonus"])
        self.loss_multiplier = float(cfg["loss_multiplier"])
        self.agents = [f"agent_{i}" for i in range(self.num_of_agents)]
        self.possible_agents = list(self.agents)
        self._vec = VecCDAEnv(cfg, num_markets=self.num_envs, device=device, order_capacity=order_capacity,
                              decimal_ledger=bool(self.config.get("decimal_ledger", True)))
        M, A = self.num_envs, self.num_of_agents
        # ONE pinned market-major action block i32[M][5][A]; the packers write straight into it
        self._blk = torch.empty((M, 5, A), dtype=torch.int32, pin_memory=True)
        b = self._blk.numpy()
        self._cat, self._price, self._off = b[:, 0], b[:, 3], b[:, 4]
        self._mean, self._sigma = b[:, 1].view(np.float32), b[:, 2].view(np.float32)
        self._step_id = 0
        self._gathered = (-1, None)
        self._ever_reset = False

    # ------------------------------------------------------------------ reset
    def reset(self, *, seed=None, options=None):
        """seed: None (OS entropy on the first reset, continue each stream afterwards), an int (sub-env m gets seed + m)
        or a sequence of num_envs ints."""
        if seed is None and not self._ever_reset:
            seed = int.from_bytes(np.random.SeedSequence().generate_state(2).tobytes(), "little") >> 1
        obs = self._vec.reset_host_window(seed=seed)
        self._ever_reset = True
        self._step_id += 1
        rows = [obs[m] for m in range(self.num_envs)]                     # ONE array per market, shared by its agents
        return [{a: o for a in self.agents} for o in rows], [{a: {} for a in self.agents} for _ in range(self.num_envs)]

    def reset_at(self, index, seed=None):
        mask = np.zeros(self.num_envs, np.uint8); mask[index] = 1
        seeds = None
        if seed is not None:
            seeds = np.zeros(self.num_envs, np.uint64); seeds[index] = np.uint64(seed)
        obs = self._vec.reset_host_window(seed=seeds, mask=mask)
        self._step_id += 1
        o = obs[index]
        return {a: o for a in self.agents}, {a: {} for a in self.agents}

    # ------------------------------------------------------------------ step
    def step(self, actions):
        if len(actions) != self.num_envs:
            raise ValueError(f"expected {self.num_envs} action dicts, got {len(actions)}")
        A = self.num_of_agents
        self._mean[:] = 0.0; self._sigma[:] = 0.0; self._price[:] = 0; self._off[:] = 1
        for m, act in enumerate(actions):
            pack_actions(act, A, self._cat, self._mean, self._sigma, self._price, self._off, m)
        obs, rew, term, trunc = self._vec.step_host_window(self._blk, market_major=True)
        self._step_id += 1
        sid = self._step_id
        rew_l = rew.tolist()
        obs_l, rew_d, te_l, tr_l, inf_l = [], [], [], [], []
        agents = self.agents
        for m in range(self.num_envs):
            o = obs[m]
            obs_l.append({a: o for a in agents})
            r = dict(zip(agents, rew_l[m]))
            rew_d.append(r)
            te = {a: False for a in agents}; te["__all__"] = bool(term[m])
            tr = {a: False for a in agents}; tr["__all__"] = bool(trunc[m])
            te_l.append(te); tr_l.append(tr)
            inf_l.append(_LazyInfos(self, sid, m, r, actions[m]))
        return obs_l, rew_d, te_l, tr_l, inf_l

    def _gather(self, step_id):
        if step_id != self._step_id:
            raise RuntimeError("lazy info read after a later step()/reset(): the device state has moved on")
        if self._gathered[0] != step_id:
            info = {k: v.cpu().numpy() for k, v in self._vec.info_all().items()}
            if int(np.bitwise_or.reduce(info["market"][:, 7])) & 29:
                self._vec.check_status()
            self._gathered = (step_id, info)
        return self._gathered[1]

    def close(self):
        self._vec.close()
