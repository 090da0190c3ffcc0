"""TEST INFRASTRUCTURE — not product code.

Stand-ins for the two imports the reference env needs and this image lacks
(`gymnasium`, `ray.rllib.env.multi_agent_env`), so that the UNMODIFIED reference
package under /root/reference can be imported in this container and used as the
behavioural oracle.  Nothing here is copied from the
reference; it only re-creates the tiny part of the gymnasium API the reference
touches:

* `gymnasium.Env.reset(seed=...)` seeding `self.np_random` with
  `Generator(PCG64(SeedSequence(seed)))`   (reference call site:
  gym_continuousDoubleAuction/envs/continuousDoubleAuction_env.py:189-190, :221)
* `gymnasium.spaces.{Box,Discrete,Dict}` constructors
  (gym_continuousDoubleAuction/envs/exchg/action_helper.py:126-138)
* `gymnasium.envs.registration.register` (gym_continuousDoubleAuction/__init__.py:18-21)
* `ray.rllib.env.multi_agent_env.MultiAgentEnv` base class
  (continuousDoubleAuction_env.py:8)

The reference tree does not exist on the GPU box, so this module is only ever
used here (golden-vector generation and the `needs_reference` CPU tests).
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("CDA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_continuousDoubleAuction"))


class _Space:
    _rng = None

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    @property
    def rng(self):
        if self._rng is None:
            self._rng = np.random.default_rng()
        return self._rng


class Box(_Space):
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    def sample(self):
        if np.isinf(self.low):
            return self.rng.normal(size=self.shape).astype(self.dtype)
        return self.rng.uniform(self.low, self.high, self.shape).astype(self.dtype)


class Discrete(_Space):
    def __init__(self, n):
        self.n = n

    def sample(self):
        return np.int64(self.rng.integers(0, self.n))


class Dict(_Space):
    def __init__(self, d):
        self.spaces = d

    def __getitem__(self, k):
        return self.spaces[k]

    def seed(self, seed=None):
        for sp, q in zip(self.spaces.values(),
                         np.random.SeedSequence(seed).spawn(len(self.spaces))):
            sp._rng = np.random.default_rng(q)

    def sample(self):
        return {k: v.sample() for k, v in self.spaces.items()}


class Env:
    _np_random = None

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.Generator(
                np.random.PCG64(np.random.SeedSequence(seed)))

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.Generator(
                np.random.PCG64(np.random.SeedSequence(None)))
        return self._np_random


def install():
    """Insert the stand-ins into sys.modules and put the reference on sys.path."""
    if "gymnasium" in sys.modules and getattr(sys.modules["gymnasium"], "_cda_stub", False):
        return
    gym = types.ModuleType("gymnasium")
    gym._cda_stub = True
    spaces = types.ModuleType("gymnasium.spaces")
    envs = types.ModuleType("gymnasium.envs")
    reg = types.ModuleType("gymnasium.envs.registration")
    reg.register = lambda **k: None
    spaces.Box, spaces.Discrete, spaces.Dict = Box, Discrete, Dict
    gym.Env, gym.spaces, gym.envs, envs.registration = Env, spaces, envs, reg
    mae = types.ModuleType("ray.rllib.env.multi_agent_env")
    mae.MultiAgentEnv = type("MultiAgentEnv", (Env,), {})
    mods = {
        "gymnasium": gym, "gymnasium.spaces": spaces, "gymnasium.envs": envs,
        "gymnasium.envs.registration": reg, "ray": types.ModuleType("ray"),
        "ray.rllib": types.ModuleType("ray.rllib"),
        "ray.rllib.env": types.ModuleType("ray.rllib.env"),
        "ray.rllib.env.multi_agent_env": mae,
    }
    for name, mod in mods.items():
        sys.modules.setdefault(name, mod)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def make_reference_env(config):
    """Build the unmodified reference env (is_render forced off unless given)."""
    install()
    os.environ.setdefault("CDA_LOG_LEVEL", "ERROR")
    from gym_continuousDoubleAuction.envs.continuousDoubleAuction_env import (
        continuousDoubleAuctionEnv,
    )
    cfg = dict(config)
    cfg.setdefault("is_render", False)
    return continuousDoubleAuctionEnv(cfg)
