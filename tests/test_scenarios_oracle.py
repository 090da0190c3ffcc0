"""Known-answer scenarios on the CPU oracle (and on the live reference where it exists)."""
import numpy as np
import pytest

import scenarios
from oracle.cda_oracle import OracleEnv


class OracleBackend:
    def __init__(self):
        self.e = None

    def make(self, cfg):
        self.e = OracleEnv(cfg, 1)

    def reset_one(self, seed):
        return self.e.reset(seeds=[seed])[0]

    def step_one(self, *a):
        return self.e.step(*[np.asarray(x)[None] for x in a])

    def dump_one(self):
        return self.e.dump(0)


class _Lazy:
    """Scenario scripts build their config first; create the backend env on reset."""
    def __init__(self, factory):
        self.factory, self.be = factory, None

    def bind(self, cfg):
        self.be = self.factory(cfg)
        return self.be


def run(scn, make_backend):
    # scripts call Script(be, cfg): patch Script to construct the backend from cfg
    orig = scenarios.Script.__init__

    def init(self, be, c):
        real = make_backend(c)
        orig(self, real, c)
    scenarios.Script.__init__ = init
    try:
        scn(None)
    finally:
        scenarios.Script.__init__ = orig


def oracle_backend(cfg):
    b = OracleBackend(); b.make(cfg); return b


@pytest.mark.parametrize("scn", scenarios.ALL, ids=[f.__name__ for f in scenarios.ALL])
def test_scenario_on_oracle(scn):
    run(scn, oracle_backend)


class RefBackend:
    def __init__(self, cfg):
        from oracle.ref_runner import ReferenceMarket
        self.m = ReferenceMarket(cfg)

    def reset_one(self, seed):
        return self.m.reset(seed=seed)

    def step_one(self, *a):
        return self.m.step(*a)

    def dump_one(self):
        return self.m.dump()


@pytest.mark.needs_reference
@pytest.mark.parametrize("scn", scenarios.ALL, ids=[f.__name__ for f in scenarios.ALL])
def test_scenario_on_live_reference(scn):
    """The hand-derived expectations hold on the UNMODIFIED reference itself."""
    run(scn, RefBackend)
