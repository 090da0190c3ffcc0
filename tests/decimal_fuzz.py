#!/usr/bin/env python
"""Wide fuzz of the device Decimal twin (decimal_ledger) against the oracle's decimal_ledger mode (which reproduces the reference's Decimal
fields exactly, tests/test_oracle_vs_reference.py): N seeded low- and mid-cash configurations x M markets x T steps; integer state compared
every `every` steps and at the end, the twins' Decimal cash / VWAP / nav at the end.  Prints one line per configuration and a summary.
usage (under gpurun): python tests/decimal_fuzz.py [N=120] [M=12] [T=160]   (a script, not a pytest module; lives under tests/ because it drives the oracle)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))   # (ROOT = repo root)
import numpy as np, torch
import gym_continuousdoubleauction_b200 as cda
from oracle.cda_oracle import OracleEnv
from parity_utils import assert_dump_equal
from test_gpu_decimal import gen

N, M, T = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 120), (2, 12), (3, 160)))
every = 20
bad = ties = resid = restarts0 = 0
L = None
t0 = time.time()
for case in range(N):
    rng = np.random.default_rng(910000 + case)
    A = int(rng.integers(2, 9))
    lo = int(rng.choice([3, 10, 37, 250, 999]))
    cfg = dict(num_of_agents=A, max_step=T + 5, tick_size=int(rng.choice([1, 1, 1, 2, 3])), n_hist=int(rng.integers(1, 5)), min_size=int(rng.integers(1, 4)),
               mkt_max_size=int(rng.choice([10, 40, 100])), limit_size_multiple=int(rng.choice([1, 3, 10])),
               init_cash=int(rng.choice([300, 1_000, 3_000, 7_777, 20_000, 100_000, 1_000_000])), initial_price_min=lo, initial_price_max=lo + int(rng.integers(0, 30)))
    mix = str(rng.choice(["uniform", "limit_market", "modify_heavy"]))
    env = cda.VecCDAEnv(cfg, num_markets=M, decimal_ledger=True, status_policy="ignore")
    L = env._L
    orc = OracleEnv(cfg, M, decimal_ledger=True, dec128=True)
    seeds = np.array([880000 + case * 100 + m for m in range(M)], np.uint64)
    env.reset(seed=seeds); orc.reset(seeds=seeds)
    per = [gen(np.random.default_rng(int(s) + 7), T, A, mix, float(rng.choice([0.0, 0.1]))) for s in seeds]
    acts = [np.stack([pm[f] for pm in per], axis=1) for f in range(5)]
    ok = True
    try:
        for t in range(T):
            og, rg, _, _ = env.step(*[torch.from_numpy(np.ascontiguousarray(a[t])).cuda() for a in acts])
            oc, rc, _, _ = orc.step(*[a[t] for a in acts])
            if t % every == every - 1 or t == T - 1:
                assert np.abs(og.cpu().numpy().astype(np.float64) - oc).max() <= 1e-6 and np.abs(rg.cpu().numpy() - rc).max() <= 1e-6
                dumps = env.dump_all()
                for m in range(M):
                    assert_dump_equal(dumps[m], orc.dump(m), ctx=f"case={case} t={t} m={m}", fills=False)
        f = env.decimal_fields()
        for m in range(M):
            d = orc.dump_decimal(m)
            assert f[m]["cash"] == d["cash"] and f[m]["VWAP"] == d["VWAP"] and f[m]["nav"] == d["nav"], (case, m)
            resid += sum(1 for c in d["cash"] if c != c.to_integral_value())
        assert int(env.status().max().item()) == 0
    except AssertionError as e:
        ok = False; bad += 1
        print("MISMATCH", case, cfg, mix, str(e)[:300], flush=True)
    r = int(L.cda_debug_restart_count())
    print(f"case {case:3d} A={A} cash={cfg['init_cash']:>8d} {mix:13s} {'ok' if ok else 'FAIL'}  tie-resolution passes so far {r}", flush=True)
    env.close()
print(f"== {N} configurations x {M} markets x {T} steps: {bad} mismatches; agents ending with a cash residue: {resid}; tie-resolution passes: {int(L.cda_debug_restart_count())}; {time.time() - t0:.0f} s")
