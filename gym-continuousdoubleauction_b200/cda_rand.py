"""Random-policy driver: the counterpart of the reference's CDA_rand.run_random
(gym_continuousDoubleAuction/CDA_rand.py:40-85) on the CUDA env — BASELINE config #1's harness.

    python -m gym_continuousdoubleauction_b200.cda_rand --num_agents 4 --max_step 1000
"""
import argparse
import time


def run_random(num_agents=4, max_step=1000, init_cash=1_000_000, seed=0):
    from .env import continuousDoubleAuctionEnv
    env = continuousDoubleAuctionEnv({"num_of_agents": num_agents, "init_cash": init_cash,
                                      "max_step": max_step, "is_render": False})
    for i, sp in enumerate(env.action_spaces.values()):
        sp.seed(seed + i)
        break   # all agents share one space object, like the reference
    obs, _ = env.reset(seed=seed)
    steps, t0 = 0, time.perf_counter()
    while True:
        actions = {a: env.action_spaces[a].sample() for a in env.agents}
        obs, rew, term, trunc, info = env.step(actions)
        steps += 1
        if term["__all__"] or trunc["__all__"]:
            break
    dt = time.perf_counter() - t0
    navs = {a: info[a]["NAV"] for a in env.agents}
    env.close()
    return {"steps": steps, "seconds": dt, "steps_per_s": steps / dt, "NAV": navs,
            "total_nav": sum(int(v) for v in navs.values())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num_agents", type=int, default=4)
    ap.add_argument("--max_step", type=int, default=1000)
    ap.add_argument("--init_cash", type=int, default=1_000_000)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    print(run_random(a.num_agents, a.max_step, a.init_cash, a.seed))


if __name__ == "__main__":
    main()
