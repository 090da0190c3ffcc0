"""The reference's OWN surface-level test files, run unmodified against cda_b200's dict adapter.

How: a subprocess runs pytest on the files under /root/reference with the plugin tests/ref_suite_plugin.py, which makes
`from gym_continuousDoubleAuction.envs.continuousDoubleAuction_env import continuousDoubleAuctionEnv` resolve to the adapter
(gym-continuousdoubleauction_b200/env.py).  There is no GPU in the container that holds /root/reference and no /root/reference on
the GPU box, so the adapter's engine is the CPU oracle here (tests/oracle_vec_shim.py); the CUDA engine is pinned to that oracle bit
for bit by the `-m gpu` tests.  What is exercised is everything above the engine: action-dict packing, spaces, seeding, flags, the
info dict, LOB_actions, lazy state attributes.

47 of the 56 tests in the five files pass.  The 9 that cannot are listed with the reason in EXPECTED_FAIL: each reaches BELOW the public
surface (calls an Action_Helper internal, injects state by assigning to env attributes, reads env.LOB) or asks for a fractional tick."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.needs_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/gym_continuousDoubleAuction/test"
FILES = ("test_env_lifecycle.py", "test_seeding.py", "test_info_dict.py", "test_observation_history.py", "test_new_action_space.py")

EXPECTED_FAIL = {
    # Action_Helper.rand_exec_seq is called directly: the execution-order shuffle happens inside the step kernel (numpy-exact stream,
    # pinned by the parity tests), it is not a host method of the adapter
    "test_seeding.py::TestExecutionOrderIsSeeded::test_shuffle_follows_the_env_seed": "calls env.rand_exec_seq (internal helper)",
    "test_seeding.py::TestExecutionOrderIsSeeded::test_shuffle_actually_shuffles": "calls env.rand_exec_seq (internal helper)",
    "test_seeding.py::TestExecutionOrderIsSeeded::test_shuffle_is_a_permutation": "calls env.rand_exec_seq (internal helper)",
    "test_seeding.py::TestExecutionOrderIsSeeded::test_explicit_seed_pins_one_shuffle": "calls env.rand_exec_seq (internal helper)",
    "test_seeding.py::TestExecutionOrderIsSeeded::test_shuffle_does_not_reorder_in_place": "calls env.rand_exec_seq (internal helper)",
    # state injection: the tests overwrite env.last_price to choose the anchor (device state is not writable through attributes) / read env.LOB.tape
    "test_new_action_space.py::TestActionSpaceRobust::test_price_offsets_bid": "assigns env.last_price to force the anchor",
    "test_new_action_space.py::TestActionSpaceRobust::test_price_offsets_ask": "assigns env.last_price to force the anchor",
    "test_new_action_space.py::TestActionSpaceRobust::test_trading_updates_anchor": "assigns env.last_price, reads env.LOB.tape",
    # fractional tick (0.25): rejected at construction — the reference does not quantise prices on such grids either
    "test_new_action_space.py::TestTickGrid::test_tick_size_config_reaches_action_layer": "tick_size 0.25 (integral ticks only)",
}


def test_reference_surface_tests_run_against_the_adapter():
    cmd = [sys.executable, "-m", "pytest", "-p", "ref_suite_plugin", "-q", "-rA", "--no-header", "-p", "no:cacheprovider"] + [os.path.join(REF_TESTS, f) for f in FILES]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests"), ROOT, os.environ.get("PYTHONPATH", "")]))
    out = subprocess.run(cmd, cwd=os.path.join(ROOT, "tests"), env=env, capture_output=True, text=True, timeout=600).stdout
    res = {}
    for status, path in re.findall(r"^(PASSED|FAILED|ERROR) \S*?/test/(\S+)", out, flags=re.M):
        res[path] = status
    assert len(res) >= 56, out[-3000:]
    failed = {k for k, v in res.items() if v != "PASSED"}
    assert failed == set(EXPECTED_FAIL), (sorted(failed - set(EXPECTED_FAIL)), sorted(set(EXPECTED_FAIL) - failed))
    assert sum(v == "PASSED" for v in res.values()) >= 47
