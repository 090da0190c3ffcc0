"""TEST INFRASTRUCTURE — pytest plugin (-p ref_suite_plugin) that lets the reference's OWN test files import
`gym_continuousDoubleAuction.envs.continuousDoubleAuction_env.continuousDoubleAuctionEnv` and get cda_b200's dict adapter instead:
gymnasium / ray stand-ins (oracle/ref_stub.py), the adapter's engine swapped for the CPU oracle (tests/oracle_vec_shim.py; this container has
no GPU), everything else of the reference package untouched.  Used by tests/test_reference_suite_on_adapter.py in a subprocess."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    from oracle import ref_stub
    ref_stub.install()
    os.environ.setdefault("CDA_LOG_LEVEL", "ERROR")
    import gym_continuousDoubleAuction.envs.continuousDoubleAuction_env as ref_mod      # the real reference module (helpers stay real)
    import gym_continuousdoubleauction_b200 as cda
    from gym_continuousdoubleauction_b200 import vec_env
    from oracle_vec_shim import OracleVec
    vec_env.VecCDAEnv = OracleVec                                                        # engine: CPU oracle (pinned == CUDA by the -m gpu tests)
    ref_mod.continuousDoubleAuctionEnv = cda.continuousDoubleAuctionEnv                  # what the reference's tests import
