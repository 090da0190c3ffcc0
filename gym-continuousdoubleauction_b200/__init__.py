"""cda_b200 — B200-native vectorised continuous-double-auction environment.

Drop-in for the per-step env hot path of ChuaCheowHuan/gym-continuousDoubleAuction
(`continuousDoubleAuctionEnv.reset/step`), as hand-written sm_100a CUDA behind the C-ABI in
include/cda_b200.h.  Importing the package does not need a GPU; constructing an env does.
"""
from . import _native, config, workloads  # noqa: F401
from .config import SNAPSHOT_DIM, K_ROWS  # noqa: F401

__all__ = ["VecCDAEnv", "VectorCDAEnv", "continuousDoubleAuctionEnv", "build"]


def build(force=False, verbose=False):
    return _native.build(force=force, verbose=verbose)


def __getattr__(name):  # lazy: torch is only imported when an env class is requested
    if name == "VecCDAEnv":
        from .vec_env import VecCDAEnv
        return VecCDAEnv
    if name == "VectorCDAEnv":
        from .vector_env import VectorCDAEnv
        return VectorCDAEnv
    if name == "continuousDoubleAuctionEnv":
        from .env import continuousDoubleAuctionEnv
        return continuousDoubleAuctionEnv
    raise AttributeError(name)
