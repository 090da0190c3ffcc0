#!/usr/bin/env python
"""Resident-CTA cap sweep (CDA_HOST_CTAS / CDA_DEV_CTAS): device time and host-synced time per window step, and the
plain device step, as a function of how many 4-market CTAs may be resident per SM.  Run under gpurun."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.e2e_timeline import CHILD  # noqa: E402  (importing runs nothing: the sweep below is guarded)

if __name__ == "__main__":
    for k in os.environ.get("SWEEP", "0 6 5 4 3 2").split():
        e = dict(os.environ); e.update(CDA_HOST_CTAS=k, CDA_DEV_CTAS=k, WSLOTS="32", TAG=f"resident CTAs/SM cap = {k}")
        subprocess.run([sys.executable, "-c", CHILD], env=e)
