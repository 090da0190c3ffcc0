#!/usr/bin/env python
"""Build kernel variants (extra nvcc -D flags) and time each with bench.py on the current GPU box.
usage: python tools/variant_bench.py "name:-DFLAG=1 -DOTHER=2" ...   (run under gpurun; nvcc is on the box)
Restores the default build at the end."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_continuousdoubleauction_b200 import _native  # noqa: E402


def build(flags):
    cmd = ["nvcc"] + _native.NVCC_FLAGS + flags + ["-I", os.path.join(ROOT, "include"), "-I", _native.CSRC,
                                                   "-o", _native.SO_PATH, os.path.join(_native.CSRC, "cda_b200.cu")]
    subprocess.check_call(cmd)


def main():
    extra = os.environ.get("VB_BENCH_ARGS", "--steps 100 --warmup 5 --no-cpu-baseline --no-e2e").split()
    results = {}
    for spec in sys.argv[1:]:
        name, _, fl = spec.partition(":")
        build(fl.split())
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            results[name] = dict(value=d["value"], ms=d["ms_per_step"], hot_ms=d["roofline"]["l2_hot_kernel_ms"],
                                 status=d["config"]["status_bits"], e2e=(d.get("e2e") or {}).get("value"))
            print(f"{name:28s} value {d['value']/1e6:8.2f} M/s  step {d['ms_per_step']*1e3:7.2f} us  L2-hot {d['roofline']['l2_hot_kernel_ms']*1e3:7.2f} us"
                  f"  e2e {((d.get('e2e') or {}).get('value') or 0)/1e6:7.2f} M/s status {d['config']['status_bits']}", flush=True)
        except Exception as e:
            print(name, "FAILED", e, out.stdout[-500:], out.stderr[-1500:], flush=True)
    build([])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
