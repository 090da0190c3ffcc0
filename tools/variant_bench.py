#!/usr/bin/env python
"""Build kernel variants (extra nvcc -D flags) and time each with bench.py.
usage: python tools/variant_bench.py --build-only "name:-DFLAG=1 -DOTHER=2" ...   here (cross-compile into tools/bin/),
then   gpurun -- python tools/variant_bench.py "name:..." ...                      on the GPU box (loads the prebuilt
libraries through CDA_B200_LIB; building on the box costs GPU-minutes)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_continuousdoubleauction_b200 import _native  # noqa: E402


VARDIR = os.path.join(ROOT, "tools", "bin")


def build(flags, name):
    """Variant libraries are built HERE (CPU container) into tools/bin/ and travel with the snapshot: `--build-only`."""
    os.makedirs(VARDIR, exist_ok=True)
    out = os.path.join(VARDIR, f"libcda_{name}.so")
    cmd = ["nvcc"] + _native.NVCC_FLAGS + flags + ["-I", os.path.join(ROOT, "include"), "-I", _native.CSRC,
                                                   "-o", out, os.path.join(_native.CSRC, "cda_b200.cu")]
    subprocess.check_call(cmd)
    return out


def main():
    extra = os.environ.get("VB_BENCH_ARGS", "--steps 100 --warmup 5 --no-cpu-baseline --no-e2e").split()
    results = {}
    args = [a for a in sys.argv[1:] if a != "--build-only"]
    build_only = "--build-only" in sys.argv
    for spec in args:
        name, _, fl = spec.partition(":")
        lib = os.path.join(VARDIR, f"libcda_{name}.so")
        if build_only or not os.path.exists(lib):
            lib = build(fl.split(), name)
        if build_only:
            continue
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, capture_output=True, text=True,
                             env=dict(os.environ, CDA_B200_LIB=lib))
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            results[name] = dict(value=d["value"], ms=d["ms_per_step"], hot_ms=d["roofline"]["l2_hot_kernel_ms"],
                                 status=d["config"]["status_bits"], e2e=(d.get("e2e") or {}).get("value"))
            print(f"{name:28s} value {d['value']/1e6:8.2f} M/s  step {d['ms_per_step']*1e3:7.2f} us  L2-hot {d['roofline']['l2_hot_kernel_ms']*1e3:7.2f} us"
                  f"  e2e {((d.get('e2e') or {}).get('value') or 0)/1e6:7.2f} M/s status {d['config']['status_bits']}", flush=True)
        except Exception as e:
            print(name, "FAILED", e, out.stdout[-500:], out.stderr[-1500:], flush=True)
    if build_only:
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
