"""bench.py output contract (the driver parses these lines): the reference arm runs on CPU and is exercised here on a
small market count; the GPU arm's line is checked on the newest committed bench line under profiles/."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _baseline_metric():
    return json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3",
                          "--markets", "64", "--ref-inner", "8"], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["unit"] == "env-steps/s" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 3
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == "cfg3_limit_market_4x4096" and line["gpu_launches"] == 0
    # both arms print the SAME config object (the driver compares them): rebuild the GPU arm's from the same helper
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.workload_config("cfg3_limit_market_4x4096", 4, 64, 1, "limit_market")
    assert cb["python_reference"]["value"] == 2181.0 and cb["python_reference"]["cores"] == 1


def test_reference_arm_other_ranks_exit_quietly():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--markets", "16"],
                         capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_gpu_bench_line_has_the_contract_keys():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_1gpu.json")))
    assert files, "no committed bench line under profiles/"
    line = json.loads(open(files[-1]).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"clocks", "roofline", "cpu_baseline"} <= set(line)
    assert line["metric"].split(" (")[0] in _baseline_metric() and line["unit"] == "env-steps/s"
    assert line["n_gpus"] == 1 and line["gpu_launches"] == line["steps"] > 0 and line["warmup"] >= 3
    assert line["config"]["workload"] == "cfg3_limit_market_4x4096" and "model" not in line["config"]
    det = line.get("details", line["config"])                                   # (r01 lines kept these under config)
    assert "flushed" in line["config"]["l2"] and det["status_bits"] == 0
    e = line["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 4096 * 4 * 20 and e["d2h_bytes_per_step"] > 4096 * 168
    assert e["value"] < line["value"]                                   # measured through the host API, not a copy of `value`
    if "launch_per_step_variant" in e:                                   # lines since the resident step server: the headline e2e is the better of the two forms
        assert e["resident_kernel_launches"] >= 1 and not e.get("resident_server_error")
        assert e["value"] >= e["launch_per_step_variant"]["value"] > 0 and "rewritten by the host" in e["inputs"]
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - 2738 * 4096 / (r["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"] and r["traffic"]
    c = line["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    k = line["clocks"]
    assert k["sm_mhz"] and k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
