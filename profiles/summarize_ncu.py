#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) of the step kernel into a small tracked text/JSON file.

usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_step_kernel.json [markets] [workload]
Also updates profiles/traffic.json (per-launch DRAM bytes per workload; bench.py reads it for
roofline.traffic)."""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "smsp__issue_active.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    markets = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    workload = sys.argv[4] if len(sys.argv) > 4 else "cfg3_limit_market_4x4096"
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            if h in KEYS or h == "Kernel Name":
                try:
                    d[h] = {"value": float(v), "unit": u}
                except ValueError:
                    d[h] = v
        launches.append(d)

    def val(d, k):
        x = d.get(k)
        if not isinstance(x, dict):
            return None
        v, u = x["value"], x["unit"].lower()
        if k.startswith("dram__bytes") or k.startswith("lts__t_bytes"):
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        if k == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
        return v
    summ = []
    for d in launches:
        s = {k: val(d, k) for k in KEYS}
        s["kernel"] = d.get("Kernel Name")
        rd, wr = s["dram__bytes_read.sum"], s["dram__bytes_write.sum"]
        s["dram_bytes_per_launch"] = (rd or 0) + (wr or 0)
        s["dram_bytes_per_market_step"] = s["dram_bytes_per_launch"] / markets
        s["warp_instructions_per_market_step"] = (s["smsp__inst_executed.sum"] or 0) / markets
        summ.append(s)
    doc = {"report": os.path.basename(rep), "markets": markets, "workload": workload,
           "note": "ncu --set full --clock-control none; times are cold-cache, serialised replays: compare shares, not absolutes",
           "launches": summ}
    json.dump(doc, open(out, "w"), indent=1)
    tpath = os.path.join(os.path.dirname(os.path.abspath(out)), "traffic.json")
    t = json.load(open(tpath)) if os.path.exists(tpath) else {}
    if summ:
        t[workload] = sum(s["dram_bytes_per_launch"] for s in summ) / len(summ)
        t[workload + "__source"] = os.path.basename(out)
    json.dump(t, open(tpath, "w"), indent=1)
    for s in summ:
        print(f"{s['kernel']}: {s['gpu__time_duration.sum']:.1f} us, DRAM {s['dram_bytes_per_launch']/1e6:.2f} MB/launch "
              f"({s['dram_bytes_per_market_step']:.0f} B/market-step), {s['warp_instructions_per_market_step']:.0f} warp-instr/market-step, "
              f"regs {s['launch__registers_per_thread']}, IPC/SMSP {s['smsp__issue_active.avg.per_cycle_active']}")


if __name__ == "__main__":
    main()
