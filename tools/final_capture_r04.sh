#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun), second half of round 2: GPU tests, smoke, both bench arms, the ncu launch list of the
# bench command and one full capture of the headline step kernel.
# usage: gpurun --timeout 900 -- bash tools/final_capture_r04.sh r04e
set -u
TAG=${1:-r04x}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/${TAG}_gputests.log; cat $O/${TAG}_gputests.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $O/${TAG}_smoke.log
timeout 300 python bench.py > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench_1gpu.err; tail -c 400 $O/${TAG}_bench_1gpu.json; echo
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > $O/${TAG}_bench_reference_arm.json 2>&1
# launch list of the bench command (cold-cache, serialised: compare SHARES).  CDA_SERVE=0: under ncu every launch is synchronous, so the
# resident step server would be one 2-ms kernel per step (it retires by its idle lease before the launch call returns); the e2e legs are
# listed through the launch-per-step path instead.
CDA_SERVE=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --prewarm 64 --no-cpu-baseline > $O/${TAG}_launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cda_step_kernel -s 300 -c 2 -o $O/prof_${TAG} \
    python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_${TAG}.log 2>&1
ls -la $O | tail -12
