#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun): tests, smoke, both bench arms, ncu launch list + full captures, sweep.
# usage: gpurun --timeout 1500 -- bash tools/final_capture.sh r02d
set -u
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/${TAG}_gputests.log; cat $O/${TAG}_gputests.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee $O/${TAG}_smoke.log
python bench.py > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench_1gpu.err; tail -c 600 $O/${TAG}_bench_1gpu.json; echo
python bench.py --impl reference --steps 20 --warmup 3 > $O/${TAG}_bench_reference_arm.json 2>&1
# launch list of the bench command (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --prewarm 64 --no-cpu-baseline > $O/${TAG}_launches_bench.log 2>&1
# full captures of the step kernel: headline workload (cfg3) and BASELINE config 4 (8 agents x 8192 markets, modify-heavy)
ncu --set full --clock-control none --import-source on -k regex:cda_step_kernel -s 300 -c 2 -o $O/prof_${TAG} \
    python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cda_step_kernel -s 300 -c 1 -o $O/prof_${TAG}_cfg4 \
    python bench.py --workload cfg4_modify_heavy_8x8192 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_${TAG}_cfg4.log 2>&1
python tools/sweep.py > $O/${TAG}_sweep.log 2>&1; cp $O/sweep.json $O/${TAG}_sweep.json; tail -14 $O/${TAG}_sweep.log | cut -c1-220
python tools/e2e_timeline.py 2>&1 | tee $O/${TAG}_e2e_timeline.txt
ls -la $O | tail -20
