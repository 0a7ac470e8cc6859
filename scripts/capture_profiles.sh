#!/bin/bash
# Runs ON the GPU box (through gpurun): plain bench first, then the ncu launch list and one `--set full` capture per hot
# kernel, each of the SAME command that just exited 0 without ncu.  Outputs under gpurun_out/r02/.
# usage: gpurun --timeout 1500 -- scripts/capture_profiles.sh
set -u
out=gpurun_out/r02; mkdir -p $out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-c3 --no-c4 --score-users 65536"
timeout -k 5 300 $B > $out/plain_c2.json 2> $out/plain_c2.err || { echo "plain c2 failed"; exit 1; }
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_c2.csv $B > $out/ncu_launch.log 2>&1
A="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-c3 --no-c4 --no-scoring --no-graphs"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:als_ws64_kernel -s 4 -c 2 -o $out/ws64 $A > $out/ncu_ws64.log 2>&1
C="python bench.py --workload c3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-c3 --no-c4 --no-scoring --no-graphs"
timeout -k 5 300 $C > $out/plain_c3.json 2> $out/plain_c3.err || { echo "plain c3 failed"; exit 1; }
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:als_ws128_kernel -s 4 -c 2 -o $out/ws128 $C > $out/ncu_ws128.log 2>&1
S="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-c3 --no-c4 --score-users 32768 --score-slab 32768"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 2 -c 2 -o $out/score $S > $out/ncu_score.log 2>&1
# the CUDA-core implicit kernel on the config-4 slice, for the comparison DESIGN.md 3.1c quotes
HALS_FORCE_SIMT=1 timeout -k 5 600 python bench.py --workload c4s --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-c3 --no-c4 --no-scoring > $out/c4s_simt.json 2> $out/c4s_simt.err
ls -la $out | tail -20
