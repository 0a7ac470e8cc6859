#!/bin/bash
# dev loop for the rank-128 kernel: build both libraries here, then the rank-128 parity tests + short c3 bench +
# instrumented run on the GPU box.   usage: scripts/dev_als128.sh <tag>
set -e
cd "$(dirname "$0")/.."
tag=$1; shift || true
python -c "import __graft_entry__ as g; g.build()" 2>&1 | grep -v "nvcc warning" || true
scripts/build_prof.sh 2>&1 | grep -v "nvcc warning" || true
/usr/local/graft/bin/gpurun --timeout 900 -- "mkdir -p gpurun_out/$tag; timeout -k 5 300 python -m pytest tests/test_gpu_als.py -x -q -m gpu -k '128' > gpurun_out/$tag/t_als.log 2>&1; echo \"tests rc=\$?\" >> gpurun_out/$tag/t_als.log; tail -4 gpurun_out/$tag/t_als.log; timeout -k 5 200 python bench.py --workload c3 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-scoring --no-c3 --no-c4 $* > gpurun_out/$tag/b_c3.log 2>&1; HALS_LIB_PATH=\$PWD/hybrid-als-twotower-recommender_b200/libhals_b200_prof.so timeout -k 5 200 python bench.py --workload c3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-scoring --no-c3 --no-graphs $* > gpurun_out/$tag/b_prof.log 2>&1; grep -v '^{' gpurun_out/$tag/b_prof.log | tail -6" 2>&1 | tail -14
python3 - <<PY
import json
d=json.loads(open('gpurun_out/$tag/b_c3.log').read().strip().splitlines()[-1])
print('ms/sweep', round(d['ms_per_step'],3), 'item', round(d['roofline']['ms_item_half'],3), 'user', round(d['roofline']['ms_user_half'],3), 'frac', round(d['roofline']['frac'],3), 'rmse', d['config']['train_rmse_after_run'])
PY
