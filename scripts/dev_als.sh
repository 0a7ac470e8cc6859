#!/bin/bash
# dev loop: build both libraries here, then ALS parity tests + short c2 bench + instrumented run on the GPU box.
# usage: scripts/dev_als.sh <tag> [extra bench args]
set -e
cd "$(dirname "$0")/.."
tag=$1; shift || true
python -c "import __graft_entry__ as g; g.build()" 2>&1 | grep -v "nvcc warning" || true
scripts/build_prof.sh 2>&1 | grep -v "nvcc warning" || true
gpurun --timeout 900 -- "mkdir -p gpurun_out/$tag; timeout -k 5 150 python -m pytest tests/test_gpu_als.py -x -q -m gpu > gpurun_out/$tag/t_als.log 2>&1; echo \"tests rc=\$?\" >> gpurun_out/$tag/t_als.log; timeout -k 5 120 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-scoring --no-c3 --no-c4 $* > gpurun_out/$tag/b_ws.log 2>&1; HALS_LIB_PATH=\$PWD/hybrid-als-twotower-recommender_b200/libhals_b200_prof.so timeout -k 5 120 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-scoring --no-c3 --no-graphs $* > gpurun_out/$tag/b_prof.log 2>&1; tail -3 gpurun_out/$tag/t_als.log" 2>&1 | tail -5
python3 - <<PY
import json,re,collections
d=json.loads(open('gpurun_out/$tag/b_ws.log').read().strip().splitlines()[-1])
print('ms/sweep', round(d['ms_per_step'],3), 'item', round(d['roofline']['ms_item_half'],3), 'user', round(d['roofline']['ms_user_half'],3), 'frac', round(d['roofline']['frac'],3), 'rmse', d['config']['train_rmse_after_run'])
lines=[l for l in open('gpurun_out/$tag/b_prof.log') if l.startswith('C ')]
per=148*4
for k in range(max(0,len(lines)-2*per),len(lines),per):
    grp=lines[k:k+per]
    dd=collections.defaultdict(list)
    for l in grp:
        m=re.match(r'C (\d+) w(\d+) role (\d+) cta (\d+)',l)
        if m: dd[int(m.group(2))].append((int(m.group(3)),int(m.group(1))))
    out=[]
    for w in sorted(dd):
        v=sorted(dd[w]); out.append(f"w{w}: min {v[0][0]/1e6:.2f}M med {v[len(v)//2][0]/1e6:.2f}M max {v[-1][0]/1e6:.2f}M (cta {v[-1][1]})")
    print(' | '.join(out))
PY
grep -v "^{\|^C " gpurun_out/$tag/b_prof.log | tail -26
