#!/bin/bash
# Runs ON the GPU box: the default bench at N GPUs, launched the way the driver launches it.  usage: scale_run.sh N
N=$1; out=gpurun_out/n$N; mkdir -p $out
if [ "$N" = 1 ]; then timeout -k 5 900 python bench.py --gpus 1 > $out/bench.json 2> $out/bench.err
else timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $out/bench.json 2> $out/bench.err; fi
tail -c 300 $out/bench.err
python - <<PY
import json
d=json.loads(open('$out/bench.json').read().strip().splitlines()[-1])
print(d['n_gpus'], 'c2 ms', round(d['ms_per_step'],3), 'e2e', d['e2e']['value'], 'rmse', d['config']['train_rmse_after_run'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_sweep'))
PY
