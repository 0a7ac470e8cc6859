#!/usr/bin/env bash
# usage: scripts/ab_bench.sh lib1.so lib2.so ...  -> one summary line per library
for lib in "$@"; do
  HALS_LIB_PATH=$PWD/$lib timeout -k 5 -s KILL 120 python scripts/debug_als.py 2>&1 | grep "^64" | cut -c1-60
  HALS_LIB_PATH=$PWD/$lib timeout -k 5 -s KILL 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'ms/sweep', round(d['ms_per_step'],3), 'item', round(d['roofline']['ms_item_half'],3), 'user', round(d['roofline']['ms_user_half'],3), 'frac', round(d['roofline']['frac'],4), 'rmse', d['config']['train_rmse_after_run'])"
done
