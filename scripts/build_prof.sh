#!/bin/bash
# builds hybrid-als-twotower-recommender_b200/libhals_b200_prof.so: same sources with -DHALS_WS_PROFILE (per-role cycle counters printed by CTA 0)
set -e
cd "$(dirname "$0")/../hybrid-als-twotower-recommender_b200/csrc"
mkdir -p /tmp/objprof
for f in *.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -DHALS_WS_PROFILE $HALS_NVCC_EXTRA -c -o /tmp/objprof/${f%.cu}.o $f &
done
wait
/usr/local/cuda/bin/nvcc -shared -o ../libhals_b200_prof.so /tmp/objprof/*.o -lcudart -ccbin /usr/bin/g++ 2>/dev/null
