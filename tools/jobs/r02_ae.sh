#!/bin/bash
# round 2, call AE: paths in flight per batch on C3 (default 2^24 = 4 batches of 16 frames per 64-spp step; each batch ends with an exposed tail)
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
Q="timeout 300 python tools/perf_probe.py --reps 4"
L=gpurun_out/ae_probe.log
for o in max_paths=16777216 max_paths=33554432 max_paths=67108864 max_paths=67108864,chains=3 max_paths=67108864,chains=4 max_paths=67108864,chain_skew=50; do
  echo -n "[$o] " >> $L; $Q --workload teapot_mc --opts $o 2>&1 | grep -v "libpng\|total light" >> $L
done
nvidia-smi --query-gpu=memory.used --format=csv >> $L
cut -c1-230 $L
