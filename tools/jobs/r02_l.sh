#!/bin/bash
# round 2, call L: one rank of an 8-way shard on one GPU: chains x tail_max sweep with the new kernels
mkdir -p gpurun_out
Q="timeout 200 python tools/perf_probe.py --reps 4 --shard 0,8"
for wl in cornell teapot_mc; do
  for ch in 1 2 3 4; do for tm in 0 4096 16384; do
    $Q --workload $wl --opts chains=$ch,tail_max=$tm 2>&1 | grep -v "libpng\|total light" | sed "s/^/[chains=$ch tail_max=$tm] /" >> gpurun_out/l_probe.log
  done; done
done
cat gpurun_out/l_probe.log
