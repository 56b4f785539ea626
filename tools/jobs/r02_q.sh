#!/bin/bash
# round 2, call Q: co-residency experiment: persistent trace / shadow kernels capped at 3 (2) CTAs per SM so that the other chain's
# shade kernel can run beside them; chains 2 / 3 / 4
mkdir -p gpurun_out
Q="timeout 200 python tools/perf_probe.py --reps 4"
for wl in cornell teapot_mc; do
  for pb in 0 3 2; do for ch in 2 3 4; do
    $Q --workload $wl --opts persist_blocks=$pb,chains=$ch 2>&1 | grep -v "libpng\|total light" | sed "s/^/[persist_blocks=$pb chains=$ch] /" >> gpurun_out/q_probe.log
  done; done
done
cut -c1-160 gpurun_out/q_probe.log
