#!/bin/bash
# round 2, call O: uneven frame split between the two chains (chain_skew) on one rank of an 8-way shard and at N = 1
mkdir -p gpurun_out
Q="timeout 200 python tools/perf_probe.py --reps 5"
for wl in cornell teapot_mc; do
  for sk in 0 55 60 67 75; do
    $Q --workload $wl --shard 0,8 --opts chain_skew=$sk 2>&1 | grep -v "libpng\|total light" | sed "s/^/[shard 0,8 skew=$sk] /" >> gpurun_out/o_probe.log
  done
  for sk in 0 60; do $Q --workload $wl --opts chain_skew=$sk 2>&1 | grep -v "libpng\|total light" | sed "s/^/[N=1 skew=$sk] /" >> gpurun_out/o_probe.log; done
done
cut -c1-150 gpurun_out/o_probe.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -3
