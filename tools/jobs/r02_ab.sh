#!/bin/bash
# round 2, call AB: tail hand-over threshold at the size of one rank of an 8-way shard (default 4096)
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
Q="timeout 200 python tools/perf_probe.py --reps 5"
L=gpurun_out/ab_probe.log
for wl in cornell teapot_mc; do
  for o in tail_max=2048 tail_max=4096 tail_max=8192 tail_max=16384 tail_max=32768 tail_max=65536; do
    echo -n "[shard 0,8 $o] " >> $L; $Q --workload $wl --shard 0,8 --opts $o 2>&1 | grep -v "libpng\|total light" >> $L
  done
done
cut -c1-230 $L
