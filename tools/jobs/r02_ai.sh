#!/bin/bash
# round 2, call AI: BDPT batch size follows max_paths (20 M -> 6 batches of 32 spp, 64 M -> 2 batches): re-measure; final BDPT / spectral bench lines
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
Q="timeout 300 python tools/perf_probe.py --reps 3"
L=gpurun_out/ai_probe.log
for o in max_paths=20971520 max_paths=33554432 max_paths=67108864 max_paths=134217728; do
  echo -n "[$o] " >> $L; $Q --workload veach_bdpt --opts $o 2>&1 | grep -v "libpng\|total light" >> $L
done
cut -c1-230 $L
timeout 300 python bench.py --workload veach_bdpt --no-cpu > gpurun_out/ai_bench_bdpt.json 2> gpurun_out/ai_bench_bdpt.err; cut -c1-200 gpurun_out/ai_bench_bdpt.json
