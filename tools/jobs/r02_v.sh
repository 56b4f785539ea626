#!/bin/bash
# round 2, call V: BDPT tests with the even batches; shadow kernel at 5 CTAs/SM; refill thresholds on the final code
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 900 python -m pytest tests/test_gpu_bdpt.py -m gpu -q 2>&1 | tail -3
Q="timeout 200 python tools/perf_probe.py --reps 5"
for lib in libtiray.so libtiray_v_sh5.so libtiray_v_r12.so libtiray_v_r6.so; do
  for wl in cornell teapot_mc; do $Q --workload $wl --lib $lib 2>&1 | grep -v "libpng\|total light" >> gpurun_out/v_probe.log; done
done
$Q --workload veach_bdpt 2>&1 | grep -v "libpng\|total light" >> gpurun_out/v_probe.log
cut -c1-170 gpurun_out/v_probe.log
