#!/bin/bash
# round 2, call S: intra-warp work stealing at the end of the queue in k_trace: parity, then A/B against the build without it
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
tail -5 gpurun_out/s_pytest.log | cut -c1-200
Q="timeout 200 python tools/perf_probe.py --reps 5"
for lib in libtiray_v_nosteal.so libtiray.so; do
  for wl in cornell teapot_mc; do
    $Q --workload $wl --lib $lib 2>&1 | grep -v "libpng\|total light" >> gpurun_out/s_probe.log
    $Q --workload $wl --lib $lib --shard 0,8 2>&1 | grep -v "libpng\|total light" | sed "s/^/[shard 0,8] /" >> gpurun_out/s_probe.log
  done
  $Q --workload veach_bdpt --lib $lib 2>&1 | grep -v "libpng\|total light" >> gpurun_out/s_probe.log
done
cut -c1-170 gpurun_out/s_probe.log
