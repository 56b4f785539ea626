#!/bin/bash
# round 2, call AC: k_shade groups the rays it emits by direction octant inside each CTA's queue reservation (coherent warps in
# k_trace / k_shadow); bit-exact tests, then with / without
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_spectral.py -m gpu -q -x 2>&1 | tail -5
Q="timeout 200 python tools/perf_probe.py --reps 5"
L=gpurun_out/ac_probe.log
for lib in libtiray_v_nosort.so libtiray.so; do
  for wl in cornell teapot_mc spectral_box; do
    $Q --workload $wl --lib $lib 2>&1 | grep -v "libpng\|total light" >> $L
  done
  for wl in cornell teapot_mc; do
    echo -n "[shard 0,8] " >> $L; $Q --workload $wl --lib $lib --shard 0,8 2>&1 | grep -v "libpng\|total light" >> $L
  done
done
cut -c1-230 $L
