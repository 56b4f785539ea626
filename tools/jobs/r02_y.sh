#!/bin/bash
# round 2, call Y: spot / laser emitters in sample_li (parity test), validation test; shade kernel before / after
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -q -x -k "spot or validation or cornell_matches or glass_env" 2>&1 | tail -5
Q="timeout 200 python tools/perf_probe.py --reps 5"
L=gpurun_out/y_probe.log
for lib in libtiray_v_base.so libtiray.so; do
  for wl in cornell teapot_mc spectral_box; do $Q --workload $wl --lib $lib 2>&1 | grep -v "libpng\|total light" >> $L; done
done
cut -c1-200 $L
