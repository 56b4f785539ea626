#!/bin/bash
# round 2, call W: per-mode refill threshold (12 for global-memory trees) + shadow kernel at 5 CTAs/SM as the defaults:
# bit-exact kernel/parity tests, then A/B against global-tree thresholds 16 / 20 and 3 node steps
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_spectral.py -m gpu -q -x 2>&1 | tail -4
Q="timeout 200 python tools/perf_probe.py --reps 5"
for lib in libtiray.so libtiray_v_rg16.so libtiray_v_rg20.so libtiray_v_rg12g3.so; do
  for wl in cornell teapot_mc; do
    if [ $lib != libtiray.so ] && [ $wl = cornell ]; then continue; fi
    $Q --workload $wl --lib $lib 2>&1 | grep -v "libpng\|total light" >> gpurun_out/w_probe.log
  done
done
cut -c1-170 gpurun_out/w_probe.log
timeout 300 python bench.py --no-cpu > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; cut -c1-600 gpurun_out/w_bench.json
