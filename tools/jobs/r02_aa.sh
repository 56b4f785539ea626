#!/bin/bash
# round 2, call AA: late-shadow schedule (one shadow launch for all thin depths at the end of the chain): bit-exact tests, thresholds
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_spectral.py -m gpu -q -x 2>&1 | tail -5
Q="timeout 200 python tools/perf_probe.py --reps 5"
L=gpurun_out/aa_probe.log
for wl in cornell teapot_mc; do
  for o in late_shadow=0 late_shadow=2 late_shadow=3 late_shadow=4 late_shadow=6; do
    echo -n "[$o] " >> $L; $Q --workload $wl --opts $o 2>&1 | grep -v "libpng\|total light" >> $L
    echo -n "[shard 0,8 $o] " >> $L; $Q --workload $wl --shard 0,8 --opts $o 2>&1 | grep -v "libpng\|total light" >> $L
  done
done
cut -c1-230 $L
