#!/bin/bash
# round 2, call X: asynchronous process_normal (tests that use it), e2e piece timings, chains / skew at one rank of an 8-way shard
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_spectral.py tests/test_gpu_bdpt.py -m gpu -q -x 2>&1 | tail -3
Q="timeout 200 python tools/perf_probe.py --reps 5"
L=gpurun_out/x_probe.log
$Q --workload teapot_mc --e2e 2>&1 | grep -v "libpng\|total light" >> $L
$Q --workload cornell --e2e 2>&1 | grep -v "libpng\|total light" >> $L
for wl in cornell teapot_mc; do
  for o in chains=1 chains=2 chains=3 chains=4 chains=2,chain_skew=50 chains=2,chain_skew=80 chains=3,chain_skew=50; do
    echo -n "[shard 0,8 $o] " >> $L
    $Q --workload $wl --shard 0,8 --opts $o 2>&1 | grep -v "libpng\|total light" >> $L
  done
done
cut -c1-260 $L
