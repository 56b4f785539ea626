#!/bin/bash
# round 2, call B: new traversal core (TrNode2 + shared-memory stack + replicated image): parity, then speed per tree mode
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -5 gpurun_out/b_pytest.log
P="timeout 200 python tools/perf_probe.py --reps 3 --counters --lib libtiray_counters.so"
Q="timeout 200 python tools/perf_probe.py --reps 3"
for o in "" "replicas=0" "smem_bvh=0"; do $Q --workload cornell --opts "$o" 2>&1 | grep -v "libpng\|total light" | sed "s/^/[$o] /" >> gpurun_out/b_probe.log; done
for o in "" "top_nodes=128" "top_nodes=256" "top_nodes=512" "stack_smem=8" "stack_smem=16" "stack_smem=32"; do $Q --workload teapot_mc16 --opts "$o" 2>&1 | grep -v "libpng\|total light" | sed "s/^/[$o] /" >> gpurun_out/b_probe.log; done
$P --workload cornell 2>&1 | grep -v "libpng\|total light" >> gpurun_out/b_probe.log
$P --workload teapot_mc16 2>&1 | grep -v "libpng\|total light" >> gpurun_out/b_probe.log
$Q --workload veach_bdpt 2>&1 | grep -v "libpng\|total light" >> gpurun_out/b_probe.log
$Q --workload spectral_box 2>&1 | grep -v "libpng\|total light" >> gpurun_out/b_probe.log
cat gpurun_out/b_probe.log
