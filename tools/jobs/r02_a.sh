#!/bin/bash
# round 2, call A: GPU tests + default bench (new record layout) + e2e pieces + tail_max sweep on the r01 kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --verbose > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err
for wl in cornell teapot_mc16; do
  timeout 120 python tools/perf_probe.py --workload $wl --e2e --reps 2 >> gpurun_out/a_probe.log 2>&1
  for tm in 0 16384 65536 262144 1048576; do
    timeout 120 python tools/perf_probe.py --workload $wl --reps 3 --opts tail_max=$tm >> gpurun_out/a_probe.log 2>&1
    timeout 120 python tools/perf_probe.py --workload $wl --reps 3 --shard 0,8 --opts tail_max=$tm >> gpurun_out/a_probe_shard8.log 2>&1
  done
done
tail -3 gpurun_out/a_pytest.log; tail -c 600 gpurun_out/a_bench.json; grep -v "^e2e\|libpng\|total light" gpurun_out/a_probe.log | tail -12; grep -v "libpng\|total light" gpurun_out/a_probe_shard8.log | tail -12
