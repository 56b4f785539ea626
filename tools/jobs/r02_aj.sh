#!/bin/bash
# round 2, call AJ: process_normal with per-vertex terms computed once (bit-exact tests), e2e pieces, BDPT in one batch
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bdpt.py tests/test_gpu_kernels.py -m gpu -q -x -k "teapot or glass_env or bdpt or normal or full_size" 2>&1 | tail -3
timeout 300 python tools/perf_probe.py --reps 2 --workload teapot_mc --e2e 2>&1 | grep "e2e pieces" | cut -c1-330
timeout 300 python bench.py --workload veach_bdpt --no-cpu > gpurun_out/aj_bench_bdpt.json 2> gpurun_out/aj_bench_bdpt.err; cut -c1-200 gpurun_out/aj_bench_bdpt.json
timeout 300 python bench.py --workload teapot_mc --no-cpu > gpurun_out/aj_bench_teapot.json 2> gpurun_out/aj_bench_teapot.err; cut -c1-200 gpurun_out/aj_bench_teapot.json
