#!/bin/bash
# round 2, call AN: BDPT light sub-paths from spot / laser emitters (Scene.sample_light's branches): parity tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bdpt.py tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -q -k "spot or sphere_emitter or vertices_and_strategies or wavefront_equals or validation" 2>&1 | tail -6
timeout 200 python tools/perf_probe.py --workload veach_bdpt --reps 2 2>&1 | grep -v "libpng\|total light" | cut -c1-200
