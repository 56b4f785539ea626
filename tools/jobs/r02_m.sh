#!/bin/bash
# round 2, call M: full GPU suite with the bit-exact spectral / BDPT tests; shade occupancy variant; default bench line at N = 1
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -25 gpurun_out/m_pytest.log | cut -c1-200
Q="timeout 200 python tools/perf_probe.py --reps 4"
for lib in libtiray.so libtiray_v_shade3.so; do for wl in cornell teapot_mc spectral_box; do $Q --workload $wl --lib $lib 2>&1 | grep -v "libpng\|total light" >> gpurun_out/m_probe.log; done; done
cat gpurun_out/m_probe.log
timeout 900 python bench.py --steps 10 --warmup 3 --verbose > gpurun_out/m_bench_n1.json 2> gpurun_out/m_bench_n1.err; echo "bench rc=$?"
grep "e2e pass" gpurun_out/m_bench_n1.err | tail -4
