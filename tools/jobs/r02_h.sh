#!/bin/bash
# round 2, call H: stress + full GPU suite + exactness report with the spill-free spectral tail
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 300 python tools/stress_spec.py 400 2>&1 | grep -v libpng | tail -3
timeout 300 python tools/stress_spec.py 200 graph=0 2>&1 | grep -v libpng | tail -3
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -15 gpurun_out/h_pytest.log
timeout 600 python tools/parity_report.py 2>&1 | grep -v libpng | tail -8
Q="timeout 200 python tools/perf_probe.py --reps 3"
for wl in cornell teapot_mc16 teapot_mc spectral_box veach_bdpt; do $Q --workload $wl 2>&1 | grep -v "libpng\|total light" >> gpurun_out/h_probe.log; done
cat gpurun_out/h_probe.log
