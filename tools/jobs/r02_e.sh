#!/bin/bash
# round 2, call E: find the illegal access of the spectral render (compute-sanitizer), full GPU suite, exactness report
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_spectral.py -q -x -k test_pt_spec_cornell_matches_oracle > gpurun_out/e_sanitizer.log 2>&1
grep -n "Invalid\|at 0x\|by thread\|Address\|=========     at\|in k_\|k_" gpurun_out/e_sanitizer.log | head -30
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -15 gpurun_out/e_pytest.log
timeout 600 python tools/parity_report.py > gpurun_out/e_parity.log 2>&1; grep -v libpng gpurun_out/e_parity.log | tail -12
