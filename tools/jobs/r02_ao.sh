#!/bin/bash
# round 2, call AO: the full GPU suite on the final tree
mkdir -p gpurun_out
timeout 110 python -m pytest tests -m gpu -q -x > gpurun_out/ao_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ao_pytest.log
tail -3 gpurun_out/ao_pytest.log
