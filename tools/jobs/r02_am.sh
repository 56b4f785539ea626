#!/bin/bash
# round 2, call AM: spot / laser emitters through PT_Spec (new parity test) and PT_RGB
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spectral.py tests/test_gpu_parity.py -m gpu -q -k "spot" 2>&1 | tail -6
