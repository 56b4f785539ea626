#!/bin/bash
# round 2, call G: stress the small spectral render (k_tail<SPEC>) in three builds to find the garbage link
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
for lib in libtiray_v_dbg.so libtiray.so libtiray_v_tail1.so; do
  echo "== $lib"; TIRAY_LIB=$lib timeout 300 python tools/stress_spec.py 400 2>&1 | grep -v libpng | head -20
  echo "== $lib graph=0"; TIRAY_LIB=$lib timeout 300 python tools/stress_spec.py 200 graph=0 2>&1 | grep -v libpng | head -20
done
echo "== rgb tail stress (PT_RGB 128^2 goes through k_tail<false> the same way)"
TIRAY_LIB=libtiray_v_dbg.so timeout 600 python -m pytest tests/test_gpu_spectral.py -q -x -k "not full_size and not sky_dome" 2>&1 | grep -v libpng | tail -15
