#!/bin/bash
# round 2, call D: shared-memory-only stack + shared libm: parity, A/B vs the round-1 kernels, ncu of the new kernels (CSV only)
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -8 gpurun_out/d_pytest.log
Q="timeout 200 python tools/perf_probe.py --reps 3"
for lib in libtiray_old.so libtiray.so libtiray_v_g3.so libtiray_v_g1.so libtiray_v_s3.so libtiray_v_s6.so; do
  for wl in cornell teapot_mc16; do $Q --workload $wl --lib $lib 2>&1 | grep -v "libpng\|total light" >> gpurun_out/d_probe.log; done
done
for wl in teapot_mc veach_bdpt spectral_box; do $Q --workload $wl 2>&1 | grep -v "libpng\|total light" >> gpurun_out/d_probe.log; done
timeout 200 python tools/perf_probe.py --reps 2 --counters --lib libtiray_counters.so --workload cornell 2>&1 | grep -v "libpng\|total light" >> gpurun_out/d_probe.log
timeout 200 python tools/perf_probe.py --reps 2 --counters --lib libtiray_counters.so --workload teapot_mc16 2>&1 | grep -v "libpng\|total light" >> gpurun_out/d_probe.log
cat gpurun_out/d_probe.log
for wl in cornell teapot_mc16; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shadow|k_shade" -c 6 -f -o /tmp/d_ncu_$wl \
     python tools/perf_probe.py --workload $wl --reps 0 --opts chains=1,graph=0 > gpurun_out/d_ncu_$wl.log 2>&1
  ncu -i /tmp/d_ncu_$wl.ncu-rep --page raw --csv > gpurun_out/d_ncu_${wl}_raw.csv 2>/dev/null
  ncu -i /tmp/d_ncu_$wl.ncu-rep --page source --csv --kernel-name regex:k_trace --launch-count 1 > gpurun_out/d_ncu_${wl}_src_trace.csv 2>/dev/null
done
ls -la gpurun_out/
