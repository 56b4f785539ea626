#!/bin/bash
# round 2, call C: remaining parity tests; A/B of the traversal-core variants; ncu --set full of old vs new trace / shadow kernels
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -8 gpurun_out/c_pytest.log
Q="timeout 200 python tools/perf_probe.py --reps 3"
for lib in libtiray_old.so libtiray.so libtiray_v_ns2.so libtiray_v_ns3.so libtiray_v_noovf.so libtiray_v_noovf_ns2.so; do
  for wl in cornell teapot_mc16; do $Q --workload $wl --lib $lib 2>&1 | grep -v "libpng\|total light" >> gpurun_out/c_probe.log; done
done
cat gpurun_out/c_probe.log
for lib in libtiray_old.so libtiray.so; do for wl in cornell teapot_mc16; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shadow" -c 4 -f -o gpurun_out/c_ncu_${wl}_${lib%.so} \
     python tools/perf_probe.py --workload $wl --lib $lib --reps 0 --opts chains=1,graph=0 > gpurun_out/c_ncu_${wl}_${lib%.so}.log 2>&1
done; done
ls -la gpurun_out/*.ncu-rep
