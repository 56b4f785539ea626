#!/bin/bash
# round 2, call K: full GPU suite, exactness report, and the ncu evidence for profiles/ (launch list, DRAM traffic, --set full)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
tail -6 gpurun_out/k_pytest.log
for o in "pdl=0" "pdl=1"; do for wl in cornell teapot_mc; do
  timeout 200 python tools/perf_probe.py --reps 4 --workload $wl --opts $o 2>&1 | grep -v "libpng\|total light" | sed "s/^/[$o] /" >> gpurun_out/k_pdl.log
  timeout 200 python tools/perf_probe.py --reps 4 --workload $wl --shard 0,8 --opts $o 2>&1 | grep -v "libpng\|total light" | sed "s/^/[$o shard 0,8] /" >> gpurun_out/k_pdl.log
done; done
cat gpurun_out/k_pdl.log
timeout 900 python tools/parity_report.py 2>&1 | grep -v libpng > gpurun_out/k_parity.log; tail -14 gpurun_out/k_parity.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/k_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/k_launches_bench.json 2> gpurun_out/k_launches_bench.err
for wl in cornell teapot_mc16 spectral_box veach_bdpt; do
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/k_traffic_$wl.csv \
     python tools/perf_probe.py --workload $wl --reps 0 --opts chains=1,graph=0 > gpurun_out/k_traffic_$wl.log 2>&1
done
for wl in cornell teapot_mc16; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shadow|k_shade|k_tail|k_generate|k_accumulate" -c 10 -f -o /tmp/k_ncu_$wl \
     python tools/perf_probe.py --workload $wl --reps 0 --opts chains=1,graph=0 > gpurun_out/k_ncu_$wl.log 2>&1
  ncu -i /tmp/k_ncu_$wl.ncu-rep --page raw --csv > gpurun_out/k_ncu_${wl}_raw.csv 2>/dev/null
done
timeout 600 ncu --set full --clock-control none -k regex:"k_trace|k_shadow|k_bdpt" -c 12 -f -o /tmp/k_ncu_bdpt python tools/perf_probe.py --workload veach_bdpt --reps 0 > gpurun_out/k_ncu_bdpt.log 2>&1
ncu -i /tmp/k_ncu_bdpt.ncu-rep --page raw --csv > gpurun_out/k_ncu_veach_bdpt_raw.csv 2>/dev/null
ls -la gpurun_out | tail -20
