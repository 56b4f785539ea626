#!/bin/bash
# round 2, call AD: final-build numbers and ncu evidence for profiles/ (bench lines, launch list, DRAM traffic, --set full)
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
timeout 600 python bench.py > gpurun_out/ad_bench_default.json 2> gpurun_out/ad_bench_default.err; cut -c1-300 gpurun_out/ad_bench_default.json
timeout 300 python bench.py --workload spectral_box --no-cpu > gpurun_out/ad_bench_spectral.json 2> gpurun_out/ad_bench_spectral.err; cut -c1-200 gpurun_out/ad_bench_spectral.json
timeout 300 python bench.py --workload veach_bdpt --no-cpu > gpurun_out/ad_bench_bdpt.json 2> gpurun_out/ad_bench_bdpt.err; cut -c1-200 gpurun_out/ad_bench_bdpt.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ad_bench_reference.json 2> gpurun_out/ad_bench_reference.err; cut -c1-300 gpurun_out/ad_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/ad_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ad_launches_bench.json 2> gpurun_out/ad_launches_bench.err
for wl in cornell teapot_mc16; do
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ad_traffic_$wl.csv \
     python tools/perf_probe.py --workload $wl --reps 0 --opts chains=1,graph=0 > gpurun_out/ad_traffic_$wl.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shadow|k_shade|k_tail|k_generate|k_accumulate" -c 10 -f -o /tmp/ad_ncu_$wl \
     python tools/perf_probe.py --workload $wl --reps 0 --opts chains=1,graph=0 > gpurun_out/ad_ncu_$wl.log 2>&1
  ncu -i /tmp/ad_ncu_$wl.ncu-rep --page raw --csv > gpurun_out/ad_ncu_${wl}_raw.csv 2>/dev/null
done
timeout 600 python tools/parity_report.py 2>&1 | grep -v libpng > gpurun_out/ad_parity.log; tail -12 gpurun_out/ad_parity.log | cut -c1-250
ls -la gpurun_out | grep " ad_"
