mkdir -p gpurun_out
python tools/perf_probe.py --workload cornell --lib libtiray_counters.so --batch 0 --reps 1 --counters 2>&1 | grep -v libpng | tail -2
python tools/perf_probe.py --workload cornell --lib libtiray.so --batch 1,4,8,16,32,64 --reps 2 2>&1 | grep -v libpng | tail -7
python tools/perf_probe.py --workload cornell --lib libtiray_fmad.so --batch 16 --reps 2 2>&1 | grep -v libpng | tail -1
python tools/perf_probe.py --workload cornell --lib libtiray.so --batch 16 --reps 2 --opts smem_bvh=0 2>&1 | grep -v libpng | tail -1
python tools/perf_probe.py --workload cornell --lib libtiray.so --batch 16 --reps 2 --opts graph=0 2>&1 | grep -v libpng | tail -1
python tools/perf_probe.py --workload teapot_mc --lib libtiray_counters.so --batch 0 --reps 1 --counters 2>&1 | grep -v libpng | tail -1
python tools/perf_probe.py --workload teapot_mc --lib libtiray.so --batch 1,2,4,8,16 --reps 2 2>&1 | grep -v libpng | tail -5
python tools/perf_probe.py --workload teapot_mc --lib libtiray_fmad.so --batch 4 --reps 2 2>&1 | grep -v libpng | tail -1
ncu --set full --clock-control none --import-source on -k regex:k_trace -c 3 -o gpurun_out/prof_trace_teapot_r01 -f python bench.py --workload teapot_mc --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_trace_teapot.log 2>&1; echo "ncu trace rc=$?"
