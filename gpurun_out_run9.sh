timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-250
python tools/perf_probe.py --workload cornell --lib libtiray_counters.so --batch 0 --reps 1 --counters 2>&1 | grep -v libpng | tail -1
python tools/perf_probe.py --workload cornell --lib libtiray.so --batch 0 --reps 2 2>&1 | grep -v libpng | tail -1
python tools/perf_probe.py --workload teapot_mc --lib libtiray_counters.so --batch 0 --reps 1 --counters 2>&1 | grep -v libpng | tail -1
python tools/perf_probe.py --workload teapot_mc --lib libtiray.so --batch 0 --reps 2 2>&1 | grep -v libpng | tail -1
