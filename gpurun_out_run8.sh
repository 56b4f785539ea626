timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-250
for ch in 1 2 4 8; do
python tools/perf_probe.py --workload cornell --lib libtiray.so --batch 0 --reps 2 --opts chains=$ch 2>&1 | grep -v libpng | tail -1
python tools/perf_probe.py --workload teapot_mc --lib libtiray.so --batch 0 --reps 2 --opts chains=$ch 2>&1 | grep -v libpng | tail -1
done
for lib in libtiray_sh3.so libtiray_sh4.so; do
python tools/perf_probe.py --workload cornell --lib $lib --batch 0 --reps 2 2>&1 | grep -v libpng | tail -1
python tools/perf_probe.py --workload teapot_mc --lib $lib --batch 0 --reps 2 2>&1 | grep -v libpng | tail -1
done
