for lib in libtiray_s4r8b4.so libtiray_s4r8b2.so libtiray_s4r8b1.so libtiray_s8r8b8.so libtiray_s6r4b6.so libtiray_s3r6b6.so; do
python tools/perf_probe.py --workload cornell --lib $lib --batch 0 --reps 2 2>&1 | grep -v libpng | tail -1
python tools/perf_probe.py --workload teapot_mc --lib $lib --batch 0 --reps 2 2>&1 | grep -v libpng | tail -1
done
