mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_trace -c 3 -o gpurun_out/prof_trace_r01 -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_trace.log 2>&1; echo "ncu trace rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_shadow -c 2 -o gpurun_out/prof_shadow_r01 -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_shadow.log 2>&1; echo "ncu shadow rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_shade -c 2 -o gpurun_out/prof_shade_r01 -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_shade.log 2>&1; echo "ncu shade rc=$?"
timeout 900 python bench.py --workload teapot_mc --steps 3 --warmup 2 > gpurun_out/bench_teapot.log 2>&1; echo "bench teapot rc=$?"; tail -2 gpurun_out/bench_teapot.log | cut -c1-3000
ls -la gpurun_out
