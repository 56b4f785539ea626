mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/prof_trace_v2 -f python tools/perf_probe.py --workload cornell --batch 0 --reps 0 > gpurun_out/ncu_v2.log 2>&1; echo "rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/prof_trace_teapot_v2 -f python tools/perf_probe.py --workload teapot_mc --batch 0 --reps 0 > gpurun_out/ncu_v2t.log 2>&1; echo "rc=$?"
