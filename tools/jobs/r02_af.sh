#!/bin/bash
# round 2, call AF: final build (one batch per step: max_paths 64 M): full GPU suite, default bench line, reference arm untouched
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/af_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/af_pytest.log
tail -4 gpurun_out/af_pytest.log
timeout 600 python bench.py > gpurun_out/af_bench_default.json 2> gpurun_out/af_bench_default.err; cut -c1-300 gpurun_out/af_bench_default.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/af_bench_default.json').read().strip().splitlines()[-1])
t=d['workloads']['teapot_mc']
print('C2', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['run'])
print('C3', t['value'], t['ms_per_step'], t['e2e']['value'], t['e2e']['ms_per_step'], t['run'], t['roofline']['stage_ms_per_step'], t['roofline']['frac'])
PY
timeout 300 python tools/perf_probe.py --reps 4 --workload teapot_mc --shard 0,2 2>&1 | grep -v "libpng\|total light" | cut -c1-200
timeout 300 python tools/perf_probe.py --reps 4 --workload teapot_mc --shard 0,4 2>&1 | grep -v "libpng\|total light" | cut -c1-200
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
