#!/bin/bash
# round 2, call AK: final build: full GPU suite, smoke(), default bench line (the driver's round-end sequence)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/ak_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ak_pytest.log
tail -3 gpurun_out/ak_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v libpng | tail -3
timeout 600 python bench.py > gpurun_out/ak_bench_default.json 2> gpurun_out/ak_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ak_bench_default.json').read().strip().splitlines()[-1]); t=d['workloads']['teapot_mc']
print('C2', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), d['clocks'], d['gpu_launches'])
print('C3', round(t['value']), round(t['ms_per_step'],2), 'e2e', round(t['e2e']['value']), round(t['e2e']['ms_per_step'],2))
PY
