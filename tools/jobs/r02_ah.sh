#!/bin/bash
# round 2, call AH (2 GPUs): NVML clock sampler + the library's NCCL reduce on the final build: two-GPU test, bench at N = 2 and N = 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "two_gpu or comm" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/ah_bench_n2.json 2> gpurun_out/ah_bench_n2.err; echo "bench n2 rc=$?"
CUDA_VISIBLE_DEVICES=1 timeout 600 python bench.py --no-cpu > gpurun_out/ah_bench_n1.json 2> gpurun_out/ah_bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/ah_bench_n2.json','gpurun_out/ah_bench_n1.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); t=d['workloads']['teapot_mc']
    print(f, d['n_gpus'], round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), '| C3', round(t['value']), round(t['ms_per_step'],2), 'e2e', round(t['e2e']['value']), d['clocks'], t.get('clocks'))
PY
tail -3 gpurun_out/ah_bench_n2.err
