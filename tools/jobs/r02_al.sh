#!/bin/bash
# round 2, call AL (4 GPUs): the default bench line at N = 4 with the final build
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/al_bench_n4.json 2> gpurun_out/al_bench_n4.err; echo "bench n4 rc=$?"
head -c 300 gpurun_out/al_bench_n4.json; echo
