#!/bin/bash
# round 2, call R (8 GPUs): the default bench line at N = 8 with the final build
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r_bench_n8.json 2> gpurun_out/r_bench_n8.err; echo "bench n8 rc=$?"
head -c 300 gpurun_out/r_bench_n8.json; echo
