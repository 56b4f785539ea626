#!/bin/bash
# round 2, call N (8 GPUs): final default bench at N = 8 + rank-0 ncu timelines (tail hand-over 4096, page-locked tables)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/n_bench_n8.json 2> gpurun_out/n_bench_n8.err; echo "bench n8 rc=$?"
head -c 300 gpurun_out/n_bench_n8.json; echo
timeout 600 $TR --master-port 29533 tools/rank0_ncu.py gpurun_out/n_ncu_rank0_cornell.csv --gpus 8 --steps 1 --warmup 1 --no-sub --no-cpu > gpurun_out/n_ncu_cornell.json 2> gpurun_out/n_ncu_cornell.err; echo "ncu cornell rc=$?"
timeout 600 $TR --master-port 29534 tools/rank0_ncu.py gpurun_out/n_ncu_rank0_teapot.csv --gpus 8 --steps 1 --warmup 1 --workload teapot_mc --no-cpu > gpurun_out/n_ncu_teapot.json 2> gpurun_out/n_ncu_teapot.err; echo "ncu teapot rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/n_bench_n4.json 2> gpurun_out/n_bench_n4.err; echo "bench n4 rc=$?"
ls -la gpurun_out | grep " n_"
