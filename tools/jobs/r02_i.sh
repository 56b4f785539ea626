#!/bin/bash
# round 2, call I (2 GPUs): the library-owned NCCL communicator: 2-process GPU test + bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "two_gpu or single_rank" 2>&1 | grep -v libpng | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/i_bench_n2.json 2> gpurun_out/i_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/i_bench_n2.json; grep -v "libpng\|total light\|warn" gpurun_out/i_bench_n2.err | tail -5
