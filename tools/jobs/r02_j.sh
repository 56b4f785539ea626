#!/bin/bash
# round 2, call J (8 GPUs): default bench at N = 8 (Cornell + mc/Teapot sub-record), reference arm under torchrun, rank-0 ncu timeline
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/j_bench_n8.json 2> gpurun_out/j_bench_n8.err; echo "bench n8 rc=$?"
head -c 400 gpurun_out/j_bench_n8.json; echo
timeout 600 $TR --master-port 29522 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > gpurun_out/j_ref_n8.json 2> gpurun_out/j_ref_n8.err; echo "ref n8 rc=$?"
head -c 300 gpurun_out/j_ref_n8.json; echo
timeout 600 $TR --master-port 29523 tools/rank0_ncu.py gpurun_out/j_ncu_rank0_cornell.csv --gpus 8 --steps 1 --warmup 1 --no-sub --no-cpu > gpurun_out/j_ncu_cornell.json 2> gpurun_out/j_ncu_cornell.err; echo "ncu cornell rc=$?"
timeout 600 $TR --master-port 29524 tools/rank0_ncu.py gpurun_out/j_ncu_rank0_teapot.csv --gpus 8 --steps 1 --warmup 1 --workload teapot_mc --no-cpu > gpurun_out/j_ncu_teapot.json 2> gpurun_out/j_ncu_teapot.err; echo "ncu teapot rc=$?"
ls -la gpurun_out
