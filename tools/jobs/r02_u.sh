#!/bin/bash
# round 2, call U (4 GPUs): BASELINE configs[4]: veach_bdpt BDPT_RGB 512x512 on 4 x B200 (asynchronous BDPT render + library reduce)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --workload veach_bdpt --steps 10 --warmup 3 --no-cpu > gpurun_out/u_bench_bdpt_n4.json 2> gpurun_out/u_bench_bdpt_n4.err; echo "bdpt n4 rc=$?"
head -c 400 gpurun_out/u_bench_bdpt_n4.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --workload spectral_box --steps 10 --warmup 3 --no-cpu > gpurun_out/u_bench_spec_n4.json 2> gpurun_out/u_bench_spec_n4.err; echo "spectral n4 rc=$?"
head -c 300 gpurun_out/u_bench_spec_n4.json; echo
