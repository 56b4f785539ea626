mkdir -p gpurun_out
nvidia-smi -L
for n in 1 2; do
if [ $n = 1 ]; then python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_cornell_$n.log 2>&1; else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_cornell_$n.log 2>&1; fi
echo "cornell n=$n rc=$?"; tail -1 gpurun_out/scale_cornell_$n.log | cut -c1-900
if [ $n = 1 ]; then python bench.py --workload teapot_mc --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_teapot_$n.log 2>&1; else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --workload teapot_mc --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_teapot_$n.log 2>&1; fi
echo "teapot n=$n rc=$?"; tail -1 gpurun_out/scale_teapot_$n.log | cut -c1-900
done
