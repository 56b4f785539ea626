#!/bin/bash
# round 2, call T: what the driver runs at round end: build(), smoke(), pytest -m gpu, bench.py (both arms)
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/t_smoke.log 2>&1; echo "smoke rc=$?"; grep -v libpng gpurun_out/t_smoke.log | tail -4
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo "bench rc=$?"; wc -l gpurun_out/t_bench.json; head -c 200 gpurun_out/t_bench.json; echo
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/t_bench_ref.json 2> gpurun_out/t_bench_ref.err; echo "ref rc=$?"; head -c 200 gpurun_out/t_bench_ref.json; echo
