#!/bin/bash
# round 2, call P: final single-GPU records: full GPU suite, default bench line, the other workloads
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p_pytest.log
tail -6 gpurun_out/p_pytest.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/p_bench_default.json 2> gpurun_out/p_bench_default.err; echo "bench default rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/p_bench_reference.json 2> gpurun_out/p_bench_reference.err; echo "bench reference rc=$?"
for wl in spectral_box veach_bdpt; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/p_bench_$wl.json 2> gpurun_out/p_bench_$wl.err; echo "bench $wl rc=$?"
done
python - <<'PY'
import json
for n in ("default","spectral_box","veach_bdpt"):
    try:
        d=json.loads([l for l in open("gpurun_out/p_bench_%s.json"%n) if l.startswith("{")][-1])
        print(n, round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "roofline", round(d["roofline"]["frac"],2), [ (k, round(v["frac"],2)) for k,v in d["roofline"].get("kernels",{}).items()])
        if "workloads" in d:
            w=d["workloads"]["teapot_mc"]; print("  teapot_mc", round(w["value"]), round(w["ms_per_step"],2), "e2e", round(w["e2e"]["value"]))
    except Exception as e: print(n, "ERR", e)
PY
