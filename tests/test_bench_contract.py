"""bench.py contract checks that need no GPU: the reference arm (the CPU restatement timed on the host cores) prints ONE JSON
line with the keys the driver reads; the native arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(args, env=None):
    e = dict(os.environ); e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, cwd=ROOT)


def test_reference_arm_json_line():
    # torchrun exports OMP_NUM_THREADS=1 to its ranks: the arm must still use every host core
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], env={"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["dtype"] == "f32"
    assert "cornell_box.py PT_RGB 512x512 64spp" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "spp" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert cb["cores"] == len(os.sched_getaffinity(0)) and "all %d host threads" % cb["cores"] in cb["sample"]
    assert d["config"]["spp_per_step"] == 64 and "64 spp per step (of 64)" in cb["sample"]        # the stated config, not a sample of it
    sub = d["workloads"]["teapot_mc"]                                                            # BASELINE configs[2] next to it
    assert "130720 tris" in sub["config"]["workload"] and sub["value"] > 0 and sub["cpu_baseline"]["cores"] == cb["cores"]


def test_reference_arm_other_ranks_stay_silent():
    """under torchrun only rank 0 runs and prints the reference arm; the other ranks exit 0 without work"""
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_native_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "0", "--no-cpu"])
    assert r.returncode != 0 and r.stdout.strip() == ""          # no JSON line from a CPU fallback


def test_clock_sampler_polls_nvml_inside_the_region(monkeypatch):
    """the `clocks` object of the bench line: NVML polled from a thread while the timed region runs (a stub NVML here: the
    median SM clock, the maximum, every throttle reason seen, at least the start and end samples even for a 1 ms region);
    without NVML and without nvidia-smi the object says so instead of inventing numbers"""
    import time
    import types
    sys.path.insert(0, ROOT)
    import bench
    calls = {"n": 0}
    nv = types.ModuleType("pynvml")
    nv.NVML_CLOCK_SM = 1
    nv.nvmlInit = lambda: None
    nv.nvmlDeviceGetHandleByIndex = lambda i: ("gpu", i)
    nv.nvmlDeviceGetMaxClockInfo = lambda h, c: 1965
    def clock(h, c):
        calls["n"] += 1
        return 1965 if calls["n"] != 2 else 1500
    nv.nvmlDeviceGetClockInfo = clock
    nv.nvmlDeviceGetCurrentClocksThrottleReasons = lambda h: 0x4 if calls["n"] == 2 else 0      # sw_power_cap once
    nv.nvmlDeviceGetPowerUsage = lambda h: 300000
    monkeypatch.setitem(sys.modules, "pynvml", nv)
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "3,5")
    s = bench.ClockSampler(1)
    assert s.h == ("gpu", 5)                                   # local rank 1 of CUDA_VISIBLE_DEVICES=3,5 is physical GPU 5
    s.start(); time.sleep(0.03); c = s.stop()
    assert c["source"] == "nvml" and c["samples"] >= 3 and c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1965.0
    assert c["reasons"] == ["sw_power_cap"] and c["power_w_max"] == 300.0
    s = bench.ClockSampler(0); s.start(); c = s.stop()         # the shortest possible region still has its two end samples
    assert c["samples"] >= 2
    broken = types.ModuleType("pynvml")
    def boom():
        raise RuntimeError("no driver")
    broken.nvmlInit = boom
    monkeypatch.setitem(sys.modules, "pynvml", broken)
    monkeypatch.setenv("PATH", "/nonexistent")
    s = bench.ClockSampler(0); s.start(); c = s.stop()
    assert c["sm_mhz"] is None and c["reasons"] == ["nvidia-smi unavailable"]
