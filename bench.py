#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: Mrays/s (and ms/spp) of PT_RGB.

  python bench.py --gpus N --steps K --warmup W [--workload cornell|teapot_mc|teapot_mc16|spectral_box|veach_bdpt] [--impl native|reference]

A step = one full pass of the hot path over one batch: clear the film, render `spp` samples per pixel
(cornell: 512x512, 64 spp = BASELINE configs[1]; teapot_mc: 1024x1024, 64 spp = configs[2]; teapot_mc16: 16 spp;
spectral_box: PT_Spec hero-wavelength 512x512, 64 spp = configs[3]; veach_bdpt: BDPT_RGB = configs[4]) with the
scene, BVH and camera resident in HBM, and (N > 1) one NCCL sum-reduce of the film (tr_film_reduce: ncclReduce enqueued by
the library on the render stream).  rays = closest-hit traversals + shadow traversals actually executed (device queue
counters).

  value      whole-job Mrays/s, timed with CUDA events on the stream the kernels run on, max over ranks
  e2e        the same metric through the public Python API with HOST buffers: scene tables H2D + LBVH
             build + render + film reduce + tone map + film D2H inside the timed region
  roofline   the closest-hit trace kernel: algorithmic bytes (48 B/ray + 32 B per internal-node
             visit + 68 B per leaf test, visit counts from the counters build of the same kernel) / mean
             kernel time (CUDA events around every launch of it) vs the measured HBM copy bandwidth;
             roofline.kernels adds the shade and shadow kernels by the same rules (224 / 176 / 112 B per vertex)
  workloads  the default run (cornell) also carries a `workloads.teapot_mc` sub-record: the 130 720-triangle scene of
             BASELINE configs[2] measured the same way in the same process, at every N
  cpu_baseline / --impl reference: the reference algorithm restated on the CPU (oracle/, -O3 -ffast-math,
             OpenMP on ALL host cores, also under torchrun) on the same workload (Taichi is not installable).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "ti-raytrace_b200")
for p in (PKG, os.path.join(PKG, "integrator"), os.path.join(PKG, "example"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    "cornell": dict(module="cornell_box", W=512, H=512, spp=64, files=["cornell_box.obj"], sphere_light=False, env_power=0.0,
                    desc="cornell_box.py PT_RGB 512x512 64spp max_depth 15 (BASELINE configs[1])", normals=False),
    "teapot_mc": dict(module="teapot_mc", W=1024, H=1024, spp=64, files=["mc.obj", "Teapot.obj"], sphere_light=True, env_power=5.0,
                      desc="single_model.py mc.obj+Teapot.obj (130720 tris) PT_RGB 1024x1024 64spp (BASELINE configs[2])", normals=True),
    "teapot_mc16": dict(module="teapot_mc", W=1024, H=1024, spp=16, files=["mc.obj", "Teapot.obj"], sphere_light=True, env_power=5.0,
                        desc="single_model.py mc.obj+Teapot.obj (130720 tris) PT_RGB 1024x1024 16spp (SURVEY 8d C3)", normals=True),
    "spectral_box": dict(module="spectral_box", W=512, H=512, spp=64, files=["cornell_box.obj"], sphere_light=False, env_power=0.0,
                         desc="spectral_box.py PT_Spec hero-wavelength 512x512 64spp max_depth 10 (BASELINE configs[3])", normals=True,
                         spectral=True, max_depth=10),
    "veach_bdpt": dict(module="veach_bdpt", W=512, H=512, spp=32, files=["bdpt.obj"], sphere_light=False, env_power=0.0,
                       desc="veach_bdpt.py BDPT_RGB 512x512 32spp MAX_DEPTH 5 (BASELINE configs[4])", normals=True, bdpt=True, fit=0.5),
}
MAX_DEPTH = 15
L2_NOTE = "GPU arm: L2 flushed (256 MiB write) between timed steps; per-step queue traffic also exceeds L2"


def workload_config(wl):
    """the `config` object: identical in the native and the reference arm (same workload, same samples per step)"""
    return {"workload": wl["desc"], "resolution": "%dx%d" % (wl["W"], wl["H"]), "spp_per_step": wl["spp"],
            "max_depth": wl.get("max_depth", 5 if wl.get("bdpt") else MAX_DEPTH), "l2": L2_NOTE}


# --------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """SM clock, throttle reasons and power DURING the timed region (B200_PROFILING.md recipe).  NVML is polled from a thread
    every 2 ms (one sample is taken when the region starts and one when it ends, so even a 50 ms region at N = 8 is covered);
    `nvidia-smi -lms` is the fallback when the NVML binding is missing (it needs ~0.3 s to start: useless for short regions)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))   # nvmlClocksThrottleReason*

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc, self.h, self.nv, self.run = gpu_index, [], None, None, None, False
        self.samples, self.smax, self.thread = [], None, None
        try:                                      # NVML start-up happens here, before the warm-up, not inside the timed region
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            phys = int(ids[gpu_index]) if ids and all(v.isdigit() for v in ids) and gpu_index < len(ids) else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys); self.nv = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _sample(self):
        nv = self.nv
        try:
            sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            try:
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
            except Exception:
                pw = None
            self.samples.append((sm, mask, pw))
        except Exception:
            pass

    def _poll(self):
        while self.run:
            self._sample()
            time.sleep(0.002)

    def start(self):
        if self.h is not None:
            self._sample()
            self.run = True
            self.thread = threading.Thread(target=self._poll, daemon=True); self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.h is not None:
            self._sample()
            self.run = False
            if self.thread is not None:
                self.thread.join(timeout=1.0)
            sm = sorted(v[0] for v in self.samples)
            reasons = sorted({name for _, mask, _ in self.samples for bit, name in self.REASONS if mask & bit})
            power = [v[2] for v in self.samples if v[2] is not None]
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "reasons": reasons,
                    "power_w_max": max(power) if power else None, "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax = float(r[2]); power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm), "source": "nvidia-smi"}


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def _traffic_table():
    for name in ("r02_dram_traffic.json", "r01_dram_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name))), name
        except Exception:
            pass
    return {}, None


def ncu_traffic(workload, kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel, from the committed ncu capture of the same
    command line (profiles/r0N_dram_traffic.json, written from `ncu --metrics dram__bytes_*` by tools/); None if absent"""
    workload = {"teapot_mc": "teapot_mc16"}.get(workload, workload)      # 64 spp = four batches of the captured 16-frame batch: same launches
    t, _ = _traffic_table()
    for k, v in t.get(workload, {}).items():
        if k.startswith(kernel_prefix) and isinstance(v, dict):
            return v.get("dram_bytes_per_launch_all")
    return None


def ncu_issue(workload):
    """issue-slot utilisation etc. of the dominant kernel from the committed --set full capture (None if absent): the kernels are
    issue-bound, not DRAM-bound, which is why the effective-bandwidth fraction can exceed 1"""
    workload = {"teapot_mc": "teapot_mc16"}.get(workload, workload)
    t, _ = _traffic_table()
    return t.get(workload, {}).get("ncu_full")


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_tables(wl):
    from oracle import objload
    shapes = [objload.sphere_light_rows()] if wl["sphere_light"] else []
    return objload.load_scene([os.path.join(PKG, "model", f) for f in wl["files"]], shapes=shapes)


_cpu_scenes = {}


def cpu_scene(wl):
    """oracle-side scene of a workload (fast build), built once per process"""
    key = wl["module"]
    if key in _cpu_scenes:
        return _cpu_scenes[key]
    from oracle import oracle
    # torchrun exports OMP_NUM_THREADS=1: the CPU baseline is defined on ALL host cores of the box
    oracle.lib(True).orc_set_num_threads(host_threads())
    t = oracle_tables(wl)
    if wl.get("spectral"):
        for k in range(3):                          # example/spectral_box.py:22-27
            t.material[k, 0] = 10.0; t.material[k, 1] = float(k)
    s = oracle.OracleScene(t, fast=True).build()
    cam = oracle.fit_camera(t, wl["W"], wl["H"], wl.get("fit", 0.8))
    s.set_camera(cam[1], cam[2], *cam[3:]); s.set_camera_view(cam[0], wl["W"], wl["H"])
    packed, w, h = oracle.load_env(os.path.join(PKG, "image", "env.png" if wl["env_power"] else "black.png"))
    s.set_env(packed, w, h, wl["env_power"])
    if wl["normals"]:
        s.process_normal()
    if wl.get("spectral"):
        from oracle import spectral
        spectral.attach(s, PKG)
    _cpu_scenes[key] = s
    return s


def cpu_reference_run(wl, spp, frame_begin=0):
    """the reference algorithm on the host cores (oracle, fast build): returns (Mrays/s, rays, seconds, threads, counters)"""
    from oracle import oracle
    s = cpu_scene(wl)
    if wl.get("spectral"):
        from oracle import spectral
        t0 = time.perf_counter()
        _, cnt = spectral.render_pt_spec(s, wl["W"], wl["H"], frame_begin, spp, wl["max_depth"], 0)
    elif wl.get("bdpt"):
        t0 = time.perf_counter()
        _, cnt = s.render_bdpt_rgb(wl["W"], wl["H"], frame_begin, spp, 0)
    else:
        t0 = time.perf_counter()
        _, cnt = s.render_pt_rgb(wl["W"], wl["H"], frame_begin, spp, MAX_DEPTH, 0)
    dt = time.perf_counter() - t0
    rays = cnt["closest"] + cnt["shadow"]
    return rays / dt / 1e6, rays, dt, int(oracle.lib(True).orc_num_threads()), cnt


# --------------------------------------------------------------------------------------------- reference arm
def reference_measure(wl, steps, warmup, spp):
    for _ in range(warmup):
        cpu_reference_run(wl, 1)
    rays = 0; secs = 0.0; cores = 1
    for _ in range(steps):
        _, r, dt, cores, _ = cpu_reference_run(wl, spp)
        rays += r; secs += dt
    v = rays / secs / 1e6
    sample = "%dx%d, %d spp per step (of %d), %d steps, all %d host threads, oracle/liboracle_fast.so" % (wl["W"], wl["H"], spp, wl["spp"], steps, cores)
    return {"value": v, "unit": "Mrays/s", "ms_per_step": secs / steps * 1e3, "ms_per_spp": secs / steps / spp * 1e3, "rays_per_step": rays // steps,
            "config": workload_config(wl), "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_reference(args, wl):
    """the reference's own CPU path (restated: Taichi is not installable) on all host cores of the box, full workload per
    step.  Rank 0 alone runs and prints; the other ranks of a torchrun launch exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the stated config: every step renders the workload's full spp (the BDPT workload is bounded to 4 of its 32 spp: ~10 s per step)
    spp = 4 if wl.get("bdpt") else wl["spp"]
    m = reference_measure(wl, args.steps, args.warmup, spp)
    out = {"impl": "reference", "metric": "Mrays/s", "value": m["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "ms_per_spp": m["ms_per_spp"],
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "reference assets (model/*.obj), counter-based RNG seed 0",
           "config": m["config"], "rays_per_step": m["rays_per_step"], "cpu_baseline": m["cpu_baseline"], "e2e": m["e2e"]}
    if args.workload == "cornell" and not args.no_sub:
        # the 130 720-triangle scene next to it: a bounded sample (16 of 64 spp, 2 steps) so the arm stays within minutes
        sub = WORKLOADS["teapot_mc"]
        sm = reference_measure(sub, max(1, min(args.steps, 2)), min(args.warmup, 1), 16)
        out["workloads"] = {"teapot_mc": sm}
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------- native arm
def build_example(wl):
    import contextlib
    import importlib
    mod = importlib.import_module(wl["module"])
    with contextlib.redirect_stdout(sys.stderr):          # the example scripts print like the reference's; stdout carries only the JSON line
        ex = mod.example(wl["W"], wl["H"], max(wl["spp"], 4))
        ex.build_scene()
    return ex


class Dist:
    """torch.distributed as plumbing: barrier, max / sum over ranks, gather of the per-rank records"""

    def __init__(self, world, local):
        self.world, self.local = world, local

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def reduce(self, values, op):
        if self.world == 1:
            return list(values)
        import torch
        import torch.distributed as dist
        t = torch.tensor(list(values), dtype=torch.float64, device="cuda:%d" % self.local)
        dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM}[op])
        return t.tolist()

    def gather(self, values):
        if self.world == 1:
            return [list(values)]
        import torch
        import torch.distributed as dist
        t = torch.tensor(list(values), dtype=torch.float64, device="cuda:%d" % self.local)
        out = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        return [o.tolist() for o in out]


def measure_native(args, name, steps, warmup, rank, world, local, D, detail=True):
    """one workload on this rank's context: device-timed steps, e2e through the host-buffer API, and (N = 1, detail) the
    roofline of the trace / shade / shadow kernels plus the CPU baseline.  Returns the record (same on every rank)."""
    import torch
    import _native
    import parallel
    import UtilsFunc as UF
    wl = WORKLOADS[name]
    ex = build_example(wl)                      # ti.init() -> fresh context on LOCAL_RANK
    ctx = _native.context()
    stream = torch.cuda.Stream(device=local)
    ctx.stream_set(stream.cuda_stream)
    parallel.comm_init(ctx)                     # library-owned NCCL communicator + tile shard of this rank
    integ, cam, scene = ex.integrator, ex.cam, ex.scene
    spp = wl["spp"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % local)     # > 126 MB L2

    def one_step():
        ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
        integ.render_frames(spp, stats=False)           # asynchronous: the film reduce is enqueued right behind the last kernel
        ctx.film_reduce()                               # ncclReduce on the render stream (no-op for one rank)

    sampler = ClockSampler(local)                # NVML start-up before the warm-up, sampling only inside the timed region
    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize(local)
    D.barrier()
    sampler.start()
    rays = 0; launches = 0; ms = 0.0
    t_wall0 = time.perf_counter()
    for _ in range(steps):
        with torch.cuda.stream(stream):
            flush.zero_()                        # L2 flush between timed iterations (outside the timed events)
        torch.cuda.synchronize(local)
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            one_step()
            e1.record(stream)
        torch.cuda.synchronize(local)
        ms += e0.elapsed_time(e1)
        st = ctx.stats()
        rays += int(st["rays_closest"]) + int(st["rays_shadow"]); launches += int(st["kernel_launches"])
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    # the film reduce on its own (N > 1): all ranks enter together, CUDA events around 5 back-to-back ncclReduce calls
    reduce_ms = 0.0
    if world > 1:
        ctx.film_reduce(); torch.cuda.synchronize(local); D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(5):
                ctx.film_reduce()
            e1.record(stream)
        torch.cuda.synchronize(local)
        reduce_ms = D.reduce([e0.elapsed_time(e1) / 5.0], "max")[0]
    per_rank = D.gather([ms / steps, rays / steps])
    ms_max = D.reduce([ms], "max")[0]
    rays_all, launches_all = D.reduce([rays, launches], "sum")
    rays_all, launches_all = int(rays_all), int(launches_all)
    value = rays_all / (ms_max * 1e-3) / 1e6
    paths_in_flight = int(st["paths_in_flight"])

    # ---- e2e: public API, host buffers, H2D + build + render + reduce + tone map + D2H inside the timed region
    h2d = scene.vertex_np.nbytes + scene.primitive_np.nbytes + scene.material_np.nbytes + scene.env.np_img.nbytes + 64 + 64 + 12
    if wl.get("spectral"):                       # sensor, rgb2spec table, four spectra, sky state
        h2d += integ.data_np.nbytes + integ.rgb2spec.table_data_np.nbytes + integ.rgb2spec.table_scale_np.nbytes + 113 * 4
        h2d += sum(sp.data_np.nbytes for sp in (integ.d65, integ.white, integ.red, integ.green))
    d2h = wl["W"] * wl["H"] * 12 if rank == 0 else 0              # only the root presents (and downloads) the tone-mapped image
    n_e2e = max(1, min(steps, 3))
    e2e_rays = 0; e2e_t = 0.0; build_ms = 0.0; checksum = 0.0
    for k in range(n_e2e + 1):
        torch.cuda.synchronize(local)
        D.barrier()
        t0 = time.perf_counter()
        if wl.get("spectral"):
            integ.setup_data_gpu()               # spectral tables H2D + white-point normalisation
        scene.setup_data_gpu()                   # tables H2D (through pinned staging) + env + LBVH build
        ta = time.perf_counter()
        if wl["normals"]:
            scene.process_normal()
        cam.dirty = True
        ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
        tb = time.perf_counter()
        integ.render_frames(spp, stats=False)
        tc = time.perf_counter()
        ctx.film_reduce()
        if rank == 0:
            UF.tone_map(0.5, integ.hdr, integ.rgb_film)
            # what the reference's loop hands to the user every pass is rgb_film (gui.set_image / ti.imwrite,
            # example/Example.py:43-53): DMA it into the context's pinned host buffer and read it on the host
            _, rgb_host = ctx.film_download(False, True, view=True)
            checksum = float(rgb_host[::37, ::41].sum())
        torch.cuda.synchronize(local)
        dt = time.perf_counter() - t0
        st2 = ctx.stats()
        build_ms = st2["ms_build"]
        if args.verbose and rank == 0:
            print("e2e pass %d (%s): upload+build %.2f ms, normals %.2f, render enqueue %.2f (device %.2f), reduce+tonemap+download (incl. waiting for the render) %.2f" %
                  (k, name, (ta - t0) * 1e3, (tb - ta) * 1e3, (tc - tb) * 1e3, st2["ms_total"], (t0 + dt - tc) * 1e3), file=sys.stderr)
        if k > 0:                                # first pass is warm-up (graph re-capture after the rebuild)
            e2e_t += dt; e2e_rays += int(st2["rays_closest"]) + int(st2["rays_shadow"])
    e2e_t = D.reduce([e2e_t], "max")[0]
    e2e_rays = int(D.reduce([e2e_rays], "sum")[0])
    h2d_all, d2h_all = (int(v) for v in D.reduce([h2d, d2h], "sum"))
    e2e_value = e2e_rays / e2e_t / 1e6

    cfg = workload_config(wl)                # identical to the reference arm's `config`: same workload, same samples per step
    out = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms_max / steps, "ms_per_spp": ms_max / steps / spp, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "reference assets (model/*.obj), counter-based RNG seed 0",
           "config": cfg, "rays_per_step": rays_all // steps,
           "run": {"paths_in_flight": paths_in_flight,
                   "tile_shard": "32x32 tiles, rank=(tx+3ty)%N, 1 ncclReduce per step (tr_film_reduce)" if world > 1 else "none"},
           "per_rank": [{"rank": r, "ms_per_step": v[0], "rays_per_step": int(v[1])} for r, v in enumerate(per_rank)],
           "wall_ms_per_step": wall / steps * 1e3, "clocks": clocks, "gpu_launches": launches_all, "nccl_reduce_ms": reduce_ms,
           "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                   "ms_per_step": e2e_t / n_e2e * 1e3, "film_checksum": checksum},
           "bvh_build": {"primitives": int(scene.primitive_count), "device_ms": build_ms,
                         "mprims_per_s": scene.primitive_count / max(build_ms, 1e-6) / 1e3,
                         "note": "Morton + 4-pass radix sort + Karras + refit + flatten, inside the e2e region of every step (SURVEY 8d)"}}

    # ---- roofline of the trace / shade / shadow kernels + cpu baseline: rank 0, N = 1 only
    if world == 1 and detail and wl.get("bdpt"):
        bdpt_roofline(out, args, name, wl, ex, ctx, local, spp)
    elif world == 1 and detail:
        pt_roofline(out, args, name, wl, ex, ctx, local, spp, paths_in_flight)
    ctx.comm_destroy()
    return out


def pt_roofline(out, args, name, wl, ex, ctx, local, spp, paths_in_flight):
    import _native
    integ, cam, scene = ex.integrator, ex.cam, ex.scene
    ctx.set_option("stage_timing", 1)
    ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
    stt = integ.render_frames(spp)
    ctx.set_option("stage_timing", 0)
    n_batches = (spp * wl["W"] * wl["H"] + paths_in_flight - 1) // paths_in_flight
    n_stage_launches = n_batches * integ.max_depth
    # visit counts of the same traversal policy from the counters flavour of the library: the same host classes
    # drive a second context
    cctx = _native.Context(local, "libtiray_counters.so")
    main_ctx, _native._ctx = _native._ctx, cctx
    try:
        integ.setup_data_cpu(); integ.setup_data_gpu(); scene.setup_data_gpu()
        if wl["normals"]:
            scene.process_normal()
        cam.dirty = True; cam.frame = 0; cam.frame_cpu[0] = 0
        integ.render_frames(spp)
        cs = cctx.stats()
    finally:
        _native._ctx = main_ctx
        cam.dirty = True
    cctx.close()
    peak, how = measured_peak_gbs()
    # SURVEY 8d accounting.  trace: 48 B/ray stream + 32 B per internal-node visit + 68 B per leaf test
    algo_bytes = 48 * cs["rays_closest"] + 32 * cs["node_visits"] + 68 * cs["leaf_tests"]
    achieved = algo_bytes / (stt["ms_trace"] * 1e-3) / 1e9
    # shade: 176 B stream per Disney / glass vertex, + 48 B when a NEE sample is emitted (= 224), 112 B per terminal vertex
    n_vert = cs["rays_closest"]; n_term = min(cs["shade_terminal"], n_vert)
    shade_bytes = 176 * (n_vert - n_term) + 112 * n_term + 48 * cs["rays_shadow"]
    shade_ach = shade_bytes / max(stt["ms_shade"] * 1e-3, 1e-9) / 1e9
    # shadow: 48 B queue record + 32 / 68 B per node visit / leaf test (+ the 16 B target leaf fetch is one of the leaf tests)
    shadow_bytes = 48 * cs["rays_shadow"] + 32 * cs["node_visits_shadow"] + 68 * cs["leaf_tests_shadow"]
    shadow_ach = shadow_bytes / max(stt["ms_shadow"] * 1e-3, 1e-9) / 1e9
    out["roofline"] = {"kernel": "k_trace (closest hit)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                       "frac": achieved / peak, "traffic": ncu_traffic(name, "k_trace"), "peak_source": how,
                       "traffic_note": "ncu dram__bytes_read+write per launch (profiles/r0N_dram_traffic.json): far below the algorithmic bytes because the BVH is cache resident",
                       "ncu": ncu_issue(name),
                       "launches_per_step": n_stage_launches, "avg_launch_ms": stt["ms_trace"] / n_stage_launches,
                       "algorithmic_bytes_per_step": int(algo_bytes), "bytes_per_ray": algo_bytes / max(1, cs["rays_closest"]),
                       "node_visits_per_ray": cs["node_visits"] / max(1, cs["rays_closest"]),
                       "leaf_tests_per_ray": cs["leaf_tests"] / max(1, cs["rays_closest"]),
                       "compulsory_dram_bytes_per_ray": 48,
                       "stage_ms_per_step": {"trace": stt["ms_trace"], "shade": stt["ms_shade"], "shadow": stt["ms_shadow"], "total": stt["ms_total"]},
                       "note": "effective bandwidth: the BVH is SMEM/L2 resident, compulsory DRAM traffic is the 48 B/ray queue stream",
                       "kernels": {
                           "k_shade": {"bound": "hbm", "achieved": shade_ach, "peak": peak, "unit": "GB/s", "frac": shade_ach / peak,
                                       "traffic": ncu_traffic(name, "k_shade"), "launches_per_step": n_stage_launches,
                                       "avg_launch_ms": stt["ms_shade"] / n_stage_launches, "algorithmic_bytes_per_step": int(shade_bytes),
                                       "bytes_per_vertex": shade_bytes / max(1, n_vert), "vertices_per_step": int(n_vert), "terminal_vertices": int(n_term),
                                       "note": "SURVEY 8d: 176 B stream per Disney / glass vertex (+48 B with a NEE sample = 224), 112 B per terminal vertex"},
                           "k_shadow": {"bound": "hbm", "achieved": shadow_ach, "peak": peak, "unit": "GB/s", "frac": shadow_ach / peak,
                                        "traffic": ncu_traffic(name, "k_shadow"), "launches_per_step": n_stage_launches,
                                        "avg_launch_ms": stt["ms_shadow"] / n_stage_launches, "algorithmic_bytes_per_step": int(shadow_bytes),
                                        "bytes_per_ray": shadow_bytes / max(1, cs["rays_shadow"]),
                                        "node_visits_per_ray": cs["node_visits_shadow"] / max(1, cs["rays_shadow"]),
                                        "leaf_tests_per_ray": cs["leaf_tests_shadow"] / max(1, cs["rays_shadow"]),
                                        "note": "48 B queue record + 32 B per node visit + 68 B per leaf test; effective bandwidth (tree on chip)"}}}
    if not args.no_cpu:
        sample_spp = 64 if name in ("cornell", "spectral_box") else 16     # a few seconds on 16 threads, ~15 s on 4
        cpu_reference_run(wl, 1)                                   # warm the pages / OpenMP pool
        v, r, dt, cores, _ = cpu_reference_run(wl, sample_spp)
        out["cpu_baseline"] = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                               "sample": "%dx%d, %d of %d spp, %.1f s, reference-algorithm CPU restatement (Taichi unavailable)" % (wl["W"], wl["H"], sample_spp, spp, dt)}


def run_native(args):
    # stdout carries exactly ONE JSON line: everything else this process (or a library it loads: NCCL prints its version banner
    # at communicator creation) writes to file descriptor 1 goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import parallel
    rank, world, local = parallel.init_process_group("nccl" if args.gpus > 1 else None)
    if world != args.gpus and args.gpus > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torchrun)" % (args.gpus, world))
    torch.cuda.set_device(local)
    D = Dist(world, local)
    out = measure_native(args, args.workload, args.steps, args.warmup, rank, world, local, D)
    if args.workload == "cornell" and not args.no_sub:
        # BASELINE configs[2] (the scene the >= 6x scaling target is defined on) in the same process, at every N
        sub = measure_native(args, "teapot_mc", max(1, min(args.steps, 5)), max(3, min(args.warmup, 3)), rank, world, local, D)
        out["workloads"] = {"teapot_mc": sub}
        out["gpu_launches_all_workloads"] = out["gpu_launches"] + sub["gpu_launches"]
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()


def bdpt_roofline(out, args, name, wl, ex, ctx, local, spp):
    """roofline of the dominant BDPT kernel + cpu baseline.  The two traversal kernels dominate the step: the persistent
    closest-hit kernel k_trace (6 launches per batch: 48 B ray / hit stream + 32 B per internal-node visit + 68 B per leaf
    test) and the connection query kernel k_shadow<QUERY> (1 launch per batch: 32 B queue entry + 4 B result + the
    same per-visit bytes); visit counts come from the counters build of the same kernels, times from CUDA events around
    every launch (stage-timing pass)."""
    import _native
    integ, cam, scene = ex.integrator, ex.cam, ex.scene
    ctx.set_option("stage_timing", 1)
    ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
    stt = integ.render_frames(spp)
    ms_ktrace, ms_kshadow = ctx.bdpt_kernel_ms()
    ctx.set_option("stage_timing", 0)
    cctx = _native.Context(local, "libtiray_counters.so")
    main_ctx, _native._ctx = _native._ctx, cctx
    try:
        integ.setup_data_cpu(); integ.setup_data_gpu(); scene.setup_data_gpu(); scene.process_normal()
        cam.dirty = True; cam.frame = 0; cam.frame_cpu[0] = 0
        integ.render_frames(spp)
        cs = cctx.stats()
    finally:
        _native._ctx = main_ctx
        cam.dirty = True
    cctx.close()
    n_batches = max(1, (spp * wl["W"] * wl["H"] + int(stt["paths_in_flight"]) - 1) // int(stt["paths_in_flight"]))
    q_bytes = 36 * cs["rays_shadow"] + 32 * cs["node_visits_shadow"] + 68 * cs["leaf_tests_shadow"]
    t_bytes = 48 * cs["rays_closest"] + 32 * cs["node_visits"] + 68 * cs["leaf_tests"]
    peak, how = measured_peak_gbs()
    kq = {"kernel": "k_shadow<QUERY> (connection visibility queries)", "achieved": q_bytes / (ms_kshadow * 1e-3) / 1e9, "launches_per_step": n_batches,
          "avg_launch_ms": ms_kshadow / n_batches, "algorithmic_bytes_per_step": int(q_bytes), "bytes_per_ray": q_bytes / max(1, cs["rays_shadow"]),
          "node_visits_per_ray": cs["node_visits_shadow"] / max(1, cs["rays_shadow"]), "leaf_tests_per_ray": cs["leaf_tests_shadow"] / max(1, cs["rays_shadow"])}
    kt = {"kernel": "k_trace (closest hit, sub-path segments)", "achieved": t_bytes / (ms_ktrace * 1e-3) / 1e9, "launches_per_step": 6 * n_batches,
          "avg_launch_ms": ms_ktrace / (6 * n_batches), "algorithmic_bytes_per_step": int(t_bytes), "bytes_per_ray": t_bytes / max(1, cs["rays_closest"]),
          "node_visits_per_ray": cs["node_visits"] / max(1, cs["rays_closest"]), "leaf_tests_per_ray": cs["leaf_tests"] / max(1, cs["rays_closest"])}
    top, other = (kq, kt) if ms_kshadow >= ms_ktrace else (kt, kq)
    out["roofline"] = dict(top, bound="hbm", peak=peak, unit="GB/s", frac=top["achieved"] / peak,
                           traffic=ncu_traffic(name, "k_trace" if top is kt else "k_shadow"), peak_source=how, ncu=ncu_issue(name),
                           second_kernel=dict(other, frac=other["achieved"] / peak),
                           stage_ms_per_step={"sub-paths (generate + 6 x (trace, vertex))": stt["ms_trace"], "of which k_trace": ms_ktrace,
                                              "connections (gen + query + eval)": stt["ms_shadow"], "of which k_shadow<QUERY>": ms_kshadow,
                                              "items + film": stt["ms_shade"], "total": stt["ms_total"]},
                           note="effective bandwidth: the 11.5 k-triangle BVH is L1/L2 resident; compulsory DRAM traffic is the ray / vertex / item / "
                                "contribution streams")
    if not args.no_cpu:
        cpu_reference_run(wl, 1)
        v, r, dt, cores, _ = cpu_reference_run(wl, 8)
        out["cpu_baseline"] = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                               "sample": "%dx%d, 8 of %d spp, %.1f s, reference-algorithm CPU restatement (Taichi unavailable)" % (wl["W"], wl["H"], spp, dt)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cornell", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sub", action="store_true", help="skip the workloads.teapot_mc sub-record of the default run")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, WORKLOADS[args.workload])
    else:
        run_native(args)


if __name__ == "__main__":
    main()
