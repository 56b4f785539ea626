#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: Mrays/s (and ms/spp) of PT_RGB.

  python bench.py --gpus N --steps K --warmup W [--workload cornell|teapot_mc|teapot_mc16|spectral_box] [--impl native|reference]

A step = one full pass of the hot path over one batch: clear the film, render `spp` samples per pixel
(cornell: 512x512, 64 spp = BASELINE configs[1]; teapot_mc: 1024x1024, 64 spp = configs[2]; teapot_mc16: 16 spp;
spectral_box: PT_Spec hero-wavelength 512x512, 64 spp = configs[3]) with the
scene, BVH and camera resident in HBM, and (N > 1) one NCCL sum-reduce of the film.  rays = closest-hit
traversals + shadow traversals actually executed (device queue counters).

  value      whole-job Mrays/s, timed with CUDA events on the stream the kernels run on, max over ranks
  e2e        the same metric through the public Python API with HOST buffers: scene tables H2D + LBVH
             build + render + tone map + film D2H inside the timed region
  roofline   the closest-hit trace kernel: algorithmic bytes (48 B/ray + 32 B per internal-node
             visit + 68 B per leaf test, visit counts from the counters build of the same kernel) / mean
             kernel time (CUDA events around every launch of it) vs the measured HBM copy bandwidth
  cpu_baseline / --impl reference: the reference algorithm restated on the CPU (oracle/, -O3 -ffast-math,
             OpenMP on all host cores) on a bounded sample of the same workload (Taichi is not installable).
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


# --------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
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
                "power_w_max": max(power) if power else None, "samples": len(sm)}


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(workload, kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel, from the committed ncu capture of the same
    command line (profiles/r01_dram_traffic.json, written from `ncu --metrics dram__bytes_*` by tools/); None if absent"""
    workload = {"teapot_mc": "teapot_mc16"}.get(workload, workload)      # 64 spp = four batches of the captured 16-frame batch: same launches
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_dram_traffic.json")))[workload]
        for k, v in t.items():
            if k.startswith(kernel_prefix):
                return v["dram_bytes_per_launch_all"]
    except Exception:
        pass
    return None


def ncu_issue(workload):
    """issue-slot utilisation etc. of the dominant kernel from the committed --set full capture (None if absent): the kernels are
    issue-bound, not DRAM-bound, which is why the effective-bandwidth fraction can exceed 1"""
    workload = {"teapot_mc": "teapot_mc16"}.get(workload, workload)
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r01_dram_traffic.json")))[workload].get("ncu_full")
    except Exception:
        return None


def oracle_tables(wl):
    from oracle import objload
    shapes = [objload.sphere_light_rows()] if wl["sphere_light"] else []
    return objload.load_scene([os.path.join(PKG, "model", f) for f in wl["files"]], shapes=shapes)


def cpu_reference_run(wl, spp, frame_begin=0):
    """the reference algorithm on the host cores (oracle, fast build): returns (Mrays/s, rays, seconds, threads)"""
    from oracle import oracle
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
def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_spp = 4 if args.workload in ("cornell", "spectral_box") else 1      # veach_bdpt: 1 spp = 1.2 M traversals, ~0.5 s
    for _ in range(args.warmup):
        cpu_reference_run(wl, 1)
    rays = 0; secs = 0.0; cores = 1
    for k in range(args.steps):
        _, r, dt, cores, _ = cpu_reference_run(wl, sample_spp)
        rays += r; secs += dt
    v = rays / secs / 1e6
    sample = "%dx%d, %d spp per step (of %d), all %d host threads, oracle/liboracle_fast.so" % (wl["W"], wl["H"], sample_spp, wl["spp"], cores)
    out = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "ms_per_spp": secs / args.steps / sample_spp * 1e3,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "reference assets (model/*.obj), seed 0",
           "config": {"workload": wl["desc"], "sample": sample},
           "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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


def run_native(args, wl):
    import numpy as np
    import torch
    import _native
    import parallel
    rank, world, local = parallel.init_process_group("nccl" if args.gpus > 1 else None)
    if world != args.gpus and args.gpus > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torchrun)" % (args.gpus, world))
    torch.cuda.set_device(local)
    import torch.distributed as dist
    ex = build_example(wl)                      # ti.init() -> context on LOCAL_RANK
    ctx = _native.context()
    ctx.set_shard(rank, world)
    stream = torch.cuda.Stream(device=local)
    ctx.stream_set(stream.cuda_stream)
    integ, cam = ex.integrator, ex.cam
    spp = wl["spp"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % local)     # > 126 MB L2
    film = parallel.film_tensor(ctx)            # torch view of the device film (zero copy), made once

    def one_step():
        ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
        integ.render_frames(spp, stats=False)           # asynchronous: the film reduce is enqueued right behind the last kernel
        if world > 1:
            with torch.cuda.stream(stream):
                parallel.reduce_film(film, dst=0)
        return ctx.stats()

    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize(local)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local); sampler.start()
    rays = 0; launches = 0; ms = 0.0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                        # L2 flush between timed iterations (outside the timed events)
        torch.cuda.synchronize(local)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            st = one_step()
            e1.record(stream)
        torch.cuda.synchronize(local)
        ms += e0.elapsed_time(e1)
        rays += int(st["rays_closest"]) + int(st["rays_shadow"]); launches += int(st["kernel_launches"])
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device="cuda:%d" % local); dist.all_reduce(tm, op=dist.ReduceOp.MAX); ms = float(tm.item())
        tr = torch.tensor([rays, launches], dtype=torch.int64, device="cuda:%d" % local); dist.all_reduce(tr, op=dist.ReduceOp.SUM)
        rays, launches = int(tr[0].item()), int(tr[1].item())
    value = rays / (ms * 1e-3) / 1e6
    paths_in_flight = int(st["paths_in_flight"])

    # ---- e2e: public API, host buffers, H2D + build + render + tone map + D2H inside the timed region
    import UtilsFunc as UF
    scene = ex.scene
    h2d = scene.vertex_np.nbytes + scene.primitive_np.nbytes + scene.material_np.nbytes + scene.env.np_img.nbytes + 64 + 64 + 12
    if wl.get("spectral"):                       # sensor, rgb2spec table, four spectra, sky state
        h2d += integ.data_np.nbytes + integ.rgb2spec.table_data_np.nbytes + integ.rgb2spec.table_scale_np.nbytes + 113 * 4
        h2d += sum(sp.data_np.nbytes for sp in (integ.d65, integ.white, integ.red, integ.green))
    d2h = 2 * wl["W"] * wl["H"] * 12
    e2e_rays = 0; e2e_t = 0.0
    for k in range(max(1, min(args.steps, 3)) + 1):
        torch.cuda.synchronize(local)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if wl.get("spectral"):
            integ.setup_data_gpu()               # spectral tables H2D + white-point normalisation
        scene.setup_data_gpu()                   # tables H2D (pageable numpy) + env + LBVH build
        ta = time.perf_counter()
        build_ms = ctx.stats()["ms_build"]       # device time of the LBVH build just done (nothing pending: no wait)
        if wl["normals"]:
            scene.process_normal()
        cam.dirty = True
        ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
        tb = time.perf_counter()
        integ.render_frames(spp, stats=False)
        tc = time.perf_counter()
        if world > 1:
            with torch.cuda.stream(stream):
                parallel.reduce_film(parallel.film_tensor(ctx), dst=0)
        UF.tone_map(0.5, integ.hdr, integ.rgb_film)
        hdr_host, rgb_host = ctx.film_download(True, True)
        torch.cuda.synchronize(local)
        dt = time.perf_counter() - t0
        st2 = ctx.stats()
        if args.verbose and rank == 0:
            print("e2e pass %d: upload+build %.2f ms, normals %.2f, render enqueue %.2f (device %.2f), reduce+tonemap+download (incl. waiting for the render) %.2f" %
                  (k, (ta - t0) * 1e3, (tb - ta) * 1e3, (tc - tb) * 1e3, st2["ms_total"], (t0 + dt - tc) * 1e3), file=sys.stderr)
        if k > 0:                                # first pass is warm-up (graph re-capture after the rebuild)
            e2e_t += dt; e2e_rays += int(st2["rays_closest"]) + int(st2["rays_shadow"])
    if world > 1:
        tm = torch.tensor([e2e_t], dtype=torch.float64, device="cuda:%d" % local); dist.all_reduce(tm, op=dist.ReduceOp.MAX); e2e_t = float(tm.item())
        tr = torch.tensor([e2e_rays], dtype=torch.int64, device="cuda:%d" % local); dist.all_reduce(tr, op=dist.ReduceOp.SUM); e2e_rays = int(tr.item())
    e2e_value = e2e_rays / e2e_t / 1e6

    out = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "ms_per_spp": ms / args.steps / spp, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "reference assets (model/*.obj), counter-based RNG seed 0",
           "config": {"workload": wl["desc"], "rays_per_step": rays // args.steps, "paths_in_flight": paths_in_flight,
                      "tile_shard": "32x32 tiles, rank=(tx+3ty)%N, 1 NCCL reduce per step" if world > 1 else "none",
                      "l2": "L2 flushed (256 MiB write) between timed steps; per-step queue traffic also exceeds L2"},
           "wall_ms_per_step": wall / args.steps * 1e3, "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": e2e_t / max(1, min(args.steps, 3)) * 1e3},
           "bvh_build": {"primitives": int(scene.primitive_count), "device_ms": build_ms,
                         "mprims_per_s": scene.primitive_count / max(build_ms, 1e-6) / 1e3, "launches": 16,
                         "note": "Morton + 4-pass radix sort + Karras + refit + flatten, inside the e2e region of every step (SURVEY 8d)"}}

    # ---- roofline of the dominant kernel (closest-hit trace) + cpu baseline: rank 0, N = 1 only
    if world == 1 and wl.get("bdpt"):
        bdpt_roofline(out, args, wl, ex, ctx, local, spp)
    elif world == 1:
        ctx.set_option("stage_timing", 1)
        ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
        stt = integ.render_frames(spp)
        ctx.set_option("stage_timing", 0)
        n_batches = (spp * wl["W"] * wl["H"] + paths_in_flight - 1) // paths_in_flight
        n_trace_launches = n_batches * integ.max_depth
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
        algo_bytes = 48 * cs["rays_closest"] + 32 * cs["node_visits"] + 68 * cs["leaf_tests"]
        trace_s = stt["ms_trace"] * 1e-3
        peak, how = measured_peak_gbs()
        achieved = algo_bytes / trace_s / 1e9
        out["roofline"] = {"kernel": "k_trace (closest hit)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                           "frac": achieved / peak, "traffic": ncu_traffic(args.workload, "k_trace"), "peak_source": how,
                           "traffic_note": "ncu dram__bytes_read+write per launch (profiles/r01_dram_traffic.json): far below the algorithmic bytes because the BVH is cache resident",
                           "ncu": ncu_issue(args.workload),
                           "launches_per_step": n_trace_launches, "avg_launch_ms": stt["ms_trace"] / n_trace_launches,
                           "algorithmic_bytes_per_step": int(algo_bytes), "bytes_per_ray": algo_bytes / max(1, cs["rays_closest"]),
                           "node_visits_per_ray": cs["node_visits"] / max(1, cs["rays_closest"]),
                           "leaf_tests_per_ray": cs["leaf_tests"] / max(1, cs["rays_closest"]),
                           "compulsory_dram_bytes_per_ray": 48,
                           "stage_ms_per_step": {"trace": stt["ms_trace"], "shade": stt["ms_shade"], "shadow": stt["ms_shadow"], "total": stt["ms_total"]},
                           "note": "effective bandwidth: the BVH is SMEM/L2 resident, compulsory DRAM traffic is the 48 B/ray queue stream"}
        if not args.no_cpu:
            sample_spp = 64 if args.workload in ("cornell", "spectral_box") else 16     # a few seconds on 16 threads, ~15 s on 4
            cpu_reference_run(wl, 1)                                   # warm the pages / OpenMP pool
            v, r, dt, cores, _ = cpu_reference_run(wl, sample_spp)
            out["cpu_baseline"] = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                   "sample": "%dx%d, %d of %d spp, %.1f s, reference-algorithm CPU restatement (Taichi unavailable)" % (wl["W"], wl["H"], sample_spp, spp, dt)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def bdpt_roofline(out, args, wl, ex, ctx, local, spp):
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
                           traffic=ncu_traffic(args.workload, "k_trace" if top is kt else "k_shadow"), peak_source=how, ncu=ncu_issue(args.workload),
                           second_kernel=dict(other, frac=other["achieved"] / peak),
                           stage_ms_per_step={"sub-paths (generate + 6 x (trace, vertex))": stt["ms_trace"], "of which k_trace": ms_ktrace,
                                              "connections (gen + query + eval)": stt["ms_shadow"], "of which k_shadow<QUERY>": ms_kshadow,
                                              "items + film": stt["ms_shade"], "total": stt["ms_total"]},
                           note="effective bandwidth: the 11.5 k-triangle BVH (1.5 MB of 64-byte nodes + 0.55 MB of leaf records) is L1/L2 "
                                "resident; compulsory DRAM traffic is the ray / vertex / item / contribution streams")
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
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_native(args, wl)


if __name__ == "__main__":
    main()
