#!/usr/bin/env python
"""Developer probe: time the PT_RGB hot path for a workload under different library flavours / options.
  python tools/perf_probe.py --workload cornell --lib libtiray.so --batch 0,4,16,64 --reps 3
Prints one line per configuration: ms/step, Mrays/s, stage split (second pass with stage timing)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ti-raytrace_b200")
for p in (PKG, os.path.join(PKG, "integrator"), os.path.join(PKG, "example"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cornell")
ap.add_argument("--lib", default="libtiray.so")
ap.add_argument("--batch", default="0")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--opts", default="", help="name=value,... passed to tr_set_option")
ap.add_argument("--counters", action="store_true")
ap.add_argument("--max-depth", type=int, default=15)
ap.add_argument("--shard", default="", help="rank,nranks: emulate one rank of a tile-sharded run")
ap.add_argument("--e2e", action="store_true", help="time the pieces of the host-buffer path")
args = ap.parse_args()
os.environ["TIRAY_LIB"] = args.lib

import bench  # noqa: E402
import _native  # noqa: E402

wl = bench.WORKLOADS[args.workload]
ex = bench.build_example(wl)
ctx = _native.context()
for kv in filter(None, args.opts.split(",")):
    k, v = kv.split("="); ctx.set_option(k, int(v))
integ, cam = ex.integrator, ex.cam
integ.max_depth = args.max_depth
if args.shard:
    r_, n_ = (int(x) for x in args.shard.split(",")); ctx.set_shard(r_, n_)
if args.e2e:
    import time
    sc = ex.scene
    for rep in range(3):
        ctx.synchronize()
        t = [time.perf_counter()]

        def mark(sync=True):
            if sync:
                ctx.synchronize()
            t.append(time.perf_counter())
        ctx.scene_upload(sc.vertex_np, sc.primitive_np, sc.material_np, sc.shape_np if sc.shape_count else None,
                         sc.light_np if sc.light_count else None, sc.minboundarynp, sc.maxboundarynp); mark(False); mark()
        sc.env.setup_data_gpu(sc.env_power); mark()
        ctx.bvh_build(); mark(False); mark()
        build_ms = ctx.stats()["ms_build"]
        if wl["normals"]:
            ctx.process_normal()
        mark()
        cam.dirty = True; ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
        integ.render_frames(wl["spp"], stats=False); mark(False); mark()
        ctx.film_reduce(); ctx.tonemap(0.5); mark()
        ctx.film_download(False, True, view=True); mark()
        names = ["scene_upload(host)", "+sync", "env_upload", "bvh_build(host)", "+sync", "process_normal", "render(enqueue)", "+sync", "reduce+tonemap", "download rgb"]
        print("e2e pieces ms:", ", ".join("%s %.2f" % (n, (b - a) * 1e3) for n, a, b in zip(names, t[:-1], t[1:])),
              "| total %.2f | bvh device ms %.3f" % ((t[-1] - t[0]) * 1e3, build_ms), flush=True)
for b in [int(x) for x in args.batch.split(",")]:
    ctx.set_option("batch_frames", b)
    best = None
    for r in range(args.reps + 1):
        ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
        st = integ.render_frames(wl["spp"])
        if (r > 0 or args.reps == 0) and (best is None or st["ms_total"] < best["ms_total"]):
            best = st
    ctx.set_option("stage_timing", 1)
    ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
    tt = integ.render_frames(wl["spp"])
    ctx.set_option("stage_timing", 0)
    rays = best["rays_closest"] + best["rays_shadow"]
    line = "%s %s batch=%d paths=%d: %.2f ms/step %.1f Mrays/s | stages trace %.2f shade %.2f shadow %.2f (total %.2f) | rays c=%d s=%d" % (
        args.workload, args.lib, b, best["paths_in_flight"], best["ms_total"], rays / best["ms_total"] / 1e3,
        tt["ms_trace"], tt["ms_shade"], tt["ms_shadow"], tt["ms_total"], best["rays_closest"], best["rays_shadow"])
    if args.counters:
        line += " | visits/ray closest n=%.2f l=%.2f shadow n=%.2f l=%.2f" % (
            best["node_visits"] / max(1, best["rays_closest"]), best["leaf_tests"] / max(1, best["rays_closest"]),
            best["node_visits_shadow"] / max(1, best["rays_shadow"]), best["leaf_tests_shadow"] / max(1, best["rays_shadow"]))
    print(line, flush=True)
