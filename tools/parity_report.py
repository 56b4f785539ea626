#!/usr/bin/env python
"""Developer report: how exact is the CUDA path against the oracle now that both share include/trmath.h?
Prints, per scenario of the GPU parity tests, the number of film words / vertices / strategies that differ at all."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: E402,F401  (sets up sys.path)
import numpy as np  # noqa: E402
import _native  # noqa: E402
from oracle import oracle, objload  # noqa: E402
from test_gpu_parity import build_gpu_scene, build_oracle_scene  # noqa: E402


def tables(name, **kw):
    shapes = [objload.sphere_light_rows()] if kw.get("sphere_light") else []

    def edit(mats):
        if kw.get("glass0"):
            mats[0][0] = 1.0; mats[0][5] = 1.3; mats[0][6] = 5.0
    return objload.load_scene([conftest.model(f) for f in conftest.SCENES[name]["files"]], shapes=shapes, material_edit=edit)


def report(tag, g, ref, st=None, cnt=None):
    diff = g != ref
    bad = diff.any(axis=-1) if g.ndim == 3 else diff
    line = "%-46s differing pixels %7d of %8d, max abs diff %.3e, max rel %.3e" % (
        tag, int(bad.sum()), bad.size, float(np.abs(g - ref).max()), float((np.abs(g - ref) / np.maximum(np.abs(ref), 1e-3)).max()))
    if st is not None:
        line += " | rays closest %d vs %d, shadow %d vs %d" % (st["rays_closest"], cnt["closest"], st["rays_shadow"], cnt["shadow"])
    print(line, flush=True)


def main():
    _native.reset_context()
    for name, W, spp, kw, env in [("cornell", 256, 4, {}, 0.0), ("cornell", 512, 64, {}, 0.0),
                                  ("sphere", 128, 4, dict(sphere_light=True, glass0=True), 5.0),
                                  ("teapot_mc", 128, 2, dict(sphere_light=True), 5.0)]:
        scene, cam, integ = build_gpu_scene(name, W, W, env_power=env, **kw)
        t = tables(name, **kw)
        if env:
            scene.process_normal()
            t = type("T", (), {})(); t.__dict__.update(tables(name, **kw).__dict__)
            o0 = build_oracle_scene(t, W, W, env_power=env); vn = o0.process_normal()
            gv = scene.vertex.to_numpy()
            nan_rows = int(np.isnan(vn).any(axis=1).sum())
            neq = ~((gv == vn) | (np.isnan(gv) & np.isnan(vn)))          # NaN for NaN counts as equal (degenerate triangles, Scene.py:754-798)
            print("%-46s vertex normals differing %d of %d (%d rows are NaN on both sides), max abs %.3e" % (
                name + " process_normal", int(neq.any(axis=1).sum()), gv.shape[0], nan_rows, float(np.nanmax(np.abs(gv - vn)))))
            t.vertex = gv
        o = build_oracle_scene(t, W, W, env_power=env)
        st = integ.render_frames(spp)
        g = integ.hdr.to_numpy()
        ref, cnt = o.render_pt_rgb(W, W, 0, spp)
        report("%s %dx%d x %d spp PT_RGB" % (name, W, W, spp), g, ref, st, cnt)
        _native.reset_context()
    # ---- PT_Spec
    import test_gpu_spectral as tgs

    def fixture_tables(name, **kw):                      # what the pytest fixture `oracle_tables` hands out
        shapes = [objload.sphere_light_rows()] if kw.get("sphere_light") else []

        def edit(mats):
            if kw.get("spectral_walls"):
                for k in range(3):
                    mats[k][0] = 10.0; mats[k][1] = float(k)
        return objload.load_scene([conftest.model(f) for f in conftest.SCENES[name]["files"]], shapes=shapes, material_edit=edit)
    try:
        from oracle import spectral as ospec
        scene, cam, integ = tgs.build_gpu_spectral(128, 128)
        st = integ.render_frames(4)
        g = integ.hdr.to_numpy()
        o = tgs.spectral_oracle(fixture_tables, 128, 128)
        ref, cnt = ospec.render_pt_spec(o, 128, 128, 0, 4)
        report("spectral_box 128x128 x 4 spp PT_Spec", g, ref, st, cnt)
    except Exception as e:          # the report is a developer aid: do not die on a fixture mismatch
        print("spectral report skipped:", repr(e))
    _native.reset_context()
    # ---- BDPT
    import test_gpu_bdpt as tgb
    for name, fit, smooth in [("cornell", 0.8, False), ("veach", 0.5, True)]:
        W = 64
        scene, cam, integ = tgb.build_gpu(name, W, W, fit, smooth)
        t = objload.load_scene([conftest.model(f) for f in conftest.SCENES[name]["files"]])
        o = tgb.build_oracle(t, W, W, fit, smooth)
        rng = np.random.RandomState(1)
        px = rng.randint(0, W, 300).astype(np.int32); py = rng.randint(0, W, 300).astype(np.int32)
        ctx = _native.context()
        ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
        integ.render()
        verts, depths, contrib = ctx.test_bdpt_dump(px, py)
        bad_d = bad_v = bad_c = flips = nstrat = 0
        for k in range(px.size):
            ov, od, oc = o.bdpt_pixel_dump(int(px[k]), int(py[k]), 0)
            if tuple(depths[k]) != od:
                bad_d += 1; continue
            for v in list(range(od[0])) + [7 + i for i in range(od[1])]:
                if not np.array_equal(verts[k, v], ov[v], equal_nan=True):
                    bad_v += 1; break
            for e in range(2, od[0] + 1):
                for l in range(0, od[1] + 1):
                    a, b = contrib[k, e - 1, l, :3], oc[e - 1, l, :3]
                    if not a.any() and not b.any():
                        continue
                    nstrat += 1
                    if a.any() != b.any():
                        flips += 1
                    elif not np.array_equal(a, b):
                        bad_c += 1
        print("%-46s pixels with other depths %d, with a differing vertex word %d of %d; strategies: %d value mismatches, %d visibility flips of %d" % (
            "BDPT " + name, bad_d, bad_v, px.size, bad_c, flips, nstrat), flush=True)
        ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
        st = integ.render_frames(2)
        g = integ.hdr.to_numpy()
        ref, cnt = o.render_bdpt_rgb(W, W, 0, 2)
        report("BDPT %s %dx%d x 2 spp film" % (name, W, W), g, ref, st, cnt)
        _native.reset_context()


if __name__ == "__main__":
    main()
