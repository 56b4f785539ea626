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
            print("%-46s vertex normals differing %d of %d, max abs %.3e" % (name + " process_normal", int((gv != vn).any(axis=1).sum()), gv.shape[0], float(np.nanmax(np.abs(gv - vn)))))
            t.vertex = gv
        o = build_oracle_scene(t, W, W, env_power=env)
        st = integ.render_frames(spp)
        g = integ.hdr.to_numpy()
        ref, cnt = o.render_pt_rgb(W, W, 0, spp)
        report("%s %dx%d x %d spp PT_RGB" % (name, W, W, spp), g, ref, st, cnt)
        _native.reset_context()


if __name__ == "__main__":
    main()
