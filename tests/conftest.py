import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ti-raytrace_b200")
for p in (PKG, os.path.join(PKG, "integrator"), os.path.join(PKG, "example"), os.path.join(PKG, "brdf"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def model(name):
    return os.path.join(PKG, "model", name)


SCENES = {
    "cornell": dict(files=["cornell_box.obj"]),
    "sphere": dict(files=["sphere.obj"]),
    "teapot": dict(files=["Teapot.obj"]),
    "teapot_mc": dict(files=["mc.obj", "Teapot.obj"]),
    "veach": dict(files=["bdpt.obj"]),
}


@pytest.fixture(scope="session")
def oracle_tables():
    """name -> oracle-side packed tables (oracle/objload.py), cached per session"""
    from oracle import objload
    cache = {}

    def get(name, sphere_light=False, glass0=False, spectral_walls=False, mirror0=False, beam_lights=False):
        key = (name, sphere_light, glass0, spectral_walls, mirror0, beam_lights)
        if key not in cache:
            shapes = [objload.sphere_light_rows()] if sphere_light else []
            if beam_lights:
                shapes += [([float(t), *pos, p0, p1, p2, *nor], [objload.MAT_LIGHT, 0.0, *col, 0, 0, 0, 0, 0])
                           for t, pos, (p0, p1, p2), nor, col in BEAM_LIGHTS]

            def edit(mats):
                if glass0:
                    mats[0][0] = 1.0; mats[0][5] = 1.3; mats[0][6] = 5.0
                if mirror0:                              # example/sky_dome.py:18-19
                    mats[0][5] = 1.0; mats[0][6] = 0.0
                if spectral_walls:                       # example/spectral_box.py:22-27
                    for k in range(3):
                        mats[k][0] = 10.0; mats[k][1] = float(k)
            cache[key] = objload.load_scene([model(f) for f in SCENES[name]["files"]], shapes=shapes, material_edit=edit)
        return cache[key]
    return get


# a laser (SHPAE_LASER = 4: radius, normal; example/prism_rainbow.py:37-50) and a spot light (SHPAE_SPOT = 3: xita1, xita2, scale,
# normal) inside the Cornell box, both pointing down: (type, pos, params 0..2, normal, colour)
BEAM_LIGHTS = [(4, (278.0, 500.0, -279.6), (60.0, 0.0, 0.0), (0.0, -1.0, 0.0), (4.0e5, 1.0e5, 1.0e5)),
               (3, (150.0, 450.0, -200.0), (0.3, 0.6, 1.0), (0.0, -1.0, 0.0), (2.0e4, 8.0e4, 2.0e4))]


def make_product_scene(name, sphere_light=False, glass0=False, env_power=0.0, spectral_walls=False, mirror0=False, beam_lights=False):
    """product-side Scene (host packing only; no device calls)"""
    import Scene
    import SceneData as SCD
    s = Scene.Scene()
    for f in SCENES[name]["files"]:
        s.add_obj("model/" + f)
    if glass0:
        m = s.material_cpu[0]; m.type = SCD.MAT_GLASS; m.setIor(1.3); m.setExtinciton(5.0)
    if mirror0:
        s.material_cpu[0].setMetal(1.0); s.material_cpu[0].setRough(0.0)
    if spectral_walls:
        for k in range(3):
            s.material_cpu[k].type = SCD.MAT_SPECTRAL; s.material_cpu[k].alebdoTex = k
    if sphere_light:
        sh = SCD.Shape(); sh.type = SCD.SHPAE_SPHERE; sh.pos = [0.0, 20.0, 0.0]; sh.setRadius(5.0)
        mt = SCD.Material(); mt.type = SCD.MAT_LIGHT; mt.setColor([50.0, 50.0, 50.0])
        s.add_shape(sh, mt)
    if beam_lights:
        for t, pos, (p0, p1, p2), nor, col in BEAM_LIGHTS:
            sh = SCD.Shape(); sh.type = t; sh.pos = list(pos)
            if t == SCD.SHPAE_LASER:
                sh.setRadius(p0)
            else:
                sh.setXita(p0, p1); sh.setScale(p2)
            sh.setNormal(list(nor))
            mt = SCD.Material(); mt.type = SCD.MAT_LIGHT; mt.setColor(list(col))
            s.add_shape(sh, mt)
    if env_power:
        s.add_env("image/env.png", env_power)
    return s


@pytest.fixture()
def gpu_ctx():
    """fresh device context (ti.init semantics)"""
    import _native
    return _native.reset_context()
