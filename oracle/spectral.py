"""ORACLE (test infrastructure): host side of the spectral integrator restated — table loading of
integrator/PT_Spec.py:55-98, spectrum/Spectrum.py:18-34, spectrum/Rgb2Spec.py:15-39 and the Hosek-Wilkie
configuration blend of sky/Sky.py:52-159 (with portable paths: the reference opens "sky\\data.csv")."""
import ctypes as C
import math
import os
import numpy as np

D65, WHITE, RED, GREEN = 0, 1, 2, 3
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def load_spectrum(path):
    rows = [l.split(",") for l in open(path) if l.strip()]
    lam = [float(r[0]) for r in rows]
    return np.asarray([float(r[1]) for r in rows], np.float32), lam[0], lam[-1]


def load_sensor(path):
    rows = [l.split(",") for l in open(path) if l.strip()]
    lam = [float(r[0]) for r in rows]
    return np.asarray([[float(r[1]), float(r[2]), float(r[3])] for r in rows], np.float32), lam[0], lam[-1]


def load_rgb2spec(path):
    """text table of the reference (spectrum/Rgb2Spec.py:15-36) or its binary image <path>.f32 (tools/make_spec_table_bin.py)"""
    if not os.path.exists(path) and os.path.exists(path + ".f32"):
        raw = np.fromfile(path + ".f32", np.float32)
        res = int(raw[:1].view(np.int32)[0])
        return raw[1:1 + res].copy(), raw[1 + res:].copy(), res
    lines = open(path).read().split("\n")
    res = int(lines[0])
    scale = np.asarray([float(x) for x in lines[1:1 + res]], np.float32)
    data = np.asarray(" ".join(lines[1 + res:]).split(), np.float64).astype(np.float32)
    assert data.size == res ** 3 * 9
    return scale, data, res


def _formula(t, A):
    return (pow(1.0 - t, 5.0) * A[0] + 5.0 * pow(1.0 - t, 4.0) * t * A[1] + 10.0 * pow(1.0 - t, 3.0) * pow(t, 2.0) * A[2]
            + 10.0 * pow(1.0 - t, 2.0) * pow(t, 3.0) * A[3] + 5.0 * (1.0 - t) * pow(t, 4.0) * A[4] + pow(t, 5.0) * A[5])


def sky_tables(sky_dir, turbidity=3.0, albedo=0.5, elevation=0.17):
    """-> configs (11,9) f32, radiances (11) f32, sun_dir (3) f32   (sky/Sky.py:107-159, 93-96)"""
    data = np.asarray([[float(x) for x in l.split(",")[:1080]] for l in open(os.path.join(sky_dir, "data.csv")) if l.strip()], np.float32)
    rad = np.asarray([[float(x) for x in l.split(",")[:120]] for l in open(os.path.join(sky_dir, "data_rad.csv")) if l.strip()], np.float32)
    it = int(turbidity); rem = turbidity - float(it)
    se = pow(elevation / (math.pi / 2.0), 1.0 / 3.0)
    cfg = np.zeros((11, 9), np.float32); rd = np.zeros(11, np.float32)
    terms = [(9 * 6 * (it - 1), (1.0 - albedo) * (1.0 - rem)), (9 * 6 * 10 + 9 * 6 * (it - 1), albedo * (1.0 - rem))]
    if it < 10:
        terms += [(9 * 6 * it, (1.0 - albedo) * rem), (9 * 6 * 10 + 9 * 6 * it, albedo * rem)]
    # NumPy 1.18 (requirements.txt) evaluates `f32_array[j, i] += python_float * formula(f32 scalars ...)` in f64 and
    # rounds to f32 on every store; spelled out here because NumPy 2 would keep f32 scalars in f32.
    for n, (index, wgt) in enumerate(terms):
        for j in range(11):
            for i in range(9):
                v = wgt * _formula(se, [float(data[j, index + i + 9 * k]) for k in range(6)])
                cfg[j, i] = v if n == 0 else float(cfg[j, i]) + v
    terms = [(6 * (it - 1), (1.0 - albedo) * (1.0 - rem)), (60 + 6 * (it - 1), albedo * (1.0 - rem))]
    if it < 10:
        terms += [(6 * it, (1.0 - albedo) * rem), (60 + 6 * it, albedo * rem)]
    for n, (index, wgt) in enumerate(terms):
        for i in range(11):
            v = wgt * _formula(se, [float(rad[i, index + k]) for k in range(6)])
            rd[i] = v if n == 0 else float(rd[i]) + v
    sun = np.asarray([0.0, math.sin(elevation), math.cos(elevation)], np.float32)
    return cfg, rd, sun


def attach(osc, pkg_root, sky=True):
    """load every spectral table into an OracleScene (PT_Spec.setup_data_cpu / setup_data_gpu incl. normalize_spec(d65))"""
    L = osc.lib
    L.orc_spec_sensor.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.c_float]
    L.orc_spec_spectrum.argtypes = [C.c_void_p, C.c_int, _f32p, C.c_int, C.c_float, C.c_float]
    L.orc_spec_rgb2spec.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int]
    L.orc_spec_sky.argtypes = [C.c_void_p, _f32p, _f32p, _f32p]
    L.orc_spec_normalize.argtypes = [C.c_void_p, C.c_int, _f32p]
    L.orc_spec_get_spectrum.argtypes = [C.c_void_p, C.c_int, _f32p]
    L.orc_render_pt_spec.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, _f32p, np.ctypeslib.ndpointer(np.uint64)]
    L.orc_spec_srgb_to_spec.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _f32p]
    L.orc_spec_sky_radiance.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _f32p, _f32p]
    sp = os.path.join(pkg_root, "spectrum")
    xyz, lo, hi = load_sensor(os.path.join(sp, "ciexyz31_1.csv")); L.orc_spec_sensor(osc.h, xyz.reshape(-1), xyz.shape[0], lo, hi)
    for which, name in ((D65, "Illuminantd65.csv"), (WHITE, "white-spec.csv"), (RED, "red-spec.csv"), (GREEN, "green-spec.csv")):
        d, lo, hi = load_spectrum(os.path.join(sp, name)); L.orc_spec_spectrum(osc.h, which, d, d.size, lo, hi)
    scale, data, res = load_rgb2spec(os.path.join(sp, "spec_table")); L.orc_spec_rgb2spec(osc.h, scale, data, res)
    if sky:
        cfg, rd, sun = sky_tables(os.path.join(pkg_root, "sky")); L.orc_spec_sky(osc.h, cfg.reshape(-1), rd, sun)
    wp = np.zeros(3, np.float32); L.orc_spec_normalize(osc.h, D65, wp)
    osc.white_point = wp
    return osc


def render_pt_spec(osc, W, H, frame_begin, n_frames, max_depth=10, seed=0, hdr=None):
    if hdr is None:
        hdr = np.zeros((W, H, 3), np.float32)
    cnt = np.zeros(4, np.uint64)
    osc.lib.orc_render_pt_spec(osc.h, W, H, frame_begin, n_frames, max_depth, seed, hdr.reshape(-1), cnt)
    return hdr, dict(closest=int(cnt[0]), shadow=int(cnt[1]))
