"""include/trmath.h — the transcendental functions shared by the CUDA kernels and the CPU oracle.
CPU: accuracy against float64 libm on the ranges the renderer uses, special values.  GPU: the device build of the same header
returns the same bits as the host build (that is what makes radiance parity exact)."""
import os
import numpy as np
import pytest
from oracle import oracle


def _ulps(got, ref64, floor=0.0):
    """error in units of the last place of max(|ref|, floor)"""
    mag = np.maximum(np.abs(ref64), max(floor, 2.0 ** -126))
    ulp = 2.0 ** (np.floor(np.log2(mag)) - 23)
    return np.abs(got.astype(np.float64) - ref64) / ulp


def _inputs(seed=7, n=400000):
    rng = np.random.RandomState(seed)
    return {
        "angle": (rng.rand(n) * 16.0 - 8.0).astype(np.float32),
        "expo": (rng.rand(n) * 108.0 - 88.0).astype(np.float32),
        "unit": (rng.rand(n) * 2.0 - 1.0).astype(np.float32),
        "y": rng.randn(n).astype(np.float32), "x": rng.randn(n).astype(np.float32),
        "base": rng.rand(n).astype(np.float32),
        "gbase": np.exp(rng.rand(n) * 16.0 - 8.0).astype(np.float32), "gexp": (rng.rand(n) * 16.0 - 8.0).astype(np.float32),
    }


def test_accuracy_against_float64():
    v = _inputs()
    a64 = v["angle"].astype(np.float64)
    # sin / cos: 2 ULP of max(|result|, 1/32) (near a zero crossing the error is absolute, ~1e-8)
    assert _ulps(oracle.math_fn(0, v["angle"]), np.sin(a64), 2.0 ** -5).max() < 2.0
    assert _ulps(oracle.math_fn(1, v["angle"]), np.cos(a64), 2.0 ** -5).max() < 2.0
    assert _ulps(oracle.math_fn(2, v["expo"]), np.exp(v["expo"].astype(np.float64))).max() < 1.5
    assert _ulps(oracle.math_fn(3, v["unit"]), np.arccos(v["unit"].astype(np.float64))).max() < 2.0
    assert _ulps(oracle.math_fn(4, v["y"], v["x"]), np.arctan2(v["y"].astype(np.float64), v["x"].astype(np.float64)), 2.0 ** -5).max() < 4.0
    for e in (2.4, 1.0 / 2.4, 5.0):                                  # srgb <-> linear, Schlick
        ee = np.full_like(v["base"], np.float32(e))
        assert _ulps(oracle.math_fn(5, v["base"], ee), v["base"].astype(np.float64) ** np.float64(np.float32(e))).max() < 2.5
    ref = v["gbase"].astype(np.float64) ** v["gexp"].astype(np.float64)
    ok = (ref > 1e-37) & (ref < 1e38)
    assert _ulps(oracle.math_fn(5, v["gbase"], v["gexp"])[ok], ref[ok]).max() < 4.0


def test_special_values():
    f = np.float32
    m = oracle.math_fn
    assert m(0, [0.0])[0] == 0.0 and m(1, [0.0])[0] == 1.0 and m(2, [0.0])[0] == 1.0
    assert m(2, [-200.0])[0] == 0.0 and np.isinf(m(2, [100.0])[0])
    sub = m(2, [-100.0])[0]; assert 0.0 < sub < 1e-43                                 # gradual underflow
    assert m(3, [1.0])[0] == 0.0 and abs(m(3, [-1.0])[0] - np.pi) < 1e-6
    assert np.isnan(m(3, [np.nextafter(f(1.0), f(2.0))])[0])                          # the reference's acos(1 + ulp) = NaN quirk (sky sample)
    assert abs(m(4, [0.0], [-1.0])[0] - np.pi) < 1e-6 and abs(m(4, [1.0], [0.0])[0] - np.pi / 2) < 1e-6 and m(4, [0.0], [1.0])[0] == 0.0
    assert m(5, [0.0], [2.4])[0] == 0.0 and m(5, [2.0], [10.0])[0] == 1024.0 and m(5, [-2.0], [3.0])[0] == -8.0
    assert np.isnan(m(5, [-2.0], [0.5])[0]) and m(5, [7.0], [0.0])[0] == 1.0 and m(5, [1.0], [123.0])[0] == 1.0


@pytest.mark.gpu
def test_device_build_returns_the_same_bits(gpu_ctx):
    v = _inputs(seed=11)
    specials = np.float32([0.0, -0.0, 1.0, -1.0, 0.5, 1e-30, 1e30, np.pi, np.pi / 2, 88.0, -88.0, -100.0, -104.0, 1e-40, np.inf, -np.inf, np.nan, 1.0000001, 262143.9, 3e5])
    cases = [(0, np.concatenate([v["angle"], specials]), None), (1, np.concatenate([v["angle"], specials]), None),
             (2, np.concatenate([v["expo"], specials]), None), (3, np.concatenate([v["unit"], specials]), None),
             (4, np.concatenate([v["y"], specials, specials[::-1]]), np.concatenate([v["x"], specials[::-1], specials])),
             (5, np.concatenate([v["gbase"], v["base"], specials]), np.concatenate([v["gexp"], np.full_like(v["base"], 2.4), specials[::-1]]))]
    for fn, a, b in cases:
        g = gpu_ctx.test_math(fn, a, b); c = oracle.math_fn(fn, a, b)
        same = (g.view(np.uint32) == c.view(np.uint32)) | (np.isnan(g) & np.isnan(c))
        assert same.all(), (fn, a[~same][:5], None if b is None else b[~same][:5], g[~same][:5], c[~same][:5])


def test_header_is_plain_c(tmp_path):
    """include/trmath.h is shared by nvcc (device), g++ (oracle) and plain C hosts: it compiles as C99 -pedantic -Werror and gives
    the oracle's bits there too"""
    import shutil, subprocess, struct
    from conftest import ROOT
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "trmath.h"\nint main(void) { float v[6]; unsigned u; int k;\n'
                   ' v[0] = tr_sinf(1.25f); v[1] = tr_cosf(-2.5f); v[2] = tr_expf(-3.75f); v[3] = tr_acosf(0.3f); v[4] = tr_atan2f(0.7f, -0.2f); v[5] = tr_powf(0.37f, 2.4f);\n'
                   ' for (k = 0; k < 6; ++k) { memcpy(&u, &v[k], 4); printf("%08x\\n", u); } return 0; }\n')
    exe = str(tmp_path / "t")
    subprocess.check_call([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe, "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    ref = [oracle.math_fn(0, [1.25])[0], oracle.math_fn(1, [-2.5])[0], oracle.math_fn(2, [-3.75])[0], oracle.math_fn(3, [0.3])[0],
           oracle.math_fn(4, [0.7], [-0.2])[0], oracle.math_fn(5, [0.37], [2.4])[0]]
    assert out == ["%08x" % struct.unpack("<I", struct.pack("<f", float(r)))[0] for r in ref]
