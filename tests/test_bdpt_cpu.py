"""CPU tests of the BDPT oracle (oracle/bdpt_core.inc): pinned against the reference's own render of example/veach_bdpt.py
(image/veach-bdpt512.png -> tests/golden), plus internal consistency of the per-pixel dump used by the GPU parity tests."""
import os
import numpy as np
import pytest
from conftest import GOLDEN
from oracle import oracle


def veach_oracle(oracle_tables, W, H, fast):
    import copy
    t = copy.copy(oracle_tables("veach"))
    t.vertex = oracle.OracleScene(oracle_tables("veach"), fast=True).build().process_normal()      # example/veach_bdpt.py:23
    s = oracle.OracleScene(t, fast=fast).build()
    cam = oracle.fit_camera(t, W, H, 0.5)                                                           # :27-29
    s.set_camera(cam[1], cam[2], *cam[3:]); s.set_camera_view(cam[0], W, H)
    return s


def test_veach_scene_tables(oracle_tables):
    """model/bdpt.obj through the PyWavefront restatement: 11 544 triangles, 6 materials in MTL order
    (Wall, Wood, Lamp: Disney; Light1, Light2: emitters; Glass: ior 1.5, extinction Ns = 100), 4 emitter triangles"""
    t = oracle_tables("veach")
    assert t.primitive.shape[0] == 11544 and t.material.shape[0] == 6 and t.light.size == 4
    assert list(t.material[:, 0]) == [0.0, 0.0, 0.0, 2.0, 2.0, 1.0]
    assert t.material[5, 5] == np.float32(1.5) and t.material[5, 6] == np.float32(100.0)
    assert np.allclose(t.material[3, 2:5], [12048.179, 8605.842, 6196.206])


def test_oracle_bdpt_vs_reference_image(oracle_tables):
    """128 x 128, 24 spp of the oracle against the reference's 512 x 512 render, area-downscaled: channel means within
    4 %, PSNR of the blurred images > 24 dB (at 256 x 256 x 24 spp: 2 % and 31.9 dB)"""
    import cv2
    W = H = 128
    s = veach_oracle(oracle_tables, W, H, fast=True)
    hdr, cnt = s.render_bdpt_rgb(W, H, 0, 24)
    assert np.isfinite(hdr).all()
    img = (np.clip(oracle.tonemap(hdr, 0.5).swapaxes(0, 1)[::-1], 0, 1) * 255).astype(np.uint8)
    ref = cv2.resize(cv2.imread(os.path.join(GOLDEN, "veach-bdpt512.png"))[:, :, ::-1], (W, H), interpolation=cv2.INTER_AREA)
    m, r = img.reshape(-1, 3).mean(0), ref.reshape(-1, 3).mean(0)
    assert np.all(np.abs(m - r) < 0.04 * r), (m, r)
    mse = ((cv2.GaussianBlur(img, (5, 5), 0).astype(np.float64) - cv2.GaussianBlur(ref, (5, 5), 0)) ** 2).mean()
    assert 10 * np.log10(255.0 ** 2 / mse) > 24.0
    # every sample traces at least its two first segments; connections need one shadow query each (<= 26 per sample)
    assert 2 * W * H * 24 <= cnt["closest"] <= 11 * W * H * 24 and 0 < cnt["shadow"] <= 26 * W * H * 24


def test_oracle_bdpt_dump_reconstructs_frame(oracle_tables):
    """film of one frame == per pixel: sum of its (e >= 2, l) strategies in loop order + the e == 1 splats aimed at it"""
    W = H = 24
    s = veach_oracle(oracle_tables, W, H, fast=False)
    for frame in (0, 3):
        hdr, _ = s.render_bdpt_rgb(W, H, frame, 1)
        hdr = hdr * (frame + 1.0)                             # undo the running-mean weight of a film that started at 0
        own = np.zeros((W, H, 3), np.float32); spl = np.zeros((W, H, 3), np.float64)
        for i in range(W):
            for j in range(H):
                v, (ed, ld), c = s.bdpt_pixel_dump(i, j, frame)
                assert 1 <= ed <= 7 and 1 <= ld <= 6
                acc = np.zeros(3, np.float32)
                for e in range(2, ed + 1):
                    for l in range(0, ld + 1):
                        if l + e - 2 <= 5:
                            acc = acc + c[e - 1, l, :3]
                own[i, j] = acc
                for l in range(2, ld + 1):
                    if l - 1 <= 5 and c[0, l, 3] >= 0:
                        pix = int(c[0, l, 3]); spl[pix >> 16, pix & 65535] += c[0, l, :3]
        assert np.allclose(own + spl, hdr, rtol=1e-4, atol=1e-6)


def test_oracle_bdpt_mask_shards_sum(oracle_tables):
    """masked renders (tile sharding: only the masked pixels trace paths, everybody receives splats) sum to the full film"""
    W = H = 32
    s = veach_oracle(oracle_tables, W, H, fast=False)
    full, _ = s.render_bdpt_rgb(W, H, 0, 2)
    m = np.zeros((W, H), np.uint8); m[: W // 2] = 1
    a, _ = s.render_bdpt_rgb(W, H, 0, 2, mask=m)
    b, _ = s.render_bdpt_rgb(W, H, 0, 2, mask=1 - m)
    assert np.allclose(a + b, full, rtol=1e-4, atol=1e-6)


def test_cpp_oracle_matches_literal_python_transliteration(oracle_tables):
    """connect_path + mis_weight restated a second time (oracle/bdpt_literal.py: dense Vertex arrays, field-by-field copy(),
    the in-place save / overwrite / walk / restore of mis_weight, -1 indices wrapping into the padding slot) reproduce the C++
    oracle's weighted contribution of every strategy and the splat pixel of the light-tracing ones: two independent
    readings of BDPT_RGB.py:258-580 agree, on the Veach room and on the Cornell box"""
    from oracle import bdpt_literal as BL, objload
    from conftest import model
    import ctypes as C
    for name, fit, smooth, W in (("veach", 0.5, True, 40), ("cornell", 0.8, False, 32)):
        H = W
        if name == "veach":
            s = veach_oracle(oracle_tables, W, H, fast=False); tables = s.t
        else:
            tables = oracle_tables("cornell"); s = oracle.OracleScene(tables).build()
        cam = oracle.fit_camera(tables, W, H, fit)
        s.set_camera(cam[1], cam[2], *cam[3:]); s.set_camera_view(cam[0], W, H)
        rng_state = np.random.RandomState(4)
        n_strat = n_weighted = 0
        for frame in (0, 2):
            for _ in range(60):
                i, j = int(rng_state.randint(0, W)), int(rng_state.randint(0, H))
                verts, depths, contrib = s.bdpt_pixel_dump(i, j, frame)

                def rng(block, i=i, j=j, frame=frame):
                    out = np.zeros(4, np.float32); s.lib.orc_rng(0, i * 65536 + j, frame, block, out); return out
                px = BL.Pixel(s, tables, cam, W, H, verts, depths, i, j, frame)
                for (e, l), (rgb, (u, v)) in px.all_strategies(rng).items():
                    ref = contrib[e - 1, l]
                    assert np.allclose(rgb, ref[:3], rtol=2e-5, atol=1e-12), (name, i, j, frame, e, l, rgb, ref)
                    if e == 1:
                        assert (u * 65536 + v if u >= 0 else -1) == int(ref[3]), (name, i, j, frame, l)
                    n_strat += 1; n_weighted += bool(rgb[0] > 0 and rgb[1] > 0 and rgb[2] > 0 and l + e != 2)
        assert n_strat > 500 and n_weighted > 50, (n_strat, n_weighted)


def test_sample_light_spot_and_laser_against_a_literal_restatement(oracle_tables):
    """Scene.sample_light (Scene.py:430-474) restated here in numpy f32 for all four emitter kinds of the beam-light Cornell box
    (two triangles of the area light, a laser, a spot light): vertex 0 of the C++ oracle's light sub-path -- position, normal,
    beta = emission / choice pdf, first direction, pdf -- agrees for every sampled pixel (numpy's sin / cos / tan differ from
    include/trmath.h by ULPs: rel 2e-5)"""
    from oracle import pt_literal as PL
    from conftest import BEAM_LIGHTS
    f = np.float32
    W = H = 48
    t = oracle_tables("cornell", beam_lights=True)
    s = oracle.OracleScene(t).build()
    cam = oracle.fit_camera(t, W, H, 0.8)
    s.set_camera(cam[1], cam[2], *cam[3:]); s.set_camera_view(cam[0], W, H)
    tr = PL.Tracer(s, t, cam)
    nl = int(t.light.size); assert nl == 4
    PI = PL.PI_REF if hasattr(PL, "PI_REF") else f(3.1415956)

    def map_to_disk(u1, u2):                                 # UtilsFunc.py:322-345
        a, b = f(2.0) * u1 - f(1.0), f(2.0) * u2 - f(1.0)
        if a > -b:
            if a > b: return a, (PI / f(4.0)) * (b / a)
            return b, (PI / f(4.0)) * (f(2.0) - a / b)
        if a < b: return -a, (PI / f(4.0)) * (f(4.0) + b / a)
        return -b, (f(0.0) if b == 0.0 else (PI / f(4.0)) * (f(6.0) - a / b))

    kinds = {1: 0, 3: 0, 4: 0}
    rng = np.random.RandomState(11)
    for i, j in zip(rng.randint(0, W, 160), rng.randint(0, H, 160)):
        frame = 3
        ov, od, _ = s.bdpt_pixel_dump(int(i), int(j), frame)
        Ra, Rb = tr.rng(int(i), int(j), frame, 40), tr.rng(int(i), int(j), frame, 41)
        index = min(int(Ra[0] * f(nl)), nl - 1); pi = int(t.light[index])
        pos, nor = tr.prim_random_point_normal(pi, Ra[1], Ra[2])
        emission = t.material[int(t.primitive[pi, 2]), 2:5].astype(f)
        choice = f(1.0) / (f(nl) * tr.prim_area(pi))
        nor = PL.normalized(nor)
        p = PL.cosine_sample_hemisphere(Rb[0], Rb[1])
        dir_pdf = max(f(0.01), p[2] / PI)
        d = PL.inverse_transform(p, nor)
        kind = 1
        if int(t.primitive[pi, 0]) != 1:
            sh = t.shape[int(t.primitive[pi, 1])].astype(f); kind = int(sh[0])
            if kind == 3:                                    # spot
                scale = sh[6]; dir_pdf = f(1.0)
                r, phi = map_to_disk(Rb[2], Rb[3])
                r1, r2 = scale * np.tan(sh[4], dtype=f), scale * np.tan(sh[5], dtype=f)
                r = r * r2
                if r > r1: emission = emission * (f(1.0) - (r - r1) / (r2 - r1))
                sp = np.array([r * np.cos(phi, dtype=f), r * np.sin(phi, dtype=f), np.sqrt(max(f(0.0), scale * scale - r * r))], f)
                d = PL.inverse_transform(sp, nor)
            elif kind == 4:                                  # laser
                choice = f(1.0) / f(nl)
                r = sh[4]; phi = Rb[2] * PI * f(2.0)
                pos = pos + PL.inverse_transform(np.array([r * np.cos(phi, dtype=f), r * np.sin(phi, dtype=f), 0.0], f), nor)
                d = nor; dir_pdf = f(1.0)
        kinds[kind] += 1
        v0 = ov[7]
        tol = dict(rtol=2e-5, atol=2e-4)
        assert np.allclose(v0[0:3], pos, **tol) and np.allclose(v0[3:6], nor, **tol), (i, j, kind)
        assert np.allclose(v0[9:12], emission / choice, rtol=2e-4) and np.allclose(v0[12:15], d, **tol), (i, j, kind)
        assert np.isclose(v0[15], choice, rtol=1e-5), (i, j, kind)
    assert min(kinds.values()) > 15, kinds                   # every emitter kind was sampled
