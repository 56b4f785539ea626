"""CPU: host-side logic of the product (no device calls) and the C-ABI surface."""
import ctypes
import math
import os
import re
import numpy as np
import pytest
from conftest import ROOT, PKG, make_product_scene
from oracle import oracle


@pytest.mark.parametrize("name", ["cornell", "sphere", "teapot", "teapot_mc", "veach"])
def test_scene_packing_matches_oracle_loader(oracle_tables, name):
    """two independent ingest implementations (product objio+Scene vs oracle objload) agree bit for bit"""
    sl = name not in ("cornell", "veach")
    s = make_product_scene(name, sphere_light=sl)
    s.setup_data_cpu()
    t = oracle_tables(name, sphere_light=sl)
    assert np.array_equal(s.vertex_np, t.vertex)
    assert np.array_equal(s.primitive_np, t.primitive)
    assert np.array_equal(s.material_np, t.material)
    assert np.array_equal(s.minboundarynp, t.bmin) and np.array_equal(s.maxboundarynp, t.bmax)
    assert np.array_equal(np.asarray(s.light_cpu, np.int32), t.light)
    if t.shape.shape[0]:
        assert np.array_equal(s.shape_np, t.shape)
    assert s.primitive_count == t.primitive.shape[0] and s.vertex_count == t.vertex.shape[0]


def test_spot_and_laser_shapes_pack_like_the_reference_rows(oracle_tables):
    """SceneData.Shape.fillStruct for SHPAE_LASER / SHPAE_SPOT (type, pos, radius | xita1 xita2 scale, normal: SceneData.py:88-131)
    through Scene.add_shape == the rows the oracle-side loader writes by hand; both are emitters (light list) and shape primitives"""
    s = make_product_scene("cornell", beam_lights=True); s.setup_data_cpu()
    t = oracle_tables("cornell", beam_lights=True)
    assert np.array_equal(s.shape_np, t.shape) and np.array_equal(s.primitive_np, t.primitive) and np.array_equal(s.material_np, t.material)
    assert s.light_cpu == [34, 35, 36, 37] and np.array_equal(np.asarray(s.light_cpu, np.int32), t.light)
    assert s.shape_np[0, 0] == 4.0 and s.shape_np[0, 4] == 60.0 and s.shape_np[0, 7:10].tolist() == [0.0, -1.0, 0.0]      # laser: radius, normal
    assert s.shape_np[1, 0] == 3.0 and np.allclose(s.shape_np[1, 4:7], [0.3, 0.6, 1.0])                                    # spot: xita1, xita2, scale
    assert s.primitive_np[36].tolist() == [2, 0, 4] and s.primitive_np[37].tolist() == [2, 1, 5]


def test_cornell_material_classes():
    s = make_product_scene("cornell"); s.setup_data_cpu()
    assert s.material_np[:, 0].tolist() == [0.0, 0.0, 0.0, 2.0]          # white, red, green disney; light
    assert s.material_np[3, 2:5].tolist() == [10.0, 10.0, 10.0] and s.light_cpu == [34, 35]
    assert s.bvh.node_count == 71 and s.bvh.primitive_pot == 64 and s.bvh.primitive_bit == 6


def test_camera_matches_oracle():
    import Camera
    cam = Camera.Camera(512, 512, 64)
    cam.scale = 768.5923
    cam.set_target(278.0, 274.4, -279.6)
    view, view_inv, eye, fx, fy, cx, cy = oracle.camera_matrices(512, 512, (278.0, 274.4, -279.6), 768.5923)
    assert np.array_equal(cam.view_np[0], view) and np.array_equal(cam.view_inv_np[0], view_inv)
    assert np.array_equal(cam.eye_np[0], eye) and (cam.fx, cam.fy, cam.cx, cam.cy) == (fx, fy, cx, cy)
    # orbit positions: yaw / pitch (pitch is clamped to +-1.57, Camera.py:71), non-square images
    rng = np.random.RandomState(5)
    for _ in range(20):
        W, H = int(rng.randint(16, 900)), int(rng.randint(16, 900))
        yaw, pitch, scale = float(rng.uniform(-3.2, 3.2)), float(rng.uniform(-2.0, 2.0)), float(rng.uniform(0.5, 2000.0))
        tgt = tuple(float(x) for x in rng.randn(3) * 100.0)
        c2 = Camera.Camera(W, H, 16)
        c2.target[:] = tgt
        c2.set_view_point(yaw, pitch, 0.0, scale)
        view, view_inv, eye, fx, fy, cx, cy = oracle.camera_matrices(W, H, tgt, scale, yaw, pitch)
        assert np.array_equal(c2.view_np[0], view) and np.array_equal(c2.view_inv_np[0], view_inv) and np.array_equal(c2.eye_np[0], eye)
        assert (c2.fx, c2.fy, c2.cx, c2.cy) == (fx, fy, cx, cy)
    cam.update_frame(); assert cam.frame == 1 and cam.frame_cpu[0] == 1
    with pytest.raises(ZeroDivisionError):          # spp < 4 raises in the reference ctor
        Camera.Camera(8, 8, 1)


def test_tile_ownership_partitions_image():
    import parallel
    for W, H, n in [(512, 512, 8), (1024, 1024, 4), (100, 70, 3), (33, 31, 2)]:
        own = parallel.tile_owner(W, H, n)
        masks = [parallel.tile_mask(W, H, r, n) for r in range(n)]
        assert np.array_equal(sum(m.astype(int) for m in masks), np.ones((W, H), int))
        assert own.min() >= 0 and own.max() < n
    own = parallel.tile_owner(512, 512, 8)
    counts = np.bincount(own.reshape(-1), minlength=8)
    assert counts.min() == counts.max()             # 256 tiles / 8 ranks, perfectly balanced


def test_cabi_exports_every_declared_symbol():
    """libtiray.so loads without a GPU and exports each function include/tiray.h declares"""
    import _native
    hdr = open(os.path.join(ROOT, "include", "tiray.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tr_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = ctypes.CDLL(_native.lib_path())
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_native.SIGNATURES)      # the ctypes table binds exactly the header


def test_no_cpu_fallback():
    """without a CUDA device the product must fail loudly (never route through the oracle)"""
    import _native
    lib = _native.load_library()
    if lib.tr_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        _native.Context(0)
    src = ""
    for dp, _, fs in os.walk(PKG):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src += open(os.path.join(dp, f)).read()
    assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src


def test_taichi_shim_imwrite_convention(tmp_path):
    import taichi as ti
    import cv2
    img = np.zeros((4, 3, 3), np.float32)          # [x][y], y up
    img[0, 2] = (1.0, 0.0, 0.0)                    # x=0, top row -> red at image row 0, col 0
    p = str(tmp_path / "o.png")
    ti.imwrite(img, p)
    back = cv2.imread(p)
    assert back.shape == (3, 4, 3) and back[0, 0].tolist() == [0, 0, 255] and back[2, 0].tolist() == [0, 0, 0]


@pytest.mark.skipif(not os.path.isdir("/root/reference/example"), reason="reference tree not present")
def test_reference_example_constructs_unchanged(monkeypatch):
    """the reference's own example/cornell_box.py imports and constructs against this package's
    modules (host path only: everything up to the first device call)"""
    import importlib.util, sys, _native
    monkeypatch.chdir(PKG)
    monkeypatch.setattr(_native, "reset_context", lambda device=None: None)     # ti.init needs a GPU
    sys.modules.pop("Example", None)
    spec_e = importlib.util.spec_from_file_location("Example", "/root/reference/example/Example.py")
    mod_e = importlib.util.module_from_spec(spec_e); sys.modules["Example"] = mod_e; spec_e.loader.exec_module(mod_e)
    spec = importlib.util.spec_from_file_location("ref_cornell_box", "/root/reference/example/cornell_box.py")
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    ex = mod.example(64, 64, 4)
    ex.scene.setup_data_cpu()
    assert ex.scene.primitive_count == 36 and ex.integrator.stack_size == 64
    # ... and so does its example/veach_bdpt.py (BDPT_RGB.BDPT, model/bdpt.obj through the native reader)
    spec = importlib.util.spec_from_file_location("ref_veach_bdpt", "/root/reference/example/veach_bdpt.py")
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    ex = mod.example(64, 64, 4)
    ex.scene.setup_data_cpu()
    assert ex.scene.primitive_count == 11544 and ex.scene.light_count == 4 and type(ex.integrator).__name__ == "BDPT"
    sys.modules.pop("Example", None)


def test_native_obj_reader_numbers_and_errors(tmp_path):
    """csrc/objparse.cpp: decimal strings become the doubles Python's float() gives (fast path and strtod fallback); polygons
    fan like PyWavefront; negative indices; malformed input raises with file:line"""
    import objio
    rng = np.random.RandomState(7)
    vals = ["0", "-0.0", "1", "-1.5", "3.14159265358979", "1e-3", "-2.5E+2", "123456789012345678", "0.1234567890123456789012",
            "1e22", "1e23", "9007199254740993", "4.9e-324", "1.7976931348623157e308", ".5", "5.", "+7.25", "00012.500"]
    vals += ["%.*g" % (rng.randint(1, 18), x) for x in rng.randn(300) * 10.0 ** rng.randint(-12, 12, 300)]
    vals += ["%.6f" % x for x in rng.rand(300) * 1000 - 500]
    while len(vals) % 3:
        vals.append("1")
    p = tmp_path / "n.obj"
    lines = ["v %s %s %s" % tuple(vals[k:k + 3]) for k in range(0, len(vals), 3)]
    n = len(lines)
    lines += ["f %d %d %d" % (k + 1, (k + 1) % n + 1, (k + 2) % n + 1) for k in range(n)]
    p.write_text("\n".join(lines) + "\n")
    m = objio.read_obj(str(p))
    assert len(m) == 1 and m[0].name == "default0" and m[0].rows.shape == (3 * n, 9)
    got = m[0].rows[0::3, 0:3].reshape(-1)
    want = np.array([float(v) for v in vals], np.float64)
    assert np.array_equal(got, want) and np.array_equal(np.signbit(got), np.signbit(want))
    # a quad and a pentagon with negative indices, vt / vn present, CRLF line ends
    q = tmp_path / "q.obj"
    q.write_bytes(b"v 0 0 0\r\nv 1 0 0\r\nv 1 1 0\r\nv 0 1 0\r\nv 0.5 2 0\r\nvt 0.25 0.75\r\nvn 0 0 1\r\n"
                  b"usemtl a\r\nf 1/1/1 2/1/1 3/1/1 4/1/1\r\nusemtl b\r\nf -5//-1 -4//-1 -3//-1 -2//-1 -1//-1\r\n")
    a, b = objio.read_obj(str(q))
    assert (a.name, a.has_vt, a.has_vn, b.name, b.has_vt, b.has_vn) == ("a", True, True, "b", False, True)
    assert a.rows[:, 0:2].tolist() == [[0, 0], [1, 0], [1, 1], [0, 1], [0, 0], [1, 1]]           # (v1 v2 v3) (v4 v1 v3)
    assert np.all(a.rows[:, 6:8] == [0.25, 0.75]) and np.all(a.rows[:, 3:6] == [0, 0, 1])
    assert b.rows[:, 0:2].tolist() == [[0, 0], [1, 0], [1, 1], [0, 1], [0, 0], [1, 1], [0.5, 2], [0, 0], [0, 1]]
    for body, msg in (("v 0 0 0\nf 1 2 3\n", "out of range"), ("v 0 0 x\n", "bad number"), ("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1/1 2/1 3/1\n", "out of range"),
                      ("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1 2 3\nf 1//1 2//1 3//1\n", "mixes vertex formats"), ("mtllib missing.mtl\n", "cannot open")):
        bad = tmp_path / "bad.obj"; bad.write_text(body)
        with pytest.raises((ValueError, FileNotFoundError), match=msg):
            objio.read_obj(str(bad))


def test_header_is_plain_c_and_usable_from_c(tmp_path):
    """include/tiray.h compiles as C99 with -Wall -Werror, and a C program linked against libtiray.so drives the OBJ reader
    and gets TR_ERR_NO_DEVICE (exit code 3) instead of a silent fallback when no GPU is present (0 on a GPU box)"""
    import shutil, subprocess, _native
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "client")
    libdir = os.path.dirname(_native.lib_path())
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cabi", "client.c"), "-o", exe, "-L", libdir, "-l:libtiray.so", "-Wl,-rpath," + libdir])
    r = subprocess.run([exe, os.path.join(PKG, "model", "cornell_box.obj")], capture_output=True, text=True)
    assert r.returncode in (0, 3), r.stderr
    out = r.stdout.splitlines()
    assert out[0].startswith("material 0 white tris 30 ") and out[3].startswith("material 3 light tris 2 ")
    assert any(l.startswith("triangles 36 devices") for l in out)
    assert ("context ok" in r.stdout) == (r.returncode == 0)
    if r.returncode == 3:
        assert "no device" in r.stdout


def test_sass_shows_tma_staging_and_sm100a():
    """the shipped binary is sm_100a code and the small-tree kernels stage the BVH with bulk async copies on an mbarrier
    (cp.async.bulk = UBLKCP in SASS, B200_PROFILING.md) -- k_trace, k_shadow, k_tail and the lock-step BDPT kernels"""
    import shutil, subprocess, _native
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    lib = _native.lib_path()
    elf = subprocess.run([cuobjdump, "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and "sm_90" not in elf
    def sass_of(fn):
        return subprocess.run([cuobjdump, "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout
    # tree modes (csrc/trace.cuh): 0 replicated shared-memory image (one bulk copy), 1 plain shared-memory tree (nodes + leaves: two),
    # 3 global memory + staged top nodes (one), 2 global memory only (none); all staged modes wait on the mbarrier
    for mode, copies in ((0, 1), (1, 2), (3, 1)):
        for fn in ("_Z7k_traceILi%dEEv6WfArgsi" % mode, "_Z8k_shadowILi%dELb0EEv6WfArgsi" % mode, "_Z6k_tailILi%dELb0EEv6WfArgsi" % mode):
            sass = sass_of(fn)
            assert sass.count("UBLKCP") == copies, (fn, sass.count("UBLKCP"))
            assert "SYNCS.ARRIVE.TRANS64" in sass and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in sass, fn
    sass0 = sass_of("_Z7k_traceILi2EEv6WfArgsi")
    assert "UBLKCP" not in sass0 and "LDG" in sass0            # large trees are walked in global memory
    assert "LDS" in sass0 and "STS" in sass0                   # ... with the traversal stack in shared memory


def test_random_obj_files_pack_like_the_oracle_loader(tmp_path, monkeypatch):
    """12 random OBJ + MTL files (polygons of 3-6 corners, v / v/vt / v//vn / v/vt/vn per material, negative indices,
    faces before any usemtl, usemtl of names the MTL does not define, emissive and transparent materials): the product
    (native reader + Scene packing) and the oracle's Python loader produce the same tables bit for bit"""
    import Scene
    from oracle import objload
    rng = np.random.RandomState(11)
    for case in range(12):
        nv, nn, nt = rng.randint(8, 60), rng.randint(1, 20), rng.randint(1, 20)
        lines = ["mtllib m%d.mtl" % case]
        lines += ["v %.6g %.6g %.6g" % tuple(rng.randn(3) * 10.0 ** rng.randint(-2, 3)) for _ in range(nv)]
        lines += ["vn %.5f %.5f %.5f" % tuple(rng.randn(3)) for _ in range(nn)]
        lines += ["vt %.4f %.4f" % tuple(rng.rand(2)) for _ in range(nt)]
        mats = ["a", "b", "c", "ghost"]                       # "ghost" is not in the MTL
        fmt_of = {}
        order = [None] + [mats[k] for k in rng.randint(0, 4, 6)]   # first group: faces before any usemtl -> "default<k>"
        for m in order:
            if m is not None:
                lines.append("usemtl " + m)
            fmt = fmt_of.setdefault(m, rng.randint(0, 4))
            for _ in range(rng.randint(1, 5)):
                corners = []
                for _c in range(rng.randint(3, 7)):
                    neg = rng.rand() < 0.3
                    a = -rng.randint(1, nv + 1) if neg else rng.randint(1, nv + 1)
                    b = -rng.randint(1, nt + 1) if neg else rng.randint(1, nt + 1)
                    c = -rng.randint(1, nn + 1) if neg else rng.randint(1, nn + 1)
                    corners.append(("%d" % a, "%d/%d" % (a, b), "%d//%d" % (a, c), "%d/%d/%d" % (a, b, c))[fmt])
                lines.append("f " + " ".join(corners))
        (tmp_path / ("r%d.obj" % case)).write_text("\n".join(lines) + "\n")
        (tmp_path / ("m%d.mtl" % case)).write_text(
            "newmtl a\nKd 0.1 0.2 0.3\nd 1.0\nNs 10\nNi 1.2\n# comment\nnewmtl b\nKd 0.9 0.9 0.9\nKe 20 30 40\n\nnewmtl c\nKd 1 1 1\nTr 0.6\nNi 1.45\nNs 3.5\n")
        path = str(tmp_path / ("r%d.obj" % case))
        s = Scene.Scene(); s.add_obj(path); s.setup_data_cpu()
        t = objload.load_scene([path])
        assert np.array_equal(s.vertex_np, t.vertex), case
        assert np.array_equal(s.primitive_np, t.primitive) and np.array_equal(s.material_np, t.material), case
        assert np.array_equal(np.asarray(s.light_cpu, np.int32), t.light), case
        assert np.array_equal(s.minboundarynp, t.bmin) and np.array_equal(s.maxboundarynp, t.bmax), case


def test_texture_loader_matches_oracle(tmp_path, monkeypatch):
    """Texture.load_image (texture/Texture.py:18-34): packed RGB i32, buf[x][H-1-row] -- product vs oracle, on the shipped
    environment map and on a random non-square image"""
    import cv2
    import Scene  # noqa: F401  (puts texture/ on sys.path like the reference's Scene.py:3-4)
    import Texture as TX
    monkeypatch.chdir(PKG)
    rng = np.random.RandomState(2)
    img = rng.randint(0, 256, (37, 91, 3)).astype(np.uint8)
    p = str(tmp_path / "t.png"); cv2.imwrite(p, img)
    for path in ("image/env.png", "image/black.png", p):
        t = TX.Texture(); t.load_image(path)
        packed, w, h = oracle.load_env(path if os.path.isabs(path) else os.path.join(PKG, path))
        assert (t.wid, t.hgt) == (w, h)
        assert np.array_equal(np.asarray(t.np_img, np.int32).reshape(w, h), packed)
    # layout: x major, y up; channels packed as R << 16 | G << 8 | B (cv2 reads BGR)
    t = TX.Texture(); t.load_image(p)
    b, g, r = (int(v) for v in img[36 - 5, 7])
    assert int(np.asarray(t.np_img).reshape(91, 37)[7, 5]) == (r << 16) | (g << 8) | b
