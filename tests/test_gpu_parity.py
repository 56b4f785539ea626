"""GPU parity tests proper: the CUDA path (through the C-ABI / the reference-named Python classes)
against the CPU oracle on the same inputs.  BIT-EXACT throughout: integer/index work, IEEE f32
arithmetic (+,-,*,/,sqrt: kernels are built with -fmad=false) and the transcendental functions, which
both sides take from include/trmath.h (one plain-f32 implementation compiled into the kernels and
into the oracle).  Films, vertices and ray counts are compared with array_equal, no tolerances."""
import os
import numpy as np
import pytest
from conftest import GOLDEN, make_product_scene
from oracle import oracle

pytestmark = pytest.mark.gpu


def build_gpu_scene(name, W, H, integrator="pt", **kw):
    """product path exactly as example/Example.py drives it"""
    import Camera, PT_RGB, Debug, math
    scene = make_product_scene(name, **kw)
    cam = Camera.Camera(W, H, 64)
    integ = (PT_RGB.PathTrace if integrator == "pt" else Debug.Debug)(W, H, cam, scene, 64)
    scene.setup_data_cpu(); integ.setup_data_cpu(); integ.setup_data_gpu(); scene.setup_data_gpu()
    lo, hi = scene.minboundarynp[0], scene.maxboundarynp[0]
    size = hi - lo
    cam.scale = math.sqrt(size[0] * size[0] + size[1] * size[1] + size[2] * size[2]) * 0.8
    c = hi + lo
    cam.set_target(c[0] * 0.5, c[1] * 0.5, c[2] * 0.5)
    cam.update()
    return scene, cam, integ


def build_oracle_scene(tables, W, H, env_power=0.0, fast=False):
    s = oracle.OracleScene(tables, fast=fast).build()
    cam = oracle.fit_camera(tables, W, H)
    s.set_camera(cam[1], cam[2], *cam[3:])
    from conftest import PKG
    packed, w, h = oracle.load_env(os.path.join(PKG, "image", "env.png" if env_power else "black.png"))
    s.set_env(packed, w, h, env_power)
    return s


# ---------------------------------------------------------------------------------- LBVH build
@pytest.mark.parametrize("name,sl", [("cornell", False), ("sphere", True), ("teapot", False), ("teapot_mc", True)])
def test_bvh_bit_exact(gpu_ctx, oracle_tables, name, sl):
    """Morton codes, sorted (code, prim) order, bvh_node (children/parents/prims/boxes) and the
    pre-order compact_node array are identical to the oracle's, word for word"""
    scene = make_product_scene(name, sphere_light=sl)
    scene.setup_data_cpu(); scene.setup_data_gpu()
    o = oracle.OracleScene(oracle_tables(name, sphere_light=sl))
    assert np.array_equal(gpu_ctx.morton_download(), o.morton_unsorted())
    o.build()
    assert np.array_equal(scene.bvh.morton_code_s.to_numpy(), o.morton)
    assert np.array_equal(scene.bvh.bvh_node.to_numpy(), o.bvh_node)
    assert np.array_equal(scene.bvh.compact_node.to_numpy(), o.compact)


def test_bvh_matches_reference_nodelist(gpu_ctx):
    """GPU-built Cornell tree printed like accel/LBvh.py:127-136 == the reference's nodelist.txt"""
    scene = make_product_scene("cornell")
    scene.setup_data_cpu(); scene.setup_data_gpu()
    gold = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "nodelist.txt"))]
    assert scene.bvh.nodelist_lines() == gold


def test_bvh_edge_cases(gpu_ctx):
    """1, 2 and 3 primitives, many identical Morton codes, sort chunk boundaries"""
    rng = np.random.RandomState(3)
    for n in (1, 2, 3, 5, 2047, 2048, 2049, 5000):
        base = rng.rand(n, 3).astype(np.float32)
        if n >= 2047:
            base[: n // 2] = base[0]                      # long duplicate run
        tri = np.zeros((n * 3, 9), np.float32)
        tri[0::3, 0:3] = base; tri[1::3, 0:3] = base + np.float32([0.01, 0, 0]); tri[2::3, 0:3] = base + np.float32([0, 0.01, 0])
        tri[:, 5] = 1.0
        prim = np.zeros((n, 3), np.int32); prim[:, 0] = 1; prim[:, 1] = 3 * np.arange(n)
        mat = np.zeros((1, 10), np.float32); mat[0, 2:5] = 0.5; mat[0, 6] = 0.5

        class T: pass
        t = T(); t.vertex = tri; t.primitive = prim; t.material = mat; t.shape = np.zeros((0, 10), np.float32)
        t.light = np.zeros(0, np.int32); t.bmin = tri[:, 0:3].min(0)[None]; t.bmax = tri[:, 0:3].max(0)[None]
        gpu_ctx.scene_upload(t.vertex, t.primitive, t.material, None, None, t.bmin, t.bmax)
        gpu_ctx.bvh_build()
        m, bn, cn = gpu_ctx.bvh_download()
        o = oracle.OracleScene(t).build()
        assert np.array_equal(m, o.morton), n
        assert np.array_equal(bn, o.bvh_node), n
        assert np.array_equal(cn, o.compact), n


# ---------------------------------------------------------------------------------- primary rays + traversal
@pytest.mark.parametrize("name,res,sl", [("cornell", 256, False), ("cornell", 512, False), ("sphere", 256, True), ("teapot_mc", 256, True)])
def test_first_hit_bit_exact(gpu_ctx, oracle_tables, name, res, sl):
    """frame-0 primary rays, closest hit (t, prim, u, v) and interpolated attributes: bit-exact.
    The GPU walk is stackless, pruned and left-first; the oracle is the reference's unpruned stack DFS."""
    scene, cam, integ = build_gpu_scene(name, res, res, "debug", sphere_light=sl)
    integ.render()
    g = integ.first_hit()
    o = build_oracle_scene(oracle_tables(name, sphere_light=sl), res, res)
    assert np.array_equal(g["dir"], o.primary_rays(res, res))
    f = o.first_hit(res, res)
    assert np.array_equal(g["prim"], f["prim"])
    assert np.array_equal(g["t"], f["t"])
    hit = f["prim"] >= 0
    assert np.array_equal(g["uv"][hit], f["uv"][hit])
    for k in ("pos", "gnormal", "normal"):
        assert np.array_equal(g[k][hit], f[k][hit]), k
    assert np.array_equal(integ.hdr.to_numpy(), o.render_debug(res, res))
    if name == "cornell" and res == 256:
        assert int(hit.sum()) == 57867 and g["prim"][128, 128] == 29


def test_trace_random_rays_bit_exact(gpu_ctx, oracle_tables):
    """incoherent rays from inside the scene: closest hit == oracle, shadow query self-consistent"""
    scene, cam, integ = build_gpu_scene("teapot_mc", 64, 64, "debug", sphere_light=True)
    t = oracle_tables("teapot_mc", sphere_light=True)
    o = oracle.OracleScene(t).build()
    rng = np.random.RandomState(11)
    n = 200000
    lo, hi = t.bmin[0], t.bmax[0]
    org = (lo + (hi - lo) * rng.rand(n, 3)).astype(np.float32)
    d = rng.randn(n, 3); d /= np.linalg.norm(d, axis=1, keepdims=True); d = d.astype(np.float32)
    d[:100, 0] = 0.0; d[100:200, 1] = 1e-7                       # parallel-axis branch of slabs
    gt, gp, guv = gpu_ctx.test_trace(org, d)
    ot, op, ouv = o.trace(org, d)
    assert np.array_equal(gp, op) and np.array_equal(gt, ot)
    assert np.array_equal(guv[op >= 0], ouv[op >= 0])
    st, sp, _ = gpu_ctx.test_trace(org, d, shadow=True)          # shadow walk must agree that the closest prim is visible
    assert np.array_equal(sp, op)


# ---------------------------------------------------------------------------------- shading functions
def _unit(a):
    return (a / np.linalg.norm(a, axis=1, keepdims=True)).astype(np.float32)


def test_brdf_hooks_match_oracle(gpu_ctx):
    rng = np.random.RandomState(5)
    n = 50000
    N, V, L = _unit(rng.randn(n, 3)), _unit(rng.randn(n, 3)), _unit(rng.randn(n, 3))
    u = rng.rand(n, 3).astype(np.float32)
    lib = oracle.lib()
    for metal, rough in [(0.0, 0.5), (1.0, 0.0), (0.3, 0.9)]:
        g = gpu_ctx.test_disney_evaluate_pdf(N, V, L, metal, rough)
        o = np.zeros((n, 2), np.float32); lib.orc_disney_evaluate_pdf(n, N.reshape(-1), V.reshape(-1), L.reshape(-1), metal, rough, o.reshape(-1))
        assert np.array_equal(g, o)                            # only +,-,*,/,sqrt: bit-exact
        gs = gpu_ctx.test_disney_sample(V, N, metal, rough, u)
        os_ = np.zeros((n, 3), np.float32); lib.orc_disney_sample(n, V.reshape(-1), N.reshape(-1), metal, rough, u.reshape(-1), os_.reshape(-1))
        assert np.array_equal(gs, os_)                         # sin / cos from the shared include/trmath.h: bit-exact too
    gg = gpu_ctx.test_glass_sample(V, N, 1.3, u[:, 0])
    og = np.zeros((n, 4), np.float32); lib.orc_glass_sample(n, V.reshape(-1), N.reshape(-1), 1.3, np.ascontiguousarray(u[:, 0]), og.reshape(-1))
    assert np.array_equal(gg, og)                              # pow in Schlick's term from the shared header: same Fresnel coins
    p = (rng.randn(n, 3) * np.array([1e-3, 1.0, 500.0])).astype(np.float32)
    go = gpu_ctx.test_offset_ray(p, N)
    oo = np.zeros((n, 3), np.float32); lib.orc_offset_ray(n, p.reshape(-1), N.reshape(-1), oo.reshape(-1))
    assert np.array_equal(go, oo)
    for args in [(0, 0, 0, 0), (12345, 7 << 16 | 9, 3, 5), (2 ** 40 + 17, 0xFFFFFFFF, 63, 30)]:
        o4 = np.zeros(4, np.float32); lib.orc_rng(*args, o4)
        assert np.array_equal(gpu_ctx.test_rng(*args), o4)


# ---------------------------------------------------------------------------------- path tracing
def _compare_radiance(g, o, rel=1e-3, outlier_budget=1e-3, floor=1.0):
    """per-pixel L-inf <= rel * max(1, value); pixels whose path took another branch at a float
    boundary (libm ulp differences) are outliers, bounded by the budget"""
    err = np.abs(g - o).max(axis=2)
    tol = rel * np.maximum(floor, np.abs(o).max(axis=2))
    bad = err > tol
    return bad.mean(), float(np.abs(g.mean((0, 1)) - o.mean((0, 1))).max() / max(1e-9, o.mean()))


def test_pt_rgb_cornell_matches_oracle(gpu_ctx, oracle_tables):
    """C1: Cornell 256^2, frames 0..3, shared counter-based RNG: same image, same ray counts"""
    W = H = 256
    scene, cam, integ = build_gpu_scene("cornell", W, H)
    for _ in range(4):
        integ.render(); cam.update_frame()
    g = integ.hdr.to_numpy()
    st = gpu_ctx.stats()
    o = build_oracle_scene(oracle_tables("cornell"), W, H)
    ref, cnt = o.render_pt_rgb(W, H, 0, 4)
    assert np.array_equal(g, ref)                              # every pixel, every bit
    # frame 3 alone: identical ray counts
    _, c3 = o.render_pt_rgb(W, H, 3, 1)
    assert st["rays_closest"] == c3["closest"] and st["rays_shadow"] == c3["shadow"]


def test_pt_rgb_batched_equals_framewise(gpu_ctx):
    """render_frames(n) in multi-frame wavefront batches == n x render(): bit-identical film"""
    W = H = 128
    scene, cam, integ = build_gpu_scene("cornell", W, H)
    for _ in range(6):
        integ.render(); cam.update_frame()
    a = integ.hdr.to_numpy()
    scene, cam, integ = build_gpu_scene("cornell", W, H)
    integ.render_frames(6)
    b = integ.hdr.to_numpy()
    assert np.array_equal(a, b)
    import _native
    ctx = _native.context(); ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
    ctx.set_option("batch_frames", 4); ctx.set_option("graph", 0)
    integ.render_frames(6)
    assert np.array_equal(integ.hdr.to_numpy(), a)
    ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
    ctx.set_option("smem_bvh", 0); ctx.set_option("graph", 1)
    integ.render_frames(6)
    assert np.array_equal(integ.hdr.to_numpy(), a)            # BVH in shared memory vs global: same result


def test_async_render_deferred_stats_and_ring_overflow(gpu_ctx):
    """tr_render_pt_rgb only enqueues; tr_stats_get folds the per-batch counter snapshots afterwards.  70 one-frame batches
    overflow the 64-entry pinned ring (the rest takes the synchronous path): ray counts and film must equal the 70 frames
    rendered in one batch, and a second render issued before any stats call must not disturb the first one's film"""
    W = H = 64
    scene, cam, integ = build_gpu_scene("cornell", W, H)
    st_one = integ.render_frames(70)
    a = integ.hdr.to_numpy()
    gpu_ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
    gpu_ctx.set_option("batch_frames", 1)
    assert integ.render_frames(70, stats=False) is None
    b = integ.hdr.to_numpy()                       # the download waits for the render
    st_many = gpu_ctx.stats()
    assert np.array_equal(a, b)
    assert (st_many["rays_closest"], st_many["rays_shadow"], st_many["frames"]) == (st_one["rays_closest"], st_one["rays_shadow"], 70)
    assert st_many["ms_total"] > 0.0
    gpu_ctx.set_option("batch_frames", 0)
    gpu_ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
    integ.render_frames(35, stats=False); integ.render_frames(35, stats=False)      # back to back, no host wait in between
    assert np.array_equal(integ.hdr.to_numpy(), a)
    st_last = gpu_ctx.stats()
    assert st_last["frames"] == 35 and 0 < st_last["rays_closest"] < st_one["rays_closest"]


def test_pt_rgb_glass_env_sphere_light_matches_oracle(gpu_ctx, oracle_tables):
    """single_model as shipped: glass sphere.obj, sphere light, env map power 5, smooth normals"""
    W = H = 128
    scene, cam, integ = build_gpu_scene("sphere", W, H, sphere_light=True, glass0=True, env_power=5.0)
    scene.process_normal()
    integ.render_frames(4)
    g = integ.hdr.to_numpy()
    t = oracle_tables("sphere", sphere_light=True, glass0=True)
    o = build_oracle_scene(t, W, H, env_power=5.0)
    vn = o.process_normal()
    gv = scene.vertex.to_numpy()
    assert np.array_equal(gv, vn, equal_nan=True)                          # acos in the angle weights: shared header
    ref, _ = o.render_pt_rgb(W, H, 0, 4)
    assert np.array_equal(g, ref)


def test_pt_rgb_spot_and_laser_emitters_match_oracle(gpu_ctx, oracle_tables):
    """Scene.sample_li's SPOT falloff and LASER radius cut-off (Scene.py:493-516): Cornell box with one of each next to the
    area light; the shapes have an empty box and are never hit (Scene.py:596-597), they only light through NEE.  (BDPT:
    test_gpu_bdpt.py::test_bdpt_spot_and_laser_emitters.)"""
    W = H = 96
    scene, cam, integ = build_gpu_scene("cornell", W, H, beam_lights=True)
    st = integ.render_frames(4)
    g = integ.hdr.to_numpy()
    o = build_oracle_scene(oracle_tables("cornell", beam_lights=True), W, H)
    ref, cnt = o.render_pt_rgb(W, H, 0, 4)
    assert np.array_equal(g, ref)
    assert int(st["rays_closest"]) == cnt["closest"] and int(st["rays_shadow"]) == cnt["shadow"]
    plain, _ = build_oracle_scene(oracle_tables("cornell"), W, H).render_pt_rgb(W, H, 0, 4)
    assert ref[..., 0].mean() > 1.05 * plain[..., 0].mean()                # the red laser really lights the scene (+12 %)


def test_pt_rgb_teapot_mc_matches_oracle(gpu_ctx, oracle_tables):
    """C3 scene (130 720 triangles) at reduced size: metal Disney, env, sphere light, smooth normals"""
    W = H = 128
    scene, cam, integ = build_gpu_scene("teapot_mc", W, H, sphere_light=True, env_power=5.0)
    scene.process_normal()
    integ.render_frames(2)
    g = integ.hdr.to_numpy()
    t = oracle_tables("teapot_mc", sphere_light=True)
    o = build_oracle_scene(t, W, H, env_power=5.0)
    vn = o.process_normal()
    gv = scene.vertex.to_numpy()
    assert np.array_equal(gv, vn, equal_nan=True)          # NaN normals: the 6 zero-area triangles of mc.obj (oracle section)
    ref, cnt = o.render_pt_rgb(W, H, 0, 2)
    assert np.array_equal(g, ref)


def test_tile_shards_sum_to_full_image(gpu_ctx):
    """rendering the 32x32-tile shards of 1, 2, 3 and 8 ranks on one GPU and summing them is
    bit-identical to the unsharded render (RNG keyed by global pixel)"""
    import parallel
    W, H = 160, 96
    scene, cam, integ = build_gpu_scene("cornell", W, H)
    integ.render_frames(3)
    full = integ.hdr.to_numpy()
    for nranks in (2, 3, 8):
        acc = np.zeros_like(full)
        for r in range(nranks):
            gpu_ctx.set_shard(r, nranks); gpu_ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
            integ.render_frames(3)
            part = integ.hdr.to_numpy()
            assert not part[~parallel.tile_mask(W, H, r, nranks)].any()
            acc += part
        assert np.array_equal(acc, full), nranks
    gpu_ctx.set_shard(0, 1)


def test_tonemap_matches_oracle(gpu_ctx):
    import UtilsFunc as UF
    W = H = 64
    scene, cam, integ = build_gpu_scene("cornell", W, H)
    integ.render_frames(2)
    UF.tone_map(0.5, integ.hdr, integ.rgb_film)
    assert np.array_equal(integ.rgb_film.to_numpy(), oracle.tonemap(integ.hdr.to_numpy(), 0.5))


# ---------------------------------------------------------------------------------- full-size properties
def test_full_size_cornell_vs_reference_image(gpu_ctx):
    """C2 size (512^2, 64 spp) through the example class; statistical pin against the reference's own
    out.png (512 spp, tone-mapped, ti.imwrite orientation): channel means within 2 %, PSNR of blurred
    images > 30 dB; plus size-independent properties (finite, non-negative, rays/pixel sane)"""
    import cv2
    import cornell_box
    ex = cornell_box.example(512, 512, 64)
    ex.build_scene()
    st = ex.integrator.render_frames(64)
    import UtilsFunc as UF
    UF.tone_map(0.5, ex.integrator.hdr, ex.integrator.rgb_film)
    hdr = ex.integrator.hdr.to_numpy(); rgb = ex.integrator.rgb_film.to_numpy()
    assert np.isfinite(hdr).all() and hdr.min() >= 0.0
    rays = st["rays_closest"] + st["rays_shadow"]
    assert 4.0 < rays / (512 * 512 * 64) < 8.0
    img = (np.clip(rgb, 0, 1) * 255.0 + 0.5).astype(np.uint8).swapaxes(0, 1)[::-1]
    ref = cv2.imread(os.path.join(GOLDEN, "out.png"))[:, :, ::-1]
    m_ours, m_ref = img.reshape(-1, 3).mean(0), ref.reshape(-1, 3).mean(0)
    assert np.all(np.abs(m_ours - m_ref) / m_ref < 0.02), (m_ours, m_ref)
    k = (9, 9)
    mse = np.mean((cv2.GaussianBlur(img, k, 0).astype(np.float64) - cv2.GaussianBlur(ref, k, 0).astype(np.float64)) ** 2)
    assert 10.0 * np.log10(255.0 ** 2 / mse) > 30.0


def test_full_size_teapot_mc_first_hit(gpu_ctx, oracle_tables):
    """C3 at full size: 1024^2 primary rays through the 261 439-node tree, bit-exact vs the oracle"""
    res = 1024
    scene, cam, integ = build_gpu_scene("teapot_mc", res, res, "debug", sphere_light=True)
    integ.render()
    g = integ.first_hit()
    o = build_oracle_scene(oracle_tables("teapot_mc", sphere_light=True), res, res)
    f = o.first_hit(res, res)
    assert np.array_equal(g["prim"], f["prim"]) and np.array_equal(g["t"], f["t"])


# ---------------------------------------------------------------------------------- edge cases
def test_ragged_image_sizes_and_single_pixel(gpu_ctx, oracle_tables):
    """image sizes that are not multiples of the 32x32 tile (partial tiles) and a 1x1 film"""
    for W, H in [(100, 70), (33, 31), (1, 1)]:
        scene, cam, integ = build_gpu_scene("cornell", W, H)
        integ.render_frames(2)
        g = integ.hdr.to_numpy()
        o = build_oracle_scene(oracle_tables("cornell"), W, H)
        ref, cnt = o.render_pt_rgb(W, H, 0, 2)
        assert g.shape == (W, H, 3) and np.array_equal(g, ref), (W, H)


def test_scene_without_lights_and_env_only(gpu_ctx, oracle_tables):
    """no emitter at all: NEE is skipped (light_count = 0); with an environment map the only light is the miss term"""
    W = H = 96
    for power in (0.0, 5.0):
        scene, cam, integ = build_gpu_scene("teapot", W, H, env_power=power)
        integ.render_frames(2)
        g = integ.hdr.to_numpy()
        o = build_oracle_scene(oracle_tables("teapot"), W, H, env_power=power)
        ref, cnt = o.render_pt_rgb(W, H, 0, 2)
        assert np.array_equal(g, ref), power
        assert cnt["shadow"] == 0 and integ_stats_shadow(gpu_ctx) == 0
        if power == 0.0:
            assert not g.any()


def integ_stats_shadow(ctx):
    return int(ctx.stats()["rays_shadow"])


def test_depth_limits_and_seed(gpu_ctx, oracle_tables):
    """max_depth 1 and 3 (the tail kernel / last-stage paths), another seed"""
    W = H = 96
    o = build_oracle_scene(oracle_tables("cornell"), W, H)
    for depth, seed in [(1, 0), (3, 7), (15, 123456789012)]:
        scene, cam, integ = build_gpu_scene("cornell", W, H)
        integ.max_depth, integ.seed = depth, seed
        integ.render_frames(3)
        ref, cnt = o.render_pt_rgb(W, H, 0, 3, max_depth=depth, seed=seed)
        assert np.array_equal(integ.hdr.to_numpy(), ref), (depth, seed)
        st = gpu_ctx.stats()
        assert st["rays_closest"] == cnt["closest"] and st["rays_shadow"] == cnt["shadow"]


def test_options_do_not_change_the_film(gpu_ctx):
    """chains, shadow overlap, tail hand-over, graph replay, batch size: bit-identical film"""
    W = H = 160
    scene, cam, integ = build_gpu_scene("sphere", W, H, sphere_light=True, glass0=True, env_power=5.0)
    ref = None
    for opts in [dict(chains=1, shadow_overlap=0, tail_max=0, graph=0), dict(chains=4, shadow_overlap=1, tail_max=16384, graph=1),
                 dict(chains=8, shadow_overlap=1, tail_max=100000000, graph=1, batch_frames=3), dict(chains=2, tail_max=64, batch_frames=1),
                 dict(chains=2, tail_max=4096, batch_frames=0, pdl=1, graph=1), dict(pdl=1, graph=0, shadow_overlap=0), dict(pdl=0)]:
        for k, v in opts.items():
            gpu_ctx.set_option(k, v)
        gpu_ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
        integ.render_frames(5)
        img = integ.hdr.to_numpy()
        if ref is None:
            ref = img
        assert np.array_equal(img, ref), opts


def test_error_paths(gpu_ctx):
    """call-order and argument errors come back as RuntimeError with the library's text, never a crash"""
    import _native
    with pytest.raises(RuntimeError, match="no scene uploaded"):
        gpu_ctx.bvh_build()
    with pytest.raises(RuntimeError):
        gpu_ctx.render_pt_rgb(0, 1)
    with pytest.raises(RuntimeError, match="bad size"):
        gpu_ctx.film_create(0, 5)
    with pytest.raises(RuntimeError, match="unknown option"):
        gpu_ctx.set_option("no_such_option", 1)
    scene = make_product_scene("cornell"); scene.setup_data_cpu(); scene.setup_data_gpu()
    gpu_ctx.film_create(8, 8)
    with pytest.raises(RuntimeError, match="camera not set"):
        gpu_ctx.render_pt_rgb(0, 1)
    with pytest.raises(RuntimeError, match="bad arguments"):
        gpu_ctx.render_pt_rgb(0, 0)
    with pytest.raises(RuntimeError, match="device"):
        _native.Context(1000)
