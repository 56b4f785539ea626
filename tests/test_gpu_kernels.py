"""GPU parity tests of the PRODUCTION traversal kernels and of the round-2 boundary additions.

The closest-hit / shadow / tail kernels that render the film (k_trace, k_shadow, k_tail: persistent warps, per-lane ray
replacement, shared-memory stacks, TMA-staged trees) are fed arbitrary rays as queue records through tr_test_trace_kernel and
compared BIT-EXACTLY with the oracle's reference walk (Scene.py:671-744 restated: explicit stack, unpruned, unordered), in
every tree mode: replicated shared-memory image, plain shared-memory tree, global memory, global memory with a staged top."""
import os
import numpy as np
import pytest
from conftest import make_product_scene
from oracle import oracle
from test_gpu_parity import build_gpu_scene, build_oracle_scene

pytestmark = pytest.mark.gpu


def _random_rays(t, n, seed):
    rng = np.random.RandomState(seed)
    lo, hi = t.bmin[0], t.bmax[0]
    ext = hi - lo
    org = (lo - 0.1 * ext + 1.2 * ext * rng.rand(n, 3)).astype(np.float32)         # inside and just outside the scene box
    d = rng.randn(n, 3); d /= np.linalg.norm(d, axis=1, keepdims=True); d = d.astype(np.float32)
    d[:100, 0] = 0.0; d[100:200, 1] = 1e-7; d[200:300, 2] = 0.0                    # "parallel axis" branch of slabs
    return org, d


def _primary_rays(o, W, H, eye):
    d = o.primary_rays(W, H).reshape(-1, 3)
    return np.repeat(np.asarray(eye, np.float32)[None], d.shape[0], 0), np.ascontiguousarray(d)


MODES = [("cornell", False, {}, "replicated shared-memory image"),
         ("cornell", False, {"replicas": 0}, "plain shared-memory tree"),
         ("cornell", False, {"smem_bvh": 0}, "global memory"),
         ("sphere", True, {}, "global memory, sphere primitive"),
         ("teapot_mc", True, {}, "global memory, 130 k triangles"),
         ("teapot_mc", True, {"top_nodes": 512}, "global memory + 512 staged top nodes"),
         ("teapot_mc", True, {"top_nodes": 1024}, "global memory + 1024 staged top nodes")]


@pytest.mark.parametrize("name,sl,opts,what", MODES)
def test_production_kernels_bit_exact(gpu_ctx, oracle_tables, name, sl, opts, what):
    scene, cam, integ = build_gpu_scene(name, 128, 128, "debug", sphere_light=sl)
    for k, v in opts.items():
        gpu_ctx.set_option(k, v)
    t = oracle_tables(name, sphere_light=sl)
    o = build_oracle_scene(t, 128, 128)
    n = 150000
    org, d = _random_rays(t, n, 17)
    po, pd = _primary_rays(o, 128, 128, cam.eye_np[0])
    org = np.concatenate([org, po]); d = np.concatenate([d, pd]); n = org.shape[0]
    ot, op, ouv = o.trace(org, d)
    hit = op >= 0
    for kernel, label in [(0, "simple walk"), (1, "k_trace"), (3, "k_tail")]:
        gt, gp, guv = gpu_ctx.test_trace(org, d, kernel=kernel)
        assert np.array_equal(gp, op), (what, label, int((gp != op).sum()))
        assert np.array_equal(gt, ot), (what, label)
        assert np.array_equal(guv[hit], ouv[hit]), (what, label)
    # shadow kernel: half of the rays must see the primitive they really hit first, the other half a random primitive
    rng = np.random.RandomState(5)
    target = np.where(rng.rand(n) < 0.5, np.maximum(op, 0), rng.randint(0, t.primitive.shape[0], n)).astype(np.int32)
    expect = np.where(op == target, target, -2).astype(np.int32)
    st, sp, _ = gpu_ctx.test_trace(org, d, kernel=2, target=target)
    assert np.array_equal(sp, expect), (what, "k_shadow", int((sp != expect).sum()))
    assert (expect >= 0).sum() > 1000


def test_degenerate_trees_trace(gpu_ctx):
    """long runs of identical Morton codes (the reference's duplicate rule builds a chain): the stack stays shallow because a
    leaf child is taken first; a single primitive is a tree without internal nodes"""
    rng = np.random.RandomState(3)
    for n in (1, 2, 3, 2049):
        base = rng.rand(n, 3).astype(np.float32)
        if n > 3:
            base[: n // 2] = base[0]
        tri = np.zeros((n * 3, 9), np.float32)
        tri[0::3, 0:3] = base; tri[1::3, 0:3] = base + np.float32([0.05, 0, 0]); tri[2::3, 0:3] = base + np.float32([0, 0.05, 0])
        tri[:, 5] = 1.0
        prim = np.zeros((n, 3), np.int32); prim[:, 0] = 1; prim[:, 1] = 3 * np.arange(n)
        mat = np.zeros((1, 10), np.float32); mat[0, 2:5] = 0.5; mat[0, 6] = 0.5

        class T: pass
        t = T(); t.vertex = tri; t.primitive = prim; t.material = mat; t.shape = np.zeros((0, 10), np.float32)
        t.light = np.zeros(0, np.int32); t.bmin = tri[:, 0:3].min(0)[None]; t.bmax = tri[:, 0:3].max(0)[None]
        gpu_ctx.scene_upload(t.vertex, t.primitive, t.material, None, None, t.bmin, t.bmax)
        gpu_ctx.bvh_build()
        o = oracle.OracleScene(t).build()
        o.lib.orc_stack_size(o.h, 8192)          # the reference's stack (64) overflows on the chain; give the oracle room to finish the walk
        m = 20000
        org = (rng.rand(m, 3) * 1.2 - 0.1).astype(np.float32); org[:, 2] = -1.0
        d = np.zeros((m, 3), np.float32); d[:, 2] = 1.0
        d[: m // 2] += (rng.randn(m // 2, 3) * 0.05).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
        org[: m // 4, 0:2] = base[0, 0:2] + np.float32([0.01, 0.01])               # straight through the stack of duplicates
        ot, op, ouv = o.trace(org, d)
        for kernel in (0, 1, 3):
            gt, gp, guv = gpu_ctx.test_trace(org, d, kernel=kernel)
            assert np.array_equal(gp, op) and np.array_equal(gt, ot), (n, kernel)
        assert (op >= 0).sum() > 0


def test_full_size_cornell_per_pixel_vs_oracle(gpu_ctx, oracle_tables):
    """C2 at full size (512^2 x 64 spp, BASELINE configs[1]) per pixel against the oracle with the shared counter-based RNG
    (the CPU port renders it in a few seconds): the 786 432 film words and both ray counts are identical"""
    W = H = 512
    scene, cam, integ = build_gpu_scene("cornell", W, H)
    st = integ.render_frames(64)
    g = integ.hdr.to_numpy()
    o = build_oracle_scene(oracle_tables("cornell"), W, H, fast=False)
    ref, cnt = o.render_pt_rgb(W, H, 0, 64)
    assert np.array_equal(g, ref)
    assert int(st["rays_closest"]) == cnt["closest"] and int(st["rays_shadow"]) == cnt["shadow"]


def test_tree_modes_render_the_same_film(gpu_ctx):
    """the four ways the kernels read the tree give bit-identical films (Cornell: replicated / plain / global; mc+Teapot: global /
    staged top)"""
    for name, sl, variants in [("cornell", False, [{}, {"replicas": 0}, {"smem_bvh": 0}]),
                               ("teapot_mc", True, [{}, {"top_nodes": 1024}])]:
        scene, cam, integ = build_gpu_scene(name, 96, 96, sphere_light=sl)
        ref = None
        for opts in variants:
            for k in ("replicas", "smem_bvh"):
                gpu_ctx.set_option(k, 1)
            gpu_ctx.set_option("top_nodes", 0)
            for k, v in opts.items():
                gpu_ctx.set_option(k, v)
            gpu_ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
            integ.render_frames(3)
            img = integ.hdr.to_numpy()
            if ref is None:
                ref = img
            assert np.array_equal(img, ref), (name, opts)


# ---------------------------------------------------------------------------------- boundary additions
def test_film_download_views_and_validation(gpu_ctx):
    import _native
    scene, cam, integ = build_gpu_scene("cornell", 64, 48)
    integ.render_frames(2)
    a, _ = gpu_ctx.film_download(True, False)
    v, r = gpu_ctx.film_download(True, True, view=True)            # zero-copy views of the pinned download buffers
    assert v.shape == (64, 48, 3) and r.shape == (64, 48, 3) and np.array_equal(a, v)
    # camera / bounds arguments that need a conversion (f64, lists) must arrive intact (temporaries kept alive)
    gpu_ctx.camera_set(np.eye(4, dtype=np.float64), np.eye(4, dtype=np.float64) * 2.0, [1.0, 2.0, 3.0], 10.0, 10.0, 32.0, 24.0)
    # out-of-range indices are rejected on the host instead of becoming out-of-bounds device reads
    s = scene
    bad = s.primitive_np.copy(); bad[3, 1] = s.vertex_np.shape[0]
    with pytest.raises(RuntimeError, match="vertex index"):
        gpu_ctx.scene_upload(s.vertex_np, bad, s.material_np, None, s.light_np, s.minboundarynp, s.maxboundarynp)
    bad = s.primitive_np.copy(); bad[0, 2] = 99
    with pytest.raises(RuntimeError, match="material index"):
        gpu_ctx.scene_upload(s.vertex_np, bad, s.material_np, None, s.light_np, s.minboundarynp, s.maxboundarynp)
    with pytest.raises(RuntimeError, match="light 0"):
        gpu_ctx.scene_upload(s.vertex_np, s.primitive_np, s.material_np, None, np.int32([4000]), s.minboundarynp, s.maxboundarynp)
    with pytest.raises(RuntimeError, match="out of range"):
        gpu_ctx.set_option("max_paths", -5)
    with pytest.raises(RuntimeError, match="out of range"):
        gpu_ctx.set_option("chains", 99)
    # a QUAD emitter has no area in the reference (Scene.get_prim_area returns 0): refused instead of rendered wrong
    import SceneData as SCD
    sc = make_product_scene("cornell")
    sh = SCD.Shape(); sh.type = SCD.SHPAE_QUAD; sh.pos = [0.0, 20.0, 0.0]; sh.setRadius(5.0)
    mt = SCD.Material(); mt.type = SCD.MAT_LIGHT; mt.setColor([50.0, 50.0, 50.0])
    sc.add_shape(sh, mt); sc.setup_data_cpu()
    with pytest.raises(RuntimeError, match="emitter shape"):
        sc.setup_data_gpu()


def test_single_rank_comm_is_a_noop(gpu_ctx):
    """tr_comm_init with one rank needs no NCCL; tr_film_reduce leaves the film alone and the reduced view is reset by a render"""
    scene, cam, integ = build_gpu_scene("cornell", 64, 64)
    gpu_ctx.comm_init(0, 1, None)
    integ.render_frames(2)
    a = integ.hdr.to_numpy()
    gpu_ctx.film_reduce()
    assert np.array_equal(integ.hdr.to_numpy(), a)
    gpu_ctx.comm_destroy()


def _two_rank_worker(rank, world, idfile, results):
    import sys
    from conftest import PKG, ROOT
    os.environ.update(LOCAL_RANK=str(rank), RANK=str(rank), WORLD_SIZE=str(world))
    import _native
    ctx = _native.reset_context(rank)
    if rank == 0:
        uid = ctx.comm_unique_id(); np.save(idfile + ".tmp.npy", uid); os.replace(idfile + ".tmp.npy", idfile)
    else:
        import time
        while not os.path.exists(idfile):
            time.sleep(0.05)
        uid = np.load(idfile)
    ctx.comm_init(rank, world, uid)
    scene, cam, integ = build_gpu_scene("cornell", 96, 64)
    ctx.set_shard(rank, world)
    out = {}
    integ.render_frames(2, stats=False)
    ctx.film_reduce()
    out["sum2"] = integ.hdr.to_numpy()
    integ.render_frames(2, stats=False)                  # progressive: two more frames, reduce again (no double counting)
    ctx.film_reduce()
    out["sum4"] = integ.hdr.to_numpy()
    ctx.film_reduce(all_ranks=True)
    out["all4"] = integ.hdr.to_numpy()
    np.savez(results % rank, **out)
    ctx.comm_destroy()


def test_two_gpu_library_reduce(tmp_path):
    """two processes, one GPU each, the library's own NCCL communicator (tr_comm_init / tr_film_reduce): rank 0 presents exactly
    the single-GPU film; the id travels through a file, no torch.distributed involved"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    idfile = str(tmp_path / "nccl_id.npy"); results = str(tmp_path / "rank%d.npz")
    mp.spawn(_two_rank_worker, args=(2, idfile, results), nprocs=2, join=True)
    import _native
    ctx = _native.reset_context(0)
    scene, cam, integ = build_gpu_scene("cornell", 96, 64)
    integ.render_frames(2); full2 = integ.hdr.to_numpy()
    integ.render_frames(2); full4 = integ.hdr.to_numpy()
    r0, r1 = np.load(results % 0), np.load(results % 1)
    assert np.array_equal(r0["sum2"], full2) and np.array_equal(r0["sum4"], full4)
    assert np.array_equal(r0["all4"], full4) and np.array_equal(r1["all4"], full4)
    assert not np.array_equal(r1["sum4"], full4)         # a non-root rank keeps presenting its partial film after a rooted reduce


def test_small_render_stress(gpu_ctx):
    """regression guard for a flaky failure seen on B200: 150 small renders each of PT_RGB and PT_Spec (128^2, every path handed
    to the tail kernel at depth 1) must give the same film every time (see the comment at k_tail in csrc/wavefront.cu)"""
    import _native
    from test_gpu_spectral import build_gpu_spectral
    for build in (lambda: build_gpu_scene("cornell", 128, 128), lambda: build_gpu_spectral(128, 128)):
        scene, cam, integ = build()
        ctx = _native.context()
        ref = None
        for i in range(150):
            ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
            for _ in range(2):
                integ.render(); cam.update_frame()
            img = integ.hdr.to_numpy()
            if ref is None:
                ref = img
            assert np.array_equal(img, ref, equal_nan=True), i


def test_full_size_teapot_mc_per_pixel_vs_oracle(gpu_ctx, oracle_tables):
    """C3 at full size (1024^2, the 130 720-triangle scene with the environment map, the sphere light and smoothed normals,
    BASELINE configs[2]) per pixel against the oracle: 2 spp, every film word and both ray counts identical"""
    W = H = 1024
    scene, cam, integ = build_gpu_scene("teapot_mc", W, H, sphere_light=True, env_power=5.0)
    scene.process_normal()
    st = integ.render_frames(2)
    g = integ.hdr.to_numpy()
    o = build_oracle_scene(oracle_tables("teapot_mc", sphere_light=True), W, H, env_power=5.0, fast=False)
    o.process_normal()
    ref, cnt = o.render_pt_rgb(W, H, 0, 2)
    assert np.array_equal(g, ref, equal_nan=True)
    assert int(st["rays_closest"]) == cnt["closest"] and int(st["rays_shadow"]) == cnt["shadow"]


def test_registered_host_tables_upload_directly(gpu_ctx):
    """page-locked caller arrays (tr_host_register, what Scene does with its packed tables) are uploaded by direct DMA: same tree,
    same film as with pageable arrays"""
    import _native
    scene, cam, integ = build_gpu_scene("teapot", 64, 64)
    assert any(p is not None for p in scene._pins)               # the 25 200-triangle vertex table is above the 1 MiB threshold
    integ.render_frames(2); a = integ.hdr.to_numpy()
    comp = scene.bvh.compact_node.to_numpy()
    scene._pins = []                                             # drop the registrations: pageable path
    import gc; gc.collect()
    scene.setup_data_gpu(); cam.dirty = True
    gpu_ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
    integ.render_frames(2)
    assert np.array_equal(integ.hdr.to_numpy(), a) and np.array_equal(scene.bvh.compact_node.to_numpy(), comp)
