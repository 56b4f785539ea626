"""CPU: pin the oracle against the reference's own golden vector and known answers (no GPU)."""
import os
import numpy as np
import pytest
from conftest import GOLDEN, model
from oracle import oracle, objload


def test_nodelist_golden(oracle_tables):
    """oracle LBVH == reference's nodelist.txt, 71/71 lines (Morton order, topology, leaf prims, boxes)"""
    s = oracle.OracleScene(oracle_tables("cornell")).build(literal_sort=True)
    gold = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "nodelist.txt"))]
    assert oracle.nodelist_lines(s.compact) == gold
    # sorted primitive order recorded in SURVEY Appendix B
    assert s.morton[:, 1].tolist() == [31, 25, 9, 24, 4, 1, 27, 5, 29, 23, 26, 22, 8, 6, 34, 20, 21, 32, 18, 17, 19, 12,
                                       2, 0, 13, 10, 16, 3, 15, 14, 11, 28, 33, 30, 35, 7]
    assert [hex(c) for c in s.morton[:4, 0]] == ["0x6186186", "0x9da4ce2", "0xa28a0ca", "0xaf0ec3d"]


@pytest.mark.parametrize("name", ["cornell", "sphere", "teapot"])
def test_literal_sort_equals_stable_sort(oracle_tables, name):
    """30 one-bit Blelloch passes (accel/LBvh.py:55-72) == stable sort on the low 30 bits"""
    t = oracle_tables(name)
    a = oracle.OracleScene(t).build(literal_sort=True)
    b = oracle.OracleScene(t).build(literal_sort=False)
    assert np.array_equal(a.morton, b.morton) and np.array_equal(a.compact, b.compact) and np.array_equal(a.bvh_node, b.bvh_node)


@pytest.mark.parametrize("name,dups", [("sphere", 1520), ("teapot", 2)])
def test_tree_validity_with_duplicates(oracle_tables, name, dups):
    """the reference's duplicate-key rule still yields a valid tree"""
    t = oracle_tables(name)
    s = oracle.OracleScene(t).build()
    n = t.primitive.shape[0]
    codes = s.morton[:, 0]
    assert (np.diff(codes) >= 0).all()
    assert int((np.diff(codes) == 0).sum()) == dups
    leaves = s.compact[(s.compact[:, 0].astype(np.int32) & 1) == 1]
    assert leaves.shape[0] == n and sorted(leaves[:, 1].astype(np.int64).tolist()) == list(range(n))
    # every internal node's box is the union of its children (left = idx+1, right = word 1)
    c = s.compact
    internal = np.nonzero((c[:, 0].astype(np.int32) & 1) == 0)[0]
    l, r = internal + 1, c[internal, 1].astype(np.int64)
    assert np.array_equal(c[internal, 2:5], np.minimum(c[l, 2:5], c[r, 2:5]))
    assert np.array_equal(c[internal, 5:8], np.maximum(c[l, 5:8], c[r, 5:8]))


def test_cornell_known_answers(oracle_tables):
    """frame-0 first hits at 256^2"""
    t = oracle_tables("cornell")
    s = oracle.OracleScene(t).build()
    cam = oracle.fit_camera(t, 256, 256)
    assert np.allclose(cam[2], [278.0, 274.4, 488.9923], atol=1e-3)
    s.set_camera(cam[1], cam[2], *cam[3:])
    fh = s.first_hit(256, 256)
    assert int((fh["t"] < 1e6).sum()) == 57867
    assert int(np.isin(fh["prim"], [34, 35]).sum()) == 366
    assert fh["prim"][128, 128] == 29 and abs(fh["t"][128, 128] - 780.9606) < 1e-3
    assert abs(fh["node_visits"] / fh["rays"] - 21.12) < 0.05 and abs(fh["leaf_tests"] / fh["rays"] - 9.04) < 0.05
    assert fh["max_stack"] == 4
    assert abs(s.total_area() - 130.0 * 105.0) < 1.0


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)"""
    L = oracle.lib()
    out = np.zeros(4, np.uint32)
    L.orc_philox_raw(0, 0, 0, 0, 0, 0, out)
    assert [hex(x) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = 0xFFFFFFFF
    L.orc_philox_raw(f, f, f, f, f, f, out)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    L.orc_philox_raw(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, out)
    assert [hex(x) for x in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_render_is_deterministic_and_mask_composable(oracle_tables):
    """same seed -> same image; rendering disjoint pixel masks and summing == rendering everything"""
    t = oracle_tables("cornell")
    s = oracle.OracleScene(t).build()
    W = H = 64
    cam = oracle.fit_camera(t, W, H)
    s.set_camera(cam[1], cam[2], *cam[3:])
    a, ca = s.render_pt_rgb(W, H, 0, 3, seed=7)
    b, cb = s.render_pt_rgb(W, H, 0, 3, seed=7)
    assert np.array_equal(a, b) and ca == cb
    m = np.zeros((W, H), bool); m[:32] = True
    p0, _ = s.render_pt_rgb(W, H, 0, 3, seed=7, mask=m)
    p1, _ = s.render_pt_rgb(W, H, 0, 3, seed=7, mask=~m)
    assert np.array_equal(p0 + p1, a)
    # frames can be rendered in pieces (running mean is sequential)
    c, _ = s.render_pt_rgb(W, H, 0, 2, seed=7)
    c, _ = s.render_pt_rgb(W, H, 2, 1, seed=7, hdr=c)
    assert np.array_equal(c, a)


def test_cornell_statistics_vs_reference_image(oracle_tables):
    """statistical pin against the reference's own out.png (512^2, 512 spp, tone-mapped):
    a 128^2 / 48 spp oracle render must have the same per-channel mean within 4 %"""
    import cv2
    t = oracle_tables("cornell")
    s = oracle.OracleScene(t, fast=True).build()
    W = H = 128
    cam = oracle.fit_camera(t, W, H)
    s.set_camera(cam[1], cam[2], *cam[3:])
    hdr, _ = s.render_pt_rgb(W, H, 0, 48)
    rgb = oracle.tonemap(hdr, 0.5)
    img = (np.clip(rgb, 0, 1) * 255.0 + 0.5).astype(np.uint8).swapaxes(0, 1)[::-1]      # ti.imwrite convention
    ref = cv2.imread(os.path.join(GOLDEN, "out.png"))[:, :, ::-1]
    ref_small = cv2.resize(ref, (W, H), interpolation=cv2.INTER_AREA)
    m_ours, m_ref = img.reshape(-1, 3).mean(0), ref_small.reshape(-1, 3).mean(0)
    assert np.all(np.abs(m_ours - m_ref) / m_ref < 0.04), (m_ours, m_ref)
    # and the picture itself agrees (PSNR on the down-sampled images)
    mse = np.mean((cv2.GaussianBlur(img, (5, 5), 0).astype(np.float64) - cv2.GaussianBlur(ref_small, (5, 5), 0).astype(np.float64)) ** 2)
    assert 10.0 * np.log10(255.0 ** 2 / mse) > 25.0


def test_process_normal_smooths_sphere(oracle_tables):
    """sphere.obj: after process_normal the vertex normals point radially (unit sphere at the origin)"""
    t = oracle_tables("sphere")
    s = oracle.OracleScene(t).build()
    v = s.process_normal()
    pos, nrm = v[:, 0:3], v[:, 3:6]
    radial = pos / np.linalg.norm(pos, axis=1, keepdims=True)
    cosang = (radial * nrm).sum(1)
    assert np.isfinite(nrm).all() and cosang.min() > 0.99


def test_cpp_oracle_pt_rgb_matches_literal_python_transliteration(oracle_tables):
    """PathTrace.render restated a second time in plain Python (oracle/pt_literal.py: hit attributes, light sampling, Disney
    sampling and evaluation, NEE with the power heuristic, throughput recursion, RNG block order) against the C++ oracle, pixel by
    pixel on the Cornell box for an unjittered and two jittered frames.  numpy's float32 sin / cos differ from glibc's by ULPs:
    rel 1e-4, and <= 1 % of the pixels may differ more (a path that changes a branch at a float boundary)"""
    from oracle import pt_literal as PL
    W = H = 24
    # second pass: the white walls / boxes become glass (ior 1.3, extinction 5); third: a laser and a spot light are added
    # (Scene.sample_li's SPOT / LASER branches, Scene.py:493-516, restated independently in pt_literal.sample_li)
    for glass0, beams, frames in ((False, False, (0, 1, 3)), (True, False, (0, 3)), (False, True, (0, 2))):
      t = oracle_tables("cornell", glass0=glass0, beam_lights=beams)
      s = oracle.OracleScene(t).build()
      cam = oracle.fit_camera(t, W, H)
      s.set_camera(cam[1], cam[2], *cam[3:])
      tr = PL.Tracer(s, t, cam)
      for frame in frames:
          hdr, cnt = s.render_pt_rgb(W, H, frame, 1)
          ref = hdr * np.float32(frame + 1)                # film started at 0: hdr = L / (frame + 1), exact for these frames
          got = np.zeros_like(ref); nc = ns = 0
          for i in range(W):
              for j in range(H):
                  got[i, j], (a, b) = tr.pixel(i, j, frame)
                  nc += a; ns += b
          bad = np.abs(got - ref).max(axis=2) > 1e-4 * np.maximum(1.0, np.abs(ref).max(axis=2))
          assert bad.mean() <= 0.01, (frame, int(bad.sum()))
          assert abs(nc - cnt["closest"]) <= 0.005 * cnt["closest"] and abs(ns - cnt["shadow"]) <= 0.005 * cnt["shadow"]
          assert ref.max() > 1.0 and (ref.max(axis=2) > 0).mean() > (0.02 if glass0 else 0.5)


def test_spot_and_laser_emitters_light_through_nee(oracle_tables):
    """Scene.sample_li's SPOT / LASER branches (Scene.py:493-516): both shapes have an empty box and are never hit, a laser of
    radius 60 pointing down lights a disc of the Cornell floor under it and nothing outside its beam"""
    from test_gpu_parity import build_oracle_scene
    W = 64
    plain, _ = build_oracle_scene(oracle_tables("cornell"), W, W).render_pt_rgb(W, W, 0, 4)
    lit, cnt = build_oracle_scene(oracle_tables("cornell", beam_lights=True), W, W).render_pt_rgb(W, W, 0, 4)
    assert np.isfinite(lit).all() and cnt["shadow"] > 0
    assert lit[..., 0].mean() > 1.05 * plain[..., 0].mean()
    assert ((lit - plain).sum(axis=2) > 0.05).sum() > 50                # directly lit pixels exist
