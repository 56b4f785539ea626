"""ORACLE (test infrastructure, not product code): ctypes front-end of oracle/tiray_oracle.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module.  See the header of tiray_oracle.cpp for what is pinned and what is not.
"""
import ctypes as C
import math
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _load(fast):
    path = os.path.join(_HERE, "liboracle_fast.so" if fast else "liboracle.so")
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    lib.orc_scene_create.restype = C.c_void_p
    lib.orc_scene_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.orc_scene_destroy.argtypes = [C.c_void_p]
    lib.orc_bvh_build.argtypes = [C.c_void_p, C.c_int]
    lib.orc_bvh_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orc_morton_unsorted.argtypes = [C.c_void_p, _i32p]
    lib.orc_refit_sweeps.argtypes = [C.c_void_p]
    lib.orc_vertex_get.argtypes = [C.c_void_p, _f32p]
    lib.orc_camera_set.argtypes = [C.c_void_p, _f32p, _f32p, C.c_float, C.c_float, C.c_float, C.c_float]
    lib.orc_env_set.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, C.c_float]
    lib.orc_stack_size.argtypes = [C.c_void_p, C.c_int]
    lib.orc_max_stack_seen.argtypes = [C.c_void_p]
    lib.orc_overflow.argtypes = [C.c_void_p]
    lib.orc_total_area.argtypes = [C.c_void_p]; lib.orc_total_area.restype = C.c_float
    lib.orc_primary_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, _f32p]
    lib.orc_first_hit.argtypes = [C.c_void_p, C.c_int, C.c_int, _f32p, _i32p, _f32p, _f32p, _f32p, _f32p, _u64p]
    lib.orc_trace.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, C.c_int, _f32p, _i32p, C.c_void_p]
    lib.orc_render_debug.argtypes = [C.c_void_p, C.c_int, C.c_int, _f32p]
    lib.orc_render_pt_rgb.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                      _f32p, C.c_void_p, _u64p]
    lib.orc_tonemap.argtypes = [C.c_int, C.c_float, _f32p, _f32p]
    lib.orc_process_normal.argtypes = [C.c_void_p]
    lib.orc_disney_evaluate_pdf.argtypes = [C.c_int, _f32p, _f32p, _f32p, C.c_float, C.c_float, _f32p]
    lib.orc_disney_sample.argtypes = [C.c_int, _f32p, _f32p, C.c_float, C.c_float, _f32p, _f32p]
    lib.orc_glass_sample.argtypes = [C.c_int, _f32p, _f32p, C.c_float, _f32p, _f32p]
    lib.orc_offset_ray.argtypes = [C.c_int, _f32p, _f32p, _f32p]
    lib.orc_rng.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _f32p]
    lib.orc_philox_raw.argtypes = [C.c_uint32] * 6 + [_u32p]
    lib.orc_slabs.argtypes = [_f32p, _f32p, _f32p, _f32p]
    lib.orc_set_num_threads.argtypes = [C.c_int]
    lib.orc_math.argtypes = [C.c_int, C.c_int, _f32p, C.c_void_p, _f32p]
    lib.orc_camera_view_set.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int]
    lib.orc_render_bdpt_rgb.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, _f32p, C.c_void_p, _u64p]
    lib.orc_bdpt_pixel_dump.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, _f32p, _i32p, _f32p]
    return lib


_libs = {}


def lib(fast=False):
    if fast not in _libs:
        _libs[fast] = _load(fast)
    return _libs[fast]


def camera_matrices(W, H, target, scale, yaw=0.0, pitch=0.0):
    """Camera.__init__ + Camera.update restated (/root/reference/Camera.py:26-34,70-93).
    Returns view (4x4 f32), view_inv (4x4 f32), eye (3 f32), fx, fy, cx, cy."""
    fx = 2.0 * W / 2.4
    pitch = max(min(pitch, 1.57), -1.57)
    eye = np.ones((1, 3), np.float32)
    target = np.asarray(target, np.float64)
    eye[0, 0] = target[0] + scale * math.cos(pitch) * math.sin(yaw)
    eye[0, 1] = target[1] + scale * math.sin(pitch)
    eye[0, 2] = target[2] + scale * math.cos(pitch) * math.cos(yaw)
    up = np.array([-math.sin(pitch) * math.sin(yaw), math.cos(pitch), -math.sin(pitch) * math.cos(yaw)])
    z = eye[0, :] - target; z = z / np.linalg.norm(z)
    x = np.cross(up, z); x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    view = np.zeros((1, 4, 4), np.float32)
    view[0] = np.array([[x[0], x[1], x[2], -np.dot(x, eye[0, :])], [y[0], y[1], y[2], -np.dot(y, eye[0, :])],
                        [z[0], z[1], z[2], -np.dot(z, eye[0, :])], [0.0, 0.0, 0.0, 1.0]])
    view_inv = np.linalg.inv(view).astype(np.float32)
    return view[0].copy(), view_inv[0].copy(), eye[0].copy(), fx, fx, W * 0.5, H * 0.5


def fit_camera(tables, W, H, factor=0.8):
    """example/cornell_box.py:26-30: scale = |size|*0.8, target = (max+min)/2"""
    centre = tables.bmax + tables.bmin
    size = tables.bmax - tables.bmin
    scale = math.sqrt(size[0, 0] * size[0, 0] + size[0, 1] * size[0, 1] + size[0, 2] * size[0, 2]) * factor
    return camera_matrices(W, H, (centre[0, 0] * 0.5, centre[0, 1] * 0.5, centre[0, 2] * 0.5), scale)


def load_env(path):
    """texture/Texture.py:18-34: packed RGB i32, buf[x][H-1-row]"""
    import cv2
    img = cv2.imread(path)
    h, w = img.shape[0], img.shape[1]
    b = img[:, :, 0].astype(np.int32); g = img[:, :, 1].astype(np.int32); r = img[:, :, 2].astype(np.int32)
    packed = (r << 16) | (g << 8) | b                      # [row][col]
    return np.ascontiguousarray(packed[::-1, :].T), w, h    # [x][H-1-row]


class OracleScene:
    def __init__(self, tables, fast=False):
        self.lib = lib(fast)
        self.t = tables
        self.n = int(tables.primitive.shape[0])
        ns = int(tables.shape.shape[0])
        self._keep = [np.ascontiguousarray(tables.vertex, np.float32), np.ascontiguousarray(tables.primitive, np.int32),
                      np.ascontiguousarray(tables.material, np.float32), np.ascontiguousarray(tables.shape, np.float32),
                      np.ascontiguousarray(tables.light, np.int32),
                      np.ascontiguousarray(tables.bmin, np.float32), np.ascontiguousarray(tables.bmax, np.float32)]
        k = self._keep
        self.h = self.lib.orc_scene_create(k[0].ctypes.data, k[0].shape[0], k[1].ctypes.data, self.n, k[2].ctypes.data,
                                           k[2].shape[0], k[3].ctypes.data if ns else None, ns,
                                           k[4].ctypes.data if k[4].size else None, int(k[4].size),
                                           k[5].ctypes.data, k[6].ctypes.data)

    def __del__(self):
        try:
            self.lib.orc_scene_destroy(self.h)
        except Exception:
            pass

    def morton_unsorted(self):
        out = np.zeros((self.n, 2), np.int32); self.lib.orc_morton_unsorted(self.h, out); return out

    def build(self, literal_sort=False):
        rc = self.lib.orc_bvh_build(self.h, int(literal_sort))
        if rc != 0:
            raise RuntimeError("aabb gen error")
        nn = 2 * self.n - 1
        self.morton = np.zeros((self.n, 2), np.int32)
        self.bvh_node = np.zeros((nn, 11), np.float32)
        self.compact = np.zeros((nn, 9), np.float32)
        self.lib.orc_bvh_get(self.h, self.morton.ctypes.data, self.bvh_node.ctypes.data, self.compact.ctypes.data)
        return self

    def set_camera(self, view_inv, eye, fx, fy, cx, cy):
        self.lib.orc_camera_set(self.h, np.ascontiguousarray(view_inv, np.float32).reshape(-1),
                                np.ascontiguousarray(eye, np.float32), fx, fy, cx, cy)

    def set_camera_view(self, view, wid, hgt):
        """Camera.view / wid / hgt, needed by BDPT (Camera.py:128-129,144-158)"""
        self.lib.orc_camera_view_set(self.h, np.ascontiguousarray(view, np.float32).reshape(-1), wid, hgt)

    def set_env(self, packed, w, h, power):
        self.lib.orc_env_set(self.h, np.ascontiguousarray(packed, np.int32).reshape(-1), w, h, power)

    def total_area(self):
        return float(self.lib.orc_total_area(self.h))

    def primary_rays(self, W, H):
        d = np.zeros((W, H, 3), np.float32); self.lib.orc_primary_rays(self.h, W, H, d.reshape(-1)); return d

    def first_hit(self, W, H):
        t = np.zeros(W * H, np.float32); prim = np.zeros(W * H, np.int32)
        uv = np.zeros(W * H * 2, np.float32); pos = np.zeros(W * H * 3, np.float32)
        gn = np.zeros(W * H * 3, np.float32); nrm = np.zeros(W * H * 3, np.float32)
        stats = np.zeros(4, np.uint64)
        self.lib.orc_first_hit(self.h, W, H, t, prim, uv, pos, gn, nrm, stats)
        return dict(t=t.reshape(W, H), prim=prim.reshape(W, H), uv=uv.reshape(W, H, 2), pos=pos.reshape(W, H, 3),
                    gnormal=gn.reshape(W, H, 3), normal=nrm.reshape(W, H, 3),
                    rays=int(stats[0]), node_visits=int(stats[1]), leaf_tests=int(stats[2]), max_stack=int(stats[3]))

    def trace(self, o, d, shadow=False):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3); d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = o.shape[0]
        t = np.zeros(n, np.float32); prim = np.zeros(n, np.int32); uv = np.zeros((n, 2), np.float32)
        self.lib.orc_trace(self.h, n, o.reshape(-1), d.reshape(-1), int(shadow), t, prim, uv.ctypes.data)
        return t, prim, uv

    def render_debug(self, W, H):
        hdr = np.zeros((W, H, 3), np.float32); self.lib.orc_render_debug(self.h, W, H, hdr.reshape(-1)); return hdr

    def render_pt_rgb(self, W, H, frame_begin, n_frames, max_depth=15, seed=0, hdr=None, mask=None):
        if hdr is None:
            hdr = np.zeros((W, H, 3), np.float32)
        cnt = np.zeros(4, np.uint64)
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask, np.uint8).reshape(-1)
        self.lib.orc_render_pt_rgb(self.h, W, H, frame_begin, n_frames, max_depth, seed, hdr.reshape(-1),
                                   m.ctypes.data if m is not None else None, cnt)
        return hdr, dict(closest=int(cnt[0]), shadow=int(cnt[1]), node_visits=int(cnt[2]), leaf_tests=int(cnt[3]))

    def render_bdpt_rgb(self, W, H, frame_begin, n_frames, seed=0, hdr=None, mask=None):
        if hdr is None:
            hdr = np.zeros((W, H, 3), np.float32)
        cnt = np.zeros(4, np.uint64)
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask, np.uint8).reshape(-1)
        self.lib.orc_render_bdpt_rgb(self.h, W, H, frame_begin, n_frames, seed, hdr.reshape(-1),
                                     m.ctypes.data if m is not None else None, cnt)
        return hdr, dict(closest=int(cnt[0]), shadow=int(cnt[1]))

    def bdpt_pixel_dump(self, i, j, frame, seed=0):
        """-> verts (13, 20) f32 [eye 0..6, light 0..5], (eye_depth, light_depth), contrib (7, 7, 4) f32 indexed [e-1][l]"""
        verts = np.zeros((13, 20), np.float32); depths = np.zeros(2, np.int32); contrib = np.zeros((7, 7, 4), np.float32)
        self.lib.orc_bdpt_pixel_dump(self.h, i, j, frame, seed, verts.reshape(-1), depths, contrib.reshape(-1))
        return verts, (int(depths[0]), int(depths[1])), contrib

    def process_normal(self):
        self.lib.orc_process_normal(self.h)
        out = np.zeros_like(self._keep[0]); self.lib.orc_vertex_get(self.h, out.reshape(-1)); return out


def math_fn(fn, a, b=None):
    """include/trmath.h on the host: fn 0 sin, 1 cos, 2 exp, 3 acos, 4 atan2(a, b), 5 pow(a, b)"""
    a = np.ascontiguousarray(a, np.float32).reshape(-1); out = np.zeros_like(a)
    bb = None if b is None else np.ascontiguousarray(b, np.float32).reshape(-1)
    lib().orc_math(int(fn), a.size, a, None if bb is None else bb.ctypes.data, out)
    return out


def tonemap(hdr, exposure=0.5):
    hdr = np.ascontiguousarray(hdr, np.float32)
    out = np.zeros_like(hdr)
    lib().orc_tonemap(hdr.size // 3, exposure, hdr.reshape(-1), out.reshape(-1))
    return out


def nodelist_lines(compact):
    """accel/LBvh.py:127-136 print_compact_info formatting (column labels shifted)"""
    out = []
    for i in range(compact.shape[0]):
        c = compact[i]
        out.append("node:%d pri:%d offset:%d min:%.2f %.2f %.2f max:%.2f %.2f %.2f" %
                   (i, int(c[1]), int(c[2]), c[3], c[4], c[5], c[6], c[7], c[8]))
    return out
