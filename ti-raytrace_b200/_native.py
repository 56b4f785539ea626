"""ctypes binding of libtiray.so (include/tiray.h) and the process-wide device context.

There is no CPU fallback: if the shared library is missing or no CUDA device is visible, every
device operation raises.  One context per process (one process per GPU), like Taichi's single program.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = os.environ.get("TIRAY_LIB", "libtiray.so")

_f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_vp = C.c_void_p


class Stats(C.Structure):
    _fields_ = [("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("node_visits", C.c_uint64),
                ("leaf_tests", C.c_uint64), ("kernel_launches", C.c_uint64), ("ms_total", C.c_float),
                ("ms_trace", C.c_float), ("ms_shade", C.c_float), ("ms_shadow", C.c_float), ("ms_build", C.c_float),
                ("frames", C.c_int32), ("paths_in_flight", C.c_int32),
                ("node_visits_shadow", C.c_uint64), ("leaf_tests_shadow", C.c_uint64),
                ("chains", C.c_int32), ("pad_", C.c_int32), ("shade_terminal", C.c_uint64)]


# every symbol include/tiray.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "tr_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "tr_ctx_destroy": (None, [_vp]),
    "tr_last_error": (C.c_char_p, [_vp]),
    "tr_device_count": (C.c_int, []),
    "tr_synchronize": (C.c_int, [_vp]),
    "tr_host_register": (C.c_int, [_vp, C.c_size_t]),
    "tr_host_unregister": (C.c_int, [_vp]),
    "tr_stream_set": (C.c_int, [_vp, _vp]),
    "tr_scene_upload": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp]),
    "tr_material_upload": (C.c_int, [_vp, _vp, C.c_int]),
    "tr_env_upload": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_float]),
    "tr_obj_open": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "tr_obj_close": (None, [_vp]),
    "tr_obj_last_error": (C.c_char_p, []),
    "tr_obj_material_count": (C.c_int, [_vp]),
    "tr_obj_material": (C.c_int, [_vp, C.c_int, C.c_char_p, C.c_int, _vp, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "tr_obj_material_vertices": (C.c_int, [_vp, C.c_int, _vp]),
    "tr_bvh_build": (C.c_int, [_vp]),
    "tr_bvh_download": (C.c_int, [_vp, _vp, _vp, _vp]),
    "tr_morton_download": (C.c_int, [_vp, _vp]),
    "tr_process_normal": (C.c_int, [_vp]),
    "tr_vertex_download": (C.c_int, [_vp, _vp]),
    "tr_total_area": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "tr_camera_set": (C.c_int, [_vp, _vp, _vp, _vp, C.c_float, C.c_float, C.c_float, C.c_float]),
    "tr_film_create": (C.c_int, [_vp, C.c_int, C.c_int]),
    "tr_film_clear": (C.c_int, [_vp]),
    "tr_film_download": (C.c_int, [_vp, _vp, _vp]),
    "tr_film_upload": (C.c_int, [_vp, _vp]),
    "tr_film_download_pinned": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp)]),
    "tr_film_device_ptr": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "tr_comm_unique_id": (C.c_int, [_vp]),
    "tr_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "tr_film_reduce": (C.c_int, [_vp, C.c_int]),
    "tr_comm_destroy": (C.c_int, [_vp]),
    "tr_set_shard": (C.c_int, [_vp, C.c_int, C.c_int]),
    "tr_render_pt_rgb": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "tr_render_pt_spec": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "tr_render_bdpt_rgb": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint64]),
    "tr_bdpt_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "tr_test_bdpt_dump": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "tr_render_debug": (C.c_int, [_vp]),
    "tr_first_hit_download": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tr_tonemap": (C.c_int, [_vp, C.c_float]),
    "tr_stats_get": (C.c_int, [_vp, C.POINTER(Stats)]),
    "tr_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "tr_test_disney_evaluate_pdf": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_float, C.c_float, _vp]),
    "tr_test_disney_sample": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_float, C.c_float, _vp, _vp]),
    "tr_test_glass_sample": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_float, _vp, _vp]),
    "tr_test_offset_ray": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "tr_test_rng": (C.c_int, [_vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "tr_test_math": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "tr_test_trace": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "tr_test_trace_kernel": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "tr_spec_sensor_upload": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_float]),
    "tr_spec_spectrum_upload": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_float, C.c_float]),
    "tr_spec_spectrum_download": (C.c_int, [_vp, C.c_int, _vp]),
    "tr_spec_rgb2spec_upload": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "tr_spec_sky_upload": (C.c_int, [_vp, _vp, _vp, _vp]),
    "tr_spec_normalize": (C.c_int, [_vp, C.c_int, _vp]),
    "tr_test_srgb_to_spec": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "tr_test_sky_radiance": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp]),
    "tr_test_spectrum_sample": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
}

_libs = {}


def lib_path(name=None):
    return os.path.join(ROOT, name or LIB_NAME)


def load_library(name=None):
    """Load libtiray.so (built in-tree by __graft_entry__.build / csrc/Makefile). Raises if absent.
    name selects another flavour of the same ABI (libtiray_counters.so: per-ray visit counters)."""
    name = name or LIB_NAME
    if name not in _libs:
        path = lib_path(name)
        if not os.path.exists(path):
            raise RuntimeError("%s not found: build it with `make -C %s` (there is no CPU fallback)"
                               % (path, os.path.join(ROOT, "csrc")))
        lib = C.CDLL(path)
        for sym, (res, args) in SIGNATURES.items():
            if os.environ.get("TIRAY_ALLOW_MISSING") and not hasattr(lib, sym):
                continue                     # developer A/B runs against an older build of the library (tools/perf_probe.py --lib)
            fn = getattr(lib, sym)           # AttributeError if the library does not export the symbol
            fn.restype, fn.argtypes = res, args
        _libs[name] = lib
    return _libs[name]


def _ptr(a):
    """address of a numpy array for a void* argument.  The CALLER must keep `a` alive across the foreign call: pass a named
    local, never a temporary expression (numpy's small-block cache would hand the freed buffer to the next argument)."""
    return None if a is None else a.ctypes.data


def _f(a, shape=None):
    a = np.ascontiguousarray(a, np.float32)
    return a if shape is None else a.reshape(shape)


def pin_array(a):
    """Best effort: page-lock the memory of a numpy array that will be uploaded repeatedly (Scene's packed tables), so that
    tr_scene_upload DMAs from it directly.  Returns a finalizer-carrying token to keep next to the array (unregisters when
    dropped), or None when there is no device / library (CPU-only use of the host classes) or the array is small."""
    import weakref
    if a is None or a.nbytes < (1 << 20) or not a.flags["C_CONTIGUOUS"]:
        return None
    try:
        lib = load_library()
        if lib.tr_device_count() <= 0 or lib.tr_host_register(a.ctypes.data, a.nbytes) != 0:
            return None
    except Exception:
        return None

    class _Pin:
        pass
    tok = _Pin(); addr = a.ctypes.data
    weakref.finalize(tok, lambda: lib.tr_host_unregister(addr))
    tok.array = a                           # the registration lives exactly as long as the array is reachable through the token
    return tok


class Context:
    """Owner of one tr_ctx. All methods raise RuntimeError with the library's error text on failure."""

    def __init__(self, device=None, lib_name=None):
        self.lib = load_library(lib_name)
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = _vp()
        rc = self.lib.tr_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise RuntimeError("tr_ctx_create(%d) failed: %s" % (device, (self.lib.tr_last_error(None) or b"").decode()))
        self.h, self.device = h, int(device)
        self.W = self.H = 0
        self.n_prims = self.n_verts = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.tr_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, (self.lib.tr_last_error(self.h) or b"").decode()))

    # ---- scene
    def scene_upload(self, vertex, prim, material, shape, light, bmin, bmax):
        v = _f(vertex); p = np.ascontiguousarray(prim, np.int32); m = _f(material)
        s = _f(shape) if shape is not None and len(shape) else None
        l = np.ascontiguousarray(light, np.int32) if light is not None and len(light) else None
        lo = _f(bmin).reshape(-1); hi = _f(bmax).reshape(-1)            # named locals: alive until the call returns
        assert lo.size == 3 and hi.size == 3
        self._ck(self.lib.tr_scene_upload(self.h, _ptr(v), v.shape[0], _ptr(p), p.shape[0], _ptr(m), m.shape[0],
                                          _ptr(s), 0 if s is None else s.shape[0], _ptr(l), 0 if l is None else l.shape[0],
                                          _ptr(lo), _ptr(hi)), "tr_scene_upload")
        self.n_prims, self.n_verts = p.shape[0], v.shape[0]

    def material_upload(self, material):
        m = _f(material); self._ck(self.lib.tr_material_upload(self.h, _ptr(m), m.shape[0]), "tr_material_upload")

    def env_upload(self, packed, w, h, power):
        a = np.ascontiguousarray(packed, np.int32)
        self._ck(self.lib.tr_env_upload(self.h, _ptr(a), int(w), int(h), float(power)), "tr_env_upload")

    def bvh_build(self):
        self._ck(self.lib.tr_bvh_build(self.h), "tr_bvh_build")

    def bvh_download(self):
        n = self.n_prims
        morton = np.zeros((n, 2), np.int32); node = np.zeros((2 * n - 1, 11), np.float32); comp = np.zeros((2 * n - 1, 9), np.float32)
        self._ck(self.lib.tr_bvh_download(self.h, _ptr(morton), _ptr(node), _ptr(comp)), "tr_bvh_download")
        return morton, node, comp

    def morton_download(self):
        out = np.zeros((self.n_prims, 2), np.int32)
        self._ck(self.lib.tr_morton_download(self.h, _ptr(out)), "tr_morton_download"); return out

    def process_normal(self):
        self._ck(self.lib.tr_process_normal(self.h), "tr_process_normal")

    def vertex_download(self):
        out = np.zeros((self.n_verts, 9), np.float32)
        self._ck(self.lib.tr_vertex_download(self.h, _ptr(out)), "tr_vertex_download"); return out

    def total_area(self):
        a = C.c_float(0.0); self._ck(self.lib.tr_total_area(self.h, C.byref(a)), "tr_total_area"); return float(a.value)

    # ---- camera / film
    def camera_set(self, view, view_inv, eye, fx, fy, cx, cy):
        v = None if view is None else _f(view).reshape(-1); vi = _f(view_inv).reshape(-1); e = _f(eye).reshape(-1)
        assert (v is None or v.size == 16) and vi.size == 16 and e.size == 3
        self._ck(self.lib.tr_camera_set(self.h, _ptr(v), _ptr(vi), _ptr(e), fx, fy, cx, cy), "tr_camera_set")

    def film_create(self, W, H):
        self._ck(self.lib.tr_film_create(self.h, int(W), int(H)), "tr_film_create"); self.W, self.H = int(W), int(H)

    def film_clear(self):
        self._ck(self.lib.tr_film_clear(self.h), "tr_film_clear")

    def film_download(self, hdr=True, rgb=False, view=False):
        """(hdr, rgb) as (W,H,3) f32 arrays.  view=True returns zero-copy views of the context's pinned download buffers (valid
        until the next download) instead of fresh arrays: the film is DMA'd once and not copied again on the host."""
        if view:
            a, b = _vp(), _vp()
            self._ck(self.lib.tr_film_download_pinned(self.h, int(hdr), int(rgb), C.byref(a), C.byref(b)), "tr_film_download_pinned")
            shape = (self.W, self.H, 3)
            mk = lambda p: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=shape) if p.value else None
            return mk(a), mk(b)
        a = np.empty((self.W, self.H, 3), np.float32) if hdr else None
        b = np.empty((self.W, self.H, 3), np.float32) if rgb else None
        self._ck(self.lib.tr_film_download(self.h, _ptr(a), _ptr(b)), "tr_film_download")
        return a, b

    def film_upload(self, hdr):
        a = _f(hdr); assert a.size == self.W * self.H * 3
        self._ck(self.lib.tr_film_upload(self.h, _ptr(a)), "tr_film_upload")

    def film_device_ptr(self):
        a, b = _vp(), _vp()
        self._ck(self.lib.tr_film_device_ptr(self.h, C.byref(a), C.byref(b)), "tr_film_device_ptr")
        return a.value, b.value

    def set_shard(self, rank, nranks):
        self._ck(self.lib.tr_set_shard(self.h, int(rank), int(nranks)), "tr_set_shard")

    # ---- multi-GPU: NCCL communicator owned by the library (one context per GPU / process)
    def comm_unique_id(self):
        buf = np.zeros(128, np.uint8)
        self._ck(self.lib.tr_comm_unique_id(_ptr(buf)), "tr_comm_unique_id")
        return buf

    def comm_init(self, rank, nranks, unique_id=None):
        uid = None if unique_id is None else np.ascontiguousarray(unique_id, np.uint8)
        assert uid is None or uid.size == 128
        self._ck(self.lib.tr_comm_init(self.h, int(rank), int(nranks), _ptr(uid)), "tr_comm_init")

    def film_reduce(self, all_ranks=False):
        """one ncclReduce (sum) of the per-rank partial films, enqueued on the context's stream; rank 0 then presents the image"""
        self._ck(self.lib.tr_film_reduce(self.h, int(all_ranks)), "tr_film_reduce")

    def comm_destroy(self):
        self._ck(self.lib.tr_comm_destroy(self.h), "tr_comm_destroy")

    # ---- integrators
    def render_pt_rgb(self, frame_begin, n_frames, max_depth=15, seed=0):
        self._ck(self.lib.tr_render_pt_rgb(self.h, int(frame_begin), int(n_frames), int(max_depth), int(seed)), "tr_render_pt_rgb")

    def render_pt_spec(self, frame_begin, n_frames, max_depth=10, seed=0):
        self._ck(self.lib.tr_render_pt_spec(self.h, int(frame_begin), int(n_frames), int(max_depth), int(seed)), "tr_render_pt_spec")

    def render_bdpt_rgb(self, frame_begin, n_frames, seed=0):
        self._ck(self.lib.tr_render_bdpt_rgb(self.h, int(frame_begin), int(n_frames), int(seed)), "tr_render_bdpt_rgb")

    def bdpt_kernel_ms(self):
        a, b = C.c_float(0), C.c_float(0)
        self._ck(self.lib.tr_bdpt_kernel_ms(self.h, C.byref(a), C.byref(b)), "tr_bdpt_kernel_ms")
        return float(a.value), float(b.value)

    def test_bdpt_dump(self, px, py):
        px = np.ascontiguousarray(px, np.int32); py = np.ascontiguousarray(py, np.int32); n = px.size
        verts = np.zeros((n, 13, 20), np.float32); depths = np.zeros((n, 2), np.int32); contrib = np.zeros((n, 7, 7, 4), np.float32)
        self._ck(self.lib.tr_test_bdpt_dump(self.h, n, _ptr(px), _ptr(py), _ptr(verts), _ptr(depths), _ptr(contrib)), "tr_test_bdpt_dump")
        return verts, depths, contrib

    # ---- spectral tables (PT_Spec)
    def spec_sensor_upload(self, xyz, lmin, lmax):
        a = _f(xyz, (-1, 3)); self._ck(self.lib.tr_spec_sensor_upload(self.h, _ptr(a), a.shape[0], float(lmin), float(lmax)), "tr_spec_sensor_upload")

    def spec_spectrum_upload(self, which, data, lmin, lmax):
        a = _f(data, (-1,)); self._ck(self.lib.tr_spec_spectrum_upload(self.h, int(which), _ptr(a), a.shape[0], float(lmin), float(lmax)), "tr_spec_spectrum_upload")

    def spec_spectrum_download(self, which, n):
        out = np.zeros(int(n), np.float32); self._ck(self.lib.tr_spec_spectrum_download(self.h, int(which), _ptr(out)), "tr_spec_spectrum_download"); return out

    def spec_rgb2spec_upload(self, scale, data, res):
        a, b = _f(scale, (-1,)), _f(data, (-1,))
        assert a.size == res and b.size == res ** 3 * 9
        self._ck(self.lib.tr_spec_rgb2spec_upload(self.h, _ptr(a), _ptr(b), int(res)), "tr_spec_rgb2spec_upload")

    def spec_sky_upload(self, configs, radiances, sun_dir):
        a, b, c = _f(configs, (-1,)), _f(radiances, (-1,)), _f(sun_dir, (-1,))
        assert a.size == 99 and b.size == 11 and c.size == 3
        self._ck(self.lib.tr_spec_sky_upload(self.h, _ptr(a), _ptr(b), _ptr(c)), "tr_spec_sky_upload")

    def spec_normalize(self, which):
        wp = np.zeros(3, np.float32); self._ck(self.lib.tr_spec_normalize(self.h, int(which), _ptr(wp)), "tr_spec_normalize"); return wp

    def test_srgb_to_spec(self, rgb, lambda0):
        a, b = _f(rgb, (-1, 3)), _f(lambda0, (-1,)); out = np.zeros((a.shape[0], 4), np.float32)
        self._ck(self.lib.tr_test_srgb_to_spec(self.h, a.shape[0], _ptr(a), _ptr(b), _ptr(out)), "hook"); return out

    def test_sky_radiance(self, theta, gamma, wl):
        a, b, c = _f(theta, (-1,)), _f(gamma, (-1,)), _f(wl, (-1,)); out = np.zeros(a.shape[0], np.float32)
        self._ck(self.lib.tr_test_sky_radiance(self.h, a.shape[0], _ptr(a), _ptr(b), _ptr(c), _ptr(out)), "hook"); return out

    def test_spectrum_sample(self, which, lam):
        a = _f(lam, (-1,)); out = np.zeros((a.shape[0], 3) if which < 0 else a.shape[0], np.float32)
        self._ck(self.lib.tr_test_spectrum_sample(self.h, int(which), a.shape[0], _ptr(a), _ptr(out)), "hook"); return out

    def render_debug(self):
        self._ck(self.lib.tr_render_debug(self.h), "tr_render_debug")

    def first_hit_download(self):
        n = self.W * self.H
        t = np.zeros(n, np.float32); prim = np.zeros(n, np.int32); uv = np.zeros((n, 2), np.float32)
        pos = np.zeros((n, 3), np.float32); gn = np.zeros((n, 3), np.float32); nr = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32)
        self._ck(self.lib.tr_first_hit_download(self.h, _ptr(t), _ptr(prim), _ptr(uv), _ptr(pos), _ptr(gn), _ptr(nr), _ptr(d)), "tr_first_hit_download")
        W, H = self.W, self.H
        return dict(t=t.reshape(W, H), prim=prim.reshape(W, H), uv=uv.reshape(W, H, 2), pos=pos.reshape(W, H, 3),
                    gnormal=gn.reshape(W, H, 3), normal=nr.reshape(W, H, 3), dir=d.reshape(W, H, 3))

    def tonemap(self, exposure):
        self._ck(self.lib.tr_tonemap(self.h, float(exposure)), "tr_tonemap")

    def stream_set(self, cuda_stream):
        """run on the caller's stream (e.g. torch.cuda.current_stream().cuda_stream); None restores"""
        self._ck(self.lib.tr_stream_set(self.h, cuda_stream), "tr_stream_set")

    def synchronize(self):
        self._ck(self.lib.tr_synchronize(self.h), "tr_synchronize")

    def stats(self):
        s = Stats(); self._ck(self.lib.tr_stats_get(self.h, C.byref(s)), "tr_stats_get")
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def set_option(self, name, value):
        self._ck(self.lib.tr_set_option(self.h, name.encode(), int(value)), "tr_set_option")

    # ---- unit hooks
    def test_disney_evaluate_pdf(self, N, V, L, metal, rough):
        N, V, L = _f(N, (-1, 3)), _f(V, (-1, 3)), _f(L, (-1, 3)); out = np.zeros((N.shape[0], 2), np.float32)
        self._ck(self.lib.tr_test_disney_evaluate_pdf(self.h, N.shape[0], _ptr(N), _ptr(V), _ptr(L), metal, rough, _ptr(out)), "hook"); return out

    def test_disney_sample(self, d, N, metal, rough, u):
        d, N, u = _f(d, (-1, 3)), _f(N, (-1, 3)), _f(u, (-1, 3)); out = np.zeros((d.shape[0], 3), np.float32)
        self._ck(self.lib.tr_test_disney_sample(self.h, d.shape[0], _ptr(d), _ptr(N), metal, rough, _ptr(u), _ptr(out)), "hook"); return out

    def test_glass_sample(self, d, N, ior, u):
        d, N, u = _f(d, (-1, 3)), _f(N, (-1, 3)), _f(u, (-1,)); out = np.zeros((d.shape[0], 4), np.float32)
        self._ck(self.lib.tr_test_glass_sample(self.h, d.shape[0], _ptr(d), _ptr(N), ior, _ptr(u), _ptr(out)), "hook"); return out

    def test_offset_ray(self, p, n):
        p, n = _f(p, (-1, 3)), _f(n, (-1, 3)); out = np.zeros_like(p)
        self._ck(self.lib.tr_test_offset_ray(self.h, p.shape[0], _ptr(p), _ptr(n), _ptr(out)), "hook"); return out

    def test_math(self, fn, a, b=None):
        a = _f(a, (-1,)); b = None if b is None else _f(b, (-1,)); out = np.zeros(a.shape[0], np.float32)
        self._ck(self.lib.tr_test_math(self.h, int(fn), a.shape[0], _ptr(a), _ptr(b), _ptr(out)), "tr_test_math"); return out

    def test_rng(self, seed, pixel, frame, block):
        out = np.zeros(4, np.float32); self._ck(self.lib.tr_test_rng(self.h, seed, pixel, frame, block, _ptr(out)), "hook"); return out

    def test_trace(self, o, d, shadow=False, kernel=0, target=None):
        """rays through the traversal code: kernel 0 the simple walk, 1 the production k_trace, 2 the production k_shadow
        (target = primitive every ray must see), 3 the production k_tail"""
        o, d = _f(o, (-1, 3)), _f(d, (-1, 3)); n = o.shape[0]
        tg = None if target is None else np.ascontiguousarray(target, np.int32)
        t = np.zeros(n, np.float32); prim = np.zeros(n, np.int32); uv = np.zeros((n, 2), np.float32)
        self._ck(self.lib.tr_test_trace_kernel(self.h, int(kernel), n, _ptr(o), _ptr(d), _ptr(tg), int(shadow), _ptr(t), _ptr(prim), _ptr(uv)),
                 "tr_test_trace_kernel")
        return t, prim, uv


_ctx = None


def context():
    """The process-wide context (created on first use or by ti.init)."""
    global _ctx
    if _ctx is None:
        _ctx = Context()
    return _ctx


def reset_context(device=None):
    """ti.init() semantics: drop all device state and start a fresh program."""
    global _ctx
    if _ctx is not None:
        _ctx.close()
    _ctx = Context(device)
    return _ctx


class Field:
    """Stand-in for a Taichi field on the reference's API surface: .to_numpy() / .from_numpy()."""

    def __init__(self, getter, setter=None):
        self._get, self._set = getter, setter

    def to_numpy(self):
        return self._get()

    def from_numpy(self, a):
        if self._set is None:
            raise RuntimeError("this field is read-only on the host")
        self._set(a)
