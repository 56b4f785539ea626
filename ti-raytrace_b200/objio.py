"""Wavefront OBJ/MTL ingest for Scene.add_obj, array based.

Produces what the reference obtains from PyWavefront 1.3.3 (`scene.materials[name].vertices`, one
de-indexed vertex soup per material, /root/reference/Scene.py:66-141) as (k, 9) f64 arrays: materials in
`newmtl` order, faces in file order inside each material, polygons fanned as (v1,v2,v3), (vj,v1,vj-1),
negative indices relative to the current element count, faces without `usemtl` in an implicit material
"default<k>".  The file is parsed by the native reader in libtiray.so (csrc/objparse.cpp); the reference spends
seconds in per-vertex Python loops on the 130 k-triangle scene, this takes tens of milliseconds.
"""
import ctypes as C
import os
import numpy as np
import _native


class ObjMaterial:
    def __init__(self, name):
        self.name = name
        self.diffuse = [0.8, 0.8, 0.8]
        self.emissive = [0.0, 0.0, 0.0]
        self.transparency = 1.0      # `d`
        self.shininess = 0.0         # `Ns`
        self.optical_density = 1.0   # `Ni`
        self.texture = None
        self.has_vt = self.has_vn = False
        self.rows = np.zeros((0, 9), np.float64)     # de-indexed vertices: pos3, normal3, tex3 (three rows per triangle)


def read_obj(path):
    """-> materials in PyWavefront order, each with its (k, 9) f64 vertex rows; parsed by the native reader
    (csrc/objparse.cpp, tr_obj_*).  Raises ValueError with the parser's message (file:line) on malformed input."""
    lib = _native.load_library()
    h = C.c_void_p()
    if lib.tr_obj_open(os.fsencode(path), C.byref(h)) != 0:
        msg = (lib.tr_obj_last_error() or b"").decode()
        raise (FileNotFoundError if msg.startswith("cannot open") else ValueError)(msg)
    try:
        out = []
        for k in range(lib.tr_obj_material_count(h)):
            name = C.create_string_buffer(256); props = np.zeros(9, np.float64)
            nv, has_vt, has_vn = C.c_int64(0), C.c_int(0), C.c_int(0)
            lib.tr_obj_material(h, k, name, 256, props.ctypes.data, C.byref(nv), C.byref(has_vt), C.byref(has_vn))
            m = ObjMaterial(name.value.decode())
            m.diffuse, m.emissive = [float(x) for x in props[0:3]], [float(x) for x in props[3:6]]
            m.transparency, m.shininess, m.optical_density = float(props[6]), float(props[7]), float(props[8])
            m.has_vt, m.has_vn = bool(has_vt.value), bool(has_vn.value)
            m.rows = np.zeros((nv.value, 9), np.float64)
            if nv.value:
                lib.tr_obj_material_vertices(h, k, m.rows.ctypes.data)
            out.append(m)
        return out
    finally:
        lib.tr_obj_close(h)


def flat_normals(rows):
    """Scene.cal_normal (Scene.py:169-179): triangles whose first vertex has a zero normal get the
    normalised (v2-v1)x(v3-v1) on all three vertices; f64 arithmetic like the reference's Python."""
    tri = rows.reshape(-1, 3, 9)
    n0 = tri[:, 0, 3:6]
    zero = np.sqrt(n0[:, 0] * n0[:, 0] + n0[:, 1] * n0[:, 1] + n0[:, 2] * n0[:, 2]) == 0.0
    if zero.any():
        a = tri[zero, 1, 0:3] - tri[zero, 0, 0:3]
        b = tri[zero, 2, 0:3] - tri[zero, 0, 0:3]
        n = np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                      a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1)
        ln = np.sqrt(n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1] + n[:, 2] * n[:, 2])
        # degenerate triangles keep their zero normal (the reference divides by zero there, Scene.py:176)
        inv = np.divide(1.0, ln, out=np.zeros_like(ln), where=ln != 0.0)
        n = n * inv[:, None]
        tri[zero, 0, 3:6] = n; tri[zero, 1, 3:6] = n; tri[zero, 2, 3:6] = n
    return tri.reshape(-1, 9)
