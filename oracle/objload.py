"""ORACLE (test infrastructure, not product code): OBJ/MTL ingest restated from the reference.

Restates, in plain Python, what `Scene.add_obj` + `Scene.cal_normal` + `Scene.setup_data_cpu`
(/root/reference/Scene.py:59-141,169-179,223-273) produce for a Wavefront OBJ, including the
behaviour of the un-vendored third-party parser they call: PyWavefront==1.3.3
(/root/reference/requirements.txt, UTF-16 line "PyWavefront==1.3.3").

PyWavefront 1.3.3 published behaviour restated here (pywavefront/obj.py, material.py):
  * materials are created in `newmtl` order of the MTL named by `mtllib`; faces seen before any
    `usemtl` go to an implicit material "default<k>" (Kd .8 .8 .8, d=1, Ke 0, Ns 0, Ni 1);
    `usemtl` of an unknown name creates that material with the defaults (non-strict mode)
  * every material owns one interleaved float list; per vertex the order is T2F, (C3F), N3F, V3F
    and the format is fixed by the first vertex of the first face of the material
  * polygon (v1..vn): (v1,v2,v3) then for each j>3 (vj, v1, v(j-1))
  * negative indices are relative to the current end of the respective array
  * `vt` keeps two components; `d` sets transparency, `Tr` sets 1-Tr, `Ni` optical_density,
    `Ns` shininess, `Ke` emissive, `Kd` diffuse
Pinned by: tests/golden/nodelist.txt (leaf prim ids + boxes depend on material/face order).
"""
import os
import numpy as np

MAT_DISNEY, MAT_GLASS, MAT_LIGHT = 0.0, 1.0, 2.0
PRIMITIVE_TRI, PRIMITIVE_SHAPE = 1, 2
INF_VALUE = 1000000.0


class _Mat:
    def __init__(self, name):
        self.name = name
        self.diffuse = [.8, .8, .8]
        self.emissive = [0., 0., 0.]
        self.transparency = 1.0
        self.shininess = 0.0
        self.optical_density = 1.0
        self.vertex_format = None
        self.vertices = []


def _parse_mtl(path, materials):
    cur = None
    with open(path, "r") as f:
        for line in f:
            v = line.split()
            if not v or v[0].startswith("#"):
                continue
            k = v[0]
            if k == "newmtl":
                cur = _Mat(v[1]); materials[cur.name] = cur
            elif cur is None:
                continue
            elif k == "Kd":
                cur.diffuse = [float(x) for x in v[1:4]]
            elif k == "Ke":
                cur.emissive = [float(x) for x in v[1:4]]
            elif k == "d":
                cur.transparency = float(v[1])
            elif k == "Tr":
                cur.transparency = 1.0 - float(v[1])
            elif k == "Ns":
                cur.shininess = float(v[1])
            elif k == "Ni":
                cur.optical_density = float(v[1])


def parse_obj(path):
    """-> ordered list of _Mat with interleaved .vertices like pywavefront's scene.materials"""
    materials = {}
    pos, nor, tex = [], [], []
    cur = None
    with open(path, "r") as f:
        for line in f:
            v = line.split()
            if not v or v[0].startswith("#"):
                continue
            k = v[0]
            if k == "v":
                pos.append([float(x) for x in v[1:]])
            elif k == "vn":
                nor.append([float(x) for x in v[1:4]])
            elif k == "vt":
                tex.append([float(v[1]), float(v[2])])
            elif k == "mtllib":
                _parse_mtl(os.path.join(os.path.dirname(path), v[1]), materials)
            elif k == "usemtl":
                name = v[1] if len(v) > 1 else ""
                cur = materials.get(name)
                if cur is None:
                    cur = _Mat(name); materials[name] = cur
            elif k == "f":
                if cur is None:
                    cur = _Mat("default%d" % len(materials)); materials[cur.name] = cur
                parts = v[1].split("/")
                has_vt = (len(parts) == 2) or (len(parts) == 3 and parts[1] != "")
                has_vn = len(parts) == 3
                vi0 = int(parts[0]); vi0 = vi0 + len(pos) if vi0 < 0 else vi0 - 1
                has_c = len(pos[vi0]) == 6
                fmt = "_".join(n for n, on in (("T2F", has_vt), ("C3F", has_c), ("N3F", has_vn), ("V3F", True)) if on)
                if cur.vertex_format and cur.vertex_format != fmt:
                    raise ValueError("inconsistent vertex format in %s" % path)
                cur.vertex_format = fmt

                def emit(tok):
                    p = tok.split("/")
                    out = []
                    if has_vt:
                        ti = int(p[1]); ti = ti + len(tex) if ti < 0 else ti - 1
                        out += tex[ti]
                    pi = int(p[0]); pi = pi + len(pos) if pi < 0 else pi - 1
                    if has_c:
                        out += pos[pi][3:6]
                    if has_vn:
                        ni = int(p[2]); ni = ni + len(nor) if ni < 0 else ni - 1
                        out += nor[ni]
                    out += pos[pi][0:3]
                    return out
                toks = v[1:]
                first = prev = None
                for i, tok in enumerate(toks):
                    e = emit(tok)
                    cur.vertices += e
                    if i >= 3:
                        cur.vertices += first
                        cur.vertices += prev
                    if i == 0:
                        first = e
                    prev = e
    return list(materials.values())


class Tables:
    """numpy tables exactly as Scene.setup_data_cpu packs them (Scene.py:225-273)."""
    pass


def load_scene(paths, shapes=(), material_edit=None):
    """paths: OBJ files in add_obj order.  shapes: list of (shape_row10, material_row10) appended
    with add_shape semantics (Scene.py:188-205).  material_edit(mats_rows) may mutate the material
    rows between add_obj and setup_data_cpu (example/single_model.py:27-29)."""
    mats, verts, prims, lights = [], [], [], []
    bmax = np.full((1, 3), -INF_VALUE, np.float32); bmin = np.full((1, 3), INF_VALUE, np.float32)
    for path in paths:
        for m in parse_obj(path):
            row = [0.0] * 10
            if m.emissive[0] > 1.0 and m.emissive[1] > 1.0 and m.emissive[2] > 1.0:
                row[0] = MAT_LIGHT; row[2:5] = m.emissive[0:3]
            elif m.transparency > 0.99:
                row[0] = MAT_DISNEY; row[5] = 0.0; row[6] = 0.5; row[2:5] = m.diffuse[0:3]
            else:
                row[0] = MAT_GLASS; row[5] = m.optical_density; row[6] = m.shininess; row[2:5] = m.diffuse[0:3]
            row[1] = -1.0
            mat_index = len(mats); mats.append(row)
            fmt = m.vertex_format
            if not m.vertices:
                continue                                 # a material no face uses: num_vert == 0, only the material row exists (Scene.py:93-97,141)
            stride = {"T2F_V3F": 5, "T2F_N3F_V3F": 8, "N3F_V3F": 6, "V3F": 3}[fmt]
            buf = m.vertices
            for k in range(0, len(buf), stride):
                p = [0.0] * 9
                if fmt == "T2F_V3F":
                    p[0:3] = buf[k + 2:k + 5]; p[6:8] = buf[k:k + 2]
                elif fmt == "T2F_N3F_V3F":
                    p[0:3] = buf[k + 5:k + 8]; p[3:6] = buf[k + 2:k + 5]; p[6:8] = buf[k:k + 2]
                elif fmt == "N3F_V3F":
                    p[0:3] = buf[k + 3:k + 6]; p[3:6] = buf[k:k + 3]
                else:
                    p[0:3] = buf[k:k + 3]
                for a in range(3):
                    bmax[0, a] = max(p[a], bmax[0, a]); bmin[0, a] = min(p[a], bmin[0, a])
                verts.append(p)
                if len(verts) % 3 == 0:
                    if row[0] == MAT_LIGHT:
                        lights.append(len(prims))
                    prims.append([PRIMITIVE_TRI, len(verts) - 3, mat_index])
    shape_rows = []
    for srow, mrow in shapes:
        if mrow[0] == MAT_LIGHT:
            lights.append(len(prims))
        prims.append([PRIMITIVE_SHAPE, len(shape_rows), len(mats)])
        shape_rows.append(list(srow)); mats.append(list(mrow))
    if material_edit is not None:
        material_edit(mats)
    # cal_normal (Scene.py:169-179): f64 python arithmetic, flat normal where the OBJ normal is zero
    for i in range(0, len(verts), 3):
        n0 = verts[i][3:6]
        if (n0[0] * n0[0] + n0[1] * n0[1] + n0[2] * n0[2]) ** 0.5 == 0.0:
            a = [verts[i + 1][k] - verts[i][k] for k in range(3)]
            b = [verts[i + 2][k] - verts[i][k] for k in range(3)]
            n = [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
            ln = (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) ** 0.5
            if ln == 0.0:
                # degenerate triangle (mc.obj has some): the reference raises ZeroDivisionError here
                # (Scene.py:176); both this oracle and the product keep the zero normal instead.
                continue
            l = 1.0 / ln
            n = [n[0] * l, n[1] * l, n[2] * l]
            for j in range(3):
                verts[i + j][3:6] = n
    t = Tables()
    t.material = np.asarray(mats, np.float32).reshape(-1, 10)
    t.vertex = np.asarray(verts, np.float32).reshape(-1, 9)
    t.primitive = np.asarray(prims, np.int32).reshape(-1, 3)
    t.shape = np.asarray(shape_rows, np.float32).reshape(-1, 10) if shape_rows else np.zeros((0, 10), np.float32)
    t.light = np.asarray(lights, np.int32)
    t.bmin, t.bmax = bmin, bmax
    return t


def sphere_light_rows(pos=(0.0, 20.0, 0.0), radius=5.0, color=(50.0, 50.0, 50.0)):
    """example/Example.py:27-36 add_sphere_light"""
    s = [1.0, pos[0], pos[1], pos[2], radius, 0, 0, 0, 0, 0]
    m = [MAT_LIGHT, 0.0, color[0], color[1], color[2], 0, 0, 0, 0, 0]
    return s, m
