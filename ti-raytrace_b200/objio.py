"""Wavefront OBJ/MTL ingest for Scene.add_obj, array based.

Produces what the reference obtains from PyWavefront 1.3.3 (`scene.materials[name].vertices`, one
de-indexed vertex soup per material, /root/reference/Scene.py:66-141) but as index arrays that are
gathered with numpy instead of per-vertex Python objects: materials in `newmtl` order, faces in file
order inside each material, polygons fanned as (v1,v2,v3), (vj,v1,vj-1), negative indices relative to
the current element count, faces without `usemtl` in an implicit material "default<k>".
"""
import os
import numpy as np


class ObjMaterial:
    def __init__(self, name):
        self.name = name
        self.diffuse = [0.8, 0.8, 0.8]
        self.emissive = [0.0, 0.0, 0.0]
        self.transparency = 1.0      # `d`
        self.shininess = 0.0         # `Ns`
        self.optical_density = 1.0   # `Ni`
        self.texture = None
        self.has_vt = self.has_vn = None
        self.corners = []            # flat list of (v, vt, vn) index triples, three per triangle


def _read_mtl(path, table):
    cur = None
    with open(path, "r") as fh:
        for raw in fh:
            tok = raw.split()
            if not tok or tok[0][0] == "#":
                continue
            key = tok[0]
            if key == "newmtl":
                cur = table.setdefault(tok[1], ObjMaterial(tok[1]))
            elif cur is not None:
                if key == "Kd": cur.diffuse = [float(t) for t in tok[1:4]]
                elif key == "Ke": cur.emissive = [float(t) for t in tok[1:4]]
                elif key == "d": cur.transparency = float(tok[1])
                elif key == "Tr": cur.transparency = 1.0 - float(tok[1])
                elif key == "Ns": cur.shininess = float(tok[1])
                elif key == "Ni": cur.optical_density = float(tok[1])


def read_obj(path):
    """-> (materials in order, positions (n,3) f64, normals (n,3) f64, texcoords (n,2) f64)"""
    table = {}
    pos, nor, tex = [], [], []
    cur = None
    with open(path, "r") as fh:
        for raw in fh:
            tok = raw.split()
            if not tok:
                continue
            key = tok[0]
            if key == "v":
                pos.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif key == "vn":
                nor.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif key == "vt":
                tex.append((float(tok[1]), float(tok[2])))
            elif key == "f":
                if cur is None:
                    cur = ObjMaterial("default%d" % len(table)); table[cur.name] = cur
                first = tok[1].split("/")
                has_vt = len(first) == 2 or (len(first) == 3 and first[1] != "")
                has_vn = len(first) == 3
                if cur.has_vt is None:
                    cur.has_vt, cur.has_vn = has_vt, has_vn
                elif (cur.has_vt, cur.has_vn) != (has_vt, has_vn):
                    raise ValueError("%s: material %s mixes vertex formats" % (path, cur.name))
                np_, nt_, nn_ = len(pos), len(tex), len(nor)
                idx = []
                for t in tok[1:]:
                    p = t.split("/")
                    a = int(p[0]); a = a + np_ if a < 0 else a - 1
                    b = c = 0
                    if has_vt:
                        b = int(p[1]); b = b + nt_ if b < 0 else b - 1
                    if has_vn:
                        c = int(p[2]); c = c + nn_ if c < 0 else c - 1
                    idx.append((a, b, c))
                out = cur.corners
                out += (idx[0], idx[1], idx[2])
                for j in range(3, len(idx)):
                    out += (idx[j], idx[0], idx[j - 1])
            elif key == "usemtl":
                name = tok[1] if len(tok) > 1 else ""
                cur = table.get(name)
                if cur is None:
                    cur = ObjMaterial(name); table[name] = cur
            elif key == "mtllib":
                _read_mtl(os.path.join(os.path.dirname(path), tok[1]), table)
    P = np.asarray(pos, np.float64).reshape(-1, 3)
    N = np.asarray(nor, np.float64).reshape(-1, 3)
    T = np.asarray(tex, np.float64).reshape(-1, 2)
    return list(table.values()), P, N, T


def material_vertices(mat, P, N, T):
    """(k,9) f64 rows pos3, normal3, tex3 for one material (k = 3 x triangles)"""
    c = np.asarray(mat.corners, np.int64).reshape(-1, 3)
    rows = np.zeros((c.shape[0], 9), np.float64)
    if c.shape[0]:
        rows[:, 0:3] = P[c[:, 0]]
        if mat.has_vn:
            rows[:, 3:6] = N[c[:, 2]]
        if mat.has_vt:
            rows[:, 6:8] = T[c[:, 1]]
    return rows


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
