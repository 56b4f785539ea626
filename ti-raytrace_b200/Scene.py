"""Scene — host side of the reference's Scene class over the libtiray C-ABI.

Same public surface as /root/reference/Scene.py:22-311 (add_obj, add_env, add_shape, mutable
material_cpu records, setup_data_cpu, setup_data_gpu, process_normal, total_area, light_area,
max/minboundarynp, env_power, bvh).  The device functions of the reference class (closet_hit,
intersect_tri, sample_li ... Scene.py:315-799) are CUDA code in csrc/trace.cuh and csrc/wavefront.cu;
the tables packed here are byte-for-byte what Taichi received through from_numpy (Scene.py:225-273).
"""
import sys
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
for _sub in ("accel", "texture"):
    _p = os.path.join(_HERE, _sub)
    if _p not in sys.path:
        sys.path.append(_p)

import SceneData as SCD
import UtilsFunc as UF
import LBvh
import Texture as TX
import objio
import _paths
import _native

MAX_STACK_SIZE = 32


class Scene:
    def __init__(self):
        self.maxboundarynp = np.full((1, 3), -UF.INF_VALUE, np.float32)
        self.minboundarynp = np.full((1, 3), UF.INF_VALUE, np.float32)
        self.light_cpu, self.material_cpu, self.shape_cpu = [], [], []
        self._vertex_blocks, self._prim_blocks = [], []      # array blocks instead of per-vertex objects
        self.material_count = self.vertex_count = self.primitive_count = 0
        self.shape_count = self.light_count = 0
        self.env = TX.Texture()
        self.env_power = 0.0
        self.bvh = None
        self._area = None
        self.light_area = _native.Field(lambda: np.array([self._area if self._area is not None else 0.0], np.float32))
        self.vertex = _native.Field(lambda: _native.context().vertex_download())

    # ------------------------------------------------------------------ ingest (Scene.py:59-141)
    def add_obj(self, filename):
        for m in objio.read_obj(_paths.resolve(filename)):
            rec = SCD.Material()
            if m.emissive[0] > 1.0 and m.emissive[1] > 1.0 and m.emissive[2] > 1.0:
                rec.type = SCD.MAT_LIGHT
                rec.setColor(m.emissive)
            elif m.transparency > 0.99:
                rec.type = SCD.MAT_DISNEY
                rec.setMetal(0.0); rec.setRough(0.5); rec.setColor(m.diffuse)
            else:
                rec.type = SCD.MAT_GLASS
                rec.setIor(m.optical_density); rec.setExtinciton(m.shininess); rec.setColor(m.diffuse)
            rec.alebdoTex = -1 if m.texture is None else float(m.texture)
            self.material_cpu.append(rec)

            rows = m.rows
            ntri = rows.shape[0] // 3
            if ntri:
                self.maxboundarynp[0, :] = np.maximum(self.maxboundarynp[0, :], rows[:, 0:3].max(axis=0).astype(np.float32))
                self.minboundarynp[0, :] = np.minimum(self.minboundarynp[0, :], rows[:, 0:3].min(axis=0).astype(np.float32))
                prim = np.empty((ntri, 3), np.int32)
                prim[:, 0] = SCD.PRIMITIVE_TRI
                prim[:, 1] = self.vertex_count + 3 * np.arange(ntri, dtype=np.int32)
                prim[:, 2] = self.material_count
                if rec.type == SCD.MAT_LIGHT:
                    self.light_cpu.extend(range(self.primitive_count, self.primitive_count + ntri))
                    self.light_count += ntri
                self._vertex_blocks.append(rows); self._prim_blocks.append(prim)
                self.vertex_count += 3 * ntri
                self.primitive_count += ntri
            self.material_count += 1

    def add_env(self, filename, env_power):
        # the reference ignores `filename` and always loads image/env.png (Scene.py:183-185)
        self.env.load_image("image/env.png")
        self.env_power = env_power

    def add_shape(self, shape, mat):
        if mat.type == SCD.MAT_LIGHT:
            self.light_cpu.append(self.primitive_count)
            self.light_count += 1
        self._prim_blocks.append(np.array([[SCD.PRIMITIVE_SHAPE, self.shape_count, self.material_count]], np.int32))
        self.primitive_count += 1
        self.shape_cpu.append(shape); self.shape_count += 1
        self.material_cpu.append(mat); self.material_count += 1

    # ------------------------------------------------------------------ packing (Scene.py:223-296)
    def setup_data_cpu(self):
        self.material_np = np.zeros((self.material_count, SCD.MAT_VEC_SIZE), np.float32)
        for i, m in enumerate(self.material_cpu):
            m.fillStruct(self.material_np, i)
        rows = np.concatenate(self._vertex_blocks, axis=0) if self._vertex_blocks else np.zeros((0, 9))
        rows = objio.flat_normals(np.ascontiguousarray(rows, np.float64))
        self.vertex_np = rows.astype(np.float32)
        self.smooth_normal_np = np.zeros((self.vertex_count, 3), np.float32)
        self.primitive_np = (np.concatenate(self._prim_blocks, axis=0) if self._prim_blocks
                             else np.zeros((0, 3), np.int32)).astype(np.int32)
        self.vertex_index_np = np.repeat(np.nonzero(self.primitive_np[:, 0] == SCD.PRIMITIVE_TRI)[0].astype(np.int32), 3)
        self.light_np = np.asarray(self.light_cpu, np.int32) if self.light_count > 0 else np.zeros(1, np.float32)
        if self.shape_count > 0:
            self.shape_np = np.zeros((self.shape_count, SCD.SHA_VEC_SIZE), np.float32)
            for i, s in enumerate(self.shape_cpu):
                s.fillStruct(self.shape_np, i)
        else:
            self.shape_np = np.zeros((1, SCD.SHA_VEC_SIZE), np.float32)
        # the big tables are re-uploaded by every setup_data_gpu(): page-lock them once so the upload is a direct DMA
        self._pins = [_native.pin_array(self.vertex_np), _native.pin_array(self.primitive_np)]
        self.bvh = LBvh.Bvh(self.primitive_count, self.minboundarynp, self.maxboundarynp)
        self.bvh.setup_data_cpu()
        if self.env_power == 0.0:
            self.env.load_image("image/black.png")

    # ------------------------------------------------------------------ upload + build (Scene.py:299-310)
    def setup_data_gpu(self):
        ctx = _native.context()
        ctx.scene_upload(self.vertex_np, self.primitive_np, self.material_np,
                         self.shape_np if self.shape_count > 0 else None,
                         self.light_np if self.light_count > 0 else None, self.minboundarynp, self.maxboundarynp)
        self.env.setup_data_gpu(self.env_power)
        self._env_power_uploaded = self.env_power
        self.bvh.setup_data_gpu(None, None, None)

    def _sync_late_scalars(self):
        """env_power may change until the first render() (example/sky_dome.py:28; Taichi bakes Python
        scalars at first launch). Called by the integrators before they launch."""
        if getattr(self, "_env_power_uploaded", None) != self.env_power and self.env.np_img is not None:
            self.env.setup_data_gpu(self.env_power)
            self._env_power_uploaded = self.env_power

    # ------------------------------------------------------------------ kernels
    def total_area(self):
        self._area = _native.context().total_area()

    def process_normal(self):
        _native.context().process_normal()
