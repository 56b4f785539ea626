"""Record layouts and host-side scene records (mirror of the reference's SceneData.py).

Same names, constants and setters as /root/reference/SceneData.py:33-174 (including its spelling of
SHPAE_*, alebdoTex and setExtinciton, which the example scripts use), so scene scripts written for
the reference run unchanged.  The tables these records are packed into are what the C-ABI
tr_scene_upload() takes (include/tiray.h).
"""
import numpy as np

# floats / ints per record (SceneData.py:33-38)
MAT_VEC_SIZE, VER_VEC_SIZE, PRI_VEC_SIZE, SHA_VEC_SIZE, NOD_VEC_SIZE, CPNOD_VEC_SIZE = 10, 9, 3, 10, 11, 9

SHPAE_NONE, SHPAE_SPHERE, SHPAE_QUAD, SHPAE_SPOT, SHPAE_LASER = 0, 1, 2, 3, 4
PRIMITIVE_NONE, PRIMITIVE_TRI, PRIMITIVE_SHAPE = 0, 1, 2
MAT_DISNEY, MAT_GLASS, MAT_LIGHT, MAT_SPECTRAL = 0.0, 1.0, 2.0, 10.0
IS_LEAF = 1


class Material:
    """row: type, albedo texture, rgb, param[5] (param0 = metal | ior, param1 = rough | extinction)"""

    def __init__(self):
        self.type, self.alebdoTex = 0, 0
        self.color = [0.0, 0.0, 0.0]
        self.param = [0.0] * 5

    def setColor(self, color): self.color = color
    def setMetal(self, metal): self.param[0] = metal
    def setRough(self, rough): self.param[1] = rough
    def setIor(self, ior): self.param[0] = ior
    def setExtinciton(self, extinction): self.param[1] = extinction

    def fillStruct(self, np_data, index):
        np_data[index, 0:2] = (float(self.type), float(self.alebdoTex))
        np_data[index, 2:5] = self.color[0:3]
        np_data[index, 5:MAT_VEC_SIZE] = self.param


class Shape:
    """row: type, pos3, param[6]"""

    def __init__(self):
        self.type = 0
        self.pos = [0.0, 0.0, 0.0]
        self.param = [0.0] * 6

    def setRadius(self, radius): self.param[0] = radius
    def getRadius(self): return self.param[0]
    def setXita(self, xita1, xita2): self.param[0:2] = [xita1, xita2]
    def setScale(self, scale): self.param[2] = scale
    def setV1(self, V1): self.param[0:3] = [V1[0], V1[1], V1[2]]
    def setV2(self, V2): self.param[3:6] = [V2[0], V2[1], V2[2]]
    def setNormal(self, normal): self.param[3:6] = [normal[0], normal[1], normal[2]]

    def fillStruct(self, np_data, index):
        np_data[index, 0] = float(self.type)
        np_data[index, 1:4] = self.pos[0:3]
        np_data[index, 4:SHA_VEC_SIZE] = self.param


class Vertex:
    """row: pos3, normal3, tex3"""

    def __init__(self):
        self.pos, self.normal, self.tex = [0.0] * 3, [0.0] * 3, [0.0] * 3

    def setPos(self, buf, offset): self.pos = [buf[offset], buf[offset + 1], buf[offset + 2]]
    def setNormal(self, buf, offset): self.normal = [buf[offset], buf[offset + 1], buf[offset + 2]]
    def setTex(self, buf, offset): self.tex = [buf[offset], buf[offset + 1], 0.0]
    def setTex3(self, buf, offset): self.tex = [buf[offset], buf[offset + 1], buf[offset + 2]]

    def fillStruct(self, np_data, index):
        np_data[index, 0:3] = self.pos
        np_data[index, 3:6] = self.normal
        np_data[index, 6:9] = self.tex


class Primitive:
    """row: type (1 triangle / 2 shape), first vertex | shape index, material index"""

    def __init__(self):
        self.type = self.vertex_shape_index = self.mat_index = 0

    def fillStruct(self, np_data, index):
        np_data[index, :] = (self.type, self.vertex_shape_index, self.mat_index)
