"""Bvh — host handle of the device LBVH (mirror of /root/reference/accel/LBvh.py:14-226).

setup_data_gpu() runs the whole build on the GPU through tr_bvh_build(): Morton-3D, stable radix
sort, Karras topology with the reference's duplicate-key rule, atomic AABB refit and pre-order
flatten (csrc/bvh_build.cu).  bvh_node / compact_node / morton_code_s expose the reference layouts
through .to_numpy(); write_nodelist() reproduces the nodelist.txt dump of LBvh.py:164-173 on request
(the reference rewrites that file in the CWD on every build; here it is opt-in via
TIRAY_WRITE_NODELIST=1)."""
import os
import numpy as np
import _native

HIT_TRI, HIT_SHA = 0.0, 1.0


class Bvh:
    def __init__(self, primitive_count, min_boundary, max_boundary):
        self.primitive_count = primitive_count
        self.minboundarynp, self.maxboundarynp = min_boundary, max_boundary
        self.leaf_node_count = 0
        self._views = None
        self.morton_code_s = _native.Field(lambda: self._download()[0])
        self.bvh_node = _native.Field(lambda: self._download()[1])
        self.compact_node = _native.Field(lambda: self._download()[2])

    @staticmethod
    def get_pot_num(num):
        m = 1
        while m < num:
            m <<= 1
        return m >> 1

    @staticmethod
    def get_pot_bit(num):
        return max(int(num) - 1, 0).bit_length()

    def setup_data_cpu(self):
        self.node_count = self.primitive_count * 2 - 1
        self.primitive_pot = self.get_pot_num(self.primitive_count) << 1
        self.primitive_bit = self.get_pot_bit(self.primitive_pot)

    def setup_data_gpu(self, vertex=None, shape=None, primitive=None):
        """The scene tables are already resident (Scene.setup_data_gpu uploaded them); the arguments
        exist for signature parity with LBvh.py:192."""
        ctx = _native.context()
        ctx.bvh_build()
        self._views = None
        self.build_ms = ctx.stats()["ms_build"]
        if os.environ.get("TIRAY_WRITE_NODELIST") == "1":
            self.write_nodelist("nodelist.txt")

    def _download(self):
        if self._views is None:
            self._views = _native.context().bvh_download()
        return self._views

    def nodelist_lines(self):
        """print_compact_info formatting (LBvh.py:127-136; its column labels are shifted by one)"""
        c = self._download()[2]
        return ["node:%d pri:%d offset:%d min:%.2f %.2f %.2f max:%.2f %.2f %.2f" %
                (i, int(r[1]), int(r[2]), r[3], r[4], r[5], r[6], r[7], r[8]) for i, r in enumerate(c)]

    def write_nodelist(self, path):
        with open(path, "w") as fo:
            for line in self.nodelist_lines():
                print(line, file=fo)
