"""Rgb2Spec — Jakob-Hanika RGB -> smooth spectrum coefficient table (mirror of
/root/reference/spectrum/Rgb2Spec.py:6-42): load_table(path), setup_data_gpu().

fetch / eval (:44-138) run on the device (csrc/spectral.cuh: rs_fetch, rs_eval).  load_table reads the reference's
text format (res, res scale values, then 3*res^3 coefficient triples, 9 numbers per line; 31 MB); the package ships
the same numbers as a sibling `<path>.f32` file (int32 res, f32 scale[res], f32 data[9*res^3],
written by tools/make_spec_table_bin.py from the reference's text), used when the text file is absent."""
import os
import numpy as np
import _native
import _paths

RGB2SPEC_N_COEFFS = 3


class Rgb2Spec:
    def __init__(self):
        self.table_res = 0
        self.table_size = 0
        self.table_scale_np = None
        self.table_data_np = None

    def load_table(self, table_path):
        path = _paths.resolve(table_path)
        if not os.path.exists(path):
            alt = _paths.resolve(table_path + ".f32")
            if os.path.exists(alt):
                path = alt
        if path.endswith(".f32"):
            raw = np.fromfile(path, np.float32)
            res = int(raw[:1].view(np.int32)[0])
            scale, data = raw[1:1 + res].copy(), raw[1 + res:].copy()
        else:
            with open(path, "r") as f:
                res = int(f.readline())
                scale = np.asarray([float(f.readline()) for _ in range(res)], np.float32)
                data = np.asarray(f.read().split(), np.float64).astype(np.float32)
        if data.size != res ** 3 * 9:
            raise ValueError("%s: expected %d coefficients, found %d" % (path, res ** 3 * 9, data.size))
        self.table_res, self.table_size = res, data.size
        self.dx, self.dy, self.dz = 3, 3 * res, 3 * res * res
        self.table_scale_np, self.table_data_np = scale, data

    def setup_data_gpu(self):
        _native.context().spec_rgb2spec_upload(self.table_scale_np, self.table_data_np, self.table_res)
