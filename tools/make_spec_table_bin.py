#!/usr/bin/env python
"""Convert the reference's text rgb2spec table (spectrum/spec_table: res, res scale values, then 3*res^3 coefficient triples,
9 numbers per line; 31 MB) into the compact binary Rgb2Spec.load_table prefers: int32 res, f32 scale[res], f32 data[9*res^3].
The values are float(text) rounded to f32, exactly what the reference's load_table stores (spectrum/Rgb2Spec.py:15-36).

  python tools/make_spec_table_bin.py /root/reference/spectrum/spec_table ti-raytrace_b200/spectrum/spec_table.f32
"""
import sys
import numpy as np


def convert(src, dst):
    with open(src, "r") as f:
        res = int(f.readline())
        scale = np.asarray([float(f.readline()) for _ in range(res)], np.float32)
        data = np.asarray(f.read().split(), np.float64).astype(np.float32)
    assert data.size == res ** 3 * 9, (data.size, res)
    with open(dst, "wb") as f:
        np.asarray([res], np.int32).tofile(f); scale.tofile(f); data.tofile(f)
    return res, scale, data


if __name__ == "__main__":
    r, s, d = convert(sys.argv[1], sys.argv[2])
    print("res %d, %d coefficients -> %s" % (r, d.size, sys.argv[2]))
