#!/usr/bin/env python
"""Developer stress loop: many small spectral renders (the configuration that hands every path to k_tail at depth 1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa
import numpy as np
import _native
from test_gpu_spectral import build_gpu_spectral
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
_native.reset_context()
scene, cam, integ = build_gpu_spectral(128, 128)
ctx = _native.context()
for k, v in [a.split("=") for a in sys.argv[2:]]:
    ctx.set_option(k, int(v))
ref = None
for i in range(n):
    ctx.film_clear(); cam.frame = 0; cam.frame_cpu[0] = 0
    try:
        for _ in range(2):
            integ.render(); cam.update_frame()
        img = integ.hdr.to_numpy()
    except RuntimeError as e:
        print("iteration %d: %s" % (i, e), flush=True); sys.exit(1)
    if ref is None:
        ref = img
    elif not np.array_equal(img, ref, equal_nan=True):
        print("iteration %d: film differs from iteration 0 in %d words" % (i, int((img != ref).sum())), flush=True); sys.exit(2)
print("stress ok: %d iterations identical" % n)
