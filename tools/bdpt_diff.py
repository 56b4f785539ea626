"""debug aid: per-strategy differences between the CUDA BDPT and the oracle (run on a GPU box)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest
import numpy as np
os.chdir(conftest.PKG)
import test_gpu_bdpt as T
import _native
from oracle import oracle, objload

name, fit, smooth = (sys.argv[1], float(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("cornell", 0.8, 0)
W = H = 64
ctx = _native.reset_context()
scene, cam, integ = T.build_gpu(name, W, H, fit, smooth)
t = objload.load_scene([conftest.model(f) for f in conftest.SCENES[name]["files"]])
o = T.build_oracle(t, W, H, fit, smooth)
xs, ys = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
px = xs.reshape(-1).astype(np.int32); py = ys.reshape(-1).astype(np.int32)
for frame in (0, 5):
    ctx.film_clear(); cam.frame = frame; cam.frame_cpu[0] = frame
    integ.render()
    verts, depths, contrib = ctx.test_bdpt_dump(px, py)
    nb = 0
    for k in range(px.size):
        ov, od, oc = o.bdpt_pixel_dump(int(px[k]), int(py[k]), frame)
        if tuple(depths[k]) != od:
            continue
        for e in range(2, 8):
            for l in range(7):
                a, b = contrib[k, e - 1, l, :3], oc[e - 1, l, :3]
                if not np.allclose(a, b, rtol=2e-3, atol=1e-6):
                    nb += 1
                    if nb < 25:
                        print("frame", frame, "pix", px[k], py[k], "depths", od, "e,l", e, l, "gpu", a, "orc", b, "ratio", a / np.where(b == 0, 1, b))
                        print("   eye types", [int(ov[v, 17]) for v in range(od[0])], "light types", [int(ov[7 + v, 17]) for v in range(od[1])],
                              "mats e", [int(ov[v, 19]) for v in range(od[0])], "mats l", [int(ov[7 + v, 19]) for v in range(od[1])])
    print("frame", frame, "bad strategies", nb)
