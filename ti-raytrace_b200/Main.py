"""Driver (counterpart of /root/reference/Main.py): pick an example module, build, render until done.
usage: python Main.py [example_module] [width height spp]      (run from this directory)"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _ROOT)
sys.path.append(os.path.join(_ROOT, "example"))

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "cornell_box"
    w, h, spp = (int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (512, 512, 512)
    ex = importlib.import_module(name).example(w, h, spp)
    ex.build_scene()
    ret = 1
    while ret == 1:
        ret = ex.render()
