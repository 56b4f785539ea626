"""Minimal `taichi` stand-in for the names the reference's example scripts touch:
ti.init / ti.gpu / ti.cpu / ti.GUI / ti.imwrite.  It is NOT a Taichi runtime: device work goes
through libtiray.so (see _native.py).  Call sites: /root/reference/example/Example.py:16,43-49,
example/cornell_box.py:15.
"""
import numpy as np

cpu, gpu, cuda, opengl, x64 = "cpu", "gpu", "cuda", "opengl", "x64"
f32, i32 = np.float32, np.int32
__version__ = (0, 7, 14)


def init(arch=None, **kwargs):
    """Start a fresh device program (Taichi semantics: ti.init resets all fields).
    arch is accepted for source compatibility; the only back end is sm_100a CUDA, also for ti.cpu."""
    import _native
    return _native.reset_context(kwargs.get("device"))


def _to_bytes(img):
    if not isinstance(img, np.ndarray) and hasattr(img, "to_numpy"):
        img = img.to_numpy()
    if img.dtype in (np.float32, np.float64):
        img = (np.clip(img, 0, 1) * 255.0 + 0.5).astype(np.uint8)
    return img


def imwrite(img, filename):
    """Taichi convention: field index [x][y] with y up -> image rows top-down."""
    import cv2
    img = _to_bytes(img)
    img = np.ascontiguousarray(img.swapaxes(0, 1)[::-1, :])
    if img.ndim == 3 and img.shape[2] == 3:
        img = img[:, :, ::-1]
    if not cv2.imwrite(filename, img):
        raise RuntimeError("could not write %s" % filename)


def imread(filename, channels=0):
    import cv2
    img = cv2.imread(filename)[:, :, ::-1]
    return np.ascontiguousarray(img[::-1, :].swapaxes(0, 1))


class GUI:
    """Headless window: keeps the last image; `running` stays True (the example loop ends by itself
    when frame == sample_count, example/Example.py:48-53)."""

    def __init__(self, name="", res=(512, 512), **kwargs):
        self.name, self.res, self.running, self.img, self.frames_shown = name, res, True, None, 0

    def set_image(self, img):
        self.img = img

    def show(self, file=None):
        self.frames_shown += 1
        if file is not None and self.img is not None:
            imwrite(self.img, file)

    def get_event(self, *a):
        return False

    def close(self):
        self.running = False
