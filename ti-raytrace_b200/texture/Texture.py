"""Environment texture (mirror of /root/reference/texture/Texture.py:18-42).  The image is packed as
RGB in one int32 per texel and stored [x][H-1-row]; the bilinear lookup (Texture.py:44-69) happens
in the shade kernel (csrc/wavefront.cu env_texture2d)."""
import numpy as np


class Texture:
    def __init__(self):
        self.wid = self.hgt = self.channel = self.size = 0
        self.np_img = None

    def load_image(self, imagePath):
        import cv2
        import _paths
        img = cv2.imread(_paths.resolve(imagePath))
        if img is None:
            raise FileNotFoundError(imagePath)
        self.hgt, self.wid, self.channel = img.shape
        self.size = self.wid * self.hgt * self.channel
        bgr = img.astype(np.int32)
        packed = (bgr[:, :, 2] << 16) | (bgr[:, :, 1] << 8) | bgr[:, :, 0]     # [row][col]
        self.np_img = np.ascontiguousarray(packed[::-1, :].T)                  # [x][H-1-row]
        import _native
        self._pin = _native.pin_array(self.np_img)                             # re-uploaded by every Scene.setup_data_gpu(): direct DMA

    def setup_data_gpu(self, power=0.0):
        import _native
        _native.context().env_upload(self.np_img, self.wid, self.hgt, power)
