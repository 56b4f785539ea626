"""Application shell shared by the example scenes: owns the (headless) window, the camera, the scene
and an integrator, like /root/reference/example/Example.py:11-59, and drives one sample per render()
call.  Scene scripts subclass `example`, add geometry in __init__ and pick an integrator."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (_ROOT, os.path.join(_ROOT, "integrator")):
    if _p not in sys.path:
        sys.path.append(_p)

import math
import taichi as ti
import Camera
import Scene
import SceneData as SCD
import UtilsFunc as UF


class example:
    exposure = 0.5
    camera_fit = 0.8       # camera distance = fit x scene diagonal

    def __init__(self, imgSizeX, imgSizeY, sample_count):
        self.imgSizeX, self.imgSizeY, self.sample_count = imgSizeX, imgSizeY, sample_count
        self.gui = ti.GUI("Render", res=(imgSizeX, imgSizeY))
        self.cam = Camera.Camera(imgSizeX, imgSizeY, sample_count)
        self.scene = Scene.Scene()
        self.integrator = None
        self.out_file = "out.png"

    def build_scene(self):
        # order matters: the BVH is built by scene.setup_data_gpu(), after the film exists
        self.scene.setup_data_cpu()
        self.integrator.setup_data_cpu()
        self.integrator.setup_data_gpu()
        self.scene.setup_data_gpu()

    def fit_camera(self):
        lo, hi = self.scene.minboundarynp[0], self.scene.maxboundarynp[0]
        diag = math.sqrt(sum(float(hi[k] - lo[k]) ** 2 for k in range(3)))
        self.cam.scale = diag * self.camera_fit
        mid = (hi + lo) * 0.5
        self.cam.set_target(float(mid[0]), float(mid[1]), float(mid[2]))
        self.cam.update()

    def add_sphere_light(self, pos=(0.0, 20.0, 0.0), radius=5.0, power=50.0):
        ball = SCD.Shape()
        ball.type = SCD.SHPAE_SPHERE
        ball.pos = list(pos)
        ball.setRadius(radius)
        emitter = SCD.Material()
        emitter.type = SCD.MAT_LIGHT
        emitter.setColor([power, power, power])
        self.scene.add_shape(ball, emitter)

    def _present(self):
        self.gui.set_image(self.integrator.rgb_film.to_numpy())
        self.gui.show()

    def render(self):
        if not self.gui.running:
            return 0
        done = int(self.cam.frame_cpu[0])
        if done < self.sample_count:
            self.integrator.render()
            UF.tone_map(self.exposure, self.integrator.hdr, self.integrator.rgb_film)
            self._present()
            self.cam.update_frame()
            return 1
        if done == self.sample_count:
            ti.imwrite(self.integrator.rgb_film, self.out_file)
            self.cam.update_frame()
            self._present()
            return 0
        self._present()
        return 1
