"""Veach's bidirectional test room, BDPT_RGB (BASELINE config C5; counterpart of /root/reference/example/veach_bdpt.py)."""
import Example
import taichi as ti
import BDPT_RGB


class example(Example.example):
    camera_fit = 0.5

    def __init__(self, imgSizeX, imgSizeY, sample_count):
        ti.init(arch=ti.gpu)
        super().__init__(imgSizeX, imgSizeY, sample_count)
        self.scene.add_obj("model/bdpt.obj")
        self.integrator = BDPT_RGB.BDPT(imgSizeX, imgSizeY, self.cam, self.scene, 64)

    def build_scene(self):
        super().build_scene()
        self.scene.process_normal()
        self.scene.total_area()
        print("********total light area:%f****" % (self.scene.light_area.to_numpy()[0]))
        self.fit_camera()
