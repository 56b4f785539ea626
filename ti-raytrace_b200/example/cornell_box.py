"""Cornell box, PT_RGB (BASELINE configs C1/C2; counterpart of /root/reference/example/cornell_box.py)."""
import Example
import taichi as ti
import PT_RGB


class example(Example.example):
    def __init__(self, imgSizeX, imgSizeY, sample_count):
        ti.init(arch=ti.gpu)
        super().__init__(imgSizeX, imgSizeY, sample_count)
        self.scene.add_obj("model/cornell_box.obj")
        self.integrator = PT_RGB.PathTrace(imgSizeX, imgSizeY, self.cam, self.scene, 64)

    def build_scene(self):
        super().build_scene()
        self.scene.total_area()
        print("********total light area:%f****" % (self.scene.light_area.to_numpy()[0]))
        self.fit_camera()
