"""Mirror sphere under the Hosek-Wilkie sky dome and a sphere light, PT_Spec (counterpart of
/root/reference/example/sky_dome.py)."""
import Example
import taichi as ti
import PT_Spec


class example(Example.example):
    camera_fit = 2.0

    def __init__(self, imgSizeX, imgSizeY, sample_count):
        ti.init(arch=ti.gpu)
        super().__init__(imgSizeX, imgSizeY, sample_count)
        self.scene.add_obj("model/sphere.obj")
        self.scene.material_cpu[0].setMetal(1.0)
        self.scene.material_cpu[0].setRough(0.0)
        self.add_sphere_light()
        self.integrator = PT_Spec.PathTrace(imgSizeX, imgSizeY, self.cam, self.scene, 64)

    def build_scene(self):
        super().build_scene()
        self.scene.process_normal()
        self.scene.total_area()
        self.scene.env_power = 0.0
        print("********total light area:%f****" % (self.scene.light_area.to_numpy()[0]))
        self.fit_camera()
