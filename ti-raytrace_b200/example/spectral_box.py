"""Cornell box with measured wall reflectances, PT_Spec (BASELINE config C4; counterpart of
/root/reference/example/spectral_box.py)."""
import Example
import taichi as ti
import PT_Spec
import SceneData as SCD


class example(Example.example):
    def __init__(self, imgSizeX, imgSizeY, sample_count):
        ti.init(arch=ti.gpu)
        super().__init__(imgSizeX, imgSizeY, sample_count)
        self.scene.add_obj("model/cornell_box.obj")
        self.integrator = PT_Spec.PathTrace(imgSizeX, imgSizeY, self.cam, self.scene, 64)
        # white, red and green walls use the measured spectra selected by alebdoTex (PT_Spec.py:119-135)
        for k in range(3):
            self.scene.material_cpu[k].type = SCD.MAT_SPECTRAL
            self.scene.material_cpu[k].alebdoTex = k

    def build_scene(self):
        super().build_scene()
        self.scene.process_normal()
        self.scene.total_area()
        print("********total light area:%f****" % (self.scene.light_area.to_numpy()[0]))
        self.fit_camera()
