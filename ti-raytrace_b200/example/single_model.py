"""One OBJ model under an environment map and a sphere light, PT_RGB (counterpart of
/root/reference/example/single_model.py as shipped: sphere.obj turned into glass)."""
import Example
import taichi as ti
import PT_RGB
import SceneData as SCD


class example(Example.example):
    models = ("model/sphere.obj",)

    def __init__(self, imgSizeX, imgSizeY, sample_count):
        ti.init(arch=ti.gpu)
        super().__init__(imgSizeX, imgSizeY, sample_count)
        for path in self.models:
            self.scene.add_obj(path)
        self.edit_materials()
        self.add_sphere_light()
        self.scene.add_env("image/env.png", 5.0)
        self.integrator = PT_RGB.PathTrace(imgSizeX, imgSizeY, self.cam, self.scene, 64)

    def edit_materials(self):
        glass = self.scene.material_cpu[0]
        glass.type = SCD.MAT_GLASS
        glass.setIor(1.3)
        glass.setExtinciton(5.0)

    def build_scene(self):
        super().build_scene()
        self.scene.process_normal()
        self.scene.total_area()
        print("********total light area:%f****" % (self.scene.light_area.to_numpy()[0]))
        self.fit_camera()
