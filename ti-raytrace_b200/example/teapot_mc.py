"""BASELINE config C3: single_model with the two commented-out add_obj lines of the reference enabled
(model/mc.obj then model/Teapot.obj, 130 720 triangles), default Disney materials, sphere light and
environment map, smooth normals."""
import single_model


class example(single_model.example):
    models = ("model/mc.obj", "model/Teapot.obj")

    def edit_materials(self):
        pass
