"""Debug integrator — primary hit -> material colour (mirror of /root/reference/integrator/Debug.py:22-66).
Deterministic (frame-0 rays, no RNG): the parity harness for ray generation + traversal.  first_hit()
additionally returns the (t, prim, uv, pos, gnormal, normal, dir) buffers of those rays."""
import _native


class Debug:
    def __init__(self, imgSizeX, imgSizeY, cam, scene, stack_size):
        self.imgSizeX, self.imgSizeY = imgSizeX, imgSizeY
        self.cam, self.scene, self.stack_size = cam, scene, stack_size
        self.hdr = _native.Field(lambda: _native.context().film_download(True, False)[0])
        self.rgb_film = _native.Field(lambda: _native.context().film_download(False, True)[1])

    def setup_data_cpu(self):
        _native.context().film_create(self.imgSizeX, self.imgSizeY)

    def setup_data_gpu(self):
        pass

    def render(self):
        ctx = _native.context()
        self.cam.push(ctx)
        ctx.render_debug()

    def first_hit(self):
        return _native.context().first_hit_download()
