"""PT_RGB.PathTrace — unidirectional path tracer (mirror of /root/reference/integrator/PT_RGB.py:21-45).

render() = one sample per pixel, like the reference's single @ti.kernel, but executed as a wavefront
of sm_100a kernels through tr_render_pt_rgb() (csrc/wavefront.cu).  render_frames(n) is an addition:
it renders n samples per pixel in one call (several frames per wavefront batch) and advances
cam.frame itself; the result is bit-identical to n x (render(); cam.update_frame())."""
import _native

MAX_DEPTH = 15


class PathTrace:
    def __init__(self, imgSizeX, imgSizeY, cam, scene, stack_size):
        self.imgSizeX, self.imgSizeY = imgSizeX, imgSizeY
        self.cam, self.scene = cam, scene
        self.stack_size = stack_size          # kept for API parity; the traversal is stackless
        self.seed = 0
        self.max_depth = MAX_DEPTH
        self.hdr = _native.Field(lambda: _native.context().film_download(True, False)[0],
                                 lambda a: _native.context().film_upload(a))
        self.rgb_film = _native.Field(lambda: _native.context().film_download(False, True)[1])

    def setup_data_cpu(self):
        _native.context().film_create(self.imgSizeX, self.imgSizeY)

    def setup_data_gpu(self):
        pass

    def _prepare(self):
        ctx = _native.context()
        self.cam.push(ctx)
        self.scene._sync_late_scalars()
        return ctx

    def render(self):
        self._prepare().render_pt_rgb(self.cam.frame, 1, self.max_depth, self.seed)

    def render_frames(self, n_frames, stats=True):
        ctx = self._prepare()
        ctx.render_pt_rgb(self.cam.frame, n_frames, self.max_depth, self.seed)
        self.cam.update_frame(n_frames)
        return ctx.stats() if stats else None      # stats() waits for the (asynchronous) render
