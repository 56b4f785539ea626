"""BDPT_RGB.BDPT — bidirectional path tracer (mirror of /root/reference/integrator/BDPT_RGB.py:33-88,600-641).

render() = one sample per pixel: eye + light sub-paths, all (e, l) connections with MIS, light-tracing splats; executed
by tr_render_bdpt_rgb() (csrc/bdpt.cuh) as sub-path / queue / connect / film kernels.  The reference's seven per-pixel
`Vertex` field groups (integrator/BDPT_Vertex.py) live in one SoA vertex buffer owned by the context.
render_frames(n) renders n samples per pixel in one call and advances cam.frame."""
import _native

STOP_DEPTH = 10000
MAX_DEPTH = 5
EYE_MAX_DEPTH = MAX_DEPTH + 2
LIGHT_MAX_DEPTH = MAX_DEPTH + 1

VERTEX_NONE, VERTEX_LIGHT, VERTEX_LENS, VERTEX_SURFACE = 0, 1, 2, 3


class BDPT:
    def __init__(self, imgSizeX, imgSizeY, cam, scene, stack_size):
        self.imgSizeX, self.imgSizeY = imgSizeX, imgSizeY
        self.cam, self.scene = cam, scene
        self.stack_size = stack_size          # kept for API parity; the traversal is stackless
        self.seed = 0
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
        self._prepare().render_bdpt_rgb(self.cam.frame, 1, self.seed)

    def render_frames(self, n_frames, stats=True):
        ctx = self._prepare()
        ctx.render_bdpt_rgb(self.cam.frame, n_frames, self.seed)
        self.cam.update_frame(n_frames)
        return ctx.stats() if stats else None      # stats() waits for the (asynchronous) render
