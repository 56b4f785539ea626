"""PT_Spec.PathTrace — hero-wavelength spectral path tracer (mirror of /root/reference/integrator/PT_Spec.py:29-107).

Same construction and set-up sequence as the reference class: setup_data_cpu() reads the CIE 1931 observer, the
rgb2spec coefficient table and the D65 / white / red / green spectra; setup_data_gpu() uploads them and the sky
model and normalises D65 to Y = 1.  render() = one sample per pixel through tr_render_pt_spec() (csrc/wavefront.cu,
SPEC instantiation of the wavefront kernels); render_frames(n) renders n samples in one call."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _sub in ("spectrum", "sky"):
    _p = os.path.join(_ROOT, _sub)
    if _p not in sys.path:
        sys.path.append(_p)

import numpy as np
import _native
import Rgb2Spec as RGB2SPEC
import Spectrum as Spec
import HeroSample as Hero
import Sky

MAX_DEPTH = 10
SLOT_D65, SLOT_WHITE, SLOT_RED, SLOT_GREEN = 0, 1, 2, 3


class PathTrace:
    def __init__(self, imgSizeX, imgSizeY, cam, scene, stack_size):
        self.imgSizeX, self.imgSizeY = imgSizeX, imgSizeY
        self.lambda_min, self.lambda_max, self.lambda_range, self.size = 10000, 0, 0, 0
        self.d65, self.white = Spec.Spectrum(SLOT_D65), Spec.Spectrum(SLOT_WHITE)
        self.red, self.green = Spec.Spectrum(SLOT_RED), Spec.Spectrum(SLOT_GREEN)
        self.rgb2spec = RGB2SPEC.Rgb2Spec()
        self.sky = Sky.Sky(3.0, 0.5, 0.17)
        self.cam, self.scene = cam, scene
        self.stack_size = stack_size          # kept for API parity; the traversal is stackless
        self.seed = 0
        self.max_depth = MAX_DEPTH
        self.hdr = _native.Field(lambda: _native.context().film_download(True, False)[0],
                                 lambda a: _native.context().film_upload(a))
        self.rgb_film = _native.Field(lambda: _native.context().film_download(False, True)[1])

    def setup_data_cpu(self):
        lam, self.data_np = Spec.read_csv_columns("spectrum/ciexyz31_1.csv", 3)
        self.size = len(lam)
        self.lambda_min, self.lambda_max = lam[0], lam[-1]
        self.lambda_range = (self.lambda_max - self.lambda_min) / (self.size - 1)
        _native.context().film_create(self.imgSizeX, self.imgSizeY)
        self.rgb2spec.load_table("spectrum/spec_table")
        self.d65.load_table("spectrum/Illuminantd65.csv")
        self.red.load_table("spectrum/red-spec.csv")
        self.green.load_table("spectrum/green-spec.csv")
        self.white.load_table("spectrum/white-spec.csv")

    def setup_data_gpu(self):
        _native.context().spec_sensor_upload(self.data_np, self.lambda_min, self.lambda_max)
        self.rgb2spec.setup_data_gpu()
        for s in (self.d65, self.red, self.green, self.white):
            s.setup_data_gpu()
        self.sky.setup_data_gpu()
        self.normalize_spec(self.d65)

    def normalize_spec(self, spec):
        """cal_white_point + Spectrum.scale(1 / Y) (PT_Spec.py:101-107,174-187), one device call"""
        spec.white_point_np[0, :] = _native.context().spec_normalize(spec.slot)

    def _prepare(self):
        ctx = _native.context()
        self.cam.push(ctx)
        return ctx

    def render(self):
        self._prepare().render_pt_spec(self.cam.frame, 1, self.max_depth, self.seed)

    def render_frames(self, n_frames, stats=True):
        ctx = self._prepare()
        ctx.render_pt_spec(self.cam.frame, n_frames, self.max_depth, self.seed)
        self.cam.update_frame(n_frames)
        return ctx.stats() if stats else None      # stats() waits for the (asynchronous) render
