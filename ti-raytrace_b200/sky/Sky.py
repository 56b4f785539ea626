"""Sky — Hosek-Wilkie spectral sky dome, host side (mirror of /root/reference/sky/Sky.py:26-159).

The constructor reads the model's coefficient files, update() blends them for (turbidity, albedo, solar elevation)
with the quintic Bezier of Sky.formula, setup_data_gpu() uploads the 11 x 9 configuration, the 11 radiances and the
sun direction.  get_solar_radiance (:258-265) runs on the device (csrc/spectral.cuh: sky_radiance); the direct-sun
disc (:217-246), commented out in the reference's get_solar_radiance, is not built.  The reference opens
"sky\\data.csv" (a Windows path); paths are resolved portably here."""
import math
import numpy as np
import _native
import _paths

MATH_PI = 3.141592653589793
LAMDDA_DIV, ALBEDO_NUM, TURB_NUM, THETA_NUM, GAMMA_NUM, PIECES, ORDER = 11, 2, 10, 9, 6, 45, 4
DATA_NUM = TURB_NUM * ALBEDO_NUM * THETA_NUM * GAMMA_NUM
RAD_NUM = TURB_NUM * ALBEDO_NUM * 6
SOLAR_NUM = TURB_NUM * PIECES * ORDER
DARK_NUM = 6
MIN_LAMBDA, MAX_LAMBDA = 320.0, 720.0


def _read(path, ncol):
    out = np.zeros((LAMDDA_DIV, ncol), np.float32)
    for i, line in enumerate(l for l in open(_paths.resolve(path), "r") if l.strip()):
        out[i, :] = [float(v) for v in line.split(",", ncol)[:ncol]]
    return out


class Sky:
    def __init__(self, turbidity=3.0, albedo=0.5, elevation=10.0 * MATH_PI / 180.0):
        self.turbidity, self.albedo, self.elevation = turbidity, albedo, elevation
        self.solar_radius = 0.51 * MATH_PI / 180.0 / 2.0
        self.configs_np = np.zeros((LAMDDA_DIV, THETA_NUM), np.float32)
        self.radiances_np = np.zeros(LAMDDA_DIV, np.float32)
        self.sun_dir_np = np.zeros((1, 3), np.float32)
        self.data_np = _read("sky/data.csv", DATA_NUM)
        self.data_rad_np = _read("sky/data_rad.csv", RAD_NUM)
        self.data_solar_np = _read("sky/data_solar.csv", SOLAR_NUM)
        self.data_dark_np = _read("sky/data_dark.csv", DARK_NUM)
        self.configs = _native.Field(lambda: self.configs_np.copy())
        self.radiances = _native.Field(lambda: self.radiances_np.copy())
        self.sun_dir = _native.Field(lambda: self.sun_dir_np.copy())

    def formula(self, t, A0, A1, A2, A3, A4, A5):
        return pow(1.0 - t, 5.0) * A0 + 5.0 * pow(1.0 - t, 4.0) * t * A1 + \
            10.0 * pow(1.0 - t, 3.0) * pow(t, 2.0) * A2 + 10.0 * pow(1.0 - t, 2.0) * pow(t, 3.0) * A3 + \
            5.0 * (1.0 - t) * pow(t, 4.0) * A4 + pow(t, 5.0) * A5

    def _blend(self, table, out, stride, block):
        """out[j, i] = sum over (albedo 0/1) x (turbidity floor / floor+1) of weight * Bezier(elevation; 6 control values)
        (Sky.py:112-156).  Arithmetic in f64 with an f32 store after every term, as NumPy 1.18 evaluates
        `f32_array[j, i] += python_float * ...`."""
        it = int(self.turbidity)
        rem = self.turbidity - float(it)
        se = pow(self.elevation / (MATH_PI / 2.0), 1.0 / 3.0)
        terms = [(block * (it - 1), (1.0 - self.albedo) * (1.0 - rem)), (block * 10 + block * (it - 1), self.albedo * (1.0 - rem))]
        if it < 10:
            terms += [(block * it, (1.0 - self.albedo) * rem), (block * 10 + block * it, self.albedo * rem)]
        for n, (index, wgt) in enumerate(terms):
            for j in range(out.shape[0]):
                for i in range(out.shape[1]):
                    A = [float(table[j, index + i + stride * k]) for k in range(6)]
                    v = wgt * self.formula(se, *A)
                    out[j, i] = v if n == 0 else float(out[j, i]) + v

    def update(self):
        self._blend(self.data_np, self.configs_np, 9, 9 * 6)
        rad = self.radiances_np.reshape(LAMDDA_DIV, 1)
        self._blend(self.data_rad_np, rad, 1, 6)

    def setup_data_gpu(self):
        self.update()
        self.sun_dir_np[0, :] = (0.0, math.sin(self.elevation), math.cos(self.elevation))
        _native.context().spec_sky_upload(self.configs_np, self.radiances_np, self.sun_dir_np[0])
