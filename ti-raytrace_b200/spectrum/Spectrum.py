"""Spectrum — one tabulated spectrum on a regular wavelength grid (mirror of
/root/reference/spectrum/Spectrum.py:8-56): load_table(csv), setup_data_gpu(), scale(coff), `data`.

Spectrum.sample (:43-51) runs on the device (csrc/spectral.cuh: spectrum_sample).  The device side has four
table slots (include/tiray.h: tr_spec_spectrum_upload); PT_Spec assigns them (0 D65, 1 white, 2 red, 3 green)."""
import numpy as np
import _native
import _paths


def read_csv_columns(path, ncol):
    """`lambda, v1[, v2, v3]` rows -> (lambda list, (n, ncol) f32); values pass through Python floats like the reference"""
    lam, rows = [], []
    for line in open(_paths.resolve(path), "r"):
        if not line.strip():
            continue
        v = line.split(",", ncol + 1)
        lam.append(float(v[0]))
        rows.append([float(v[1 + k]) for k in range(ncol)])
    return lam, np.asarray(rows, np.float32)


class Spectrum:
    def __init__(self, slot=None):
        self.slot = slot
        self.lambda_min, self.lambda_max, self.lambda_range, self.size = 10000, 0, 0, 0
        self.data_np = None
        self.white_point_np = np.zeros((1, 3), np.float32)
        self.data = _native.Field(lambda: _native.context().spec_spectrum_download(self._slot(), self.size))
        self.white_point = _native.Field(lambda: self.white_point_np.copy())

    def _slot(self):
        if self.slot is None:
            raise RuntimeError("Spectrum has no device slot (PT_Spec assigns 0 D65, 1 white, 2 red, 3 green)")
        return self.slot

    def load_table(self, table_path):
        lam, rows = read_csv_columns(table_path, 1)
        self.size = len(lam)
        self.lambda_min, self.lambda_max = lam[0], lam[-1]
        self.data_np = np.ascontiguousarray(rows[:, 0])
        self.lambda_range = (self.lambda_max - self.lambda_min) / (self.size - 1)

    def setup_data_gpu(self):
        _native.context().spec_spectrum_upload(self._slot(), self.data_np, self.lambda_min, self.lambda_max)

    def scale(self, coff):
        """Spectrum.scale (:53-56); coff is rounded to f32 like a ti.f32 kernel argument"""
        ctx = _native.context()
        d = ctx.spec_spectrum_download(self._slot(), self.size) * np.float32(coff)
        ctx.spec_spectrum_upload(self._slot(), d, self.lambda_min, self.lambda_max)
