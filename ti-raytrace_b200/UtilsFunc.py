"""Host-visible part of the reference's UtilsFunc.py: the constants scene scripts read and the
`tone_map` kernel entry (UtilsFunc.py:583-586).  The device math of that module (accessors, slabs,
Morton, sampling, offset_ray ...) lives in csrc/common.cuh and csrc/trace.cuh."""
import _native

AXIS_X, AXIS_Y, AXIS_Z = 0, 1, 2
EPS = 0.00001
M_PIf = 3.1415956          # (sic) UtilsFunc.py:37
INF_VALUE = 1000000.0


def tone_map(exposure, input, output):
    """output = srgb(clamp(ACES(input * exposure))) on the device film (input/output are the
    integrator's hdr / rgb_film fields; there is one film per context)."""
    _native.context().tonemap(exposure)
