"""Glass BSDF (mirror of /root/reference/brdf/Glass.py:9-34,68-78); array form over the GPU hook."""
import _native


def sample(dir, N, ior, u):
    """u = Fresnel probability per row -> (n,4): next direction, f_or_b (-1 on refraction)"""
    return _native.context().test_glass_sample(dir, N, ior, u)


def evaluate_pdf(*args):
    return 1.0, 1.0
