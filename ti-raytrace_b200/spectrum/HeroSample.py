"""Hero-wavelength sampling constants (mirror of /root/reference/spectrum/HeroSample.py:5-8).

The device functions of the reference module (sample, sample_xyz, get_rnd_hero, srgb_to_spec, sky_sample,
HeroSample.py:10-70) are CUDA code in csrc/spectral.cuh; a path carries SAMPLE_WAVELENGTHS wavelengths
Lambda0 + i * LAMBDA_STEP."""
SAMPLE_WAVELENGTHS = 4
LAMBDA_MIN = 360.0
LAMBDA_MAX = 760.0
LAMBDA_STEP = (LAMBDA_MAX - LAMBDA_MIN) / SAMPLE_WAVELENGTHS
