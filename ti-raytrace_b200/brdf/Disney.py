"""Disney BRDF (mirror of /root/reference/brdf/Disney.py).  The reference's functions are @ti.func
device code; here they are CUDA device functions (csrc/common.cuh) used by the shade kernel.  The
array forms below run them on the GPU through the unit hooks of include/tiray.h, with the material
given as (metal, rough) instead of (mat_buf, mat_id)."""
import _native


def evaluate_pdf(N, V, L, metal, rough):
    """-> (n,2) array of (brdf, pdf); (0,-1) where NdotL<=0 or NdotV<=0 (Disney.py:65-108)"""
    return _native.context().test_disney_evaluate_pdf(N, V, L, metal, rough)


def sample(dir, N, metal, rough, u):
    """u = (lobe probability, r1, r2) per row (Disney.py:17-40) -> next directions"""
    return _native.context().test_disney_sample(dir, N, metal, rough, u)
