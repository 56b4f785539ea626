/* trmath.h — the transcendental functions of the hot path, written once in plain IEEE f32 arithmetic.
 *
 * The reference calls ti.sin / ti.cos / ti.pow / ti.exp / ts.acos / ts.atan (Taichi 0.7.14 with fast_math=True: whatever the LLVM
 * back end of the day lowers them to, parity unpinned).  CUDA's libdevice and glibc's libm round these functions
 * differently by an ULP or two, which is enough to flip a `t > 0` test or a Fresnel coin once in 10^5 vertices.  This header is
 * compiled into BOTH the CUDA kernels (nvcc -fmad=false, IEEE division / square root) and the CPU oracle (g++ -ffp-contract=off
 * -fno-fast-math): only +, -, *, /, sqrt, int<->float conversions and bit casts in a fixed order, so the two sides produce the same
 * bits and radiance parity is exact instead of "within libm".
 *
 * Algorithms: the classic single-precision Cody-Waite reductions and minimax polynomials of the Cephes math library
 * (sinf / cosf: octant reduction with a three-part pi/4; expf: ln2 split + degree-5 polynomial; logf: sqrt(1/2)-centred mantissa +
 * degree-8 polynomial; asinf / atanf: the usual interval folding).  Accuracy against correctly rounded results is checked in
 * tests/test_trmath.py (sin, cos, exp, acos, atan2 within 2 ULP on the ranges the renderer uses; pow = exp2(y log2 x) evaluated
 * with a two-float logarithm, within 4 ULP there).
 */
#ifndef TRMATH_H
#define TRMATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#ifdef __CUDACC__
#define TRM_FN __host__ __device__ __forceinline__
#else
#define TRM_FN static inline
#endif

TRM_FN int32_t trm_f2i(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int32_t i; memcpy(&i, &f, 4); return i;
#endif
}
TRM_FN float trm_i2f(int32_t i) {
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}
TRM_FN float trm_abs(float x) { return trm_i2f(trm_f2i(x) & 0x7fffffff); }
TRM_FN float trm_nan(void) { return trm_i2f(0x7fc00000); }
/* sqrtf is IEEE (correctly rounded) on both sides; spelled out so that no fast-math macro can replace it in this header */
TRM_FN float trm_sqrt(float x) { return sqrtf(x); }

/* ---- sin / cos ------------------------------------------------------------------------------------------------------------
 * |x| < 2^18: octant j = trunc(|x| * 4/pi), made even; r = ((|x| - j*DP1) - j*DP2) - j*DP3 with DP1 + DP2 + DP3 = pi/4 and
 * DP1, DP2 carrying few enough mantissa bits that j*DP1, j*DP2 are exact for j < 2^13 (beyond that the result degrades
 * gracefully; the renderer's arguments are below 8).  Larger arguments and non-finite ones return NaN / 0 deterministically. */
#define TRM_FOPI 1.27323954473516f
#define TRM_DP1 0.78515625f
#define TRM_DP2 2.4187564849853515625e-4f
#define TRM_DP3 3.77489497744594108e-8f

TRM_FN float trm_sin_poly(float r) {       /* |r| <= pi/4 */
    float z = r * r;
    float p = -1.9515295891e-4f * z + 8.3321608736e-3f;
    p = p * z - 1.6666654611e-1f;
    return (p * z) * r + r;
}
TRM_FN float trm_cos_poly(float r) {       /* |r| <= pi/4 */
    float z = r * r;
    float p = 2.443315711809948e-5f * z - 1.388731625493765e-3f;
    p = p * z + 4.166664568298827e-2f;
    return (p * z) * z - 0.5f * z + 1.0f;
}
/* octant and reduced argument of |x| */
TRM_FN float trm_reduce(float ax, int32_t* octant) {
    int32_t j = (int32_t)(ax * TRM_FOPI);
    j = (j + 1) & ~1;                        /* map odd octants up: r in [-pi/4, pi/4] */
    float y = (float)j;
    *octant = j;
    return ((ax - y * TRM_DP1) - y * TRM_DP2) - y * TRM_DP3;
}
TRM_FN float tr_sinf(float x) {
    float ax = trm_abs(x);
    if (!(ax < 262144.0f)) return (ax == ax && ax < 3.0e38f) ? 0.0f : trm_nan();
    int32_t j; float r = trm_reduce(ax, &j);
    int32_t q = (j >> 1) & 3;                /* quadrant: 0 sin, 1 cos, 2 -sin, 3 -cos */
    float v = (q & 1) ? trm_cos_poly(r) : trm_sin_poly(r);
    if (q & 2) v = -v;
    return (trm_f2i(x) < 0) ? -v : v;
}
TRM_FN float tr_cosf(float x) {
    float ax = trm_abs(x);
    if (!(ax < 262144.0f)) return (ax == ax && ax < 3.0e38f) ? 1.0f : trm_nan();
    int32_t j; float r = trm_reduce(ax, &j);
    int32_t q = (j >> 1) & 3;                /* quadrant: 0 cos, 1 -sin, 2 -cos, 3 sin */
    float v = (q & 1) ? trm_sin_poly(r) : trm_cos_poly(r);
    if (q == 1 || q == 2) v = -v;
    return v;
}

/* tan = sin / cos (one IEEE division): only the spot light's cone radius uses it (Scene.py:456-457) */
TRM_FN float tr_tanf(float x) { return tr_sinf(x) / tr_cosf(x); }

/* ---- exp ------------------------------------------------------------------------------------------------------------------
 * n = floor(x*log2(e) + 0.5); r = (x - n*C1) - n*C2 (C1 + C2 = ln 2, n*C1 exact); e^r by a degree-5 polynomial; scaled by 2^n in
 * two exact power-of-two multiplications so that gradual underflow rounds once. */
TRM_FN float trm_pow2i(int32_t n) { return trm_i2f((n + 127) << 23); }        /* 2^n, -126 <= n <= 127 */
TRM_FN float trm_scale2(float p, int32_t n) {                                  /* p * 2^n, any n in [-300, 300] */
    int32_t h = n / 2;
    return (p * trm_pow2i(h)) * trm_pow2i(n - h);
}
TRM_FN float tr_expf(float x) {
    if (x != x) return x;
    if (x > 88.72283905206835f) return trm_i2f(0x7f800000);
    if (x < -103.972084045410f) return 0.0f;
    float fn = floorf(x * 1.44269504088896341f + 0.5f);
    int32_t n = (int32_t)fn;
    float r = x - fn * 0.693359375f;
    r = r - fn * -2.12194440e-4f;
    float z = r * r;
    float p = 1.9875691500e-4f * r + 1.3981999507e-3f;
    p = p * r + 8.3334519073e-3f;
    p = p * r + 4.1665795894e-2f;
    p = p * r + 1.6666665459e-1f;
    p = p * r + 5.0000001201e-1f;
    p = p * z + r;
    p = p + 1.0f;
    return trm_scale2(p, n);
}

/* ---- pow -----------------------------------------------------------------------------------------------------------------
 * x^y = 2^(y log2 x).  log2 x = e + log2(m), m in [sqrt(1/2), sqrt(2)); log2(m) is produced as an unevaluated sum hi + lo of two
 * floats (the polynomial part is small, its leading term (m-1)/ln2 is split exactly), so that y*log2(x) keeps ~30 bits and the
 * result stays within a few ULP although |y log2 x| reaches 20.  The renderer's calls: srgb <-> linear (y = 2.4, 1/2.4), Schlick
 * (y = 5), ACES / sky terms; x >= 0 there.  x < 0 returns NaN for non-integer y like C, x == 0 returns 0 (y > 0), 1 (y == 0), inf. */
TRM_FN void trm_split(float a, float* hi, float* lo) {                         /* Veltkamp split: a = hi + lo, hi has 12 bits */
    float c = 4097.0f * a;
    float h = c - (c - a);
    *hi = h; *lo = a - h;
}
TRM_FN void trm_two_prod(float a, float b, float* p, float* e) {               /* a*b = p + e exactly (Dekker, no FMA) */
    float ah, al, bh, bl;
    *p = a * b;
    trm_split(a, &ah, &al); trm_split(b, &bh, &bl);
    *e = ((ah * bh - *p) + ah * bl + al * bh) + al * bl;
}
/* log2(x) for finite x > 0 as hi + lo */
TRM_FN void trm_log2_2(float x, float* hi, float* lo) {
    int32_t ix = trm_f2i(x), e = 0;
    if (ix < 0x00800000) { x = x * 8388608.0f; ix = trm_f2i(x); e = -23; }     /* subnormal */
    e += (ix >> 23) - 127;
    float m = trm_i2f((ix & 0x007fffff) | 0x3f800000);                         /* [1, 2) */
    if (m > 1.41421356237f) { m = m * 0.5f; e += 1; }
    float f = m - 1.0f;                                                        /* exact, [-0.2929, 0.4143) */
    /* ln(1+f) = f - f^2/2 + f^3 * P(f)   (Cephes logf) */
    float z = f * f;
    float p = 7.0376836292e-2f * f - 1.1514610310e-1f;
    p = p * f + 1.1676998740e-1f;
    p = p * f - 1.2420140846e-1f;
    p = p * f + 1.4249322787e-1f;
    p = p * f - 1.6668057665e-1f;
    p = p * f + 2.0000714765e-1f;
    p = p * f - 2.4999993993e-1f;
    p = p * f + 3.3333331174e-1f;
    float t = (p * f) * z - 0.5f * z;                                          /* ln(1+f) - f, small */
    /* log2 = (f + t) * log2(e), log2(e) = L1 + L2 with L1 = 2954 / 2048 (12 bits: fh * L1 is exact) */
    const float L1 = 1.4423828125f, L2 = 3.1222838896341e-4f;
    float fh, fl; trm_split(f, &fh, &fl);
    float a = fh * L1;                                                         /* exact (12 x 12 bits) */
    float b = ((fl * L1 + f * L2) + t * L1) + t * L2;
    float s = (float)e + a;                                                    /* |e| <= 150, a has <= 24 bits: rounding absorbed in lo */
    float r = ((float)e - s) + a;
    *hi = s; *lo = r + b;
}
TRM_FN float trm_exp2(float h, float l) {                                      /* 2^(h + l), |l| tiny */
    if (h > 128.0f) return trm_i2f(0x7f800000);
    if (h < -151.0f) return 0.0f;
    float fn = floorf(h + 0.5f);
    int32_t n = (int32_t)fn;
    float r = (h - fn) + l;                                                    /* [-0.5, 0.5] */
    r = r * 0.693147180559945f;                                                /* e^(r ln2) with the exp polynomial */
    float z = r * r;
    float p = 1.9875691500e-4f * r + 1.3981999507e-3f;
    p = p * r + 8.3334519073e-3f;
    p = p * r + 4.1665795894e-2f;
    p = p * r + 1.6666665459e-1f;
    p = p * r + 5.0000001201e-1f;
    p = p * z + r;
    p = p + 1.0f;
    return trm_scale2(p, n);
}
TRM_FN float tr_powf(float x, float y) {
    if (y == 0.0f) return 1.0f;
    if (x != x || y != y) return trm_nan();
    if (x == 1.0f) return 1.0f;
    float ax = trm_abs(x);
    int neg = 0;
    if (trm_f2i(x) < 0 && x != 0.0f) {                                         /* negative base: integer exponents only */
        float yi = floorf(y);
        if (yi != y) return trm_nan();
        neg = (trm_abs(y) < 16777216.0f) && (((int32_t)yi) & 1);
    }
    float r;
    if (ax == 0.0f) r = (y > 0.0f) ? 0.0f : trm_i2f(0x7f800000);
    else if (ax > 3.0e38f) r = (y > 0.0f) ? trm_i2f(0x7f800000) : 0.0f;
    else {
        float lh, ll; trm_log2_2(ax, &lh, &ll);
        float ph, pe; trm_two_prod(lh, y, &ph, &pe);
        float l2 = pe + ll * y;                                                /* lo is not tiny (it carries ln(1+f) - f): renormalise */
        float sh = ph + l2;
        float sl = (trm_abs(ph) >= trm_abs(l2)) ? (ph - sh) + l2 : (l2 - sh) + ph;
        r = trm_exp2(sh, sl);
    }
    return neg ? -r : r;
}

/* ---- acos ----------------------------------------------------------------------------------------------------------------
 * asin on [0, 0.5] by a degree-4 polynomial in x^2; |x| > 0.5 folded through acos(x) = 2 asin(sqrt((1-x)/2)).  |x| > 1 -> NaN
 * (the reference relies on it: acos(1 + ulp) poisons a sky sample). */
TRM_FN float trm_asin_poly(float x) {      /* 0 <= x <= 0.5 */
    float z = x * x;
    float p = 4.2163199048e-2f * z + 2.4181311049e-2f;
    p = p * z + 4.5470025998e-2f;
    p = p * z + 7.4953002686e-2f;
    p = p * z + 1.6666752422e-1f;
    return (p * z) * x + x;
}
TRM_FN float tr_acosf(float x) {
    if (x != x) return x;
    float ax = trm_abs(x);
    if (ax > 1.0f) return trm_nan();
    if (ax > 0.5f) {
        float s = trm_sqrt(0.5f * (1.0f - ax));
        float a = 2.0f * trm_asin_poly(s);
        return (trm_f2i(x) < 0) ? 3.14159265358979f - a : a;
    }
    float a = trm_asin_poly(ax);
    return (trm_f2i(x) < 0) ? 1.5707963267948966f + a : 1.5707963267948966f - a;
}

/* ---- atan2 --------------------------------------------------------------------------------------------------------------- */
TRM_FN float trm_atan_pos(float x) {       /* x >= 0 */
    float y;
    if (x > 2.414213562373095f) { y = 1.5707963267948966f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483f; x = (x - 1.0f) / (x + 1.0f); }
    else y = 0.0f;
    float z = x * x;
    float p = 8.05374449538e-2f * z - 1.38776856032e-1f;
    p = p * z + 1.99777106478e-1f;
    p = p * z - 3.33329491539e-1f;
    return y + ((p * z) * x + x);
}
TRM_FN float tr_atan2f(float y, float x) {
    if (x != x || y != y) return trm_nan();
    const float PI = 3.14159265358979f, PIO2 = 1.5707963267948966f;
    if (y == 0.0f) {
        float r = (trm_f2i(x) < 0) ? PI : 0.0f;                                /* atan2(+-0, -x) = +-pi, atan2(+-0, +x) = +-0 */
        return (trm_f2i(y) < 0) ? -r : r;
    }
    if (x == 0.0f) return (y > 0.0f) ? PIO2 : -PIO2;
    float ay = trm_abs(y), ax = trm_abs(x);
    float a;
    if (ay > 3.0e38f) a = (ax > 3.0e38f) ? 0.7853981633974483f : PIO2;
    else if (ax > 3.0e38f) a = 0.0f;
    else a = trm_atan_pos(ay / ax);
    if (trm_f2i(x) < 0) a = PI - a;
    return (trm_f2i(y) < 0) ? -a : a;
}

#endif /* TRMATH_H */
