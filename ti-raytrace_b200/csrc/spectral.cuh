// spectral.cuh — device functions of the hero-wavelength spectral path (integrator/PT_Spec.py) shared by the shade,
// tail and accumulate kernels: tabulated spectra (spectrum/Spectrum.py), the CIE observer (PT_Spec.sample), the
// Jakob-Hanika RGB -> spectrum model (spectrum/Rgb2Spec.py), hero sampling (spectrum/HeroSample.py), the Sellmeier
// glass (UtilsFunc.py:481-484) and the Hosek-Wilkie sky dome (sky/Sky.py:206-265).  Same arithmetic contract as
// common.cuh: operation order follows the reference expressions, only libm calls may differ from the CPU oracle.
#pragma once
#include "common.cuh"

#define HERO_N 4                      // spectrum/HeroSample.py:5
#define HERO_LAMBDA_MIN 360.0f        // :6
#define HERO_LAMBDA_MAX 760.0f        // :7
#define HERO_STEP ((HERO_LAMBDA_MAX - HERO_LAMBDA_MIN) / (float)HERO_N)
#define TR_MAT_SPECTRAL 10            // SceneData.py:53
#define TR_SPEC_D65 0
#define TR_SPEC_WHITE 1
#define TR_SPEC_RED 2
#define TR_SPEC_GREEN 3
#define TR_SKY_FLOATS (11 * 9 + 11 + 3)

struct SpecTable { const float* data; int size; float lmin, lmax, lrange; };

// everything the spectral kernels read; lives by value inside the kernel arguments
struct SpecDev {
    const float4* sensor; int s_size; float s_lmin, s_lmax, s_lrange;   // CIE 1931 colour matching functions (x, y, z, 0)
    SpecTable sp[4];                                                     // d65, white, red, green
    const float4* matspec;     // per material: (reflectance coefficients c0 c1 c2, kind) (tint coefficients c0 c1 c2, |colour|)
    const float* sky;          // configs 11 x 9, radiances 11, sun direction 3
    int sky_on; int pad_;
};

struct V4 { float v[4]; };
__device__ __forceinline__ V4 mk4(float a) { V4 r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = a; return r; }
__device__ __forceinline__ V4 ld4(float4 a) { V4 r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; return r; }
__device__ __forceinline__ float4 st4(V4 a) { return make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }
__device__ __forceinline__ V4 operator*(V4 a, V4 b) { V4 r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] * b.v[i]; return r; }
__device__ __forceinline__ V4 operator*(V4 a, float s) { V4 r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] * s; return r; }
__device__ __forceinline__ V4 operator*(float s, V4 a) { return a * s; }
__device__ __forceinline__ V4 operator/(V4 a, float s) { V4 r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] / s; return r; }
__device__ __forceinline__ V4 operator+(V4 a, V4 b) { V4 r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] + b.v[i]; return r; }
__device__ __forceinline__ float sum4(V4 a) { return ((a.v[0] + a.v[1]) + a.v[2]) + a.v[3]; }
__device__ __forceinline__ float fractf_(float x) { return x - floorf(x); }            // taichi_glsl fract

// spectrum/Spectrum.py:43-51.  The blend weight is fract(offset), not fract(offset / step): kept.  At Lambda ==
// lambda_max the reference reads one entry past the table; the index is clamped here.
__device__ __forceinline__ float spectrum_sample(const SpecTable& s, float L) {
    float ret = 0.0f;
    if (L >= s.lmin && L <= s.lmax) {
        float off = L - s.lmin; int idx = (int)(off / s.lrange); float w = fractf_(off);
        int i1 = min(idx + 1, s.size - 1);
        ret = mixf(__ldg(s.data + idx), __ldg(s.data + i1), w);
    }
    return ret;
}
// spectrum/HeroSample.py:10-16
__device__ __forceinline__ V4 hero_sample(const SpecTable& s, float L0) {
    V4 r;
#pragma unroll
    for (int i = 0; i < HERO_N; ++i) r.v[i] = spectrum_sample(s, L0 + (float)i * HERO_STEP);
    return r;
}
// integrator/PT_Spec.py:138-146
__device__ __forceinline__ V3 sensor_sample(const SpecDev& sd, float L) {
    V3 r = mk3(0.0f, 0.0f, 0.0f);
    if (L >= sd.s_lmin && L <= sd.s_lmax) {
        float off = L - sd.s_lmin; int idx = (int)(off / sd.s_lrange); float w = fractf_(off);
        int i1 = min(idx + 1, sd.s_size - 1);
        float4 a = __ldg(sd.sensor + idx), b = __ldg(sd.sensor + i1);
        r = mk3(mixf(a.x, b.x, w), mixf(a.y, b.y, w), mixf(a.z, b.z, w));
    }
    return r;
}

// ---- spectrum/Rgb2Spec.py:44-138
__device__ __forceinline__ float rs_fma(float a, float b, float c) { return a * b + c; }     // compiled with -fmad=false: two roundings, like the oracle
__device__ __forceinline__ int rs_find_interval(const float* __restrict__ scale, int size, float x) {
    int left = 0, last_interval = size - 2; size = last_interval;
    while (size > 0) {
        int half = size >> 1, middle = left + half + 1;
        if (scale[middle] <= x) { left = middle; size -= half + 1; } else size = half;
    }
    return min(left, last_interval);
}
__device__ __forceinline__ V3 rs_fetch(const float* __restrict__ scale_t, const float* __restrict__ d, int res, V3 rgb) {
    float c[3] = {clampf(rgb.x, 0.0f, 1.0f), clampf(rgb.y, 0.0f, 1.0f), clampf(rgb.z, 0.0f, 1.0f)};
    int index = 0; float xyz[3] = {c[0], c[1], c[2]};
    if (c[1] > c[0]) {
        if (c[2] > c[1]) index = 2;
        else { index = 1; xyz[0] = c[2]; xyz[1] = c[0]; xyz[2] = c[1]; }
    } else {
        if (c[2] > c[0]) index = 2;
        else { index = 0; xyz[0] = c[1]; xyz[1] = c[2]; xyz[2] = c[0]; }
    }
    xyz[2] = fmaxf(0.00001f, xyz[2]);
    float scale = (float)(res - 1) / xyz[2];
    float x = xyz[0] * scale, y = xyz[1] * scale, z = xyz[2];
    const int dx = 3, dy = 3 * res, dz = 3 * res * res;
    int xi = (int)fminf(x, (float)(res - 2)), yi = (int)fminf(y, (float)(res - 2)), zi = rs_find_interval(scale_t, res, z);
    int offset = (((index * res + zi) * res + yi) * res + xi) * 3;
    float x0 = x - (float)xi, y0 = y - (float)yi, z0 = (z - scale_t[zi]) / (scale_t[zi + 1] - scale_t[zi]);
    float out[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        int i = offset + j;
        out[j] = mixf(mixf(mixf(d[i], d[i + dx], x0), mixf(d[i + dy], d[i + dy + dx], x0), y0),
                      mixf(mixf(d[i + dz], d[i + dz + dx], x0), mixf(d[i + dy + dz], d[i + dx + dy + dz], x0), y0), z0);
    }
    return mk3(out[0], out[1], out[2]);
}
__device__ __forceinline__ float rs_eval(float c0, float c1, float c2, float L) {
    float x = rs_fma(rs_fma(c0, L, c1), L, c2);
    float y = 1.0f / sqrtf(rs_fma(x, x, 1.0f));
    return rs_fma(0.5f * x, y, 0.5f);
}
// spectrum/HeroSample.py:46-57 with the coefficient fetch hoisted to a per-material table (it depends on the
// material colour only; same device function, same bits)
__device__ __forceinline__ V4 rs_eval_hero(float4 c, float L0) {
    V4 r;
#pragma unroll
    for (int i = 0; i < HERO_N; ++i) r.v[i] = rs_eval(c.x, c.y, c.z, L0 + (float)i * HERO_STEP);
    return r;
}
// integrator/PT_Spec.py:119-135: kind 0 = RGB colour through rgb2spec, 1..3 = measured white / red / green table,
// -1 = MAT_SPECTRAL with an unknown table index (zero reflectance)
__device__ __forceinline__ V4 get_spec_power(const SpecDev& sd, int mat, float L0) {
    float4 c = __ldg(sd.matspec + 2 * mat);
    int kind = __float_as_int(c.w);
    if (kind == 0) return rs_eval_hero(c, L0);
    if (kind > 0) return hero_sample(sd.sp[kind], L0);
    return mk4(0.0f);
}
// integrator/PT_Spec.py:110-117
__device__ __forceinline__ V4 emission_to_rad(const SpecDev& sd, int mat, float L0) {
    float4 c = __ldg(sd.matspec + 2 * mat + 1);
    V4 r = mk4(0.0f);
    if (c.w > 0.0f) r = rs_eval_hero(c, L0);
    return r * c.w;
}
// UtilsFunc.py:481-484
__device__ __forceinline__ float get_glass_ior(float L) {
    L = L / 1000.0f; float L2 = L * L;
    return sqrtf(1.0f + 1.03961212f * L2 / (L2 - 0.00600069867f) + 0.231792344f * L2 / (L2 - 0.0200179144f) + 1.01046945f * L2 / (L2 - 103.560653f));
}

// ---- sky/Sky.py:206-215,248-265 (sky-dome radiance; the direct-sun term is commented out in the reference)
__device__ __forceinline__ float sky_internal(const float* __restrict__ sky, int wl, float theta, float gamma) {
    const float* c = sky + wl * 9;
    float c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3), c4 = __ldg(c + 4), c5 = __ldg(c + 5), c6 = __ldg(c + 6), c7 = __ldg(c + 7), c8 = __ldg(c + 8);
    float cg = tr_cosf(gamma), ct = tr_cosf(theta);
    float expM = tr_expf(c4 * gamma), rayM = cg * cg;
    float mieM = (1.0f + cg * cg) / tr_powf((1.0f + c8 * c8 - 2.0f * c8 * cg), 1.5f);
    float zenith = sqrtf(ct);
    return (1.0f + c0 * tr_expf(c1 / (ct + 0.01f))) * (c2 + c3 * expM + c5 * rayM + c6 * mieM + c7 * zenith);
}
__device__ __forceinline__ float sky_radiance(const SpecDev& sd, float theta, float gamma, float wl) {
    float ret = 0.0f;
    if (sd.sky_on && wl >= 320.0f && wl <= 720.0f) {
        int low = (int)((wl - 320.0f) / 40.0f);
        if (low >= 0 && low < 11) {
            const float* rad = sd.sky + 99;
            float interp = fractf_((wl - 320.0f) / 40.0f);
            float val_low = sky_internal(sd.sky, low, theta, gamma) * __ldg(rad + low);
            if (interp < 1e-6f) ret = val_low;
            else { ret = (1.0f - interp) * val_low; if (low + 1 < 11) ret += interp * sky_internal(sd.sky, low + 1, theta, gamma) * __ldg(rad + low + 1); }
        }
    }
    return ret;
}

// integrator/PT_Spec.py:148-165: four radiance lanes -> CIE XYZ (Monte Carlo over the sensor range) -> linear sRGB,
// blended into the running mean
__device__ __forceinline__ void add_splat(const SpecDev& sd, V4 spec, float Lambda0, float coff, float& r, float& g, float& b) {
    V4 xf, yf, zf;
#pragma unroll
    for (int k = 0; k < HERO_N; ++k) { V3 s = sensor_sample(sd, Lambda0 + (float)k * HERO_STEP); xf.v[k] = s.x; yf.v[k] = s.y; zf.v[k] = s.z; }
    xf = xf * spec; yf = yf * spec; zf = zf * spec;
    float range = sd.s_lmax - sd.s_lmin;
    float X = sum4((xf * range) / (float)HERO_N), Y = sum4((yf * range) / (float)HERO_N), Z = sum4((zf * range) / (float)HERO_N);
    // UF.xyz_to_srgb (UtilsFunc.py:42)
    float lr = (3.240479f * X + -1.537150f * Y) + -0.498535f * Z;
    float lg = (-0.969256f * X + 1.875991f * Y) + 0.041556f * Z;
    float lb = (0.055648f * X + -0.204043f * Y) + 1.057311f * Z;
    r = mixf(r, lr, coff); g = mixf(g, lg, coff); b = mixf(b, lb, coff);
}
