// common.cuh — device math shared by the libtiray kernels (sm_100a).
//
// Arithmetic contract: every kernel file is compiled with -fmad=false, IEEE division and sqrt
// (nvcc defaults -prec-div=true -prec-sqrt=true -ftz=false), so that +,-,*,/,sqrt produce the same
// f32 bits as the strict-IEEE CPU oracle; only libm calls (sin/cos/pow/exp/atan2/acos) may differ
// by their documented ULP error.  Operation ORDER follows the reference expressions cited at each
// function (file:line in the reference tree).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/trmath.h"

#define TR_INF 1000000.0f        // UtilsFunc.py:38 INF_VALUE
#define TR_PI_REF 3.1415956f     // UtilsFunc.py:37 M_PIf (sic)
#define TR_PI_ENV 3.1415926f     // Scene.py:319,343; integrator/PT_RGB.py:129-130

#define TR_MAT_DISNEY 0
#define TR_MAT_GLASS 1
#define TR_MAT_LIGHT 2
#define TR_PRIM_TRI 1
#define TR_SHAPE_SPHERE 1
#define TR_SHAPE_SPOT 3
#define TR_SHAPE_LASER 4

struct V3 { float x, y, z; };
__host__ __device__ __forceinline__ V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float length3(V3 a) { return sqrtf(dot3(a, a)); }
// Taichi 0.7.14 Vector.normalized(): (1/norm) * v
__device__ __forceinline__ V3 normalize3(V3 a) { float inv = 1.0f / length3(a); return inv * a; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ V3 mix3(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float signf_(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }
__device__ __forceinline__ V3 f4xyz(float4 a) { return mk3(a.x, a.y, a.z); }

// ---- Philox4x32-10, counter (pixel, frame, block, 0), key = seed (shared spec with the oracle)
struct U4 { uint32_t x, y, z, w; };
__device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        U4 n;
        n.x = hi1 ^ c.y ^ k0; n.y = lo1; n.z = hi0 ^ c.w ^ k1; n.w = lo0;
        c = n; k0 += W0; k1 += W1;
    }
    return c;
}
__device__ __forceinline__ float u01(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ float4 rng4(uint64_t seed, uint32_t pixel, uint32_t frame, uint32_t block) {
    U4 c; c.x = pixel; c.y = frame; c.z = block; c.w = 0u;
    U4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return make_float4(u01(r.x), u01(r.y), u01(r.z), u01(r.w));
}

// ---- sampling / BRDF helpers
// UtilsFunc.py:348-350
__device__ __forceinline__ float cosine_hemisphere_pdf(float c) { return fmaxf(0.01f, c / TR_PI_REF); }
// UtilsFunc.py:352-360
__device__ __forceinline__ V3 cosine_sample_hemisphere(float u1, float u2) {
    float r = sqrtf(u1), phi = 2.0f * TR_PI_REF * u2;
    V3 p; p.x = r * tr_cosf(phi); p.y = r * tr_sinf(phi);
    p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
    return normalize3(p);
}
// UtilsFunc.py:373-387
__device__ __forceinline__ V3 inverse_transform(V3 dir, V3 N) {
    V3 Nn = normalize3(N), B;
    if (fabsf(Nn.x) > fabsf(Nn.z)) B = mk3(-Nn.y, Nn.x, 0.0f); else B = mk3(0.0f, -Nn.z, Nn.y);
    B = normalize3(B);
    V3 T = normalize3(cross3(B, Nn));
    return (dir.x * T + dir.y * B) + dir.z * Nn;
}
__device__ __forceinline__ float schlick_fresnel(float u) { float m = clampf(1.0f - u, 0.0f, 1.0f); float m2 = m * m; return m2 * m2 * m; }
__device__ __forceinline__ float gtr2(float NDotH, float a) { float a2 = a * a; float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH; return a2 / (TR_PI_REF * t * t); }
__device__ __forceinline__ float smithg_ggx(float NDotv, float alphaG) { float a = alphaG * alphaG, b = NDotv * NDotv; return 1.0f / (NDotv + sqrtf(a + b - a * b)); }
__device__ __forceinline__ V3 reflect3(V3 I, V3 N) { return I - 2.0f * dot3(N, I) * N; }
// UtilsFunc.py:417-426
__device__ __forceinline__ V3 refract3(V3 I, V3 N, float eta, float& suc) {
    suc = -1.0f; float NI = dot3(N, I); float k = 1.0f - eta * eta * (1.0f - NI * NI);
    V3 R = mk3(0.0f, 0.0f, 0.0f);
    if (k > 0.0f) { R = eta * I - (eta * NI + sqrtf(k)) * N; suc = 1.0f; }
    return R;
}
// UtilsFunc.py:428-438
__device__ __forceinline__ float schlick_r(float cosine, float ior) { float r0 = (1.0f - ior) / (1.0f + ior); r0 = r0 * r0; return r0 + (1.0f - r0) * tr_powf(1.0f - cosine, 5.0f); }
__device__ __forceinline__ float power_heuristic(float a, float b) { float t = a * a; return t / (b * b + t); }
// UtilsFunc.py:440-461: integer-ULP nudge along the normal
__device__ __forceinline__ V3 offset_ray(V3 p, V3 n) {
    const float int_scale = 256.0f, float_scale = 1.0f / 2048.0f, origin = 1.0f / 256.0f;
    float pp[3] = {p.x, p.y, p.z}, nn[3] = {n.x, n.y, n.z}, r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int i_of = (int)(int_scale * nn[k]);
        int i_p = __float_as_int(pp[k]);
        i_p = (pp[k] < 0.0f) ? i_p - i_of : i_p + i_of;
        r[k] = (fabsf(pp[k]) < origin) ? pp[k] + float_scale * nn[k] : __int_as_float(i_p);
    }
    return mk3(r[0], r[1], r[2]);
}
// UtilsFunc.py:76-94,113-120
__device__ __forceinline__ float srgb_to_lrgb1(float c) { return c < 0.04045f ? c / 12.92f : tr_powf((c + 0.055f) / 1.055f, 2.4f); }
__device__ __forceinline__ V3 srgb_to_lrgb(V3 c) { return mk3(srgb_to_lrgb1(c.x), srgb_to_lrgb1(c.y), srgb_to_lrgb1(c.z)); }
__device__ __forceinline__ float lrgb_to_srgb1(float c) { float r = c < 0.0031308f ? c * 12.92f : 1.055f * tr_powf(c, 1.0f / 2.4f) - 0.055f; return clampf(r, 0.0f, 1.0f); }
__device__ __forceinline__ float tone_aces1(float x) { const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f; return clampf((x * (a * x + b)) / (x * (c * x + d) + e), 0.0f, 1.0f); }

// brdf/Disney.py:65-108
__device__ __forceinline__ void disney_evaluate_pdf(V3 N, V3 V, V3 L, float metal, float rough, float& out, float& pdf) {
    out = 0.0f; pdf = -1.0f;
    float NDotL = dot3(N, L), NDotV = dot3(N, V);
    if (NDotL > 0.0f && NDotV > 0.0f) {
        V3 H = normalize3(L + V);
        float NDotH = dot3(H, N), LDotH = dot3(H, L);
        float Cspec0 = mixf(0.04f, 1.0f, metal), Csheen = 0.5f;
        float FL = schlick_fresnel(NDotL), FV = schlick_fresnel(NDotV);
        float Fd90 = 0.5f + 2.0f * LDotH * LDotH * rough;
        float Fd = mixf(1.0f, Fd90, FL) * mixf(1.0f, Fd90, FV);
        float alpha = fmaxf(0.001f, rough);
        float Ds = gtr2(NDotH, alpha);
        float FH = schlick_fresnel(LDotH);
        float Fs = mixf(Cspec0, 1.0f, FH);
        float rg = rough * 0.5f + 0.5f; rg = rg * rg;
        float Gs = smithg_ggx(NDotL, rg) * smithg_ggx(NDotV, rg);
        float Fsheen = FH * Csheen;
        out = (Fsheen + 1.0f / TR_PI_REF) * Fd * (1.0f - metal) + Gs * Fs * Ds;
        float dr = 0.5f * (1.0f - metal), sr = 1.0f - dr;
        float pdfGTR2 = Ds * NDotH, pdfSpec = pdfGTR2 / (4.0f * fabsf(LDotH)), pdfDiff = 1.0f / TR_PI_REF;
        pdf = dr * pdfDiff + sr * pdfSpec;
    }
}
// brdf/Disney.py:42-63
__device__ __forceinline__ float disney_pdf(V3 N, V3 V, V3 L, float metal, float rough) {
    float pdf = 0.0f;
    float NDotL = dot3(N, L), NDotV = dot3(N, V);
    if (NDotL > 0.0f && NDotV > 0.0f) {
        V3 H = normalize3(L + V);
        float NDotH = dot3(H, N), LDotH = dot3(H, L);
        float alpha = fmaxf(0.001f, rough);
        float Ds = gtr2(NDotH, alpha);
        float dr = 0.5f * (1.0f - metal), sr = 1.0f - dr;
        float pdfGTR2 = Ds * NDotH, pdfSpec = pdfGTR2 / (4.0f * fabsf(LDotH)), pdfDiff = 1.0f / TR_PI_REF;
        pdf = dr * pdfDiff + sr * pdfSpec;
    }
    return pdf;
}
// brdf/Disney.py:17-40 (randoms: lobe probability, r1, r2)
__device__ __forceinline__ V3 disney_sample(V3 dir, V3 N, float metal, float rough, float prob, float r1, float r2) {
    float dr = 0.5f * (1.0f - metal), alpha = fmaxf(0.001f, rough);
    V3 next;
    if (prob < dr) {
        next = inverse_transform(cosine_sample_hemisphere(r1, r2), N);
    } else {
        float phi = r1 * 2.0f * TR_PI_REF;
        float cosT = sqrtf((1.0f - r2) / (1.0f + (alpha * alpha - 1.0f) * r2));
        float sinT = sqrtf(1.0f - (cosT * cosT));
        float sinP = tr_sinf(phi), cosP = tr_cosf(phi);
        V3 half = inverse_transform(mk3(sinT * cosP, sinT * sinP, cosT), N);
        next = reflect3(dir, half);
    }
    return next;
}
// brdf/Glass.py:9-34 (random: Fresnel probability)
__device__ __forceinline__ V3 glass_sample(V3 dir, V3 N, float ior, float prob, float& f_or_b) {
    float cos_i = dot3(dir, N), eta = ior; f_or_b = 1.0f; float R = prob + 1.0f;
    if (cos_i > 0.0f) N = -N; else { cos_i = -cos_i; eta = 1.0f / ior; }
    float suc; V3 next = refract3(dir, N, eta, suc);
    if (suc > 0.0f) R = schlick_r(cos_i, ior);
    if (prob < R) next = reflect3(dir, N); else f_or_b = -1.0f;
    return next;
}
// Scene.py:315-322
__device__ __forceinline__ V3 uniform_sample_sphere(float u1, float u2) {
    float z = 1.0f - 2.0f * u1;
    float r = sqrtf(clampf(1.0f - z * z, 0.0f, 1.0f));
    float phi = 2.0f * TR_PI_ENV * u2;
    return mk3(r * tr_cosf(phi), r * tr_sinf(phi), z);
}

// ---- Morton (UtilsFunc.py:538-580)
__device__ __forceinline__ int expand_bits(int x) {
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}
__device__ __forceinline__ int morton3d(float x, float y, float z) {
    x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    y = fminf(fmaxf(y * 1024.0f, 0.0f), 1023.0f);
    z = fminf(fmaxf(z * 1024.0f, 0.0f), 1023.0f);
    return expand_bits((int)x) | (expand_bits((int)y) << 1) | (expand_bits((int)z) << 2);
}
// UtilsFunc.py:555-566: 32 - (bit length of a^b); 32 for identical codes. Codes are < 2^30.
__device__ __forceinline__ int common_upper_bits(int a, int b) { return __clz(a ^ b); }
