// spectral.cu — tables of the hero-wavelength spectral integrator behind the C-ABI (include/tiray.h, "spectral" block):
// uploads that replace the from_numpy calls of PT_Spec.setup_data_gpu (integrator/PT_Spec.py:89-98), the white-point
// normalisation (:101-107,174-187), the per-material coefficient table used by the shade kernels, and unit hooks.
#include <string.h>
#include <vector>
#include "ctx.h"
#include "spectral.cuh"

static inline int cdiv_(int a, int b) { return (a + b - 1) / b; }

extern "C" int tr_spec_sensor_upload(tr_ctx* ctx, const float* xyz, int n, float lmin, float lmax) {
    if (!ctx || !xyz || n < 2 || !(lmax > lmin)) return tr_fail(ctx, TR_ERR_INVALID, "tr_spec_sensor_upload: bad table (n=%d, %g..%g nm)", n, lmin, lmax);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc; if ((rc = tr_realloc(ctx, &ctx->d_sensor, (size_t)n))) return rc;
    std::vector<float4> h((size_t)n);
    for (int i = 0; i < n; ++i) h[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0f);
    if ((rc = tr_stage_h2d(ctx, ctx->d_sensor, h.data(), (size_t)n * sizeof(float4)))) return rc;
    ctx->spec.sensor = ctx->d_sensor; ctx->spec.s_size = n; ctx->spec.s_lmin = lmin; ctx->spec.s_lmax = lmax;
    ctx->spec.s_lrange = (lmax - lmin) / (float)(n - 1);          // integrator/PT_Spec.py:74
    ctx->gen++;
    return TR_OK;
}

extern "C" int tr_spec_spectrum_upload(tr_ctx* ctx, int which, const float* data, int n, float lmin, float lmax) {
    if (!ctx || which < 0 || which > 3 || !data || n < 2 || !(lmax > lmin))
        return tr_fail(ctx, TR_ERR_INVALID, "tr_spec_spectrum_upload: bad table %d (n=%d, %g..%g nm)", which, n, lmin, lmax);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc; if ((rc = tr_realloc(ctx, &ctx->d_spectrum[which], (size_t)n))) return rc;
    if ((rc = tr_stage_h2d(ctx, ctx->d_spectrum[which], data, (size_t)n * 4))) return rc;
    SpecTable& t = ctx->spec.sp[which];
    t.data = ctx->d_spectrum[which]; t.size = n; t.lmin = lmin; t.lmax = lmax; t.lrange = (lmax - lmin) / (float)(n - 1);   // spectrum/Spectrum.py:33
    ctx->gen++;
    return TR_OK;
}

extern "C" int tr_spec_spectrum_download(tr_ctx* ctx, int which, float* data) {
    if (!ctx || which < 0 || which > 3 || !data || !ctx->d_spectrum[which]) return tr_fail(ctx, TR_ERR_INVALID, "tr_spec_spectrum_download: table %d not uploaded", which);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    TR_CUDA(ctx, cudaMemcpyAsync(data, ctx->d_spectrum[which], (size_t)ctx->spec.sp[which].size * 4, cudaMemcpyDeviceToHost, ctx->stream));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

extern "C" int tr_spec_rgb2spec_upload(tr_ctx* ctx, const float* scale, const float* data, int res) {
    if (!ctx || !scale || !data || res < 2 || res > 1024) return tr_fail(ctx, TR_ERR_INVALID, "tr_spec_rgb2spec_upload: bad table (res=%d)", res);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)res * res * res * 9;
    int rc;
    if ((rc = tr_realloc(ctx, &ctx->d_rs_scale, (size_t)res))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_rs_data, n))) return rc;
    if ((rc = tr_stage_h2d(ctx, ctx->d_rs_scale, scale, (size_t)res * 4))) return rc;
    if ((rc = tr_stage_h2d(ctx, ctx->d_rs_data, data, n * 4))) return rc;
    ctx->rs_res = res; ctx->matspec_ready = false; ctx->gen++;
    return TR_OK;
}

extern "C" int tr_spec_sky_upload(tr_ctx* ctx, const float* configs, const float* radiances, const float sun_dir[3]) {
    if (!ctx || !configs || !radiances || !sun_dir) return tr_fail(ctx, TR_ERR_INVALID, "tr_spec_sky_upload: NULL table");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc; if ((rc = tr_realloc(ctx, &ctx->d_sky, (size_t)TR_SKY_FLOATS))) return rc;
    float h[TR_SKY_FLOATS];
    memcpy(h, configs, 99 * 4); memcpy(h + 99, radiances, 11 * 4); memcpy(h + 110, sun_dir, 12);
    if ((rc = tr_stage_h2d(ctx, ctx->d_sky, h, sizeof(h)))) return rc;
    ctx->spec.sky = ctx->d_sky; ctx->spec.sky_on = 1; ctx->gen++;
    return TR_OK;
}

// PT_Spec.cal_white_point (integrator/PT_Spec.py:174-187): Simpson 3/8 weights over the sensor grid.  The reference
// sums with atomic adds in scheduling order; here one thread sums in index order (470 terms), like the oracle.
__global__ void k_white_point(SpecDev sd, int which, float* __restrict__ wp) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    float X = 0.0f, Y = 0.0f, Z = 0.0f;
    for (int i = 0; i < sd.s_size; ++i) {
        float L = sd.s_lmin + (float)i * sd.s_lrange;
        float hh = (sd.s_lmax - sd.s_lmin) / (float)(sd.s_size - 1);
        float weight = 3.0f / 8.0f * hh;
        if (i == 0 || i == sd.s_size - 1) weight = weight;
        else if ((i - 1) % 3 == 2) weight = weight * 2.0f;
        else weight = weight * 3.0f;
        float v = spectrum_sample(sd.sp[which], L);
        float4 s = sd.sensor[i];
        X += (s.x * v) * weight; Y += (s.y * v) * weight; Z += (s.z * v) * weight;
    }
    wp[0] = X; wp[1] = Y; wp[2] = Z;
}
// Spectrum.scale (spectrum/Spectrum.py:53-56)
__global__ void k_scale(float* __restrict__ d, int n, float coff) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] *= coff;
}

extern "C" int tr_spec_normalize(tr_ctx* ctx, int which, float white_point[3]) {
    if (!ctx || which < 0 || which > 3 || !ctx->d_spectrum[which] || !ctx->d_sensor)
        return tr_fail(ctx, TR_ERR_INVALID, "tr_spec_normalize: sensor or spectrum %d not uploaded", which);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc; if ((rc = tr_realloc(ctx, &ctx->d_white_point, (size_t)4))) return rc;
    k_white_point<<<1, 32, 0, ctx->stream>>>(ctx->spec, which, ctx->d_white_point);
    TR_CHECK_LAUNCH(ctx);
    float wp[3];
    TR_CUDA(ctx, cudaMemcpyAsync(wp, ctx->d_white_point, 12, cudaMemcpyDeviceToHost, ctx->stream));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!(wp[1] > 0.0f)) return tr_fail(ctx, TR_ERR_INVALID, "tr_spec_normalize: white point Y = %g", wp[1]);
    const float coff = (float)(1.0 / (double)wp[1]);              // host f64 division, passed as ti.f32 (:104-105)
    k_scale<<<cdiv_(ctx->spec.sp[which].size, 256), 256, 0, ctx->stream>>>(ctx->d_spectrum[which], ctx->spec.sp[which].size, coff);
    TR_CHECK_LAUNCH(ctx);
    if (white_point) { white_point[0] = wp[0]; white_point[1] = wp[1]; white_point[2] = wp[2]; }
    return TR_OK;
}

// per-material spectral coefficients: what Hero.srgb_to_spec (spectrum/HeroSample.py:46-57) fetches for the material
// colour in get_spec_power (integrator/PT_Spec.py:119-135) and emission_to_rad (:110-117)
__global__ void k_matspec(const float* __restrict__ material, int nm, const float* __restrict__ rs_scale, const float* __restrict__ rs_data, int res,
                          float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nm) return;
    const float* m = material + (size_t)i * 10;
    int mt = (int)m[0], tex = (int)m[1];
    V3 col = mk3(m[2], m[3], m[4]);
    float4 refl;
    if (mt == TR_MAT_SPECTRAL) refl = make_float4(0.0f, 0.0f, 0.0f, __int_as_float((tex >= 0 && tex <= 2) ? 1 + tex : -1));
    else { V3 c = rs_fetch(rs_scale, rs_data, res, srgb_to_lrgb(col)); refl = make_float4(c.x, c.y, c.z, __int_as_float(0)); }
    float scale = length3(col);
    float4 tint = make_float4(0.0f, 0.0f, 0.0f, scale);
    if (scale > 0.0f) { V3 c = rs_fetch(rs_scale, rs_data, res, srgb_to_lrgb(col / scale)); tint = make_float4(c.x, c.y, c.z, scale); }
    out[2 * i] = refl; out[2 * i + 1] = tint;
}

int tr_spec_prepare(tr_ctx* ctx) {
    if (!ctx->d_sensor) return tr_fail(ctx, TR_ERR_INVALID, "PT_Spec: sensor table not uploaded (tr_spec_sensor_upload)");
    for (int k = 0; k < 4; ++k) if (!ctx->d_spectrum[k]) return tr_fail(ctx, TR_ERR_INVALID, "PT_Spec: spectrum %d not uploaded (tr_spec_spectrum_upload)", k);
    if (!ctx->d_rs_data) return tr_fail(ctx, TR_ERR_INVALID, "PT_Spec: rgb2spec table not uploaded (tr_spec_rgb2spec_upload)");
    if (!ctx->d_sky) return tr_fail(ctx, TR_ERR_INVALID, "PT_Spec: sky tables not uploaded (tr_spec_sky_upload)");
    if (!ctx->matspec_ready) {
        int rc; if ((rc = tr_realloc(ctx, &ctx->d_matspec, (size_t)ctx->nm * 2))) return rc;
        k_matspec<<<cdiv_(ctx->nm, 64), 64, 0, ctx->stream>>>(ctx->d_material, ctx->nm, ctx->d_rs_scale, ctx->d_rs_data, ctx->rs_res, ctx->d_matspec);
        TR_CHECK_LAUNCH(ctx);
        ctx->matspec_ready = true; ctx->gen++;
    }
    ctx->spec.matspec = ctx->d_matspec;
    return TR_OK;
}

// ------------------------------------------------------------------ unit hooks (host pointers in, host pointers out)
__global__ void k_test_srgb_to_spec(const float* __restrict__ rs_scale, const float* __restrict__ rs_data, int res, int n,
                                    const float* __restrict__ rgb, const float* __restrict__ lambda0, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3 c = rs_fetch(rs_scale, rs_data, res, srgb_to_lrgb(mk3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2])));
    V4 r = rs_eval_hero(make_float4(c.x, c.y, c.z, 0.0f), lambda0[i]);
    for (int k = 0; k < 4; ++k) out[4 * i + k] = r.v[k];
}
__global__ void k_test_sky(SpecDev sd, int n, const float* __restrict__ theta, const float* __restrict__ gamma, const float* __restrict__ wl, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sky_radiance(sd, theta[i], gamma[i], wl[i]);
}
__global__ void k_test_spectrum(SpecDev sd, int which, int n, const float* __restrict__ lambda, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (which >= 0) out[i] = spectrum_sample(sd.sp[which], lambda[i]);
    else { V3 s = sensor_sample(sd, lambda[i]); out[3 * i] = s.x; out[3 * i + 1] = s.y; out[3 * i + 2] = s.z; }
}

namespace {
struct DevBuf {
    float* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t up(const float* h, size_t n, cudaStream_t s) {
        cudaError_t e = cudaMalloc((void**)&p, n * 4);
        if (e == cudaSuccess && h) e = cudaMemcpyAsync(p, h, n * 4, cudaMemcpyHostToDevice, s);
        return e;
    }
};
}

extern "C" int tr_test_srgb_to_spec(tr_ctx* ctx, int n, const float* rgb, const float* lambda0, float* out4) {
    if (!ctx || n <= 0 || !rgb || !lambda0 || !out4 || !ctx->d_rs_data) return tr_fail(ctx, TR_ERR_INVALID, "tr_test_srgb_to_spec: bad arguments or rgb2spec table missing");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf a, b, o; cudaStream_t s = ctx->stream;
    TR_CUDA(ctx, a.up(rgb, (size_t)n * 3, s)); TR_CUDA(ctx, b.up(lambda0, (size_t)n, s)); TR_CUDA(ctx, o.up(nullptr, (size_t)n * 4, s));
    k_test_srgb_to_spec<<<cdiv_(n, 128), 128, 0, s>>>(ctx->d_rs_scale, ctx->d_rs_data, ctx->rs_res, n, a.p, b.p, o.p);
    TR_CHECK_LAUNCH(ctx);
    TR_CUDA(ctx, cudaMemcpyAsync(out4, o.p, (size_t)n * 16, cudaMemcpyDeviceToHost, s));
    TR_CUDA(ctx, cudaStreamSynchronize(s));
    return TR_OK;
}

extern "C" int tr_test_sky_radiance(tr_ctx* ctx, int n, const float* theta, const float* gamma, const float* wavelength, float* out) {
    if (!ctx || n <= 0 || !theta || !gamma || !wavelength || !out || !ctx->d_sky) return tr_fail(ctx, TR_ERR_INVALID, "tr_test_sky_radiance: bad arguments or sky tables missing");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf a, b, c, o; cudaStream_t s = ctx->stream;
    TR_CUDA(ctx, a.up(theta, (size_t)n, s)); TR_CUDA(ctx, b.up(gamma, (size_t)n, s)); TR_CUDA(ctx, c.up(wavelength, (size_t)n, s)); TR_CUDA(ctx, o.up(nullptr, (size_t)n, s));
    k_test_sky<<<cdiv_(n, 128), 128, 0, s>>>(ctx->spec, n, a.p, b.p, c.p, o.p);
    TR_CHECK_LAUNCH(ctx);
    TR_CUDA(ctx, cudaMemcpyAsync(out, o.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    TR_CUDA(ctx, cudaStreamSynchronize(s));
    return TR_OK;
}

// which 0..3: Spectrum.sample of that table (out n); which -1: the CIE sensor (out n x 3)
extern "C" int tr_test_spectrum_sample(tr_ctx* ctx, int which, int n, const float* lambda, float* out) {
    if (!ctx || n <= 0 || !lambda || !out || which < -1 || which > 3 || (which >= 0 ? !ctx->d_spectrum[which] : !ctx->d_sensor))
        return tr_fail(ctx, TR_ERR_INVALID, "tr_test_spectrum_sample: bad arguments or table %d missing", which);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t no = (size_t)n * (which < 0 ? 3 : 1);
    DevBuf a, o; cudaStream_t s = ctx->stream;
    TR_CUDA(ctx, a.up(lambda, (size_t)n, s)); TR_CUDA(ctx, o.up(nullptr, no, s));
    k_test_spectrum<<<cdiv_(n, 128), 128, 0, s>>>(ctx->spec, which, n, a.p, o.p);
    TR_CHECK_LAUNCH(ctx);
    TR_CUDA(ctx, cudaMemcpyAsync(out, o.p, no * 4, cudaMemcpyDeviceToHost, s));
    TR_CUDA(ctx, cudaStreamSynchronize(s));
    return TR_OK;
}
