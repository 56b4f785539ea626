// hooks.cu — unit hooks: the shading device functions run over arrays so that tests can compare them
// with the oracle one function at a time (brdf/Disney.py, brdf/Glass.py, UtilsFunc.offset_ray, RNG).
#include <vector>
#include "ctx.h"
#include "common.cuh"

enum { H_DISNEY_EVAL = 0, H_DISNEY_SAMPLE = 1, H_GLASS_SAMPLE = 2, H_OFFSET_RAY = 3, H_RNG = 4, H_MATH = 5 };

__global__ void k_hook(int op, int n, const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                       float p0, float p1, unsigned long long seed, unsigned u0, unsigned u1, unsigned u2, float* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (op == H_DISNEY_EVAL) {
        float o, p; disney_evaluate_pdf(mk3(a[k * 3], a[k * 3 + 1], a[k * 3 + 2]), mk3(b[k * 3], b[k * 3 + 1], b[k * 3 + 2]),
                                        mk3(c[k * 3], c[k * 3 + 1], c[k * 3 + 2]), p0, p1, o, p);
        out[k * 2] = o; out[k * 2 + 1] = p;
    } else if (op == H_DISNEY_SAMPLE) {
        V3 r = disney_sample(mk3(a[k * 3], a[k * 3 + 1], a[k * 3 + 2]), mk3(b[k * 3], b[k * 3 + 1], b[k * 3 + 2]), p0, p1,
                             c[k * 3], c[k * 3 + 1], c[k * 3 + 2]);
        out[k * 3] = r.x; out[k * 3 + 1] = r.y; out[k * 3 + 2] = r.z;
    } else if (op == H_GLASS_SAMPLE) {
        float fb; V3 r = glass_sample(mk3(a[k * 3], a[k * 3 + 1], a[k * 3 + 2]), mk3(b[k * 3], b[k * 3 + 1], b[k * 3 + 2]), p0, c[k], fb);
        out[k * 4] = r.x; out[k * 4 + 1] = r.y; out[k * 4 + 2] = r.z; out[k * 4 + 3] = fb;
    } else if (op == H_OFFSET_RAY) {
        V3 r = offset_ray(mk3(a[k * 3], a[k * 3 + 1], a[k * 3 + 2]), mk3(b[k * 3], b[k * 3 + 1], b[k * 3 + 2]));
        out[k * 3] = r.x; out[k * 3 + 1] = r.y; out[k * 3 + 2] = r.z;
    } else if (op == H_MATH) {
        // include/trmath.h: the function is selected by u0 (0 sin, 1 cos, 2 exp, 3 acos, 4 atan2(a, b), 5 pow(a, b))
        const float x = a[k], y = b ? b[k] : 0.0f;
        out[k] = u0 == 0 ? tr_sinf(x) : u0 == 1 ? tr_cosf(x) : u0 == 2 ? tr_expf(x) : u0 == 3 ? tr_acosf(x) : u0 == 4 ? tr_atan2f(x, y) : tr_powf(x, y);
    } else if (op == H_RNG) {
        float4 r = rng4(seed, u0, u1, u2);
        out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
    }
}

static int run_hook(tr_ctx* ctx, int op, int n, const float* a, size_t na, const float* b, size_t nb, const float* c, size_t nc,
                    float p0, float p1, unsigned long long seed, unsigned u0, unsigned u1, unsigned u2, float* out, size_t nout) {
    if (!ctx || n <= 0 || !out) return tr_fail(ctx, TR_ERR_INVALID, "test hook: bad arguments");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    float *da = nullptr, *db = nullptr, *dc = nullptr, *dout = nullptr;
    cudaStream_t s = ctx->stream;
    if (a) { TR_CUDA(ctx, cudaMalloc((void**)&da, na * 4)); cudaMemcpyAsync(da, a, na * 4, cudaMemcpyHostToDevice, s); }
    if (b) { TR_CUDA(ctx, cudaMalloc((void**)&db, nb * 4)); cudaMemcpyAsync(db, b, nb * 4, cudaMemcpyHostToDevice, s); }
    if (c) { TR_CUDA(ctx, cudaMalloc((void**)&dc, nc * 4)); cudaMemcpyAsync(dc, c, nc * 4, cudaMemcpyHostToDevice, s); }
    TR_CUDA(ctx, cudaMalloc((void**)&dout, nout * 4));
    k_hook<<<(n + 127) / 128, 128, 0, s>>>(op, n, da, db, dc, p0, p1, seed, u0, u1, u2, dout);
    cudaMemcpyAsync(out, dout, nout * 4, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(da); cudaFree(db); cudaFree(dc); cudaFree(dout);
    if (e != cudaSuccess) return tr_fail(ctx, TR_ERR_CUDA, "test hook: %s", cudaGetErrorString(e));
    TR_CHECK_LAUNCH(ctx);
    return TR_OK;
}

extern "C" {
int tr_test_disney_evaluate_pdf(tr_ctx* ctx, int n, const float* N, const float* V, const float* L, float metal, float rough, float* out) {
    return run_hook(ctx, H_DISNEY_EVAL, n, N, (size_t)n * 3, V, (size_t)n * 3, L, (size_t)n * 3, metal, rough, 0, 0, 0, 0, out, (size_t)n * 2);
}
int tr_test_disney_sample(tr_ctx* ctx, int n, const float* dir, const float* N, float metal, float rough, const float* u, float* out) {
    return run_hook(ctx, H_DISNEY_SAMPLE, n, dir, (size_t)n * 3, N, (size_t)n * 3, u, (size_t)n * 3, metal, rough, 0, 0, 0, 0, out, (size_t)n * 3);
}
int tr_test_glass_sample(tr_ctx* ctx, int n, const float* dir, const float* N, float ior, const float* u, float* out) {
    return run_hook(ctx, H_GLASS_SAMPLE, n, dir, (size_t)n * 3, N, (size_t)n * 3, u, (size_t)n, ior, 0.0f, 0, 0, 0, 0, out, (size_t)n * 4);
}
int tr_test_offset_ray(tr_ctx* ctx, int n, const float* p, const float* nrm, float* out) {
    return run_hook(ctx, H_OFFSET_RAY, n, p, (size_t)n * 3, nrm, (size_t)n * 3, nullptr, 0, 0.0f, 0.0f, 0, 0, 0, 0, out, (size_t)n * 3);
}
int tr_test_math(tr_ctx* ctx, int fn, int n, const float* a, const float* b, float* out) {
    if (fn < 0 || fn > 5 || !a || (fn >= 4 && !b)) return tr_fail(ctx, TR_ERR_INVALID, "tr_test_math: bad arguments");
    return run_hook(ctx, H_MATH, n, a, (size_t)n, fn >= 4 ? b : nullptr, (size_t)n, nullptr, 0, 0.0f, 0.0f, 0, (unsigned)fn, 0, 0, out, (size_t)n);
}
int tr_test_rng(tr_ctx* ctx, uint64_t seed, uint32_t pixel, uint32_t frame, uint32_t block, float* out4) {
    return run_hook(ctx, H_RNG, 1, nullptr, 0, nullptr, 0, nullptr, 0, 0.0f, 0.0f, seed, pixel, frame, block, out4, 4);
}
}
