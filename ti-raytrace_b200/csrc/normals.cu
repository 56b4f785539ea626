// normals.cu — Scene.process_normal (Scene.py:754-798) and Scene.total_area (Scene.py:747-750).
//
// process_normal is a BVH point query per vertex that accumulates angle x area weighted normals of
// coincident vertices.  The float sum depends on the visiting order, so this kernel keeps the
// reference's order (explicit stack, right child popped first, 32-entry bound of Scene.stack) on the
// pre-order node array; the stack lives in registers/local memory instead of a global field.
#include "ctx.h"
#include "common.cuh"

#define PN_STACK 32   // Scene.py:19 MAX_STACK_SIZE

__device__ __forceinline__ V3 ldpos(const float* __restrict__ vertex, int i) { const float* p = vertex + (size_t)i * 9; return mk3(p[0], p[1], p[2]); }
__device__ __forceinline__ V3 ldnor(const float* __restrict__ vertex, int i) { const float* p = vertex + (size_t)i * 9; return mk3(p[3], p[4], p[5]); }

// Scene.py:324-350 (triangle branch; shapes never reach process_normal's area call with a vertex)
__device__ __forceinline__ float tri_area(const float* __restrict__ vertex, int vi) {
    V3 v1 = ldpos(vertex, vi), v2 = ldpos(vertex, vi + 1), v3 = ldpos(vertex, vi + 2);
    float a = length3(v1 - v2), b = length3(v1 - v3), c = length3(v3 - v2);
    float sum = ((a + b) + c) * 0.5f;
    return sqrtf(sum * (sum - a) * (sum - b) * (sum - c));
}
// Scene.py:353-377
__device__ __forceinline__ float tri_angle(const float* __restrict__ vertex, int vi, V3 v) {
    V3 v1 = ldpos(vertex, vi), v2 = ldpos(vertex, vi + 1), v3 = ldpos(vertex, vi + 2);
    float ret;
    if (length3(v1 - v) < 0.00001f)      ret = dot3(normalize3(v2 - v1), normalize3(v3 - v1));
    else if (length3(v2 - v) < 0.00001f) ret = dot3(normalize3(v1 - v2), normalize3(v3 - v2));
    else                                 ret = dot3(normalize3(v1 - v3), normalize3(v2 - v3));
    return tr_acosf(ret);
}

// Per vertex, what every query that finds it adds: w = (normalize(n) * angle at the vertex) * area of its triangle -- the
// expression of Scene.py:778-790, which depends on the found vertex only.  Computing it once per vertex (coherent, streaming)
// instead of inside every divergent tree walk that reaches the vertex keeps the bits and takes the acos / sqrt chains out of
// the walk.
__global__ void k_normal_terms(const float* __restrict__ vertex, int nv, float4* __restrict__ w4, float4* __restrict__ n4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    int vi = (i / 3) * 3;                           // vertex soup: 3 vertices per triangle, in primitive order
    V3 nn = normalize3(ldnor(vertex, i));
    float ang = tri_angle(vertex, vi, ldpos(vertex, i));
    V3 w = (nn * ang) * tri_area(vertex, vi);
    w4[i] = make_float4(w.x, w.y, w.z, 0.0f); n4[i] = make_float4(nn.x, nn.y, nn.z, 0.0f);
}

__global__ void k_smooth_normals(const float* __restrict__ vertex, const int* __restrict__ prim, int nv,
                                 const TrNode* __restrict__ nodes, const TrLeaf* __restrict__ leaves,
                                 const float4* __restrict__ w4, const float4* __restrict__ n4, float* __restrict__ smooth) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    V3 v = ldpos(vertex, i), n = f4xyz(n4[i]);
    V3 sm = f4xyz(w4[i]);
    int stack[PN_STACK + 2];
    stack[0] = 0; int sp = 0;
    while (sp >= 0 && sp < PN_STACK) {
        int ni = stack[sp]; sp -= 1;
        float4 lo = nodes[ni].lo, hi = nodes[ni].hi;
        int link = __float_as_int(hi.w);
        if (link < 0) {
            float4 la = leaves[-link - 1].a, lb = leaves[-link - 1].b;
            if (__float_as_int(lb.w) == 0) {
                int pi = __float_as_int(la.w);
                int vi = prim[pi * 3 + 1];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    int nb = vi + j;
                    if (i != nb) {
                        V3 nvp = ldpos(vertex, nb), nn = f4xyz(n4[nb]);
                        if (length3(v - nvp) < 0.000001f && dot3(nn, n) > 0.5f) sm = sm + f4xyz(w4[nb]);
                    }
                }
            }
        } else {
            if (v.x >= lo.x && v.y >= lo.y && v.z >= lo.z && v.x <= hi.x && v.y <= hi.y && v.z <= hi.z) {
                stack[++sp] = ni + 1; stack[++sp] = link;
            }
        }
    }
    smooth[(size_t)i * 3] = sm.x; smooth[(size_t)i * 3 + 1] = sm.y; smooth[(size_t)i * 3 + 2] = sm.z;
}

__global__ void k_write_normals(float* __restrict__ vertex, const float* __restrict__ smooth, int nv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    V3 n = normalize3(mk3(smooth[(size_t)i * 3], smooth[(size_t)i * 3 + 1], smooth[(size_t)i * 3 + 2]));
    float* p = vertex + (size_t)i * 9; p[3] = n.x; p[4] = n.y; p[5] = n.z;
}

// Scene.py:747-750; a single thread keeps the sum order deterministic (the reference's atomic adds do not)
__global__ void k_total_area(const float* __restrict__ vertex, const int* __restrict__ prim, const float* __restrict__ shape,
                             const int* __restrict__ light, int nl, float* out) {
    float a = 0.0f;
    for (int k = 0; k < nl; ++k) {
        int pi = light[k];
        if (prim[pi * 3] == TR_PRIM_TRI) a += tri_area(vertex, prim[pi * 3 + 1]);
        else {
            const float* sp = shape + (size_t)prim[pi * 3 + 1] * 10; int st = (int)sp[0];
            if (st == TR_SHAPE_SPHERE || st == TR_SHAPE_SPOT || st == TR_SHAPE_LASER) a += sp[4] * sp[4] * TR_PI_ENV;
        }
    }
    *out = a;
}

extern "C" int tr_process_normal(tr_ctx* ctx) {
    if (!ctx || !ctx->bvh_ready) return tr_fail(ctx, TR_ERR_INVALID, "tr_process_normal: BVH not built");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    // the vertex soup must be 3 consecutive vertices per triangle primitive (Scene.py:94-141)
    // scratch: smooth normals (nv x 3) | per-vertex terms w4 (nv float4) | normalised normals n4 (nv float4)
    const size_t off4 = ((size_t)ctx->nv * 3 + 3) & ~(size_t)3;
    int rc; if ((rc = tr_realloc(ctx, &ctx->d_smooth, off4 + (size_t)ctx->nv * 8))) return rc;
    float* d_smooth = ctx->d_smooth;
    float4* w4 = reinterpret_cast<float4*>(d_smooth + off4); float4* n4 = w4 + ctx->nv;
    int g = (ctx->nv + 127) / 128;
    k_normal_terms<<<g, 128, 0, ctx->stream>>>(ctx->d_vertex, ctx->nv, w4, n4);
    k_smooth_normals<<<g, 128, 0, ctx->stream>>>(ctx->d_vertex, ctx->d_prim, ctx->nv, ctx->d_nodes, ctx->d_leaves, w4, n4, d_smooth);
    k_write_normals<<<g, 128, 0, ctx->stream>>>(ctx->d_vertex, d_smooth, ctx->nv);
    TR_CHECK_LAUNCH(ctx);              // asynchronous like the build: a fault surfaces at the next synchronising call
    ctx->shade_ready = false; ctx->fh_ready = false; ctx->gen++;    // shading records hold normals
    return TR_OK;
}

extern "C" int tr_vertex_download(tr_ctx* ctx, float* vertex) {
    if (!ctx || !ctx->d_vertex || !vertex) return tr_fail(ctx, TR_ERR_INVALID, "tr_vertex_download: no scene");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    TR_CUDA(ctx, cudaMemcpyAsync(vertex, ctx->d_vertex, (size_t)ctx->nv * 36, cudaMemcpyDeviceToHost, ctx->stream));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

extern "C" int tr_total_area(tr_ctx* ctx, float* area) {
    if (!ctx || !ctx->d_vertex || !area) return tr_fail(ctx, TR_ERR_INVALID, "tr_total_area: no scene");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    float* d = (float*)(ctx->d_build_status + 8);      // scratch word of the status block
    k_total_area<<<1, 1, 0, ctx->stream>>>(ctx->d_vertex, ctx->d_prim, ctx->d_shape, ctx->d_light, ctx->nl, d);
    cudaMemcpyAsync(area, d, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return tr_fail(ctx, TR_ERR_CUDA, "tr_total_area: %s", cudaGetErrorString(e));
    return TR_OK;
}
