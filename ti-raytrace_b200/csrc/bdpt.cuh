// bdpt.cuh — BDPT_RGB (integrator/BDPT_RGB.py + integrator/BDPT_Vertex.py) as a staged pipeline; included at the end of
// wavefront.cu (same translation unit: it re-uses WfArgs, the traversal, surface_at / sample_li and the tile mapping).
//
// The reference runs one thread per pixel that builds an eye sub-path (<= 7 vertices) and a light sub-path (<= 6), then
// loops over every (e, l) prefix pair, tracing one shadow ray per connection and rewriting the end-point vertices in place
// for the MIS weight (mis_weight, :258-434).  Here one frame batch is five launches over SoA buffers in HBM:
//   k_bdpt_paths    one lane per (sample, sub-path): warps are homogeneous (all-eye or all-light), vertices go to the
//                   vertex buffer  vb[(v * 5 + c) * cap + s]  (v = 0..6 eye, 7..12 light; c = 5 float4 words, 80 B / vertex)
//   k_bdpt_items    enumerates the valid strategies of every sample, strategy-major, into a compact connection queue
//   k_bdpt_connect  one lane per connection: geometric term, shadow query (the early-exit nearest-hit query of trace.cuh),
//                   BSDF terms and the MIS weight computed on register copies of the <= 4 vertices the reference
//                   overwrites; e >= 2 results go to contrib[strategy][s], e == 1 results are splatted with float atomics
//   k_bdpt_film     per pixel and frame: own strategies summed in the reference's loop order + splats, running mean
// Quirks kept literally are listed in oracle/bdpt_core.inc (the restatement this is tested against); RNG blocks as there.
#pragma once

#define BD_MAX_DEPTH 5
#define BD_EYE_MAX (BD_MAX_DEPTH + 2)
#define BD_LIGHT_MAX (BD_MAX_DEPTH + 1)
#define BD_NVERT (BD_EYE_MAX + BD_LIGHT_MAX)
#define BD_NCONTRIB 21                 // strategies with e >= 2:  sum_{e=2..7} (8 - e)
#define BD_NSTRAT 26                   // + the five e == 1 (light tracing) strategies
#define BD_VERTEX_NONE 0
#define BD_VERTEX_LIGHT 1
#define BD_VERTEX_LENS 2
#define BD_VERTEX_SURFACE 3
#define BD_EPS 0.00001f                // UtilsFunc.py:36

struct BdArgs {
    float4* vb; int* depths; float4* contrib; float* splat; unsigned* items;
    const int* tile_slot;
    unsigned long long* ctr;           // [0] closest-hit traversals, [1] shadow traversals, [2] connection queue size (low 32 bits)
    size_t cap;                        // samples per batch
    float view[16];
};

// contrib row of strategy (e >= 2, l)
__device__ __forceinline__ int bd_row(int e, int l) { return (e - 2) * 6 - ((e - 2) * (e - 3)) / 2 + l; }

struct BV { V3 pos, normal, snormal, beta, wo; float fpdf, rpdf; int type, prim, mat, delta; };

__device__ __forceinline__ void bv_store(const BdArgs& b, size_t s, int v, const BV& x) {
    float4* p = b.vb + (size_t)(v * 5) * b.cap + s;
    p[0] = make_float4(x.pos.x, x.pos.y, x.pos.z, x.fpdf);
    p[b.cap] = make_float4(x.normal.x, x.normal.y, x.normal.z, x.rpdf);
    p[2 * b.cap] = make_float4(x.snormal.x, x.snormal.y, x.snormal.z, __int_as_float(x.type | (x.delta << 4) | (x.mat << 8)));
    p[3 * b.cap] = make_float4(x.beta.x, x.beta.y, x.beta.z, __int_as_float(x.prim));
    p[4 * b.cap] = make_float4(x.wo.x, x.wo.y, x.wo.z, 0.0f);
}
__device__ __forceinline__ BV bv_load(const BdArgs& b, size_t s, int v) {
    const float4* p = b.vb + (size_t)(v * 5) * b.cap + s;
    float4 a0 = p[0], a1 = p[b.cap], a2 = p[2 * b.cap], a3 = p[3 * b.cap], a4 = p[4 * b.cap];
    BV x; x.pos = f4xyz(a0); x.fpdf = a0.w; x.normal = f4xyz(a1); x.rpdf = a1.w; x.snormal = f4xyz(a2);
    int fl = __float_as_int(a2.w); x.type = fl & 15; x.delta = (fl >> 4) & 15; x.mat = fl >> 8;
    x.beta = f4xyz(a3); x.prim = __float_as_int(a3.w); x.wo = f4xyz(a4);
    return x;
}
__device__ __forceinline__ void bv_set_rpdf(const BdArgs& b, size_t s, int v, float r) { ((float*)(b.vb + (size_t)(v * 5 + 1) * b.cap + s))[3] = r; }
__device__ __forceinline__ void bv_scalars(const BdArgs& b, size_t s, int v, float& fpdf, float& rpdf, int& delta) {
    fpdf = ((const float*)(b.vb + (size_t)(v * 5) * b.cap + s))[3];
    rpdf = ((const float*)(b.vb + (size_t)(v * 5 + 1) * b.cap + s))[3];
    delta = (__float_as_int(((const float*)(b.vb + (size_t)(v * 5 + 2) * b.cap + s))[3]) >> 4) & 15;
}

// One sub-path (BDPT_RGB.py:104-187 eye_path / :189-250 light_path), one lane per sample, the warp walks in lock step.
// The two loops differ where the reference differs: the eye path stores the emitter vertex it hits and measures `to`
// from the offset ray origin with a clamped distance; the light path stops in front of emitters, measures from the
// stored previous position, and multiplies fpdf in a different order.
template <bool SMEM, bool LIGHT>
__device__ __forceinline__ void bdpt_subpath(const WfArgs& a, const BdArgs& b, const BatchParams& bp, const TrNode* nodes, const TrLeaf* leaves,
                                             const TrNodeX* nodesx, size_t s, bool active, unsigned pix, unsigned frame, int x, int y,
                                             unsigned long long& n_closest) {
    const int vbase = LIGHT ? BD_EYE_MAX : 0, maxd = LIGHT ? BD_LIGHT_MAX : BD_EYE_MAX;
    V3 origin = mk3(0.f, 0.f, 0.f), dir = mk3(1.f, 1.f, 1.f), beta = mk3(1.f, 1.f, 1.f), prev_pos = origin, prev_normal = dir;
    float pdfFwd = 1.0f, pdfRev = 0.0f;
    int depth = 1;
    if (active) {
        BV v0; v0.snormal = mk3(0.f, 0.f, 0.f); v0.wo = mk3(0.f, 0.f, 0.f); v0.rpdf = 0.0f; v0.prim = 0; v0.mat = 0; v0.delta = 0;
        if (!LIGHT) {
            float jx = 0.0f, jy = 0.0f;
            if (frame != 0) { float4 r = rng4(bp.seed, pix, frame, 0u); jx = r.x - 0.5f; jy = r.y - 0.5f; }
            origin = mk3(a.cam.eye[0], a.cam.eye[1], a.cam.eye[2]); dir = camera_dir(a.cam, x, y, jx, jy);
            v0.pos = origin; v0.normal = dir; v0.beta = mk3(1.f, 1.f, 1.f); v0.fpdf = 1.0f; v0.type = BD_VERTEX_LENS;
        } else {
            // Scene.sample_light (Scene.py:430-474): point as in sample_li, cosine-hemisphere direction
            float4 Ra = rng4(bp.seed, pix, frame, 40u), Rb = rng4(bp.seed, pix, frame, 41u);
            LightSample ls = sample_li(a, mk3(0.f, 0.f, 0.f), Ra.x, Ra.y, Ra.z);
            V3 p = cosine_sample_hemisphere(Rb.x, Rb.y);
            pdfFwd = cosine_hemisphere_pdf(p.z);
            dir = inverse_transform(p, ls.normal);
            const float light_pdf = ls.choice_pdf;
            origin = ls.pos;
            v0.pos = ls.pos; v0.normal = ls.normal; v0.beta = ls.emission / light_pdf; v0.fpdf = light_pdf; v0.wo = dir; v0.type = BD_VERTEX_LIGHT;
            beta = (ls.emission / light_pdf) * fabsf(dot3(ls.normal, dir));
        }
        prev_pos = v0.pos; prev_normal = v0.normal;
        bv_store(b, s, vbase, v0);
    }
    bool alive = active;
    for (int it = 1; it < maxd; ++it) {
        if (__ballot_sync(0xffffffffu, alive) == 0u) break;
        RayPre r = make_ray(origin, dir);
        HitRec h = trace_closest<SMEM>(nodes, leaves, nodesx, a.nnodes, r, alive, a.ctr->visits);
        if (!alive) continue;
        ++n_closest;
        if (h.prim < 0) { alive = false; continue; }
        Surf sf = surface_at(a, h.prim, h.u, h.v, origin, dir, h.t);
        V3 fn = signf_(dot3(-dir, sf.gn)) * sf.n;
        const float* m = a.material + (size_t)sf.mat * 10;
        const int mt = (int)__ldg(m);
        const float p0 = __ldg(m + 5), p1 = __ldg(m + 6);
        if (LIGHT && mt == TR_MAT_LIGHT) { alive = false; continue; }
        V3 to; float dist;
        if (!LIGHT) { to = sf.pos - origin; dist = fmaxf(length3(to), 0.01f); }
        else { to = sf.pos - prev_pos; dist = length3(to); }
        const float inv_dist2 = 1.0f / (dist * dist);
        to = to / dist;
        BV v; v.pos = sf.pos; v.normal = sf.n; v.snormal = fn; v.wo = dir; v.rpdf = 0.0f; v.prim = h.prim; v.mat = sf.mat; v.delta = 0;
        if (!LIGHT) v.fpdf = pdfFwd * fabsf(dot3(to, prev_normal)) * inv_dist2;
        else v.fpdf = pdfFwd * (fabsf(dot3(to, prev_normal)) * inv_dist2);
        if (!LIGHT && mt == TR_MAT_LIGHT) {
            V3 mcol = mk3(__ldg(m + 2), __ldg(m + 3), __ldg(m + 4));
            v.beta = (beta * mcol) * fabsf(dot3(sf.n, dir)); v.type = BD_VERTEX_LIGHT;
            bv_store(b, s, vbase + depth, v);
            depth += 1; alive = false; continue;
        }
        v.beta = beta * fabsf(dot3(dir, sf.n)); v.type = BD_VERTEX_SURFACE;
        V3 rc = f4xyz(__ldg(a.matlin + sf.mat));
        float4 R0 = rng4(bp.seed, pix, frame, (LIGHT ? 42u : 1u) + 2u * (unsigned)(depth - 1));
        float4 R1 = rng4(bp.seed, pix, frame, (LIGHT ? 43u : 2u) + 2u * (unsigned)(depth - 1));
        V3 next_dir; float brdf, f_or_b = 1.0f;
        if (mt == TR_MAT_GLASS) { next_dir = glass_sample(dir, sf.n, p0, R0.w, f_or_b); brdf = 1.0f; pdfFwd = 1.0f; v.delta = 1; }
        else { next_dir = disney_sample(dir, fn, p0, p1, R0.w, R1.x, R1.y); disney_evaluate_pdf(fn, -dir, next_dir, p0, p1, brdf, pdfFwd); }
        bv_store(b, s, vbase + depth, v);
        if (!(pdfFwd > 0.0f)) { alive = false; continue; }
        if (mt == TR_MAT_GLASS) { pdfRev = 0.0f; pdfFwd = 0.0f; beta = beta * (brdf * rc); }
        else {
            beta = beta * (((brdf * rc) * fabsf(dot3(sf.n, next_dir))) / pdfFwd);
            pdfRev = disney_pdf(fn, next_dir, -dir, p0, p1);
        }
        bv_set_rpdf(b, s, vbase + depth - 1, pdfRev * fabsf(dot3(to, sf.n)) * inv_dist2);
        if (f_or_b < 0.0f) { float Rr = expf(-h.t / p1); if (R1.z >= Rr) { alive = false; continue; } }
        depth += 1;
        origin = offset_ray(sf.pos, signf_(f_or_b) * fn);
        dir = next_dir; prev_pos = sf.pos; prev_normal = sf.n;
    }
    if (active) b.depths[(LIGHT ? b.cap : 0) + s] = depth;
}

template <bool SMEM>
__global__ void __launch_bounds__(WF_THREADS) k_bdpt_paths(WfArgs a, BdArgs b) {
    const TrNode* nodes; const TrLeaf* leaves; const TrNodeX* nodesx;
    bvh_view<SMEM>(a, nodes, leaves, nodesx);
    const BatchParams bp = *a.bp;
    const int nsamp = bp.n_frames * a.npix, nsamp_r = (nsamp + 31) & ~31;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long n_closest = 0;
    for (int base = warp * 32; base < 2 * nsamp_r; base += nwarps * 32) {
        const bool light = base >= nsamp_r;                         // warp-uniform
        const int s = base + lane - (light ? nsamp_r : 0);
        int x = 0, y = 0; bool active = false; unsigned frame = 0;
        if (s < nsamp) { int f = s / a.npix, p = s - f * a.npix; frame = (unsigned)(bp.frame_begin + f); active = slot_to_pixel(a, p, x, y); }
        const unsigned pix = ((unsigned)x << 16) | (unsigned)y;
        if (s < nsamp && !active) b.depths[(light ? b.cap : 0) + s] = 0;
        if (light) bdpt_subpath<SMEM, true>(a, b, bp, nodes, leaves, nodesx, (size_t)s, active, pix, frame, x, y, n_closest);
        else bdpt_subpath<SMEM, false>(a, b, bp, nodes, leaves, nodesx, (size_t)s, active, pix, frame, x, y, n_closest);
    }
    if (n_closest) atomicAdd(b.ctr, n_closest);
}

// Connection queue: item = sample | e << 26 | l << 29, emitted strategy-major per warp so that the connect kernel's warps
// are (mostly) homogeneous in strategy type.  Loop bounds of BDPT_RGB.py:625-637.
__global__ void __launch_bounds__(WF_THREADS) k_bdpt_items(WfArgs a, BdArgs b) {
    const BatchParams bp = *a.bp;
    const int nsamp = bp.n_frames * a.npix, nsamp_r = (nsamp + 31) & ~31;
    const int stride = gridDim.x * blockDim.x;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nsamp_r; s += stride) {
        int ed = 0, ld = 0;
        if (s < nsamp) { ed = b.depths[s]; ld = b.depths[b.cap + s]; }
        for (int e = 1; e <= BD_EYE_MAX; ++e)
            for (int l = 0; l <= BD_LIGHT_MAX; ++l) {
                const int depth = l + e - 2;
                if ((l == 1 && e == 1) || depth < 0 || depth > BD_MAX_DEPTH) continue;      // warp-uniform
                const bool valid = e <= ed && l <= ld;
                int q = warp_append((int*)(b.ctr + 2), valid);
                if (valid) b.items[q] = (unsigned)s | ((unsigned)e << 26) | ((unsigned)l << 29);
            }
    }
}

// mis_weight (BDPT_RGB.py:258-434).  The reference saves <= 4 end-point vertices, overwrites them in the per-pixel vertex
// fields, walks the pdf ratios and restores them; here the same four vertices are register copies and the walk reads the
// untouched (fpdf, rpdf, delta) scalars of the others from the vertex buffer.  smp_*: the freshly sampled emitter vertex
// that replaces light[0] for l == 1 (the e == 1 replacement of eye[0] by the lens sample changes nothing that is read).
__device__ __forceinline__ float bdpt_mis_weight(const WfArgs& a, const BdArgs& b, size_t s, int e, int l,
                                                 V3 smp_pos, V3 smp_normal, float smp_fpdf) {
    if (l + e == 2) return 1.0f;
    BV E1, E2, L1, L2;
    E1 = bv_load(b, s, e - 1);
    if (e > 1) E2 = bv_load(b, s, e - 2); else E2 = E1;
    if (l > 0) L1 = bv_load(b, s, BD_EYE_MAX + l - 1); else L1 = E1;
    if (l > 1) L2 = bv_load(b, s, BD_EYE_MAX + l - 2); else L2 = L1;
    if (l == 1) { L1.pos = smp_pos; L1.normal = smp_normal; L1.snormal = smp_normal; L1.fpdf = smp_fpdf; L1.type = BD_VERTEX_LIGHT; }
    float e1_rpdf = E1.rpdf, e2_rpdf = E2.rpdf, l1_rpdf = L1.rpdf, l2_rpdf = L2.rpdf;
    const V3 axis = mk3(b.view[8], b.view[9], b.view[10]);                // Camera.get_optical_axis (Camera.py:128-129)
    // eye[e-1].rpdf (:290-316)
    if (l == 0) {
        float pdfPos = 1.0f / __ldg(&a.shade[E1.prim].q[2].w), pdfChoice = 1.0f / (float)a.nl;
        e1_rpdf = pdfPos * pdfChoice;
    } else if (l == 1) {
        if (E1.type == BD_VERTEX_SURFACE) {
            V3 to = E1.pos - L1.pos; float dist = length3(to); to = to / dist;
            float c = fabsf(dot3(to, L1.normal));
            e1_rpdf = cosine_hemisphere_pdf(c) * c / (dist * dist);
        } else e1_rpdf = 1.0f;
    } else {
        V3 wi = L2.pos - L1.pos, wo = E1.pos - L1.pos; float dist = length3(wo);
        wi = normalize3(wi); wo = normalize3(wo);
        float pdf = 1.0f;
        if (L1.mat == TR_MAT_DISNEY) { const float* m = a.material + (size_t)L1.mat * 10; pdf = disney_pdf(L1.snormal, wi, wo, __ldg(m + 5), __ldg(m + 6)); }   // material INDEX == 0 (sic, :312)
        e1_rpdf = pdf * fabsf(dot3(L1.normal, wo)) / (dist * dist);
    }
    // light[l-1].rpdf (:317-343)
    if (l > 0) {
        if (e > 1) {
            if (E1.type == BD_VERTEX_SURFACE) {
                V3 wi = E2.pos - E1.pos, wo = L1.pos - E1.pos; float dist = length3(wo);
                wi = normalize3(wi); wo = normalize3(wo);
                float pdf = 1.0f;
                if (E1.mat == TR_MAT_DISNEY) { const float* m = a.material + (size_t)E1.mat * 10; pdf = disney_pdf(E1.snormal, wi, wo, __ldg(m + 5), __ldg(m + 6)); }
                l1_rpdf = pdf * fabsf(dot3(E1.normal, wo)) / (dist * dist);
            } else l1_rpdf = 1.0f;
        } else {
            V3 to = E1.pos - L1.pos; float dist = length3(to); to = to / dist;      // e == 1: eye[0] = the lens
            l1_rpdf = dot3(to, axis) / (dist * dist);
        }
    }
    // eye[e-2].rpdf (:345-369)
    if (e > 1) {
        if (l == 0) {
            V3 to = E2.pos - E1.pos; float dist = length3(to); to = to / dist;
            float pdfDir = cosine_hemisphere_pdf(fabsf(dot3(to, E1.normal)));
            float LdotN = dot3(to, E1.normal);
            e2_rpdf = fabsf(pdfDir * LdotN) / (dist * dist);
        } else if (E1.type == BD_VERTEX_SURFACE) {
            V3 wi = L1.pos - E1.pos, wo = E2.pos - E1.pos; float dist = length3(wo);
            wi = normalize3(wi); wo = normalize3(wo);
            const float* m = a.material + (size_t)E1.mat * 10;
            float pdf = disney_pdf(E1.snormal, wi, wo, __ldg(m + 5), __ldg(m + 6));
            e2_rpdf = pdf / (dist * dist);
            if (E2.type == BD_VERTEX_SURFACE) e2_rpdf *= fabsf(dot3(E1.normal, wo));
        } else e2_rpdf = 1.0f;
    }
    // light[l-2].rpdf (:371-388)
    if (l > 1) {
        if (E1.type != BD_VERTEX_LIGHT) {
            V3 wi = E1.pos - L1.pos, wo = L2.pos - L1.pos; float dist = length3(wo);
            wi = normalize3(wi); wo = normalize3(wo);
            float pdf = 1.0f;
            if (L1.mat == TR_MAT_DISNEY) { const float* m = a.material + (size_t)L1.mat * 10; pdf = disney_pdf(L1.normal, wi, wo, __ldg(m + 5), __ldg(m + 6)); }
            l2_rpdf = pdf / (dist * dist);
            if (L2.type == BD_VERTEX_SURFACE) l2_rpdf *= fabsf(dot3(L1.normal, wo));
        } else l2_rpdf = 1.0f;
    }
    // the walks (:394-416); delta of eye[e-1] and light[l-1] is forced to 0 (:283-286)
    float weight_sum = 0.0f, weight = 1.0f;
    int dk = 0;                                                           // delta of vertex k (from the previous iteration: k+1 -> k)
    for (int k = e - 1; k > 0; --k) {
        float f, r; int d0, d1;
        if (k == e - 1) { f = E1.fpdf; r = e1_rpdf; d0 = 0; }
        else { bv_scalars(b, s, k, f, r, d0); d0 = dk; if (k == e - 2) r = e2_rpdf; }
        { float f1, r1; bv_scalars(b, s, k - 1, f1, r1, d1); }
        weight *= (r == 0.0f ? 1.0f : r) / (f == 0.0f ? 1.0f : f);
        if (d0 == 0 && d1 == 0) weight_sum += weight;
        dk = d1;
    }
    weight = 1.0f; dk = 0;
    for (int k = l - 1; k >= 0; --k) {
        float f, r; int d0, d1 = 0;
        if (k == l - 1) { f = L1.fpdf; r = l1_rpdf; d0 = 0; }
        else { bv_scalars(b, s, BD_EYE_MAX + k, f, r, d0); d0 = dk; if (k == l - 2) r = l2_rpdf; }
        if (k > 0) { float f1, r1; bv_scalars(b, s, BD_EYE_MAX + k - 1, f1, r1, d1); }
        weight *= (r == 0.0f ? 1.0f : r) / (f == 0.0f ? 1.0f : f);
        if (d0 == 0 && d1 == 0) weight_sum += weight;
        dk = d1;
    }
    return 1.0f / (1.0f + weight_sum);
}

// Camera.get_image_point (Camera.py:144-158)
__device__ __forceinline__ void bd_image_point(const WfArgs& a, const BdArgs& b, V3 p, int& u, int& v, V3& wi) {
    const float* m = b.view;
    float px = ((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3] * 1.0f;
    float py = ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7] * 1.0f;
    float pz = ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11] * 1.0f;
    float fu = -px / pz * a.cam.fx + a.cam.cx, fv = -py / pz * a.cam.fy + a.cam.cy;
    u = (fu > -2.0e9f && fu < 2.0e9f) ? (int)fu : -1; v = (fv > -2.0e9f && fv < 2.0e9f) ? (int)fv : -1;
    wi = mk3(0.f, 0.f, 0.f);
    if (u < 0 || u >= a.W || v < 0 || v >= a.H || pz > 0.0f) { u = -1; v = -1; }
    else wi = p - mk3(a.cam.eye[0], a.cam.eye[1], a.cam.eye[2]);
    wi = normalize3(wi);
}

// connect_path (BDPT_RGB.py:436-580), one lane per (sample, e, l)
template <bool SMEM>
__global__ void __launch_bounds__(WF_THREADS) k_bdpt_connect(WfArgs a, BdArgs b) {
    const TrNode* nodes; const TrLeaf* leaves; const TrNodeX* nodesx;
    bvh_view<SMEM>(a, nodes, leaves, nodesx);
    const BatchParams bp = *a.bp;
    const int n = (int)b.ctr[2], n_r = (n + 31) & ~31;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long n_shadow = 0;
    for (int base = warp * 32; base < n_r; base += nwarps * 32) {
        const int it = base + lane;
        const bool active = it < n;
        unsigned item = active ? b.items[it] : 0u;
        const size_t s = item & 0x3ffffffu; const int e = (item >> 26) & 7, l = (int)(item >> 29);
        V3 radiance = mk3(0.f, 0.f, 0.f), ro = mk3(0.f, 0.f, 0.f), rd = mk3(1.f, 1.f, 1.f);
        V3 smp_pos = ro, smp_normal = ro; float smp_fpdf = 1.0f;
        bool need = false; int target = 0, nu = -1, nv = -1, kind = 0;       // kind: 1 e == 1, 2 l == 1, 3 general
        BV ev, lv; LightSample ls; float c0 = 0.0f, c1 = 0.0f, dist = 1.0f;
        if (active) {
            if (l == 0) {
                ev = bv_load(b, s, e - 1);
                if (ev.type == BD_VERTEX_LIGHT) radiance = ev.beta;
            } else if (e == 1) {
                lv = bv_load(b, s, BD_EYE_MAX + l - 1);
                bd_image_point(a, b, lv.pos, nu, nv, rd);
                c0 = dot3(rd, lv.snormal);                                   // NdotL
                if (nu >= 0 && lv.delta != 1 && c0 < 0.0f && lv.type == BD_VERTEX_SURFACE) {
                    need = true; kind = 1; target = lv.prim; ro = mk3(a.cam.eye[0], a.cam.eye[1], a.cam.eye[2]);
                }
            } else if (l == 1) {
                ev = bv_load(b, s, e - 1);
                if (ev.delta != 1) {
                    int x, y; slot_to_pixel(a, (int)(s % (size_t)a.npix), x, y);
                    unsigned pix = ((unsigned)x << 16) | (unsigned)y, frame = (unsigned)bp.frame_begin + (unsigned)(s / (size_t)a.npix);
                    float4 R0 = rng4(bp.seed, pix, frame, 1u + 2u * (unsigned)(e - 2));
                    ro = offset_ray(ev.pos, ev.snormal);
                    ls = sample_li(a, ro, R0.x, R0.y, R0.z);
                    c0 = dot3(ls.dir, ls.normal); c1 = dot3(ls.dir, ev.snormal);        // NdotLl, NdotLe
                    need = true; kind = 2; target = ls.prim; rd = -ls.dir;
                }
            } else {
                ev = bv_load(b, s, e - 1); lv = bv_load(b, s, BD_EYE_MAX + l - 1);
                if (lv.delta != 1 && ev.delta != 1 && ev.type == BD_VERTEX_SURFACE && lv.type == BD_VERTEX_SURFACE) {
                    V3 d = ev.pos - lv.pos; dist = length3(d); d = d / dist;
                    c0 = dot3(d, lv.snormal); c1 = dot3(d, ev.snormal);                 // NdotLl, NdotLe
                    need = true; kind = 3; target = ev.prim; ro = lv.pos; rd = d;
                }
            }
        }
        if (__ballot_sync(0xffffffffu, need) != 0u) {
            RayPre r = make_ray(ro, rd);
            int tleaf = need ? __ldg(a.leaf_of_prim + target) : 0;
            float tt = TR_INF;
            bool vis = trace_shadow_visible<SMEM>(nodes, leaves, nodesx, a.nnodes, r, need, tleaf, a.ctr->visits + 2, &tt);
            if (need) {
                ++n_shadow;
                if (kind == 1 && vis) {
                    const float* m = a.material + (size_t)lv.mat * 10;
                    float brdf, pdf; disney_evaluate_pdf(lv.snormal, -lv.wo, -rd, __ldg(m + 5), __ldg(m + 6), brdf, pdf);
                    if (pdf > 0.0f) { float G = fabsf(c0) / (tt * tt); radiance = (((G * lv.beta) * f4xyz(__ldg(a.matlin + lv.mat))) * brdf) / pdf; }
                } else if (kind == 2 && vis && tt > BD_EPS) {
                    const float* m = a.material + (size_t)ev.mat * 10;
                    const float light_pdf = ls.choice_pdf;
                    float brdf, pdf; disney_evaluate_pdf(ev.snormal, -ev.wo, -ls.dir, __ldg(m + 5), __ldg(m + 6), brdf, pdf);
                    if (pdf > 0.0f) {
                        float G = fabsf(c1 * c0) / (tt * tt);
                        radiance = (((((G * ev.beta) * brdf) / pdf) * f4xyz(__ldg(a.matlin + ev.mat))) * ls.emission) / light_pdf;
                    }
                    smp_pos = ls.pos; smp_normal = ls.normal; smp_fpdf = light_pdf;
                } else if (kind == 3 && vis && tt > BD_EPS) {
                    const float* mE = a.material + (size_t)ev.mat * 10; const float* mL = a.material + (size_t)lv.mat * 10;
                    float brdfL, lpdf, brdfE, epdf;
                    disney_evaluate_pdf(lv.snormal, -lv.wo, rd, __ldg(mL + 5), __ldg(mL + 6), brdfL, lpdf);
                    disney_evaluate_pdf(ev.snormal, -ev.wo, -rd, __ldg(mE + 5), __ldg(mE + 6), brdfE, epdf);
                    if (brdfL > 0.0f && brdfE > 0.0f) {
                        float G = fabsf(c1 * c0) / (dist * dist);
                        radiance = (((((((G * ev.beta) * lv.beta) * brdfL) / lpdf) * brdfE) / epdf) * f4xyz(__ldg(a.matlin + ev.mat))) * f4xyz(__ldg(a.matlin + lv.mat));
                    }
                }
            }
        }
        if (active) {
            if (radiance.x > 0.0f && radiance.y > 0.0f && radiance.z > 0.0f) radiance = radiance * bdpt_mis_weight(a, b, s, e, l, smp_pos, smp_normal, smp_fpdf);
            if (e == 1) {
                if (nu >= 0 && (radiance.x != 0.0f || radiance.y != 0.0f || radiance.z != 0.0f)) {
                    float* o = b.splat + ((size_t)(s / (size_t)a.npix) * a.W * a.H + (size_t)nu * a.H + nv) * 3;
                    atomicAdd(o, radiance.x); atomicAdd(o + 1, radiance.y); atomicAdd(o + 2, radiance.z);
                }
            } else b.contrib[(size_t)bd_row(e, l) * b.cap + s] = make_float4(radiance.x, radiance.y, radiance.z, 0.0f);
        }
    }
    if (n_shadow) atomicAdd(b.ctr + 1, n_shadow);
}

// render() tail (BDPT_RGB.py:625-641): radiance = own strategies in loop order + splats, then the running mean, frame by frame.
// Runs over the FULL frame: pixels of other ranks' tiles still receive this rank's splats (the film reduce is a sum).
__global__ void __launch_bounds__(WF_THREADS) k_bdpt_film(WfArgs a, BdArgs b) {
    const BatchParams bp = *a.bp;
    const int ntx = (a.W + TR_TILE - 1) / TR_TILE;
    const int stride = gridDim.x * blockDim.x;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < a.W * a.H; g += stride) {
        const int x = g / a.H, y = g - x * a.H;
        const int lt = b.tile_slot[(y / TR_TILE) * ntx + x / TR_TILE];
        int p = -1;
        if (lt >= 0) { int xi = x % TR_TILE, yi = y % TR_TILE; p = lt * 1024 + (((yi >> 3) * 8 + (xi >> 2)) << 5) + ((xi & 3) << 3) + (yi & 7); }   // inverse of slot_to_pixel
        float* o = a.hdr + (size_t)g * 3;
        float r = o[0], gg = o[1], bb = o[2];
        for (int f = 0; f < bp.n_frames; ++f) {
            V3 rad = mk3(0.f, 0.f, 0.f);
            if (p >= 0) {
                const size_t s = (size_t)f * a.npix + p;
                const int ed = b.depths[s], ld = b.depths[b.cap + s];
                for (int e = 2; e <= ed; ++e)
                    for (int l = 0; l <= ld && l + e - 2 <= BD_MAX_DEPTH; ++l) rad = rad + f4xyz(b.contrib[(size_t)bd_row(e, l) * b.cap + s]);
            }
            const float* sp = b.splat + ((size_t)f * a.W * a.H + g) * 3;
            rad = rad + mk3(sp[0], sp[1], sp[2]);
            float coff = 1.0f / ((float)(bp.frame_begin + f) + 1.0f);
            r = rad.x * coff + r * (1.0f - coff); gg = rad.y * coff + gg * (1.0f - coff); bb = rad.z * coff + bb * (1.0f - coff);
        }
        o[0] = r; o[1] = gg; o[2] = bb;
    }
}

static int render_bdpt(tr_ctx* ctx, int frame_begin, int n_frames, uint64_t seed) {
    if (!ctx || n_frames <= 0 || frame_begin < 0) return tr_fail(ctx, TR_ERR_INVALID, "tr_render_bdpt_rgb: bad arguments (frames %d+%d)", frame_begin, n_frames);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    WfArgs a; int rc;
    if ((rc = fill_args(ctx, a, false))) return rc;
    if (ctx->nl <= 0) return tr_fail(ctx, TR_ERR_INVALID, "tr_render_bdpt_rgb: the scene has no emitter (Scene.sample_light needs one)");
    // frames per batch: ~1.5 KB of vertex / contribution / queue records per sample
    const size_t per_sample = (size_t)BD_NVERT * 80 + BD_NCONTRIB * 16 + BD_NSTRAT * 4 + 8;
    size_t budget = ctx->opt_max_paths * 252 / per_sample;
    const size_t npix = (size_t)(a.npix > 0 ? a.npix : 1);
    int F = ctx->opt_batch_frames > 0 ? ctx->opt_batch_frames : (int)(budget / npix);
    if (F < 1) F = 1; if (F > n_frames) F = n_frames;
    const size_t cap = (size_t)F * npix;
    if (cap >= ((size_t)1 << 26)) return tr_fail(ctx, TR_ERR_INVALID, "tr_render_bdpt_rgb: %zu samples per batch exceed the queue encoding (2^26)", cap);
    if (cap > ctx->bd_cap) {
        if ((rc = tr_realloc(ctx, &ctx->d_bd_vb, cap * BD_NVERT * 5))) return rc;
        if ((rc = tr_realloc(ctx, &ctx->d_bd_depths, cap * 2))) return rc;
        if ((rc = tr_realloc(ctx, &ctx->d_bd_contrib, cap * BD_NCONTRIB))) return rc;
        if ((rc = tr_realloc(ctx, &ctx->d_bd_items, cap * BD_NSTRAT))) return rc;
        ctx->bd_cap = cap;
    }
    if (F > ctx->bd_splat_frames || !ctx->d_bd_splat) { if ((rc = tr_realloc(ctx, &ctx->d_bd_splat, (size_t)F * ctx->W * ctx->H * 3))) return rc; ctx->bd_splat_frames = F; }
    if (!ctx->d_bd_ctr) TR_CUDA(ctx, cudaMalloc((void**)&ctx->d_bd_ctr, 4 * sizeof(unsigned long long)));
    BdArgs b; memset(&b, 0, sizeof(b));
    b.vb = ctx->d_bd_vb; b.depths = ctx->d_bd_depths; b.contrib = ctx->d_bd_contrib; b.splat = ctx->d_bd_splat; b.items = ctx->d_bd_items;
    b.tile_slot = ctx->d_bd_tile_slot; b.ctr = ctx->d_bd_ctr; b.cap = ctx->bd_cap; memcpy(b.view, ctx->view, 64);
    LaunchCfg cfg; memset(&cfg, 0, sizeof(cfg)); if ((rc = launch_cfg(ctx, a, cfg))) return rc;
    int bp_ = 1, bc_ = 1;
    if (cfg.use_smem) {
        TR_CUDA(ctx, cudaFuncSetAttribute(k_bdpt_paths<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
        TR_CUDA(ctx, cudaFuncSetAttribute(k_bdpt_connect<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
        TR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bp_, k_bdpt_paths<true>, WF_THREADS, cfg.smem));
        TR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bc_, k_bdpt_connect<true>, WF_THREADS, cfg.smem));
    } else {
        TR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bp_, k_bdpt_paths<false>, WF_THREADS, 0));
        TR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bc_, k_bdpt_connect<false>, WF_THREADS, 0));
    }
    if (bp_ < 1) bp_ = 1; if (bc_ < 1) bc_ = 1;
    cudaStream_t s = ctx->stream;
    uint64_t launches = 0, rays_c = 0, rays_s = 0;
    TR_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    for (int f0 = 0; f0 < n_frames; f0 += F) {
        const int nf = (n_frames - f0 < F) ? n_frames - f0 : F;
        BatchParams bp; bp.frame_begin = frame_begin + f0; bp.n_frames = nf; bp.seed = seed; bp.max_depth = BD_MAX_DEPTH; bp.pad = 0;
        TR_CUDA(ctx, cudaMemcpyAsync(ctx->d_batch_params, &bp, sizeof(bp), cudaMemcpyHostToDevice, s));
        TR_CUDA(ctx, cudaMemsetAsync(ctx->d_ctr, 0, sizeof(TrCounters), s));
        TR_CUDA(ctx, cudaMemsetAsync(ctx->d_bd_ctr, 0, 4 * sizeof(unsigned long long), s));
        TR_CUDA(ctx, cudaMemsetAsync(ctx->d_bd_splat, 0, (size_t)nf * ctx->W * ctx->H * 3 * sizeof(float), s));
        if (a.npix > 0) {
            if (cfg.use_smem) k_bdpt_paths<true><<<ctx->num_sms * bp_, WF_THREADS, cfg.smem, s>>>(a, b);
            else k_bdpt_paths<false><<<ctx->num_sms * bp_, WF_THREADS, 0, s>>>(a, b);
            k_bdpt_items<<<cfg.grid_simple, WF_THREADS, 0, s>>>(a, b);
            if (cfg.use_smem) k_bdpt_connect<true><<<ctx->num_sms * bc_, WF_THREADS, cfg.smem, s>>>(a, b);
            else k_bdpt_connect<false><<<ctx->num_sms * bc_, WF_THREADS, 0, s>>>(a, b);
            launches += 3;
        }
        k_bdpt_film<<<cfg.grid_simple, WF_THREADS, 0, s>>>(a, b); ++launches;
        TR_CHECK_LAUNCH(ctx);
        unsigned long long hc[4];
        TR_CUDA(ctx, cudaMemcpyAsync(hc, ctx->d_bd_ctr, sizeof(hc), cudaMemcpyDeviceToHost, s));
        TR_CUDA(ctx, cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(TrCounters), cudaMemcpyDeviceToHost, s));
        TR_CUDA(ctx, cudaStreamSynchronize(s));
        rays_c += hc[0]; rays_s += hc[1];
    }
    TR_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    TR_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0.0f; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->stats.rays_closest = rays_c; ctx->stats.rays_shadow = rays_s;
    ctx->stats.node_visits = ctx->h_ctr[0].visits[0]; ctx->stats.leaf_tests = ctx->h_ctr[0].visits[1];
    ctx->stats.node_visits_shadow = ctx->h_ctr[0].visits[2]; ctx->stats.leaf_tests_shadow = ctx->h_ctr[0].visits[3];
    ctx->stats.kernel_launches = launches; ctx->stats.ms_total = ms; ctx->stats.frames = n_frames; ctx->stats.paths_in_flight = (int)cap; ctx->stats.chains = 1;
    ctx->stats.ms_trace = ctx->stats.ms_shade = ctx->stats.ms_shadow = 0.0f;
    return TR_OK;
}

extern "C" int tr_render_bdpt_rgb(tr_ctx* ctx, int frame_begin, int n_frames, uint64_t seed) { return render_bdpt(ctx, frame_begin, n_frames, seed); }

// ---- unit hook: vertices, depths and per-strategy contributions of the samples of the LAST batch (frame 0 of it), for the
// pixel list given; layouts as oracle orc_bdpt_pixel_dump: verts n x 13 x 20, depths n x 2, contrib n x 7 x 7 x 4 (e == 1 rows are
// not kept on the device: they are splatted, and stay 0 here)
extern "C" int tr_test_bdpt_dump(tr_ctx* ctx, int n, const int32_t* px, const int32_t* py, float* verts, int32_t* depths, float* contrib) {
    if (!ctx || n <= 0 || !px || !py || !ctx->d_bd_vb) return tr_fail(ctx, TR_ERR_INVALID, "tr_test_bdpt_dump: render a BDPT batch first");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int ntx = (ctx->W + TR_TILE - 1) / TR_TILE;
    std::vector<int> slot_of((size_t)ntx * ((ctx->H + TR_TILE - 1) / TR_TILE));
    TR_CUDA(ctx, cudaMemcpy(slot_of.data(), ctx->d_bd_tile_slot, slot_of.size() * 4, cudaMemcpyDeviceToHost));
    const size_t cap = ctx->bd_cap;
    for (int k = 0; k < n; ++k) {
        int x = px[k], y = py[k];
        if (x < 0 || y < 0 || x >= ctx->W || y >= ctx->H) return tr_fail(ctx, TR_ERR_INVALID, "tr_test_bdpt_dump: pixel out of range");
        int lt = slot_of[(size_t)(y / TR_TILE) * ntx + x / TR_TILE];
        if (lt < 0) return tr_fail(ctx, TR_ERR_INVALID, "tr_test_bdpt_dump: pixel belongs to another rank");
        int xi = x % TR_TILE, yi = y % TR_TILE;
        size_t s = (size_t)lt * 1024 + (((yi >> 3) * 8 + (xi >> 2)) << 5) + ((xi & 3) << 3) + (yi & 7);
        int d[2];
        TR_CUDA(ctx, cudaMemcpy(&d[0], ctx->d_bd_depths + s, 4, cudaMemcpyDeviceToHost));
        TR_CUDA(ctx, cudaMemcpy(&d[1], ctx->d_bd_depths + cap + s, 4, cudaMemcpyDeviceToHost));
        depths[k * 2] = d[0]; depths[k * 2 + 1] = d[1];
        for (int v = 0; v < BD_NVERT; ++v) {
            float4 w[5];
            for (int c = 0; c < 5; ++c) TR_CUDA(ctx, cudaMemcpy(&w[c], ctx->d_bd_vb + (size_t)(v * 5 + c) * cap + s, 16, cudaMemcpyDeviceToHost));
            float* o = verts + ((size_t)k * BD_NVERT + v) * 20;
            const bool valid = v < BD_EYE_MAX ? v < d[0] || (v == d[0] && false) : (v - BD_EYE_MAX) < d[1];
            if (!valid) { for (int c = 0; c < 20; ++c) o[c] = 0.0f; continue; }
            int fl; memcpy(&fl, &w[2].w, 4); int prim; memcpy(&prim, &w[3].w, 4);
            o[0] = w[0].x; o[1] = w[0].y; o[2] = w[0].z; o[3] = w[1].x; o[4] = w[1].y; o[5] = w[1].z; o[6] = w[2].x; o[7] = w[2].y; o[8] = w[2].z;
            o[9] = w[3].x; o[10] = w[3].y; o[11] = w[3].z; o[12] = w[4].x; o[13] = w[4].y; o[14] = w[4].z; o[15] = w[0].w; o[16] = w[1].w;
            o[17] = (float)((fl & 15) + 16 * ((fl >> 4) & 15)); o[18] = (float)prim; o[19] = (float)(fl >> 8);
        }
        float* co = contrib + (size_t)k * 7 * 7 * 4;
        for (int c = 0; c < 7 * 7 * 4; ++c) co[c] = 0.0f;
        for (int e = 2; e <= d[0]; ++e) for (int l = 0; l <= d[1] && l + e - 2 <= BD_MAX_DEPTH; ++l) {
            float4 w; int row = (e - 2) * 6 - ((e - 2) * (e - 3)) / 2 + l;
            TR_CUDA(ctx, cudaMemcpy(&w, ctx->d_bd_contrib + (size_t)row * cap + s, 16, cudaMemcpyDeviceToHost));
            float* o = co + ((e - 1) * 7 + l) * 4; o[0] = w.x; o[1] = w.y; o[2] = w.z; o[3] = (float)(x * 65536 + y);
        }
    }
    return TR_OK;
}
