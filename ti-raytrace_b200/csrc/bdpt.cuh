// bdpt.cuh — BDPT_RGB (integrator/BDPT_RGB.py + integrator/BDPT_Vertex.py) as a staged pipeline; included at the end of
// wavefront.cu (same translation unit: it re-uses WfArgs, the traversal, surface_at / sample_li and the tile mapping).
//
// The reference runs one thread per pixel that builds an eye sub-path (<= 7 vertices) and a light sub-path (<= 6), then
// loops over every (e, l) prefix pair, tracing one shadow ray per connection and rewriting the end-point vertices in place
// for the MIS weight (mis_weight, :258-434).  Here one frame batch is a wavefront over SoA buffers in HBM:
//   k_bdpt_generate        vertex 0 of both sub-paths of every sample + their first rays into the path queue (eye and
//                          light sub-paths share the queue: 2 entries per sample)
//   6 x { k_trace, k_bdpt_vertex }   the persistent closest-hit kernel of PT_RGB (per-lane ray replacement, material-
//                          sorted output queues), then one thread per hit: build the vertex record
//                          vb[(v * 5 + c) * cap + s]  (v = 0..6 eye, 7..12 light; c = 5 float4 words, 80 B / vertex),
//                          patch the previous vertex's reverse pdf, sample the BSDF, append the next ray
//   k_bdpt_items           enumerates the valid strategies of every sample, strategy-major, into a connection queue
//   k_bdpt_connect_gen     one thread per connection: geometric term -> shadow-queue entry when a visibility query is needed
//   k_shadow<QUERY>        the persistent early-exit nearest-hit shadow kernel of PT_RGB, writing (visible ? t : -1) per item
//   k_bdpt_connect_eval    BSDF terms; connections that carry light (three positive channels) are compacted into a MIS queue
//   k_bdpt_mis             the MIS weight computed on register copies of the <= 4 vertices the reference overwrites;
//                          e >= 2 results go to contrib[strategy][s], e == 1 results are splatted (float atomics)
//   k_bdpt_film            per pixel and frame: own strategies summed in the reference's loop order + splats, running mean
// The first version (option "bdpt_wavefront" = 0) is kept as a cross-check: k_bdpt_paths / k_bdpt_connect run one lane per
// sub-path / connection with the warp walking the BVH in lock step; same device functions, same film.
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
    unsigned long long* ctr;           // [0] closest-hit traversals, [1] shadow traversals, [2] connection queue size, [3] MIS queue size (low 32 bits)
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
// Out-of-line copies of the Disney evaluations for the connection / MIS kernels: they are called from up to six places per
// kernel, and the inlined code (3 600 SASS instructions for k_bdpt_mis) made instruction fetch a top stall reason.
__device__ __noinline__ float bd_disney_pdf(V3 N, V3 V, V3 L, float metal, float rough) { return disney_pdf(N, V, L, metal, rough); }
__device__ __noinline__ float2 bd_disney_evaluate_pdf(V3 N, V3 V, V3 L, float metal, float rough) { float o, p; disney_evaluate_pdf(N, V, L, metal, rough, o, p); return make_float2(o, p); }

// One hit of a sub-path (the loop bodies of eye_path, BDPT_RGB.py:118-186, and light_path, :213-249): build vertex `depth`,
// patch the reverse pdf of vertex depth-1, sample the continuation.  counted: the vertex joins the sub-path (depth += 1);
// cont: the walk goes on with (origin, dir, beta, pdfFwd) and (pos, normal) as the new previous vertex.
// The two variants differ where the reference differs: the eye path stores the emitter vertex it hits and measures `to`
// from the offset ray origin with a clamped distance; the light path stops in front of emitters, measures from the
// stored previous position, and multiplies fpdf in a different order.
struct BdStep { bool counted, cont; V3 origin, dir, beta, pos, normal; float pdfFwd; };
__device__ __forceinline__ BdStep bd_vertex(const bool LIGHT, const WfArgs& a, const BdArgs& b, const BatchParams& bp, size_t s, unsigned pix, unsigned frame, int depth,
                                            V3 origin, V3 dir, V3 beta, float pdfFwd, V3 prev_pos, V3 prev_normal, int prim, float hu, float hv, float ht) {
    const int vbase = LIGHT ? BD_EYE_MAX : 0;
    BdStep st; st.counted = false; st.cont = false;
    Surf sf = surface_at(a, prim, hu, hv, origin, dir, ht);
    V3 fn = signf_(dot3(-dir, sf.gn)) * sf.n;
    const float* m = a.material + (size_t)sf.mat * 10;
    const int mt = (int)__ldg(m);
    const float p0 = __ldg(m + 5), p1 = __ldg(m + 6);
    if (LIGHT && mt == TR_MAT_LIGHT) return st;
    V3 to; float dist;
    if (!LIGHT) { to = sf.pos - origin; dist = fmaxf(length3(to), 0.01f); }
    else { to = sf.pos - prev_pos; dist = length3(to); }
    const float inv_dist2 = 1.0f / (dist * dist);
    to = to / dist;
    BV v; v.pos = sf.pos; v.normal = sf.n; v.snormal = fn; v.wo = dir; v.rpdf = 0.0f; v.prim = prim; v.mat = sf.mat; v.delta = 0;
    if (!LIGHT) v.fpdf = pdfFwd * fabsf(dot3(to, prev_normal)) * inv_dist2;
    else v.fpdf = pdfFwd * (fabsf(dot3(to, prev_normal)) * inv_dist2);
    if (!LIGHT && mt == TR_MAT_LIGHT) {
        V3 mcol = mk3(__ldg(m + 2), __ldg(m + 3), __ldg(m + 4));
        v.beta = (beta * mcol) * fabsf(dot3(sf.n, dir)); v.type = BD_VERTEX_LIGHT;
        bv_store(b, s, vbase + depth, v);
        st.counted = true;
        return st;
    }
    v.beta = beta * fabsf(dot3(dir, sf.n)); v.type = BD_VERTEX_SURFACE;
    V3 rc = f4xyz(__ldg(a.matlin + sf.mat));
    float4 R0 = rng4(bp.seed, pix, frame, (LIGHT ? 42u : 1u) + 2u * (unsigned)(depth - 1));
    float4 R1 = rng4(bp.seed, pix, frame, (LIGHT ? 43u : 2u) + 2u * (unsigned)(depth - 1));
    V3 next_dir; float brdf, f_or_b = 1.0f, pdfRev;
    if (mt == TR_MAT_GLASS) { next_dir = glass_sample(dir, sf.n, p0, R0.w, f_or_b); brdf = 1.0f; pdfFwd = 1.0f; v.delta = 1; }
    else { next_dir = disney_sample(dir, fn, p0, p1, R0.w, R1.x, R1.y); disney_evaluate_pdf(fn, -dir, next_dir, p0, p1, brdf, pdfFwd); }
    bv_store(b, s, vbase + depth, v);
    if (!(pdfFwd > 0.0f)) return st;
    if (mt == TR_MAT_GLASS) { pdfRev = 0.0f; pdfFwd = 0.0f; beta = beta * (brdf * rc); }
    else {
        beta = beta * (((brdf * rc) * fabsf(dot3(sf.n, next_dir))) / pdfFwd);
        pdfRev = disney_pdf(fn, next_dir, -dir, p0, p1);
    }
    bv_set_rpdf(b, s, vbase + depth - 1, pdfRev * fabsf(dot3(to, sf.n)) * inv_dist2);
    if (f_or_b < 0.0f) { float Rr = tr_expf(-ht / p1); if (R1.z >= Rr) return st; }
    st.counted = true; st.cont = true;
    st.origin = offset_ray(sf.pos, signf_(f_or_b) * fn); st.dir = next_dir; st.beta = beta; st.pdfFwd = pdfFwd; st.pos = sf.pos; st.normal = sf.n;
    return st;
}

// UtilsFunc.py:322-345
__device__ __forceinline__ void map_to_disk(float u1, float u2, float& r, float& phi) {
    phi = 0.0f; r = 0.0f;
    const float a = 2.0f * u1 - 1.0f, b = 2.0f * u2 - 1.0f;
    if (a > -b) {
        if (a > b) { r = a; phi = (TR_PI_REF / 4.0f) * (b / a); }
        else { r = b; phi = (TR_PI_REF / 4.0f) * (2.0f - a / b); }
    } else {
        if (a < b) { r = -a; phi = (TR_PI_REF / 4.0f) * (4.0f + b / a); }
        else { r = -b; phi = (b == 0.0f) ? 0.0f : (TR_PI_REF / 4.0f) * (6.0f - a / b); }
    }
}

// vertex 0 of a sub-path and its first ray: the lens (eye_path :108-116) or a sampled emitter point with a cosine-hemisphere
// direction (light_path :194-211, Scene.sample_light Scene.py:430-474: point as in sample_li)
template <bool LIGHT>
__device__ __forceinline__ void bd_vertex0(const WfArgs& a, const BdArgs& b, const BatchParams& bp, size_t s, unsigned pix, unsigned frame, int x, int y,
                                           V3& origin, V3& dir, V3& beta, float& pdfFwd) {
    BV v0; v0.snormal = mk3(0.f, 0.f, 0.f); v0.wo = mk3(0.f, 0.f, 0.f); v0.rpdf = 0.0f; v0.prim = 0; v0.mat = 0; v0.delta = 0;
    if (!LIGHT) {
        float jx = 0.0f, jy = 0.0f;
        if (frame != 0) { float4 r = rng4(bp.seed, pix, frame, 0u); jx = r.x - 0.5f; jy = r.y - 0.5f; }
        origin = mk3(a.cam.eye[0], a.cam.eye[1], a.cam.eye[2]); dir = camera_dir(a.cam, x, y, jx, jy);
        v0.pos = origin; v0.normal = dir; v0.beta = mk3(1.f, 1.f, 1.f); v0.fpdf = 1.0f; v0.type = BD_VERTEX_LENS;
        beta = mk3(1.f, 1.f, 1.f); pdfFwd = 1.0f;
    } else {
        float4 Ra = rng4(bp.seed, pix, frame, 40u), Rb = rng4(bp.seed, pix, frame, 41u);
        LightSample ls = sample_li(a, mk3(0.f, 0.f, 0.f), Ra.x, Ra.y, Ra.z, false);
        V3 p = cosine_sample_hemisphere(Rb.x, Rb.y);
        pdfFwd = cosine_hemisphere_pdf(p.z);
        dir = inverse_transform(p, ls.normal);
        if (ls.kind == 2) {                          // spot light (Scene.py:449-463): a direction through a disk at distance `scale`
            const float scale = ls.p2;
            pdfFwd = 1.0f;
            float r, phi; map_to_disk(Rb.z, Rb.w, r, phi);
            const float r1 = scale * tr_tanf(ls.p0), r2 = scale * tr_tanf(ls.p1);
            r *= r2;
            if (r > r1) ls.emission = ls.emission * (1.0f - (r - r1) / (r2 - r1));
            dir = inverse_transform(mk3(r * tr_cosf(phi), r * tr_sinf(phi), sqrtf(fmaxf(0.0f, scale * scale - r * r))), ls.normal);
        } else if (ls.kind == 3) {                   // laser (Scene.py:464-472): along the normal from a point of the rim
            ls.choice_pdf = 1.0f / (float)a.nl;
            const float r = ls.p0, phi = (Rb.z * TR_PI_REF) * 2.0f;
            ls.pos = ls.pos + inverse_transform(mk3(r * tr_cosf(phi), r * tr_sinf(phi), 0.0f), ls.normal);
            dir = ls.normal; pdfFwd = 1.0f;
        }
        const float light_pdf = ls.choice_pdf;
        origin = ls.pos;
        v0.pos = ls.pos; v0.normal = ls.normal; v0.beta = ls.emission / light_pdf; v0.fpdf = light_pdf; v0.wo = dir; v0.type = BD_VERTEX_LIGHT;
        beta = (ls.emission / light_pdf) * fabsf(dot3(ls.normal, dir));
    }
    bv_store(b, s, LIGHT ? BD_EYE_MAX : 0, v0);
}

// One sub-path (BDPT_RGB.py:104-187 eye_path / :189-250 light_path), one lane per sample, the warp walks in lock step.
// The two loops differ where the reference differs: the eye path stores the emitter vertex it hits and measures `to`
// from the offset ray origin with a clamped distance; the light path stops in front of emitters, measures from the
// stored previous position, and multiplies fpdf in a different order.
template <bool LIGHT>
__device__ __forceinline__ void bdpt_subpath(const WfArgs& a, const BdArgs& b, const BatchParams& bp, const TreeView& tv,
                                             size_t s, bool active, unsigned pix, unsigned frame, int x, int y,
                                             unsigned long long& n_closest) {
    const int maxd = LIGHT ? BD_LIGHT_MAX : BD_EYE_MAX;
    V3 origin = mk3(0.f, 0.f, 0.f), dir = mk3(1.f, 1.f, 1.f), beta = mk3(1.f, 1.f, 1.f), prev_pos = origin, prev_normal = dir;
    float pdfFwd = 1.0f;
    int depth = 1;
    if (active) { bd_vertex0<LIGHT>(a, b, bp, s, pix, frame, x, y, origin, dir, beta, pdfFwd); prev_pos = origin; prev_normal = LIGHT ? f4xyz(b.vb[(size_t)(BD_EYE_MAX * 5 + 1) * b.cap + s]) : dir; }
    bool alive = active;
    for (int it = 1; it < maxd; ++it) {
        if (__ballot_sync(0xffffffffu, alive) == 0u) break;
        RayPre r = make_ray(origin, dir);
        HitRec h = trace_closest(tv, a.root, r, alive, a.ctr->visits);
        if (!alive) continue;
        ++n_closest;
        if (h.prim < 0) { alive = false; continue; }
        BdStep st = bd_vertex(LIGHT, a, b, bp, s, pix, frame, depth, origin, dir, beta, pdfFwd, prev_pos, prev_normal, h.prim, h.u, h.v, h.t);
        if (st.counted) depth += 1;
        alive = st.cont;
        if (st.cont) { prev_pos = st.pos; prev_normal = st.normal; origin = st.origin; dir = st.dir; beta = st.beta; pdfFwd = st.pdfFwd; }
    }
    if (active) b.depths[(LIGHT ? b.cap : 0) + s] = depth;
}

// global-memory view of the tree for the simple (one lane = one ray) walks of the lock-step cross-check pipeline
__device__ __forceinline__ TreeView global_tree(const WfArgs& a) {
    TreeView tv; tv.snodes = tv.sleaves = nullptr; tv.gnodes = a.nodes2; tv.gleaves = a.leaves4; tv.top = 0;
    return tv;
}

__global__ void __launch_bounds__(WF_THREADS) k_bdpt_paths(WfArgs a, BdArgs b) {
    const TreeView tv = global_tree(a);
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
        if (light) bdpt_subpath<true>(a, b, bp, tv, (size_t)s, active, pix, frame, x, y, n_closest);
        else bdpt_subpath<false>(a, b, bp, tv, (size_t)s, active, pix, frame, x, y, n_closest);
    }
    if (n_closest) atomicAdd(b.ctr, n_closest);
}

// A strategy of a sample is worth a queue item when both prefixes exist -- and, for l == 0 (the eye sub-path hit an emitter,
// BDPT_RGB.py:447-451), only when e - 1 is the LAST eye vertex: eye_path stops at the first emitter it hits, so every shorter
// l == 0 prefix ends on a surface and contributes exactly 0.  k_bdpt_film and the dump hook apply the same rule.
__host__ __device__ __forceinline__ bool bd_strategy_live(int e, int l, int ed, int ld) { return e <= ed && l <= ld && (l != 0 || e == ed); }

// Connection queue: item = sample | e << 26 | l << 29, emitted strategy-major per warp so that the connect kernel's warps
// are (mostly) homogeneous in strategy type.  Loop bounds of BDPT_RGB.py:625-637.
__global__ void __launch_bounds__(WF_THREADS) k_bdpt_items(WfArgs a, BdArgs b) {
    const BatchParams bp = *a.bp;
    const int nsamp = bp.n_frames * a.npix, nsamp_r = (nsamp + 31) & ~31;
    const int stride = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nsamp_r; s += stride) {
        int ed = 0, ld = 0;
        if (s < nsamp) { ed = b.depths[s]; ld = b.depths[b.cap + s]; }
        // one queue reservation per warp: count the warp's strategies first, then write them strategy-major
        int total = 0;
        for (int e = 1; e <= BD_EYE_MAX; ++e)
            for (int l = 0; l <= BD_LIGHT_MAX; ++l) {
                const int depth = l + e - 2;
                if ((l == 1 && e == 1) || depth < 0 || depth > BD_MAX_DEPTH) continue;      // warp-uniform
                total += __popc(__ballot_sync(0xffffffffu, bd_strategy_live(e, l, ed, ld)));
            }
        int off = 0;
        if (lane == 0 && total > 0) off = atomicAdd((int*)(b.ctr + 2), total);
        off = __shfl_sync(0xffffffffu, off, 0);
        for (int e = 1; e <= BD_EYE_MAX; ++e)
            for (int l = 0; l <= BD_LIGHT_MAX; ++l) {
                const int depth = l + e - 2;
                if ((l == 1 && e == 1) || depth < 0 || depth > BD_MAX_DEPTH) continue;
                const bool valid = bd_strategy_live(e, l, ed, ld);
                const unsigned m = __ballot_sync(0xffffffffu, valid);
                if (valid) b.items[off + __popc(m & ((1u << lane) - 1u))] = (unsigned)s | ((unsigned)e << 26) | ((unsigned)l << 29);
                off += __popc(m);
            }
    }
}

// mis_weight (BDPT_RGB.py:258-434).  The reference saves <= 4 end-point vertices, overwrites them in the per-pixel vertex
// fields, walks the pdf ratios and restores them; here the end points are the register copies the connection already holds
// (E1 = eye[e-1], L1 = light[l-1]), of eye[e-2] / light[l-2] only position and type are fetched, and the walk reads each
// remaining vertex's (fpdf, rpdf, delta) scalars once.  For l == 1 the freshly sampled emitter vertex (smp_*) stands in for
// light[0]; for e == 1 the lens sample that replaces eye[0] changes nothing that is read (same position, fpdf = 1).
__device__ __forceinline__ float bdpt_mis_weight(const WfArgs& a, const BdArgs& b, size_t s, int e, int l, const BV& Ein, const BV& Lin,
                                                 V3 smp_pos, V3 smp_normal, float smp_fpdf) {
    if (l + e == 2) return 1.0f;
    // end points: position, normals, type, material, fpdf
    V3 E1pos, E1normal = mk3(0.f, 0.f, 0.f), E1snormal = E1normal; int E1type, E1mat = 0, E1prim = 0; float E1fpdf;
    if (e == 1) { E1pos = mk3(a.cam.eye[0], a.cam.eye[1], a.cam.eye[2]); E1type = BD_VERTEX_LENS; E1fpdf = 1.0f; }
    else { E1pos = Ein.pos; E1normal = Ein.normal; E1snormal = Ein.snormal; E1type = Ein.type; E1mat = Ein.mat; E1prim = Ein.prim; E1fpdf = Ein.fpdf; }
    V3 L1pos = smp_pos, L1normal = smp_normal, L1snormal = smp_normal; int L1mat = 0; float L1fpdf = smp_fpdf;
    if (l > 1) { L1pos = Lin.pos; L1normal = Lin.normal; L1snormal = Lin.snormal; L1mat = Lin.mat; L1fpdf = Lin.fpdf; }
    V3 E2pos = E1pos, L2pos = L1pos; int E2type = 0, L2type = 0;
    if (e > 1) { E2pos = f4xyz(b.vb[(size_t)((e - 2) * 5) * b.cap + s]); E2type = __float_as_int(((const float*)(b.vb + (size_t)((e - 2) * 5 + 2) * b.cap + s))[3]) & 15; }
    if (l > 1) { const int v = BD_EYE_MAX + l - 2; L2pos = f4xyz(b.vb[(size_t)(v * 5) * b.cap + s]); L2type = __float_as_int(((const float*)(b.vb + (size_t)(v * 5 + 2) * b.cap + s))[3]) & 15; }
    float e1_rpdf, e2_rpdf = 0.0f, l1_rpdf = 0.0f, l2_rpdf = 0.0f;
    const V3 axis = mk3(b.view[8], b.view[9], b.view[10]);                // Camera.get_optical_axis (Camera.py:128-129)
    // eye[e-1].rpdf (:290-316)
    if (l == 0) {
        float pdfPos = 1.0f / __ldg(&a.shade[E1prim].q[2].w), pdfChoice = 1.0f / (float)a.nl;
        e1_rpdf = pdfPos * pdfChoice;
    } else if (l == 1) {
        if (E1type == BD_VERTEX_SURFACE) {
            V3 to = E1pos - L1pos; float dist = length3(to); to = to / dist;
            float c = fabsf(dot3(to, L1normal));
            e1_rpdf = cosine_hemisphere_pdf(c) * c / (dist * dist);
        } else e1_rpdf = 1.0f;
    } else {
        V3 wi = L2pos - L1pos, wo = E1pos - L1pos; float dist = length3(wo);
        wi = normalize3(wi); wo = normalize3(wo);
        float pdf = 1.0f;
        if (L1mat == TR_MAT_DISNEY) { const float* m = a.material + (size_t)L1mat * 10; pdf = bd_disney_pdf(L1snormal, wi, wo, __ldg(m + 5), __ldg(m + 6)); }   // material INDEX == 0 (sic, :312)
        e1_rpdf = pdf * fabsf(dot3(L1normal, wo)) / (dist * dist);
    }
    // light[l-1].rpdf (:317-343)
    if (l > 0) {
        if (e > 1) {
            if (E1type == BD_VERTEX_SURFACE) {
                V3 wi = E2pos - E1pos, wo = L1pos - E1pos; float dist = length3(wo);
                wi = normalize3(wi); wo = normalize3(wo);
                float pdf = 1.0f;
                if (E1mat == TR_MAT_DISNEY) { const float* m = a.material + (size_t)E1mat * 10; pdf = bd_disney_pdf(E1snormal, wi, wo, __ldg(m + 5), __ldg(m + 6)); }
                l1_rpdf = pdf * fabsf(dot3(E1normal, wo)) / (dist * dist);
            } else l1_rpdf = 1.0f;
        } else {
            V3 to = E1pos - L1pos; float dist = length3(to); to = to / dist;      // e == 1: eye[0] = the lens
            l1_rpdf = dot3(to, axis) / (dist * dist);
        }
    }
    // eye[e-2].rpdf (:345-369)
    if (e > 1) {
        if (l == 0) {
            V3 to = E2pos - E1pos; float dist = length3(to); to = to / dist;
            float pdfDir = cosine_hemisphere_pdf(fabsf(dot3(to, E1normal)));
            float LdotN = dot3(to, E1normal);
            e2_rpdf = fabsf(pdfDir * LdotN) / (dist * dist);
        } else if (E1type == BD_VERTEX_SURFACE) {
            V3 wi = L1pos - E1pos, wo = E2pos - E1pos; float dist = length3(wo);
            wi = normalize3(wi); wo = normalize3(wo);
            const float* m = a.material + (size_t)E1mat * 10;
            float pdf = bd_disney_pdf(E1snormal, wi, wo, __ldg(m + 5), __ldg(m + 6));
            e2_rpdf = pdf / (dist * dist);
            if (E2type == BD_VERTEX_SURFACE) e2_rpdf *= fabsf(dot3(E1normal, wo));
        } else e2_rpdf = 1.0f;
    }
    // light[l-2].rpdf (:371-388)
    if (l > 1) {
        if (E1type != BD_VERTEX_LIGHT) {
            V3 wi = E1pos - L1pos, wo = L2pos - L1pos; float dist = length3(wo);
            wi = normalize3(wi); wo = normalize3(wo);
            float pdf = 1.0f;
            if (L1mat == TR_MAT_DISNEY) { const float* m = a.material + (size_t)L1mat * 10; pdf = bd_disney_pdf(L1normal, wi, wo, __ldg(m + 5), __ldg(m + 6)); }
            l2_rpdf = pdf / (dist * dist);
            if (L2type == BD_VERTEX_SURFACE) l2_rpdf *= fabsf(dot3(L1normal, wo));
        } else l2_rpdf = 1.0f;
    }
    // the walks (:394-416); delta of eye[e-1] and light[l-1] is forced to 0 (:283-286); delta of eye[0] (lens) and light[0] is 0
    float weight_sum = 0.0f, weight = 1.0f;
    int d_above = 0;                                                      // delta of vertex k (k = e-1 first)
    for (int k = e - 1; k > 0; --k) {
        float f, r; int d_below = 0;
        if (k == e - 1) { f = E1fpdf; r = e1_rpdf; }
        else { f = ((const float*)(b.vb + (size_t)(k * 5) * b.cap + s))[3]; r = (k == e - 2) ? e2_rpdf : ((const float*)(b.vb + (size_t)(k * 5 + 1) * b.cap + s))[3]; }
        if (k - 1 > 0) d_below = (__float_as_int(((const float*)(b.vb + (size_t)((k - 1) * 5 + 2) * b.cap + s))[3]) >> 4) & 15;
        weight *= (r == 0.0f ? 1.0f : r) / (f == 0.0f ? 1.0f : f);
        if (d_above == 0 && d_below == 0) weight_sum += weight;
        d_above = d_below;
    }
    weight = 1.0f; d_above = 0;
    for (int k = l - 1; k >= 0; --k) {
        float f, r; int d_below = 0;
        const int v = BD_EYE_MAX + k;
        if (k == l - 1) { f = L1fpdf; r = l1_rpdf; }
        else { f = ((const float*)(b.vb + (size_t)(v * 5) * b.cap + s))[3]; r = (k == l - 2) ? l2_rpdf : ((const float*)(b.vb + (size_t)(v * 5 + 1) * b.cap + s))[3]; }
        if (k - 1 > 0) d_below = (__float_as_int(((const float*)(b.vb + (size_t)((v - 1) * 5 + 2) * b.cap + s))[3]) >> 4) & 15;
        weight *= (r == 0.0f ? 1.0f : r) / (f == 0.0f ? 1.0f : f);
        if (d_above == 0 && d_below == 0) weight_sum += weight;
        d_above = d_below;
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

// connect_path (BDPT_RGB.py:436-580) in three phases shared by both pipelines:
//   bd_connect_geom   which vertices, whether a visibility query is needed and which ray / target primitive it uses
//   (shadow query)    lock-step in k_bdpt_connect, the persistent k_shadow<QUERY> kernel in the wavefront pipeline
//   bd_connect_finish BSDF terms, geometric factor, MIS weight, contribution / splat write
struct BdConn {
    bool need; int kind, target, nu, nv;          // kind: 1 e == 1 (light tracing), 2 l == 1 (next-event), 3 general
    V3 ro, rd, radiance; float c0, c1, dist;
    BV ev, lv; LightSample ls;
};
__device__ __forceinline__ void bd_connect_geom(const WfArgs& a, const BdArgs& b, const BatchParams& bp, size_t s, int e, int l, BdConn& c) {
    c.need = false; c.kind = 0; c.target = 0; c.nu = -1; c.nv = -1; c.c0 = c.c1 = 0.0f; c.dist = 1.0f;
    c.radiance = mk3(0.f, 0.f, 0.f); c.ro = mk3(0.f, 0.f, 0.f); c.rd = mk3(1.f, 1.f, 1.f);
    if (l == 0) {
        c.ev = bv_load(b, s, e - 1);
        if (c.ev.type == BD_VERTEX_LIGHT) c.radiance = c.ev.beta;
    } else if (e == 1) {
        c.lv = bv_load(b, s, BD_EYE_MAX + l - 1);
        bd_image_point(a, b, c.lv.pos, c.nu, c.nv, c.rd);
        c.c0 = dot3(c.rd, c.lv.snormal);                                   // NdotL
        if (c.nu >= 0 && c.lv.delta != 1 && c.c0 < 0.0f && c.lv.type == BD_VERTEX_SURFACE) {
            c.need = true; c.kind = 1; c.target = c.lv.prim; c.ro = mk3(a.cam.eye[0], a.cam.eye[1], a.cam.eye[2]);
        }
    } else if (l == 1) {
        c.ev = bv_load(b, s, e - 1);
        if (c.ev.delta != 1) {
            int x, y; slot_to_pixel(a, (int)(s % (size_t)a.npix), x, y);
            unsigned pix = ((unsigned)x << 16) | (unsigned)y, frame = (unsigned)bp.frame_begin + (unsigned)(s / (size_t)a.npix);
            float4 R0 = rng4(bp.seed, pix, frame, 1u + 2u * (unsigned)(e - 2));
            c.ro = offset_ray(c.ev.pos, c.ev.snormal);
            c.ls = sample_li(a, c.ro, R0.x, R0.y, R0.z);
            c.c0 = dot3(c.ls.dir, c.ls.normal); c.c1 = dot3(c.ls.dir, c.ev.snormal);        // NdotLl, NdotLe
            c.need = true; c.kind = 2; c.target = c.ls.prim; c.rd = -c.ls.dir;
        }
    } else {
        c.ev = bv_load(b, s, e - 1); c.lv = bv_load(b, s, BD_EYE_MAX + l - 1);
        if (c.lv.delta != 1 && c.ev.delta != 1 && c.ev.type == BD_VERTEX_SURFACE && c.lv.type == BD_VERTEX_SURFACE) {
            V3 d = c.ev.pos - c.lv.pos; c.dist = length3(d); d = d / c.dist;
            c.c0 = dot3(d, c.lv.snormal); c.c1 = dot3(d, c.ev.snormal);                 // NdotLl, NdotLe
            c.need = true; c.kind = 3; c.target = c.ev.prim; c.ro = c.lv.pos; c.rd = d;
        }
    }
}
// vis / tt: the target primitive is the nearest hit of the query ray, at distance tt (closet_hit_shadow, Scene.py:671-699).
// Returns the unweighted contribution of the strategy.
__device__ __forceinline__ V3 bd_connect_radiance(const WfArgs& a, const BdConn& c, bool vis, float tt) {
    V3 radiance = c.radiance;
    const BV& ev = c.ev; const BV& lv = c.lv;
    if (c.need) {
        if (c.kind == 1 && vis) {
            const float* m = a.material + (size_t)lv.mat * 10;
            float2 bp_ = bd_disney_evaluate_pdf(lv.snormal, -lv.wo, -c.rd, __ldg(m + 5), __ldg(m + 6)); const float brdf = bp_.x, pdf = bp_.y;
            if (pdf > 0.0f) { float G = fabsf(c.c0) / (tt * tt); radiance = (((G * lv.beta) * f4xyz(__ldg(a.matlin + lv.mat))) * brdf) / pdf; }
        } else if (c.kind == 2 && vis && tt > BD_EPS) {
            const float* m = a.material + (size_t)ev.mat * 10;
            const float light_pdf = c.ls.choice_pdf;
            float2 bp_ = bd_disney_evaluate_pdf(ev.snormal, -ev.wo, -c.ls.dir, __ldg(m + 5), __ldg(m + 6)); const float brdf = bp_.x, pdf = bp_.y;
            if (pdf > 0.0f) {
                float G = fabsf(c.c1 * c.c0) / (tt * tt);
                radiance = (((((G * ev.beta) * brdf) / pdf) * f4xyz(__ldg(a.matlin + ev.mat))) * c.ls.emission) / light_pdf;
            }
        } else if (c.kind == 3 && vis && tt > BD_EPS) {
            const float* mE = a.material + (size_t)ev.mat * 10; const float* mL = a.material + (size_t)lv.mat * 10;
            const float2 bl = bd_disney_evaluate_pdf(lv.snormal, -lv.wo, c.rd, __ldg(mL + 5), __ldg(mL + 6)), be = bd_disney_evaluate_pdf(ev.snormal, -ev.wo, -c.rd, __ldg(mE + 5), __ldg(mE + 6));
            const float brdfL = bl.x, lpdf = bl.y, brdfE = be.x, epdf = be.y;
            if (brdfL > 0.0f && brdfE > 0.0f) {
                float G = fabsf(c.c1 * c.c0) / (c.dist * c.dist);
                radiance = (((((((G * ev.beta) * lv.beta) * brdfL) / lpdf) * brdfE) / epdf) * f4xyz(__ldg(a.matlin + ev.mat))) * f4xyz(__ldg(a.matlin + lv.mat));
            }
        }
    }
    return radiance;
}
// mis_weight is only evaluated for contributions with three positive channels (BDPT_RGB.py:578-579)
__device__ __forceinline__ bool bd_needs_mis(V3 r) { return r.x > 0.0f && r.y > 0.0f && r.z > 0.0f; }
// the freshly sampled emitter vertex of the l == 1 strategy stands in for light[0] in the weight (sample.* of :522-528)
__device__ __forceinline__ float bd_connect_weight(const WfArgs& a, const BdArgs& b, size_t s, int e, int l, const BdConn& c) {
    V3 z = mk3(0.f, 0.f, 0.f);
    if (c.kind == 2) return bdpt_mis_weight(a, b, s, e, l, c.ev, c.lv, c.ls.pos, c.ls.normal, c.ls.choice_pdf);
    return bdpt_mis_weight(a, b, s, e, l, c.ev, c.lv, z, z, 1.0f);
}
// e >= 2: the strategy's slot of the per-sample contribution table; e == 1: splat onto the pixel the light vertex projects to
__device__ __forceinline__ void bd_connect_write(const WfArgs& a, const BdArgs& b, size_t s, int e, int l, const BdConn& c, V3 radiance) {
    if (e == 1) {
        if (c.nu >= 0 && (radiance.x != 0.0f || radiance.y != 0.0f || radiance.z != 0.0f)) {
            float* o = b.splat + ((size_t)(s / (size_t)a.npix) * a.W * a.H + (size_t)c.nu * a.H + c.nv) * 3;
            atomicAdd(o, radiance.x); atomicAdd(o + 1, radiance.y); atomicAdd(o + 2, radiance.z);
        }
    } else b.contrib[(size_t)bd_row(e, l) * b.cap + s] = make_float4(radiance.x, radiance.y, radiance.z, 0.0f);
}
__device__ __forceinline__ void bd_connect_finish(const WfArgs& a, const BdArgs& b, size_t s, int e, int l, BdConn& c, bool vis, float tt) {
    V3 radiance = bd_connect_radiance(a, c, vis, tt);
    if (bd_needs_mis(radiance)) radiance = radiance * bd_connect_weight(a, b, s, e, l, c);
    bd_connect_write(a, b, s, e, l, c, radiance);
}

// lock-step pipeline: one lane per (sample, e, l), the warp walks the BVH together
__global__ void __launch_bounds__(WF_THREADS) k_bdpt_connect(WfArgs a, BdArgs b) {
    const TreeView tv = global_tree(a);
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
        BdConn c; c.need = false;
        if (active) bd_connect_geom(a, b, bp, s, e, l, c);
        bool vis = false; float tt = TR_INF;
        if (__ballot_sync(0xffffffffu, c.need) != 0u) {
            RayPre r = make_ray(c.need ? c.ro : mk3(0.f, 0.f, 0.f), c.need ? c.rd : mk3(1.f, 1.f, 1.f));
            int tleaf = c.need ? __ldg(a.leaf_of_prim + c.target) : 0;
            vis = trace_shadow_visible(tv, a.root, r, c.need, tleaf, a.ctr->visits + 2, &tt);
            if (c.need) ++n_shadow;
        }
        if (active) bd_connect_finish(a, b, s, e, l, c, vis, tt);
    }
    if (n_shadow) atomicAdd(b.ctr + 1, n_shadow);
}

// ------------------------------------------------------------------ wavefront pipeline
// Path-queue record (the PT_RGB layout, so k_trace is shared):  A = (o.xyz, d.x)  B = (d.y, d.z, pdfFwd, sample | light << 31)
// C = (beta.rgb, -).  Stage d traces the segment that ends in vertex d + 1 of either sub-path.
__global__ void __launch_bounds__(WF_THREADS) k_bdpt_generate(WfArgs a, BdArgs b) {
    const BatchParams bp = *a.bp;
    const int nsamp = bp.n_frames * a.npix, nsamp_r = (nsamp + WF_THREADS - 1) / WF_THREADS * WF_THREADS;    // block-uniform halves (block_append)
    const int stride = gridDim.x * blockDim.x;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < 2 * nsamp_r; w += stride) {
        const bool light = w >= nsamp_r;                                // warp-uniform
        const int s = w - (light ? nsamp_r : 0);
        int x = 0, y = 0; bool active = false; unsigned frame = 0;
        if (s < nsamp) { int f = s / a.npix, p = s - f * a.npix; frame = (unsigned)(bp.frame_begin + f); active = slot_to_pixel(a, p, x, y); }
        const unsigned pix = ((unsigned)x << 16) | (unsigned)y;
        if (s < nsamp) b.depths[(light ? b.cap : 0) + s] = active ? 1 : 0;
        V3 origin = mk3(0.f, 0.f, 0.f), dir = mk3(1.f, 1.f, 1.f), beta = dir; float pdfFwd = 1.0f;
        if (active) { if (light) bd_vertex0<true>(a, b, bp, (size_t)s, pix, frame, x, y, origin, dir, beta, pdfFwd); else bd_vertex0<false>(a, b, bp, (size_t)s, pix, frame, x, y, origin, dir, beta, pdfFwd); }
        int q = block_append(&a.ctr->nq[0], active);
        if (active) {
            a.pa[0][q] = make_float4(origin.x, origin.y, origin.z, dir.x);
            a.pb[0][q] = make_float4(dir.y, dir.z, pdfFwd, __uint_as_float((unsigned)s | (light ? SPEC_BIT : 0u)));
            a.pc[0][q] = make_float4(beta.x, beta.y, beta.z, 0.0f);
        }
    }
}

// one thread per traced segment of stage d, walking the material-sorted queues k_trace filled (terminal | Disney | glass)
__global__ void __launch_bounds__(WF_THREADS, 4) k_bdpt_vertex(WfArgs a, BdArgs b, int d) {
    const BatchParams bp = *a.bp;
    const int n0 = a.ctr->ncls[d][0], n1 = a.ctr->ncls[d][1], n2 = a.ctr->ncls[d][2];
    const int n = n0 + n1 + n2, n_r = (n + WF_THREADS - 1) / WF_THREADS * WF_THREADS;       // block-uniform trip count (block_append)
    const int pp = d & 1, np_ = pp ^ 1, depth = d + 1;
    const int stride = gridDim.x * blockDim.x;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_r; w += stride) {
        BdStep st; st.cont = false; unsigned sw = 0; bool light = false;
        if (w < n) {
            int cl = (w < n0) ? 0 : (w < n0 + n1 ? 1 : 2);
            int q = a.cls[(size_t)cl * a.cap + (w - (cl == 0 ? 0 : (cl == 1 ? n0 : n0 + n1)))];
            float4 A = a.pa[pp][q], B = a.pb[pp][q], C = a.pc[pp][q], Hh = a.hit[q];
            const int prim = __float_as_int(Hh.y);
            sw = __float_as_uint(B.w); light = (sw & SPEC_BIT) != 0;
            const size_t s = sw & ~SPEC_BIT;
            if (prim >= 0) {
                int x, y; slot_to_pixel(a, (int)(s % (size_t)a.npix), x, y);
                const unsigned pix = ((unsigned)x << 16) | (unsigned)y, frame = (unsigned)bp.frame_begin + (unsigned)(s / (size_t)a.npix);
                const int vprev = (light ? BD_EYE_MAX : 0) + depth - 1;
                V3 prev_pos = f4xyz(b.vb[(size_t)(vprev * 5) * b.cap + s]), prev_normal = f4xyz(b.vb[(size_t)(vprev * 5 + 1) * b.cap + s]);
                V3 o = mk3(A.x, A.y, A.z), dir = mk3(A.w, B.x, B.y), beta = mk3(C.x, C.y, C.z);
                st = bd_vertex(light, a, b, bp, s, pix, frame, depth, o, dir, beta, B.z, prev_pos, prev_normal, prim, Hh.z, Hh.w, Hh.x);   // one shared copy of the code for both sub-paths
                if (st.counted) b.depths[(light ? b.cap : 0) + s] = depth + 1;
                if (depth + 1 >= (light ? BD_LIGHT_MAX : BD_EYE_MAX)) st.cont = false;         // the sub-path is full
            }
        }
        int qn = block_append(&a.ctr->nq[d + 1], st.cont);
        if (st.cont) {
            a.pa[np_][qn] = make_float4(st.origin.x, st.origin.y, st.origin.z, st.dir.x);
            a.pb[np_][qn] = make_float4(st.dir.y, st.dir.z, st.pdfFwd, __uint_as_float(sw));
            a.pc[np_][qn] = make_float4(st.beta.x, st.beta.y, st.beta.z, 0.0f);
        }
    }
}

// one thread per connection: geometry, then a shadow-queue entry  sa = (o.xyz, d.x)  sb = (d.y, d.z, target prim, item)
__global__ void __launch_bounds__(WF_THREADS) k_bdpt_connect_gen(WfArgs a, BdArgs b) {
    const BatchParams bp = *a.bp;
    const int n = (int)b.ctr[2], n_r = (n + WF_THREADS - 1) / WF_THREADS * WF_THREADS;       // block-uniform trip count (block_append)
    const int stride = gridDim.x * blockDim.x;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < n_r; it += stride) {
        BdConn c; c.need = false;
        if (it < n) { unsigned item = b.items[it]; bd_connect_geom(a, b, bp, item & 0x3ffffffu, (item >> 26) & 7, (int)(item >> 29), c); }
        int q = block_append(&a.ctr->nshadow[0], c.need);
        if (c.need) {
            a.sa[0][q] = make_float4(c.ro.x, c.ro.y, c.ro.z, c.rd.x);
            a.sb[0][q] = make_float4(c.rd.y, c.rd.z, __int_as_float(c.target), __int_as_float(it));
        }
    }
}
// one thread per connection: geometry again (cheaper than spilling BdConn to HBM), the query result, BSDF terms.  Only a
// fraction of the connections carries light and needs the MIS weight: those are compacted into their own queue
// (radiance.rgb, item) -- it lives in the connection shadow queue, which is free once k_shadow<QUERY> is done -- so that the
// weight kernel runs with full warps instead of a third of the lanes.
__global__ void __launch_bounds__(WF_THREADS, 3) k_bdpt_connect_eval(WfArgs a, BdArgs b) {
    const BatchParams bp = *a.bp;
    const int n = (int)b.ctr[2], n_r = (n + WF_THREADS - 1) / WF_THREADS * WF_THREADS;       // block-uniform trip count (block_append)
    const int stride = gridDim.x * blockDim.x;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < n_r; it += stride) {
        bool defer = false; V3 radiance = mk3(0.f, 0.f, 0.f);
        if (it < n) {
            unsigned item = b.items[it];
            const size_t s = item & 0x3ffffffu; const int e = (item >> 26) & 7, l = (int)(item >> 29);
            BdConn c; bd_connect_geom(a, b, bp, s, e, l, c);
            float tt = c.need ? a.vis[it] : -1.0f;
            radiance = bd_connect_radiance(a, c, tt >= 0.0f, tt);
            defer = bd_needs_mis(radiance) && (l + e != 2);
            if (!defer) bd_connect_write(a, b, s, e, l, c, radiance);
        }
        int q = block_append((int*)(b.ctr + 3), defer);
        if (defer) a.sa[0][q] = make_float4(radiance.x, radiance.y, radiance.z, __int_as_float(it));
    }
}
// one thread per light-carrying connection: mis_weight (BDPT_RGB.py:258-434), then the contribution / splat write
__global__ void __launch_bounds__(WF_THREADS, 3) k_bdpt_mis(WfArgs a, BdArgs b) {
    const BatchParams bp = *a.bp;
    const int n = (int)b.ctr[3];
    const int stride = gridDim.x * blockDim.x;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
        const float4 Q = a.sa[0][q];
        const unsigned item = b.items[__float_as_int(Q.w)];
        const size_t s = item & 0x3ffffffu; const int e = (item >> 26) & 7, l = (int)(item >> 29);
        BdConn c; bd_connect_geom(a, b, bp, s, e, l, c);                   // end-point vertices, light sample, splat pixel: as in the eval pass
        V3 radiance = mk3(Q.x, Q.y, Q.z) * bd_connect_weight(a, b, s, e, l, c);
        bd_connect_write(a, b, s, e, l, c, radiance);
    }
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
                    for (int l = 0; l <= ld && l + e - 2 <= BD_MAX_DEPTH; ++l)
                        if (bd_strategy_live(e, l, ed, ld)) rad = rad + f4xyz(b.contrib[(size_t)bd_row(e, l) * b.cap + s]);     // the others are exactly 0
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
    if ((rc = tr_stats_resolve(ctx))) return rc;
    ctx->present_sum = false;
    if (!ctx->h_ring) TR_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_ring, sizeof(TrCounters) * TR_MAX_CHAINS * TR_RING_BATCHES, cudaHostAllocDefault));
    if ((rc = fill_args(ctx, a, false))) return rc;
    if (ctx->nl <= 0) return tr_fail(ctx, TR_ERR_INVALID, "tr_render_bdpt_rgb: the scene has no emitter (Scene.sample_light needs one)");
    if (!ctx->view_set) return tr_fail(ctx, TR_ERR_INVALID, "tr_render_bdpt_rgb: tr_camera_set was called without the view matrix (Camera.get_image_point needs it)");
    const bool wave = ctx->opt_bdpt_wavefront != 0;
    // bytes per sample: 13 vertex records + 21 contributions + 26 queue items + depths; the wavefront pipeline adds two path-queue
    // slots (PT_RGB's 252 B each) and a worst-case connection shadow queue (26 x (32 B + 4 B))
    const size_t per_sample = (size_t)BD_NVERT * 80 + BD_NCONTRIB * 16 + BD_NSTRAT * 4 + 8 + (wave ? 2 * 252 + BD_NSTRAT * 36 : 0);
    size_t budget = ctx->opt_max_paths * 2 * 252 / per_sample;     // twice PT's byte budget (2.9 KB per sample): 32 spp of a 512^2 image = 8.4 M samples = 25 GB = ONE batch (C5: 6 batches 70.2, 2: 67.4, 1: 66.4 ms/step)
    const size_t npix = (size_t)(a.npix > 0 ? a.npix : 1);
    int F = ctx->opt_batch_frames > 0 ? ctx->opt_batch_frames : (int)(budget / npix);
    if (F < 1) F = 1; if (F > n_frames) F = n_frames;
    if (ctx->opt_batch_frames <= 0) { const int nb = (n_frames + F - 1) / F; F = (n_frames + nb - 1) / nb; }     // equal batches
    const size_t cap = (size_t)F * npix;
    if (cap >= ((size_t)1 << 26)) return tr_fail(ctx, TR_ERR_INVALID, "tr_render_bdpt_rgb: %zu samples per batch exceed the queue encoding (2^26)", cap);
    if (cap > ctx->bd_cap) {
        if ((rc = tr_realloc(ctx, &ctx->d_bd_vb, cap * BD_NVERT * 5))) return rc;
        if ((rc = tr_realloc(ctx, &ctx->d_bd_depths, cap * 2))) return rc;
        if ((rc = tr_realloc(ctx, &ctx->d_bd_contrib, cap * BD_NCONTRIB))) return rc;
        if ((rc = tr_realloc(ctx, &ctx->d_bd_items, cap * BD_NSTRAT))) return rc;
        ctx->bd_cap = cap;
    }
    if (wave) {
        if ((rc = ensure_wavefront(ctx, 2 * ctx->bd_cap + 64))) return rc;
        for (int k = 0; k < 2; ++k) if ((rc = tr_realloc(ctx, &ctx->d_bd_sq[k], ctx->bd_cap * BD_NSTRAT))) return rc;
        if ((rc = tr_realloc(ctx, &ctx->d_bd_vis, ctx->bd_cap * BD_NSTRAT))) return rc;
        if ((rc = fill_args(ctx, a, false))) return rc;
        a.sa[0] = ctx->d_bd_sq[0]; a.sb[0] = ctx->d_bd_sq[1]; a.vis = ctx->d_bd_vis; a.tail_max = 0;
    }
    if (F > ctx->bd_splat_frames || !ctx->d_bd_splat) { if ((rc = tr_realloc(ctx, &ctx->d_bd_splat, (size_t)F * ctx->W * ctx->H * 3))) return rc; ctx->bd_splat_frames = F; }
    if (!ctx->d_bd_ctr) TR_CUDA(ctx, cudaMalloc((void**)&ctx->d_bd_ctr, 4 * sizeof(unsigned long long)));
    BdArgs b; memset(&b, 0, sizeof(b));
    b.vb = ctx->d_bd_vb; b.depths = ctx->d_bd_depths; b.contrib = ctx->d_bd_contrib; b.splat = ctx->d_bd_splat; b.items = ctx->d_bd_items;
    b.tile_slot = ctx->d_bd_tile_slot; b.ctr = ctx->d_bd_ctr; b.cap = ctx->bd_cap; memcpy(b.view, ctx->view, 64);
    LaunchCfg cfg; memset(&cfg, 0, sizeof(cfg)); if ((rc = launch_cfg(ctx, a, cfg))) return rc;
    int bp_ = 1, bc_ = 1, bq_ = 1;
    TR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bp_, k_bdpt_paths, WF_THREADS, 0));
    TR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bc_, k_bdpt_connect, WF_THREADS, 0));
    TR_MODE_SWITCH(cfg.mode, { if (!rc) rc = kernel_blocks(ctx, k_shadow<M, true>, cfg.smem, bq_); });
    if (rc) return rc;
    if (bp_ < 1) bp_ = 1; if (bc_ < 1) bc_ = 1; if (bq_ < 1) bq_ = 1;
    cudaStream_t s = ctx->stream;
    uint64_t launches = 0, rays_c = 0, rays_s = 0, vis[4] = {0, 0, 0, 0};
    const bool timing = ctx->opt_stage_timing != 0;
    if (timing && ctx->stage_ev.empty()) {
        ctx->stage_ev.resize(4 * TR_MAX_DEPTH_CAP);
        for (auto& e : ctx->stage_ev) TR_CUDA(ctx, cudaEventCreate(&e));
    }
    cudaEvent_t* ev = timing ? ctx->stage_ev.data() : nullptr;
    float ms_paths = 0.0f, ms_connect = 0.0f, ms_other = 0.0f, ms_ktrace = 0.0f, ms_kshadow = 0.0f;
    TR_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    // auto mode: the batches differ by at most one frame (32 frames at 6 per batch: 6 6 5 5 5 5 instead of 6 6 6 6 6 2)
    const int nbatch = (n_frames + F - 1) / F, fbase = n_frames / nbatch, fextra = n_frames % nbatch;
    for (int f0 = 0, bi = 0, nf = 0; f0 < n_frames; f0 += nf, ++bi) {
        nf = ctx->opt_batch_frames > 0 ? ((n_frames - f0 < F) ? n_frames - f0 : F) : fbase + (bi < fextra ? 1 : 0);
        BatchParams bp; bp.frame_begin = frame_begin + f0; bp.n_frames = nf; bp.seed = seed; bp.max_depth = BD_MAX_DEPTH; bp.pad = 0;
        TR_CUDA(ctx, cudaMemcpyAsync(ctx->d_batch_params, &bp, sizeof(bp), cudaMemcpyHostToDevice, s));
        TR_CUDA(ctx, cudaMemsetAsync(ctx->d_ctr, 0, sizeof(TrCounters), s));
        TR_CUDA(ctx, cudaMemsetAsync(ctx->d_bd_ctr, 0, 4 * sizeof(unsigned long long), s));
        TR_CUDA(ctx, cudaMemsetAsync(ctx->d_bd_splat, 0, (size_t)nf * ctx->W * ctx->H * 3 * sizeof(float), s));
        if (ev) cudaEventRecord(ev[0], s);
        if (a.npix > 0 && wave) {
            // stage events (timing mode): ev[8 + 2d], ev[9 + 2d] bracket k_trace(d); ev[5], ev[6] bracket the shadow-query kernel
            k_bdpt_generate<<<cfg.grid_simple, WF_THREADS, 0, s>>>(a, b); ++launches;
            for (int d = 0; d < BD_EYE_MAX - 1; ++d) {
                if (ev) cudaEventRecord(ev[8 + 2 * d], s);
                TR_MODE_SWITCH(cfg.mode, (k_trace<M><<<cfg.grid_trace, WF_THREADS, cfg.smem, s>>>(a, d)));
                if (ev) cudaEventRecord(ev[9 + 2 * d], s);
                k_bdpt_vertex<<<cfg.grid_simple, WF_THREADS, 0, s>>>(a, b, d);
                launches += 2;
            }
            if (ev) cudaEventRecord(ev[1], s);
            k_bdpt_items<<<cfg.grid_simple, WF_THREADS, 0, s>>>(a, b);
            if (ev) cudaEventRecord(ev[2], s);
            k_bdpt_connect_gen<<<cfg.grid_simple, WF_THREADS, 0, s>>>(a, b);
            if (ev) cudaEventRecord(ev[5], s);
            TR_MODE_SWITCH(cfg.mode, (k_shadow<M, true><<<ctx->num_sms * bq_, WF_THREADS, cfg.smem, s>>>(a, 0)));
            if (ev) cudaEventRecord(ev[6], s);
            k_bdpt_connect_eval<<<cfg.grid_simple, WF_THREADS, 0, s>>>(a, b);
            k_bdpt_mis<<<cfg.grid_simple, WF_THREADS, 0, s>>>(a, b);
            if (ev) cudaEventRecord(ev[3], s);
            launches += 5;
        } else if (a.npix > 0) {
            k_bdpt_paths<<<ctx->num_sms * bp_, WF_THREADS, 0, s>>>(a, b);
            if (ev) cudaEventRecord(ev[1], s);
            k_bdpt_items<<<cfg.grid_simple, WF_THREADS, 0, s>>>(a, b);
            if (ev) cudaEventRecord(ev[2], s);
            k_bdpt_connect<<<ctx->num_sms * bc_, WF_THREADS, 0, s>>>(a, b);
            if (ev) cudaEventRecord(ev[3], s);
            launches += 3;
        }
        k_bdpt_film<<<cfg.grid_simple, WF_THREADS, 0, s>>>(a, b); ++launches;
        if (ev) cudaEventRecord(ev[4], s);
        TR_CHECK_LAUNCH(ctx);
        // ray counters of this batch.  Asynchronous like PT_RGB: snapshot into the pinned ring (slot 0: the wavefront counters,
        // slot 1: the four BDPT counters) and go on; tr_stats_get folds the snapshots.  Stage timing stays synchronous.
        if (!timing && bi < TR_RING_BATCHES) {
            TR_CUDA(ctx, cudaMemcpyAsync(ctx->h_ring + (size_t)bi * TR_MAX_CHAINS, ctx->d_ctr, sizeof(TrCounters), cudaMemcpyDeviceToHost, s));
            TR_CUDA(ctx, cudaMemcpyAsync(ctx->h_ring + (size_t)bi * TR_MAX_CHAINS + 1, ctx->d_bd_ctr, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
            ctx->ring_batches = bi + 1;
            if (bi + 1 == TR_RING_BATCHES && f0 + nf < n_frames) TR_CUDA(ctx, cudaStreamSynchronize(s));     // ring full: later batches take the synchronous path
            continue;
        }
        unsigned long long hc[4];
        TR_CUDA(ctx, cudaMemcpyAsync(hc, ctx->d_bd_ctr, sizeof(hc), cudaMemcpyDeviceToHost, s));
        TR_CUDA(ctx, cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(TrCounters), cudaMemcpyDeviceToHost, s));
        TR_CUDA(ctx, cudaStreamSynchronize(s));
        if (wave) { for (int d = 0; d < BD_EYE_MAX - 1; ++d) rays_c += (uint64_t)ctx->h_ctr[0].nq[d]; rays_s += (uint64_t)ctx->h_ctr[0].nshadow[0]; }
        else { rays_c += hc[0]; rays_s += hc[1]; }
        for (int k = 0; k < 4; ++k) vis[k] += ctx->h_ctr[0].visits[k];
        if (ev && a.npix > 0) {
            float t0 = 0, t1 = 0, t2 = 0, t3 = 0;
            cudaEventElapsedTime(&t0, ev[0], ev[1]); cudaEventElapsedTime(&t1, ev[1], ev[2]); cudaEventElapsedTime(&t2, ev[2], ev[3]); cudaEventElapsedTime(&t3, ev[3], ev[4]);
            ms_paths += t0; ms_connect += t2; ms_other += t1 + t3;
            if (wave) {
                for (int d = 0; d < BD_EYE_MAX - 1; ++d) { float t = 0; cudaEventElapsedTime(&t, ev[8 + 2 * d], ev[9 + 2 * d]); ms_ktrace += t; }
                float t = 0; cudaEventElapsedTime(&t, ev[5], ev[6]); ms_kshadow += t;
            }
        }
    }
    TR_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    ctx->stats.rays_closest = rays_c; ctx->stats.rays_shadow = rays_s; ctx->stats.shade_terminal = 0;
    ctx->stats.node_visits = vis[0]; ctx->stats.leaf_tests = vis[1]; ctx->stats.node_visits_shadow = vis[2]; ctx->stats.leaf_tests_shadow = vis[3];
    ctx->stats.kernel_launches = launches; ctx->stats.ms_total = 0.0f; ctx->stats.frames = n_frames; ctx->stats.paths_in_flight = (int)cap; ctx->stats.chains = 1;
    ctx->ring_mode = wave ? 1 : 2; ctx->ring_depth = BD_EYE_MAX - 1; ctx->stats_pending = true;            // tr_stats_resolve waits for ev1 and folds the ring
    if (timing && (rc = tr_stats_resolve(ctx))) return rc;
    // stage timing: ms_trace = the sub-path stage, ms_shadow = the connection stage, ms_shade = items + film; the traversal
    // kernels alone (wavefront pipeline) are reported through tr_bdpt_kernel_ms
    ctx->stats.ms_trace = ms_paths; ctx->stats.ms_shade = ms_other; ctx->stats.ms_shadow = ms_connect;
    ctx->bd_ms_ktrace = ms_ktrace; ctx->bd_ms_kshadow = ms_kshadow;
    return TR_OK;
}

extern "C" int tr_render_bdpt_rgb(tr_ctx* ctx, int frame_begin, int n_frames, uint64_t seed) { return render_bdpt(ctx, frame_begin, n_frames, seed); }
extern "C" int tr_bdpt_kernel_ms(tr_ctx* ctx, float* ms_trace_kernels, float* ms_shadow_kernel) {
    if (!ctx) return TR_ERR_INVALID;
    if (ms_trace_kernels) *ms_trace_kernels = ctx->bd_ms_ktrace;
    if (ms_shadow_kernel) *ms_shadow_kernel = ctx->bd_ms_kshadow;
    return TR_OK;
}

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
            const bool valid = v < BD_EYE_MAX ? v < d[0] : (v - BD_EYE_MAX) < d[1];
            if (!valid) { for (int c = 0; c < 20; ++c) o[c] = 0.0f; continue; }
            int fl; memcpy(&fl, &w[2].w, 4); int prim; memcpy(&prim, &w[3].w, 4);
            o[0] = w[0].x; o[1] = w[0].y; o[2] = w[0].z; o[3] = w[1].x; o[4] = w[1].y; o[5] = w[1].z; o[6] = w[2].x; o[7] = w[2].y; o[8] = w[2].z;
            o[9] = w[3].x; o[10] = w[3].y; o[11] = w[3].z; o[12] = w[4].x; o[13] = w[4].y; o[14] = w[4].z; o[15] = w[0].w; o[16] = w[1].w;
            o[17] = (float)((fl & 15) + 16 * ((fl >> 4) & 15)); o[18] = (float)prim; o[19] = (float)(fl >> 8);
        }
        float* co = contrib + (size_t)k * 7 * 7 * 4;
        for (int c = 0; c < 7 * 7 * 4; ++c) co[c] = 0.0f;
        for (int e = 2; e <= d[0]; ++e) for (int l = 0; l <= d[1] && l + e - 2 <= BD_MAX_DEPTH; ++l) {
            if (!bd_strategy_live(e, l, d[0], d[1])) continue;          // never queued: identically zero
            float4 w; int row = (e - 2) * 6 - ((e - 2) * (e - 3)) / 2 + l;
            TR_CUDA(ctx, cudaMemcpy(&w, ctx->d_bd_contrib + (size_t)row * cap + s, 16, cudaMemcpyDeviceToHost));
            float* o = co + ((e - 1) * 7 + l) * 4; o[0] = w.x; o[1] = w.y; o[2] = w.z; o[3] = (float)(x * 65536 + y);
        }
    }
    return TR_OK;
}
