// trace.cuh — stackless (threaded pre-order) LBVH traversal + Moller-Trumbore, device side.
//
// Replaces Scene.closet_hit / closet_hit_shadow (Scene.py:671-744), which walk the same pre-order
// node array with an explicit per-pixel stack in global memory, unordered and unpruned.  Here the
// walk follows escape links (no stack; one link per node for small trees in shared memory, eight
// direction-octant links per node for a front-to-back walk of large trees), prunes sub-trees whose slab
// entry lies beyond the best hit (with a relative guard band so exact/near ties are still tested) and keeps
// the reference's result: the reference pops the right child first and accepts strict t < best, so among
// equal-t hits the leaf with the LARGEST sorted position wins; closer() applies exactly that rule.
#pragma once
#include "ctx.h"
#include "common.cuh"

#define TR_PRUNE_GUARD 1.0001f

struct RayPre {
    int oct;               // direction-sign octant: bit a set <=> d[a] < 0 (selects the front-to-back threading)
    V3 o, d;
    float ix, iy, iz;      // 1/d per axis (UtilsFunc.py:510), unused when the axis is "parallel"
    bool px, py, pz;       // |d| < 1e-6 (UtilsFunc.py:506)
};

__device__ __forceinline__ RayPre make_ray(V3 o, V3 d) {
    RayPre r; r.o = o; r.d = d;
    r.oct = (d.x < 0.0f ? 1 : 0) | (d.y < 0.0f ? 2 : 0) | (d.z < 0.0f ? 4 : 0);
    r.px = fabsf(d.x) < 0.000001f; r.py = fabsf(d.y) < 0.000001f; r.pz = fabsf(d.z) < 0.000001f;
    r.ix = 1.0f / d.x; r.iy = 1.0f / d.y; r.iz = 1.0f / d.z;
    return r;
}

// UtilsFunc.py:494-523, plus tmin returned for pruning
__device__ __forceinline__ bool slabs(const RayPre& r, float4 lo, float4 hi, float& tmin_out) {
    bool ret = true; float tmin = 0.0f, tmax = TR_INF;
    if (r.px) { if (r.o.x < lo.x || r.o.x > hi.x) ret = false; }
    else { float t1 = (lo.x - r.o.x) * r.ix, t2 = (hi.x - r.o.x) * r.ix; tmin = fmaxf(tmin, fminf(t1, t2)); tmax = fminf(tmax, fmaxf(t1, t2)); }
    if (r.py) { if (r.o.y < lo.y || r.o.y > hi.y) ret = false; }
    else { float t1 = (lo.y - r.o.y) * r.iy, t2 = (hi.y - r.o.y) * r.iy; tmin = fmaxf(tmin, fminf(t1, t2)); tmax = fminf(tmax, fmaxf(t1, t2)); }
    if (r.pz) { if (r.o.z < lo.z || r.o.z > hi.z) ret = false; }
    else { float t1 = (lo.z - r.o.z) * r.iz, t2 = (hi.z - r.o.z) * r.iz; tmin = fmaxf(tmin, fminf(t1, t2)); tmax = fminf(tmax, fmaxf(t1, t2)); }
    tmin_out = tmin;
    return ret && !(tmin > tmax);
}

// Scene.py:603-638 with E1/E2 precomputed at build time (same single-rounding subtractions)
__device__ __forceinline__ void intersect_tri(V3 o, V3 d, V3 v0, V3 E1, V3 E2, float& t, float& u, float& v) {
    t = TR_INF; u = 0.0f; v = 0.0f;
    V3 P = cross3(d, E2);
    float det = dot3(E1, P);
    V3 T;
    if (det > 0.0f) T = o - v0; else { T = v0 - o; det = -det; }
    if (det > 0.0f) {
        u = dot3(T, P);
        if (u >= 0.0f && u <= det) {
            V3 Q = cross3(T, E1);
            v = dot3(d, Q);
            if (v >= 0.0f && u + v <= det) {
                t = dot3(E2, Q);
                float inv = 1.0f / det;
                t *= inv; u *= inv; v *= inv;
            }
        }
    }
}

// Scene.py:565-596 / 653-665: analytic sphere, nearest root only; returns the reference's scalar c
__device__ __forceinline__ bool intersect_sphere(V3 o, V3 d, V3 ce, float r, float& t, float& c_out) {
    V3 oc = ce - o; float oc2 = dot3(oc, oc), op = dot3(d, oc);
    float cp = sqrtf(oc2 - op * op);
    t = TR_INF; c_out = 0.0f;
    if (cp < r) {
        float a = dot3(d, d), b = -2.0f * op, c = oc2 - r * r;
        t = (-b - sqrtf(b * b - 4.0f * a * c)) / 2.0f / a;
        c_out = c;
        return true;
    }
    return false;
}

__device__ __forceinline__ float intersect_leaf(const RayPre& r, float4 la, float4 lb, float4 lc, float& u, float& v) {
    int kind = __float_as_int(lb.w);
    float t;
    if (kind == 0) {
        intersect_tri(r.o, r.d, f4xyz(la), f4xyz(lb), f4xyz(lc), t, u, v);
    } else if (kind == 1) {
        float c; u = 0.0f; v = 0.0f;
        intersect_sphere(r.o, r.d, f4xyz(la), lb.x, t, c);
    } else { t = TR_INF; u = v = 0.0f; }
    return t;
}

// hint the next node's line into L1 while the slab test of the current node is still in flight
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

struct HitRec { float t, u, v; int prim; int mat; int leaf; };

// next node of the walk. ORDERED (trees read from global memory): near child first on a box hit, octant escape link
// (TrNodeX::next) otherwise = stackless front-to-back order, which halves the node visits on the 130 k-triangle scene.
// Unordered (small trees staged in shared memory, where every box overlaps every ray and order buys nothing):
// left child first, single escape link stored in the node.
template <bool ORDERED>
__device__ __forceinline__ int next_node(int idx, int link, bool hit, int esc, int oct) {
    int first = idx + 1;
    if (ORDERED) first = ((oct >> (link >> 29)) & 1) ? (link & 0x1fffffff) : idx + 1;
    return (hit && link >= 0) ? first : esc;
}
// the reference pops the right child first and accepts strict t < best: among equal-t hits the LARGEST sorted leaf wins
__device__ __forceinline__ bool closer(float t, int k, float best_t, int best_k) {
    return t > 0.0f && t < TR_INF && (t < best_t || (t == best_t && k > best_k));
}

#ifdef TR_COUNTERS
#define TR_COUNT_NODE() (++cnt_nodes)
#define TR_COUNT_LEAF() (++cnt_leaves)
#else
#define TR_COUNT_NODE()
#define TR_COUNT_LEAF()
#endif

// Fast slab test for rays without a "parallel" axis (the common case): same arithmetic, no flags.
__device__ __forceinline__ bool slabs_fast(const RayPre& r, float4 lo, float4 hi, float& tmin_out) {
    float t1 = (lo.x - r.o.x) * r.ix, t2 = (hi.x - r.o.x) * r.ix;
    float tmin = fmaxf(0.0f, fminf(t1, t2)), tmax = fminf(TR_INF, fmaxf(t1, t2));
    t1 = (lo.y - r.o.y) * r.iy; t2 = (hi.y - r.o.y) * r.iy;
    tmin = fmaxf(tmin, fminf(t1, t2)); tmax = fminf(tmax, fmaxf(t1, t2));
    t1 = (lo.z - r.o.z) * r.iz; t2 = (hi.z - r.o.z) * r.iz;
    tmin = fmaxf(tmin, fminf(t1, t2)); tmax = fminf(tmax, fmaxf(t1, t2));
    tmin_out = tmin;
    return !(tmin > tmax);
}

// Warp-cooperative traversal schedule.
// Every node visit is the same instruction stream for internal nodes and leaves: one slab test on the
// node's box (leaf nodes carry their triangle's box, grown by a guard band at build time so that a
// Moller-Trumbore hit can never be culled by it; the reference tests every leaf under a hit parent).  A lane
// that reaches a leaf whose box is hit parks the leaf in `pend`; the Moller-Trumbore block then runs for the
// parked lanes once per round of node steps.  TR_LEAF_BATCH > 1 would delay it until that many lanes are
// parked; measured on B200 (sweep 1..24, one and two parked slots per lane): 1 is fastest on both scenes.
// Hits are compared with closer(): the reference's tie rule made explicit, so the visiting order is free.
#ifndef TR_LEAF_BATCH
#define TR_LEAF_BATCH 1
#endif

// Closest hit (Scene.py:702-744 semantics).  nodes/leaves may point to shared or global memory.
// Warp-synchronous: all 32 lanes must call it together; lanes without a ray pass active = false.
template <bool SMEM>
__device__ __forceinline__ HitRec trace_closest(const TrNode* __restrict__ nodes, const TrLeaf* __restrict__ leaves,
                                                 const TrNodeX* __restrict__ nodesx, int nnodes,
                                                 const RayPre& r, bool active, unsigned long long* cnt) {
    HitRec h; h.t = TR_INF; h.u = 0.0f; h.v = 0.0f; h.prim = -1; h.mat = 0; h.leaf = -1;
#ifdef TR_COUNTERS
    unsigned cnt_nodes = 0, cnt_leaves = 0;
#endif
    const bool anypar = r.px || r.py || r.pz;
    int idx = active ? 0 : nnodes, pend = -1;
    while (true) {
        if (pend < 0 && idx < nnodes) {
            float4 lo, hi; int esc;
            if (SMEM) { lo = nodes[idx].lo; hi = nodes[idx].hi; esc = __float_as_int(lo.w); }
            else { lo = nodesx[idx].lo; hi = nodesx[idx].hi; esc = nodesx[idx].next[r.oct]; }
            int link = __float_as_int(hi.w);
            float tmin;
            bool hit = (anypar ? slabs(r, lo, hi, tmin) : slabs_fast(r, lo, hi, tmin)) && !(tmin > h.t * TR_PRUNE_GUARD);
            if (link < 0) { if (hit) pend = -link - 1; } else { TR_COUNT_NODE(); }
            idx = next_node<!SMEM>(idx, link, hit, esc, r.oct);
        }
        const unsigned parked = __ballot_sync(0xffffffffu, pend >= 0);
        const unsigned walking = __ballot_sync(0xffffffffu, pend < 0 && idx < nnodes);
        if (parked == 0u && walking == 0u) break;
        if (__popc(parked) >= TR_LEAF_BATCH || walking == 0u) {
            if (pend >= 0) {
                TR_COUNT_LEAF();
                const TrLeaf* lf = leaves + pend;
                float4 la = lf->a, lb = lf->b, lc = lf->c;
                float u, v, t = intersect_leaf(r, la, lb, lc, u, v);
                if (closer(t, pend, h.t, h.leaf)) { h.t = t; h.u = u; h.v = v; h.prim = __float_as_int(la.w); h.mat = __float_as_int(lc.w); h.leaf = pend; }
                pend = -1;
            }
        }
    }
#ifdef TR_COUNTERS
    if (cnt && active) { atomicAdd(cnt, (unsigned long long)cnt_nodes); atomicAdd(cnt + 1, (unsigned long long)cnt_leaves); }
#endif
    return h;
}

// Shadow query.  The reference (integrator/PT_RGB.py:104-105, Scene.py:671-699) finds the nearest
// hit of the light->surface ray and tests `shadow_prim == prim_id`.  Equivalent early-exit form:
// intersect the target primitive first (t_t), then walk the tree (bounded by t_t) looking for ANY other
// primitive that would have won the reference's comparison (t < t_t, or t == t_t at a later leaf
// position); the target only counts if the walk reaches its leaf (every ancestor passes the slab test).
// Warp-synchronous like trace_closest.
template <bool SMEM>
__device__ __forceinline__ bool trace_shadow_visible(const TrNode* __restrict__ nodes, const TrLeaf* __restrict__ leaves,
                                                     const TrNodeX* __restrict__ nodesx, int nnodes,
                                                     const RayPre& r, bool active, int target_leaf, unsigned long long* cnt,
                                                     float* tt_out = nullptr) {
#ifdef TR_COUNTERS
    unsigned cnt_nodes = 0, cnt_leaves = 0;
#endif
    bool visible = false, found = false;
    float tt = TR_INF;
    if (active) {
        const TrLeaf* lf = leaves + target_leaf;
        float u, v; tt = intersect_leaf(r, lf->a, lf->b, lf->c, u, v);
        TR_COUNT_LEAF();
        visible = (tt > 0.0f && tt < TR_INF);
    }
    const bool anypar = r.px || r.py || r.pz;
    int idx = visible ? 0 : nnodes, pend = -1;
    while (true) {
        if (pend < 0 && idx < nnodes) {
            float4 lo, hi; int esc;
            if (SMEM) { lo = nodes[idx].lo; hi = nodes[idx].hi; esc = __float_as_int(lo.w); }
            else { lo = nodesx[idx].lo; hi = nodesx[idx].hi; esc = nodesx[idx].next[r.oct]; }
            int link = __float_as_int(hi.w);
            float tmin;
            bool hit = (anypar ? slabs(r, lo, hi, tmin) : slabs_fast(r, lo, hi, tmin)) && !(tmin > tt * TR_PRUNE_GUARD);
            if (link < 0) {
                int k = -link - 1;
                if (k == target_leaf) found = true; else if (hit) pend = k;
            } else { TR_COUNT_NODE(); }
            idx = next_node<!SMEM>(idx, link, hit, esc, r.oct);
        }
        const unsigned parked = __ballot_sync(0xffffffffu, pend >= 0);
        const unsigned walking = __ballot_sync(0xffffffffu, pend < 0 && idx < nnodes);
        if (parked == 0u && walking == 0u) break;
        if (__popc(parked) >= TR_LEAF_BATCH || walking == 0u) {
            if (pend >= 0) {
                TR_COUNT_LEAF();
                const TrLeaf* l2 = leaves + pend;
                float u, v, t = intersect_leaf(r, l2->a, l2->b, l2->c, u, v);
                if (t > 0.0f && t < TR_INF && (t < tt || (t == tt && pend > target_leaf))) { visible = false; idx = nnodes; }
                pend = -1;
            }
        }
    }
#ifdef TR_COUNTERS
    if (cnt && active) { atomicAdd(cnt, (unsigned long long)cnt_nodes); atomicAdd(cnt + 1, (unsigned long long)cnt_leaves); }
#endif
    if (tt_out) *tt_out = tt;          // distance to the target primitive = the reference's hit_t when the target is the nearest hit
    return visible && found;
}

// ---- TMA bulk staging of the whole BVH into shared memory (small scenes, e.g. the Cornell box)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// One thread issues cp.async.bulk (global -> shared, mbarrier completion); everybody waits.
// bytes must be a multiple of 16 and both pointers 16-byte aligned.
__device__ __forceinline__ void tma_stage_to_smem(void* smem_dst, const void* gsrc, unsigned bytes, void* smem_dst2,
                                                  const void* gsrc2, unsigned bytes2, unsigned long long* bar) {
    const unsigned bar_a = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes + bytes2) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(bar_a) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem_dst2)), "l"(gsrc2), "r"(bytes2), "r"(bar_a) : "memory");
    }
    // wait for phase 0
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar_a) : "memory");
}
