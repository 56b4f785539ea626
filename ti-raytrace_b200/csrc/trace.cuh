// trace.cuh — LBVH traversal + Moller-Trumbore, device side.
//
// Replaces Scene.closet_hit / closet_hit_shadow (Scene.py:671-744), which walk the pre-order node array with an explicit
// per-pixel stack in GLOBAL memory (integrator/PT_RGB.py:37), unordered and unpruned, one 36-byte node per dependent load.
// Here:
//   * traversal nodes (TrNode2, 64 B) hold the boxes of BOTH children of an internal node, so one dependent load serves two
//     slab tests and the chain of dependent loads per ray is halved;
//   * the per-lane stack lives in SHARED memory ([entry][thread]: conflict free), sized at launch to what the tree can need;
//     a leaf child is always taken before an internal one, so chains of duplicate Morton codes (the reference's duplicate rule
//     builds them) need one entry, and the build rejects trees that need more than TR_STACK_MAX entries with the reference's
//     "overflow, need larger stack" (Scene.py:741-742);
//   * of two hit children the one whose slab entry is nearer is visited first, and sub-trees whose slab entry lies beyond the
//     best hit (relative guard band, so exact / near ties are still tested) are pruned;
//   * small trees are staged whole into shared memory by TMA (cp.async.bulk) as a bank-conflict-free image: every 16-byte
//     word is replicated 8 times in a 128-byte row and lane l reads column l & 7, so the 8 lanes of a quarter-warp (the unit
//     a 128-bit shared load is served in) always touch 8 different bank quads whatever nodes they are at; medium trees are
//     staged once (plain); large trees are read from global memory (L1 / L2), optionally with their breadth-first top staged
//     (option "top_nodes": built and measured, slower than L1 on B200, off by default).
// The result is the reference's: the slab arithmetic (UtilsFunc.py:494-523) and Moller-Trumbore (Scene.py:603-638) are
// restated operation by operation, every internal node's box is tested before its children are entered, and among equal-t
// hits the reference's winner is kept explicitly (it pops the right child first and accepts strict t < best, so the LARGEST
// sorted leaf position wins; closer() applies exactly that rule), which makes the visiting order free.
#pragma once
#include "ctx.h"
#include "common.cuh"

#define TR_PRUNE_GUARD 1.0001f
#define TR_DONE 0x7fffffff          // link value of a lane whose walk is finished (also what an empty stack pops)

struct RayPre {
    V3 o, d;
    float ix, iy, iz;      // 1/d per axis (UtilsFunc.py:510), unused when the axis is "parallel"
    bool px, py, pz;       // |d| < 1e-6 (UtilsFunc.py:506)
};

__device__ __forceinline__ RayPre make_ray(V3 o, V3 d) {
    RayPre r; r.o = o; r.d = d;
    r.px = fabsf(d.x) < 0.000001f; r.py = fabsf(d.y) < 0.000001f; r.pz = fabsf(d.z) < 0.000001f;
    r.ix = 1.0f / d.x; r.iy = 1.0f / d.y; r.iz = 1.0f / d.z;
    return r;
}

// UtilsFunc.py:494-523, plus tmin returned for ordering / pruning
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
__device__ __forceinline__ bool slab_any(const RayPre& r, bool anypar, float4 lo, float4 hi, float& tmin) {
    return anypar ? slabs(r, lo, hi, tmin) : slabs_fast(r, lo, hi, tmin);
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

struct HitRec { float t, u, v; int prim; int mat; int leaf; };
__device__ __forceinline__ void hit_reset(HitRec& h) { h.t = TR_INF; h.u = h.v = 0.0f; h.prim = -1; h.mat = 0; h.leaf = -1; }

// the reference pops the right child first and accepts strict t < best: among equal-t hits the LARGEST sorted leaf wins
__device__ __forceinline__ bool closer(float t, int k, float best_t, int best_k) {
    return t > 0.0f && t < TR_INF && (t < best_t || (t == best_t && k > best_k));
}

// ---- tree views ----------------------------------------------------------------------------------------------------------
// How a kernel reads the tree (template parameter MODE of the traversal kernels):
enum { TM_REP = 0,      // whole tree in shared memory as the 8-way replicated, bank-conflict-free image
       TM_SMEM = 1,     // whole tree in shared memory, plain
       TM_GLOBAL = 2,   // tree in global memory (L1 / L2)
       TM_GTOP = 3 };   // global memory, the first `top` nodes (breadth-first top of the tree) staged in shared memory

struct TreeView {
    const float4* snodes;    // shared-memory nodes (REP image / plain / staged top)
    const float4* sleaves;   // shared-memory leaves (REP / SMEM)
    const float4* gnodes;    // global TrNode2 array (4 float4 per node)
    const float4* gleaves;   // global TrLeaf array (3 float4 per leaf)
    int top;                 // TM_GTOP: nodes [0, top) are in snodes
};

template <int MODE>
__device__ __forceinline__ void node_fetch(const TreeView& tv, int idx, int lane8, float4& A, float4& B, float4& C, float4& D) {
    if (MODE == TM_REP) { const float4* p = tv.snodes + idx * 32 + lane8; A = p[0]; B = p[8]; C = p[16]; D = p[24]; }
    else if (MODE == TM_SMEM) { const float4* p = tv.snodes + idx * 4; A = p[0]; B = p[1]; C = p[2]; D = p[3]; }
    else if (MODE == TM_GLOBAL) { const float4* p = tv.gnodes + (size_t)idx * 4; A = __ldg(p); B = __ldg(p + 1); C = __ldg(p + 2); D = __ldg(p + 3); }
    else { const float4* p = (idx < tv.top ? tv.snodes : tv.gnodes) + (size_t)idx * 4; A = p[0]; B = p[1]; C = p[2]; D = p[3]; }   // generic loads
}
template <int MODE>
__device__ __forceinline__ void leaf_fetch(const TreeView& tv, int k, int lane8, float4& la, float4& lb, float4& lc) {
    if (MODE == TM_REP) { const float4* p = tv.sleaves + k * 24 + lane8; la = p[0]; lb = p[8]; lc = p[16]; }
    else if (MODE == TM_SMEM) { const float4* p = tv.sleaves + k * 3; la = p[0]; lb = p[1]; lc = p[2]; }
    else { const float4* p = tv.gleaves + (size_t)k * 3; la = __ldg(p); lb = __ldg(p + 1); lc = __ldg(p + 2); }
}

// ---- per-lane traversal stack -------------------------------------------------------------------------------------------
// Render kernels: all entries in shared memory at s[entry * WF_THREADS] (s points at this thread's column: consecutive lanes hit
// consecutive banks); the kernel is launched with as many entries per lane as the tree needs (computed by the build, at most
// TR_STACK_MAX, else the build fails with TR_ERR_STACK).  Measured on B200: a local-memory overflow path behind the shared
// entries cost 20 % of the trace kernel although it was never taken (C2 9.96 -> 7.84 ms, C3 5.55 -> 4.60 ms without it).
// Simple one-lane-per-ray walks (Debug integrator, test hooks, lock-step BDPT cross-check): a local-memory array.
#define WF_THREADS 256
struct SmemStack {
    int* s; int sp;
    __device__ __forceinline__ void init(int* column) { s = column; sp = 0; }
    __device__ __forceinline__ void reset() { sp = 0; }
    __device__ __forceinline__ void push(int v) { s[sp * WF_THREADS] = v; ++sp; }
    __device__ __forceinline__ int pop() { if (sp == 0) return TR_DONE; --sp; return s[sp * WF_THREADS]; }
};
struct LocalStack {
    int a[TR_STACK_MAX]; int sp;
    __device__ __forceinline__ void reset() { sp = 0; }
    __device__ __forceinline__ void push(int v) { if (sp < TR_STACK_MAX) a[sp] = v; ++sp; }
    __device__ __forceinline__ int pop() { if (sp == 0) return TR_DONE; --sp; return sp < TR_STACK_MAX ? a[sp] : TR_DONE; }
};

#ifdef TR_COUNTERS
#define TR_COUNT(x) (++(x))
#else
#define TR_COUNT(x)
#endif

// link of the first node of a ray's walk: the root box is tested like the reference tests it at its first pop; a tree of
// one primitive is a single leaf, which the reference intersects without any box test
struct TreeRoot { float4 lo, hi; int link; };
__device__ __forceinline__ int root_enter(const TreeRoot& rt, const RayPre& r, bool anypar) {
    if (rt.link < 0) return rt.link;
    float tmin;
    return slab_any(r, anypar, rt.lo, rt.hi, tmin) ? rt.link : TR_DONE;
}

// One node step of a lane: fetch internal node `cur` (both child boxes), test them, choose the next link.
//   link >= 0: internal node index; link < 0: leaf with sorted position -link-1; TR_DONE: walk finished.
//   bound: best hit so far (closest hit) or the distance of the target (shadow query): children whose slab entry lies beyond
//   bound x guard are pruned.  SHADOW: the target's own leaf (tlink) is never entered; reaching it sets `found`.
// Order: a leaf child before an internal one (keeps the stack at one entry along chains), else the nearer slab entry first.
template <int MODE, bool SHADOW, typename Stack>
__device__ __forceinline__ int node_step(const TreeView& tv, const RayPre& r, bool anypar, float bound, int tlink, bool& found,
                                         Stack& st, int cur, int lane8) {
    float4 A, B, C, D;
    node_fetch<MODE>(tv, cur, lane8, A, B, C, D);
    const int l0 = __float_as_int(A.w), l1 = __float_as_int(B.w);
    float t0, t1;
    bool h0 = slab_any(r, anypar, A, B, t0) && !(t0 > bound * TR_PRUNE_GUARD);
    bool h1 = slab_any(r, anypar, C, D, t1) && !(t1 > bound * TR_PRUNE_GUARD);
    if (SHADOW) {
        if (l0 == tlink) { found = true; h0 = false; }
        if (l1 == tlink) { found = true; h1 = false; }
    }
    if (h0 && h1) {
        const bool first1 = (l1 < 0 && l0 >= 0) || (((l0 < 0) == (l1 < 0)) && t1 < t0);
        st.push(first1 ? l0 : l1);
        return first1 ? l1 : l0;
    }
    return h0 ? l0 : (h1 ? l1 : st.pop());
}

// ---- simple (non-persistent) walks: Debug integrator, test hooks, lock-step BDPT cross-check -----------------------------
// One lane = one ray from start to end; the stack is a local-memory array.  Same node_step / closer as the persistent
// kernels, global-memory tree.
__device__ __forceinline__ HitRec trace_closest(const TreeView& tv, const TreeRoot& rt, const RayPre& r, bool active, unsigned long long* cnt) {
    HitRec h; hit_reset(h);
#ifdef TR_COUNTERS
    unsigned cnt_nodes = 0, cnt_leaves = 0;
#endif
    const bool anypar = r.px || r.py || r.pz;
    LocalStack st; st.reset();
    bool found = false;
    int cur = active ? root_enter(rt, r, anypar) : TR_DONE;
    while (cur != TR_DONE) {
        if (cur >= 0) { TR_COUNT(cnt_nodes); cur = node_step<TM_GLOBAL, false>(tv, r, anypar, h.t, 0, found, st, cur, 0); }
        else {
            TR_COUNT(cnt_leaves);
            const int k = -cur - 1;
            float4 la, lb, lc; leaf_fetch<TM_GLOBAL>(tv, k, 0, la, lb, lc);
            float u, v, t = intersect_leaf(r, la, lb, lc, u, v);
            if (closer(t, k, h.t, h.leaf)) { h.t = t; h.u = u; h.v = v; h.prim = __float_as_int(la.w); h.mat = __float_as_int(lc.w); h.leaf = k; }
            cur = st.pop();
        }
    }
#ifdef TR_COUNTERS
    if (cnt && active) { atomicAdd(cnt, 2ull * cnt_nodes); atomicAdd(cnt + 1, (unsigned long long)cnt_leaves); }     // 2 boxes (2 x 32 B) per node visit
#endif
    return h;
}

// Shadow query.  The reference (integrator/PT_RGB.py:104-105, Scene.py:671-699) finds the nearest hit of the light->surface
// ray and tests `shadow_prim == prim_id`.  Equivalent early-exit form: intersect the target primitive first (t_t), then walk
// the tree (bounded by t_t) looking for ANY other primitive that would have won the reference's comparison (t < t_t, or
// t == t_t at a later leaf position); the target only counts if the walk reaches its leaf (every ancestor passes the slab test).
__device__ __forceinline__ bool blocks_target(float t, int k, float tt, int tleaf) {
    return t > 0.0f && t < TR_INF && (t < tt || (t == tt && k > tleaf));
}
// start of a shadow walk: distance of the target, visibility so far, first link
__device__ __forceinline__ int shadow_enter(const TreeView& tv, const TreeRoot& rt, const RayPre& r, bool anypar, int tleaf,
                                            float4 la, float4 lb, float4 lc, float& tt, bool& visible, bool& found) {
    float u, v; tt = intersect_leaf(r, la, lb, lc, u, v);
    visible = (tt > 0.0f && tt < TR_INF); found = false;
    if (!visible) return TR_DONE;
    int cur = root_enter(rt, r, anypar);
    if (cur == -tleaf - 1) { found = true; cur = TR_DONE; }          // single-leaf tree: the root is the target
    return cur;
}
__device__ __forceinline__ bool trace_shadow_visible(const TreeView& tv, const TreeRoot& rt, const RayPre& r, bool active, int target_leaf,
                                                     unsigned long long* cnt, float* tt_out = nullptr) {
#ifdef TR_COUNTERS
    unsigned cnt_nodes = 0, cnt_leaves = 0;
#endif
    const bool anypar = r.px || r.py || r.pz;
    bool visible = false, found = false; float tt = TR_INF;
    LocalStack st; st.reset();
    int cur = TR_DONE;
    const int tlink = -target_leaf - 1;
    if (active) {
        float4 la, lb, lc; leaf_fetch<TM_GLOBAL>(tv, target_leaf, 0, la, lb, lc);
        TR_COUNT(cnt_leaves);
        cur = shadow_enter(tv, rt, r, anypar, target_leaf, la, lb, lc, tt, visible, found);
    }
    while (cur != TR_DONE) {
        if (cur >= 0) { TR_COUNT(cnt_nodes); cur = node_step<TM_GLOBAL, true>(tv, r, anypar, tt, tlink, found, st, cur, 0); }
        else {
            TR_COUNT(cnt_leaves);
            const int k = -cur - 1;
            float4 la, lb, lc; leaf_fetch<TM_GLOBAL>(tv, k, 0, la, lb, lc);
            float u, v, t = intersect_leaf(r, la, lb, lc, u, v);
            if (blocks_target(t, k, tt, target_leaf)) { visible = false; cur = TR_DONE; } else cur = st.pop();
        }
    }
#ifdef TR_COUNTERS
    if (cnt && active) { atomicAdd(cnt, 2ull * cnt_nodes); atomicAdd(cnt + 1, (unsigned long long)cnt_leaves); }
#endif
    if (tt_out) *tt_out = tt;          // distance to the target primitive = the reference's hit_t when the target is the nearest hit
    return visible && found;
}

// ---- TMA bulk staging into shared memory --------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// One thread issues up to two cp.async.bulk (global -> shared, mbarrier completion); tma_stage_wait() makes everybody wait.
// bytes must be multiples of 16 and all pointers 16-byte aligned.
__device__ __forceinline__ void tma_stage_issue(void* smem_dst, const void* gsrc, unsigned bytes, void* smem_dst2,
                                                const void* gsrc2, unsigned bytes2, unsigned long long* bar) {
    const unsigned bar_a = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes + bytes2) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(bar_a) : "memory");
        if (bytes2)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem_dst2)), "l"(gsrc2), "r"(bytes2), "r"(bar_a) : "memory");
    }
}
__device__ __forceinline__ void tma_stage_wait(unsigned long long* bar) {
    __syncthreads();                    // the barrier was initialised by thread 0
    const unsigned bar_a = smem_u32(bar);
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
