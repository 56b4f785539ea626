// wavefront.cu — the PT_RGB hot path as a wavefront of persistent-thread kernels.
//
// Replaces PathTrace.render (integrator/PT_RGB.py:44-136), a one-thread-per-pixel mega-kernel with a
// per-pixel traversal stack in global memory, by
//     generate -> [ trace -> shade{terminal | disney | glass} -> shadow ] x max_depth -> accumulate
// over compacted SoA queues that live in HBM:
//   path queue (ping-pong)  A=(o.xyz, d.x)  B=(d.y, d.z, brdf_pdf, slot|spec<<31)  C=(T.rgb, pixel id)
//   hit queue               (t, prim, u, v)                                   16 B
//   shadow queue            (p.xyz, d.x) (d.y, d.z, target prim, slot) (contribution.rgb, -)   48 B
//   sample radiance L       float4 per (frame-in-batch, pixel)
// Queue sizes are device counters, so the whole batch is one fixed launch sequence (CUDA-graph
// replayable); trace kernels are persistent and fetch 32-ray chunks with one atomic per warp; live
// paths are compacted with warp-ballot appends; shade work is sorted by material class.
// Per-path RNG is counter based (Philox4x32-10 keyed by seed; pixel, frame, block), see common.cuh.
#include "ctx.h"
#include "common.cuh"
#include "trace.cuh"
#include "spectral.cuh"

#define SPEC_BIT 0x80000000u

struct BatchParams { int frame_begin; int n_frames; unsigned long long seed; int max_depth; int pad; };

struct WfArgs {
    // scene
    const float4* nodes2; const float4* leaves4; const float4* small_img;   // TrNode2 array, TrLeaf array, replicated image of a small tree
    int nint, nleaves, top, stack_cap;             // internal nodes, leaves, staged top nodes (TM_GTOP), stack entries kept in shared memory
    TreeRoot root;                                 // root box + link (0: internal root, -1: the tree is a single leaf)
    unsigned stage_bytes;                          // bytes of shared memory in front of the stack (tree image / top nodes), multiple of 128
    const TrShade* shade; const float* material; const float4* matlin; const int* light; int nl; const int* leaf_of_prim;
    const int* env; int env_w, env_h; float env_power;
    TrCamera cam;
    // film / tiles
    int W, H; const int* tiles; int npix;          // npix = local pixel slots = n_local_tiles * 1024
    float* hdr;
    // queues
    float4* pa[2]; float4* pb[2]; float4* pc[2];
    float4* hit; int* cls; size_t cap;
    float4* sa[2]; float4* sb[2]; float4* sc[2];   // shadow queue, ping-pong by depth parity (shadow(d) overlaps trace/shade(d+1))
    float4* L;                                     // terminal term (emitter / environment) per sample
    float4* Lnee;                                  // sum of the NEE terms per sample, in depth order
    float* vis;                                    // BDPT connection queries (k_shadow<.., QUERY>): t of the visible target or -1, per item
    TrCounters* ctr;
    const BatchParams* bp;
    int tail_max;                                  // hand the chain to k_tail once its live paths drop to this (0 = never)
    int tail_chunk;                                // k_tail: at least this many paths per warp
    int frame_off, sub_frames;                     // this chain renders local frames [frame_off, frame_off + sub_frames) of the batch
    int probe;                                     // test hooks: k_tail records the hit of each path's first walk in hit[] and stops there
    SpecDev spec;                                  // tables of the spectral integrator (PT_Spec), zero for PT_RGB
};

// local pixel slot p -> pixel coordinates.  32x32 tiles, inside a tile 4(x) x 8(y) pixel blocks per warp
__device__ __forceinline__ bool slot_to_pixel(const WfArgs& a, int p, int& x, int& y) {
    int tile = a.tiles[p >> 10];
    int ntx = (a.W + TR_TILE - 1) / TR_TILE;
    int tx = tile % ntx, ty = tile / ntx;
    int w = p & 1023, b = w >> 5, l = w & 31;
    x = tx * TR_TILE + (b & 7) * 4 + (l >> 3);
    y = ty * TR_TILE + (b >> 3) * 8 + (l & 7);
    return x < a.W && y < a.H;
}

// Camera.py:122-142
__device__ __forceinline__ V3 camera_dir(const TrCamera& c, int i, int j, float jx, float jy) {
    float x = ((float)i + jx - c.cx) / c.fx, y = ((float)j + jy - c.cy) / c.fy, z = -1.0f;
    const float* m = c.view_inv;
    V3 w = mk3((m[0] * x + m[1] * y) + m[2] * z, (m[4] * x + m[5] * y) + m[6] * z, (m[8] * x + m[9] * y) + m[10] * z);
    return normalize3(w);
}

// warp-aggregated append: returns the queue position for lanes with pred, -1 otherwise
__device__ __forceinline__ int warp_append(int* counter, bool pred) {
    unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return -1;
    int lane = threadIdx.x & 31, leader = __ffs(m) - 1, base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return pred ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}

// block-aggregated append: one atomic per CTA instead of one per warp (thousands of warps appending to ONE counter serialise in
// the L2 atomic unit and every warp waits for its round trip).  All threads of the block must call it together (block-uniform
// loops); returns the queue position for threads with pred, -1 otherwise.
__device__ __forceinline__ int block_append(int* counter, bool pred) {
    __shared__ int s_cnt[WF_THREADS / 32];
    __shared__ int s_base;
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_cnt[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < WF_THREADS / 32; ++w) { int c = s_cnt[w]; s_cnt[w] = tot; tot += c; }
        s_base = tot ? atomicAdd(counter, tot) : 0;
    }
    __syncthreads();
    const int pos = s_base + s_cnt[warp] + __popc(m & ((1u << lane) - 1u));
    __syncthreads();                                   // the shared slots are rewritten by the next call
    return pred ? pos : -1;
}

// two appends (to two queues) with one set of barriers: k_shade emits a continuation and a shadow ray per vertex
__device__ __forceinline__ void block_append2(int* counter_a, bool pred_a, int* counter_b, bool pred_b, int& pos_a, int& pos_b) {
    __shared__ int s_cnt2[2][WF_THREADS / 32];
    __shared__ int s_base2[2];
    const unsigned ma = __ballot_sync(0xffffffffu, pred_a), mb = __ballot_sync(0xffffffffu, pred_b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_cnt2[0][warp] = __popc(ma); s_cnt2[1][warp] = __popc(mb); }
    __syncthreads();
    if (threadIdx.x < 2) {                                 // thread 0: queue a, thread 1: queue b -- one atomic instruction for both
        int tot = 0;
#pragma unroll
        for (int w = 0; w < WF_THREADS / 32; ++w) { int c = s_cnt2[threadIdx.x][w]; s_cnt2[threadIdx.x][w] = tot; tot += c; }
        s_base2[threadIdx.x] = tot ? atomicAdd(threadIdx.x == 0 ? counter_a : counter_b, tot) : 0;
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1u;
    pos_a = pred_a ? s_base2[0] + s_cnt2[0][warp] + __popc(ma & lt) : -1;
    pos_b = pred_b ? s_base2[1] + s_cnt2[1][warp] + __popc(mb & lt) : -1;
    __syncthreads();                                   // the shared slots are rewritten by the next call
}

// true when an earlier stage handed the remaining paths of this chain to the tail kernel (tail_from: 0 = no hand-over,
// d >= 1 = k_tail owns the paths from depth d on)
__device__ __forceinline__ bool tail_took_over(const WfArgs& a, int depth) {
    const int tf = a.ctr->tail_from;
    return tf > 0 && depth >= tf;
}

// ------------------------------------------------------------------ generate
// SPEC (PT_Spec): B.z carries the pixel id (brdf_pdf / perfect_spec are unused there: the emitter term is never MIS
// weighted, integrator/PT_Spec.py:219-231) and C the four hero-wavelength throughputs.
template <bool SPEC>
__global__ void __launch_bounds__(WF_THREADS) k_generate(WfArgs a) {
    const BatchParams bp = *a.bp;
    const int nf = max(0, min(a.sub_frames, bp.n_frames - a.frame_off));
    const int total = nf * a.npix;
    const int stride = gridDim.x * blockDim.x;
    const int total_r = (total + WF_THREADS - 1) / WF_THREADS * WF_THREADS;   // block-uniform trip count (block_append)
    for (int sl = blockIdx.x * blockDim.x + threadIdx.x; sl < total_r; sl += stride) {
        bool valid = false; int x = 0, y = 0, frame = 0, s = 0;
        if (sl < total) {
            int f = sl / a.npix, p = sl - f * a.npix;
            f += a.frame_off;
            s = f * a.npix + p;                                   // sample slot within the whole batch
            frame = bp.frame_begin + f;
            valid = slot_to_pixel(a, p, x, y);
            a.L[s] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); a.Lnee[s] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        int q = block_append(&a.ctr->nq[0], valid);
        if (valid) {
            unsigned pix = ((unsigned)x << 16) | (unsigned)y;
            float jx = 0.0f, jy = 0.0f;
            if (frame != 0) { float4 r = rng4(bp.seed, pix, (unsigned)frame, 0u); jx = r.x - 0.5f; jy = r.y - 0.5f; }
            V3 d = camera_dir(a.cam, x, y, jx, jy);
            a.pa[0][q] = make_float4(a.cam.eye[0], a.cam.eye[1], a.cam.eye[2], d.x);
            if (SPEC) {
                a.pb[0][q] = make_float4(d.y, d.z, __uint_as_float(pix), __uint_as_float((unsigned)s | SPEC_BIT));
                a.pc[0][q] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
            } else {
                a.pb[0][q] = make_float4(d.y, d.z, 1.0f, __uint_as_float((unsigned)s | SPEC_BIT));
                a.pc[0][q] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(pix));
            }
        }
    }
}

// ------------------------------------------------------------------ BVH staging
// Dynamic shared memory of the traversal kernels: [tree image (stage_bytes)] [stack: stack_cap x WF_THREADS ints].
extern __shared__ __align__(128) unsigned char wf_smem[];

// Thread 0 issues the TMA bulk copies of the part of the tree this MODE keeps in shared memory (nothing for TM_GLOBAL).
// The tree is static for the whole batch, so this may run before griddepcontrol.wait (programmatic dependent launch).
template <int MODE>
__device__ __forceinline__ void tree_stage_issue(const WfArgs& a, unsigned long long* bar) {
    if (MODE == TM_REP) tma_stage_issue(wf_smem, a.small_img, a.stage_bytes, nullptr, nullptr, 0u, bar);
    else if (MODE == TM_SMEM) {
        const unsigned nb = (unsigned)a.nint * 64u;
        tma_stage_issue(wf_smem, a.nodes2, nb, wf_smem + nb, a.leaves4, (unsigned)a.nleaves * 48u, bar);
    } else if (MODE == TM_GTOP) tma_stage_issue(wf_smem, a.nodes2, (unsigned)a.top * 64u, nullptr, nullptr, 0u, bar);
}
template <int MODE>
__device__ __forceinline__ void tree_stage_wait(const WfArgs& a, unsigned long long* bar, TreeView& tv, int*& stack_base) {
    if (MODE != TM_GLOBAL) tma_stage_wait(bar);
    const float4* sm = (const float4*)wf_smem;
    tv.gnodes = a.nodes2; tv.gleaves = a.leaves4; tv.top = a.top; tv.snodes = sm;
    tv.sleaves = (MODE == TM_REP) ? sm + (size_t)a.nint * 32 : sm + (size_t)a.nint * 4;
    stack_base = (int*)(wf_smem + a.stage_bytes) + threadIdx.x;
}
// programmatic dependent launch: everything above this call overlaps the previous kernel of the stream; nothing the previous
// kernel wrote (queues, counters) may be read before it.  A no-op for kernels launched without the attribute.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------ trace (closest hit)
// Persistent warps with per-lane ray replacement: a warp takes WF_CHUNK consecutive rays of the queue with
// one atomic, hands them to lanes as they become free, and never waits for its slowest ray.  Inside the
// loop every lane runs the same three blocks (refill / node step / batched leaf step, see trace.cuh), so
// the issue slots are spent with most lanes active although the per-ray walk lengths differ by 10x.
#ifndef WF_NODE_STEPS
#define WF_NODE_STEPS 3             // node steps per round, tree in shared memory (measured on C2: 3 beats 2, 4 and 6)
#endif
#ifndef WF_NODE_STEPS_GLOBAL
#define WF_NODE_STEPS_GLOBAL 2      // ... tree in global memory (measured on C3: 2 and 3 beat 1 and 4)
#endif
template <int MODE> struct NodeSteps { static constexpr int value = (MODE == TM_GLOBAL || MODE == TM_GTOP) ? WF_NODE_STEPS_GLOBAL : WF_NODE_STEPS; };
#ifndef WF_REFILL_MIN
#define WF_REFILL_MIN 8             // idle lanes of a warp before it fetches new rays, tree in shared memory
#endif
#ifndef WF_REFILL_MIN_GLOBAL
#define WF_REFILL_MIN_GLOBAL 12     // ... tree in global memory (measured on C3: 12 beats 8 by 3 %, 6 loses 3 %; on C2 8 beats 6 and 12)
#endif
template <int MODE> struct RefillMin { static constexpr int value = (MODE == TM_GLOBAL || MODE == TM_GTOP) ? WF_REFILL_MIN_GLOBAL : WF_REFILL_MIN; };

struct WarpFeed { int cb, ce, chunk; bool more; };

// rays a warp takes per atomic: large enough to amortise the atomic, small enough that every resident
// warp gets work when the queue is short (deep bounces) and that the end-of-kernel tail stays short
__device__ __forceinline__ WarpFeed make_feed(int n) {
    WarpFeed f; f.cb = f.ce = 0; f.more = true;
    int blocks = min((int)gridDim.x, (n + WF_THREADS - 1) / WF_THREADS);       // surplus CTAs have left (see k_trace)
    int warps = max(1, blocks) * (blockDim.x >> 5);
    int c = n / (warps * 8);
    f.chunk = min(256, max(32, (c + 31) & ~31));
    return f;
}

// hand the next queue entries to the idle lanes (idle = warp-uniform mask); returns this lane's new queue
// index or -1 and removes the served lanes from `idle`
__device__ __forceinline__ int feed_lanes(WarpFeed& f, int* cursor, int n, unsigned& idle, int lane) {
    if (f.cb >= f.ce && f.more) {
        int b = 0;
        if (lane == 0) b = atomicAdd(cursor, f.chunk);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= n) { f.more = false; f.cb = f.ce = 0; } else { f.cb = b; f.ce = min(b + f.chunk, n); }
    }
    const int avail = f.ce - f.cb;
    int q = -1;
    if (avail > 0) {
        const int rank = __popc(idle & ((1u << lane) - 1u));
        const bool mine = ((idle >> lane) & 1u) && rank < avail;
        if (mine) q = f.cb + rank;
        f.cb += min(__popc(idle), avail);
        idle &= ~__ballot_sync(0xffffffffu, mine);
    }
    return q;
}

// append the top `take` (<= 32) buffered (queue index | class << 30) entries to the material-sorted shade queues.
// The three class counters are bumped by ONE atomic instruction (lanes 0..2, one address each): a single round trip to the
// L2 atomic unit per flush instead of three back-to-back ones.
__device__ __forceinline__ void flush_retired(const WfArgs& a, int depth, const int* rbuf, int rcount, int take, int lane) {
    unsigned e = (lane < take) ? (unsigned)rbuf[rcount - take + lane] : 0u;
    int c = (lane < take) ? (int)(e >> 30) : -1, qq = (int)(e & 0x3fffffffu);
    const unsigned m0 = __ballot_sync(0xffffffffu, c == 0), m1 = __ballot_sync(0xffffffffu, c == 1), m2 = __ballot_sync(0xffffffffu, c == 2);
    int base = 0;
    if (lane < 3) {
        const unsigned mk = lane == 0 ? m0 : (lane == 1 ? m1 : m2);
        if (mk) base = atomicAdd(&a.ctr->ncls[depth][lane], __popc(mk));
    }
    const int b0 = __shfl_sync(0xffffffffu, base, 0), b1 = __shfl_sync(0xffffffffu, base, 1), b2 = __shfl_sync(0xffffffffu, base, 2);
    if (c >= 0) {
        const unsigned mk = c == 0 ? m0 : (c == 1 ? m1 : m2);
        const int pos = (c == 0 ? b0 : (c == 1 ? b1 : b2)) + __popc(mk & ((1u << lane) - 1u));
        a.cls[(size_t)c * a.cap + pos] = qq;
    }
    __syncwarp();
}

#ifndef WF_TRACE_MIN_BLOCKS
#define WF_TRACE_MIN_BLOCKS 4
#endif
#ifndef WF_TRACE_MIN_BLOCKS_GLOBAL
#define WF_TRACE_MIN_BLOCKS_GLOBAL 4
#endif
template <int MODE>
__global__ void __launch_bounds__(WF_THREADS, (MODE == TM_GLOBAL || MODE == TM_GTOP) ? WF_TRACE_MIN_BLOCKS_GLOBAL : WF_TRACE_MIN_BLOCKS)
k_trace(WfArgs a, int depth) {
    grid_dep_wait(); grid_dep_launch();
    if (tail_took_over(a, depth)) return;
    // surplus CTAs of a short queue leave BEFORE staging the tree (deep bounces, small shards: the per-stage floor)
    const int n = a.ctr->nq[depth];
    if ((long long)blockIdx.x * WF_THREADS >= (long long)n) return;
    __shared__ unsigned long long bar;
    TreeView tv; int* stack_base;
    tree_stage_issue<MODE>(a, &bar);
    tree_stage_wait<MODE>(a, &bar, tv, stack_base);
    __shared__ int retire_buf[WF_THREADS / 32][64];      // finished (queue index | class << 30), flushed 32 at a time
    const int pp = depth & 1;
    const float4* __restrict__ pa = a.pa[pp]; const float4* __restrict__ pb = a.pb[pp];
    int* cursor = &a.ctr->wf_trace[depth];
    const int lane = threadIdx.x & 31, lane8 = lane & 7;
    int* rbuf = retire_buf[threadIdx.x >> 5];
    int rcount = 0;
    WarpFeed feed = make_feed(n);
    unsigned idle = 0xffffffffu;                        // warp-uniform: lanes without a ray
    bool anypar = false, found = false; int q = 0, cur = TR_DONE;
    SmemStack st; st.init(stack_base);
    RayPre r = make_ray(mk3(0.f, 0.f, 0.f), mk3(1.f, 1.f, 1.f));
    HitRec h; hit_reset(h);
#ifdef TR_COUNTERS
    unsigned long long cnt_nodes = 0, cnt_leaves = 0;
#endif
    while (true) {
        // ---- refill: only when enough lanes are free (the ray set-up runs at the utilisation of the idle set)
        if ((__popc(idle) >= RefillMin<MODE>::value || idle == 0xffffffffu) && (feed.more || feed.cb < feed.ce)) {
            int nq = feed_lanes(feed, cursor, n, idle, lane);
            if (nq >= 0) {
                float4 A = pa[nq], B = pb[nq];
                r = make_ray(mk3(A.x, A.y, A.z), mk3(A.w, B.x, B.y));
                anypar = r.px || r.py || r.pz;
                hit_reset(h);
                q = nq; st.reset(); cur = root_enter(a.root, r, anypar);
            }
        }
        if (idle == 0xffffffffu) { if (!feed.more && feed.cb >= feed.ce) break; else continue; }
        const bool has = !((idle >> lane) & 1u);
        // ---- node steps of a round; the warp-level bookkeeping below is paid once per round.  A lane that reaches a leaf waits
        //      (parked, cur < 0) for the leaf step of the round.  (Measured and rejected: putting the leaf aside and walking on
        //      from the stack within the round -- fewer idle slots, but the extra instructions per step cost more: C2 trace
        //      7.48 -> 7.57 ms, C3 4.76 -> 4.97 ms.)
#pragma unroll
        for (int step = 0; step < NodeSteps<MODE>::value; ++step) {
            if (has && (unsigned)cur < (unsigned)TR_DONE) {
                TR_COUNT(cnt_nodes);
                cur = node_step<MODE, false>(tv, r, anypar, h.t, 0, found, st, cur, lane8);
            }
        }
        // ---- leaf step of the lanes parked at a leaf
        if (has && cur < 0) {
            TR_COUNT(cnt_leaves);
            const int k = -cur - 1;
            float4 la, lb, lc; leaf_fetch<MODE>(tv, k, lane8, la, lb, lc);
            float u, v, t = intersect_leaf(r, la, lb, lc, u, v);
            if (closer(t, k, h.t, h.leaf)) { h.t = t; h.u = u; h.v = v; h.prim = __float_as_int(la.w); h.mat = __float_as_int(lc.w); h.leaf = k; }
            cur = st.pop();
        }
        // ---- retire finished rays: hit record now, material-sorted queue entries through the warp's buffer
        const unsigned finm = __ballot_sync(0xffffffffu, has && cur == TR_DONE);
        if (finm != 0u) {
            if ((finm >> lane) & 1u) {
                a.hit[q] = make_float4(h.t, __int_as_float(h.prim), h.u, h.v);
                int c = 0;
                if (h.prim >= 0) {
                    int mt = (int)__ldg(a.material + (size_t)h.mat * 10);
                    c = (mt == TR_MAT_LIGHT) ? 0 : (mt == TR_MAT_GLASS ? 2 : 1);
                }
                rbuf[rcount + __popc(finm & ((1u << lane) - 1u))] = q | (c << 30);
            }
            rcount += __popc(finm);
            idle |= finm;
            __syncwarp();
        }
        if (rcount >= 32) { flush_retired(a, depth, rbuf, rcount, 32, lane); rcount -= 32; }
    }
    if (rcount > 0) flush_retired(a, depth, rbuf, rcount, rcount, lane);
#ifdef TR_COUNTERS
    atomicAdd(a.ctr->visits, 2ull * cnt_nodes); atomicAdd(a.ctr->visits + 1, cnt_leaves);      // 2 child boxes (2 x 32 B) per node visit
#endif
}

// ------------------------------------------------------------------ shading helpers
struct Surf { V3 pos, gn, n; int mat; float area; };

// attribute reconstruction of Scene.intersect_prim (Scene.py:537-600) from (prim, u, v)
__device__ __forceinline__ Surf surface_at(const WfArgs& a, int prim, float u, float v, V3 o, V3 d, float t) {
    const float4* q = a.shade[prim].q;
    float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
    Surf s; s.mat = __float_as_int(q0.w); s.area = q2.w;
    int kind = __float_as_int(q1.w);
    if (kind == 0) {
        float4 q3 = __ldg(q + 3), q4 = __ldg(q + 4), q5 = __ldg(q + 5);
        V3 v1 = f4xyz(q0), v2 = f4xyz(q1), v3 = f4xyz(q2);
        float aa = 1.0f - u - v, bb = u, cc = v;
        s.gn = mk3(q3.w, q4.w, q5.w);                      // normalize((v2 - v1) x (v3 - v1)), precomputed per primitive by k_shade_table
        s.pos = (aa * v1 + bb * v2) + cc * v3;
        s.n = normalize3((aa * f4xyz(q3) + bb * f4xyz(q4)) + cc * f4xyz(q5));
    } else {
        // sphere: hit_nor = hit_pos - c with the scalar quadratic coefficient c (sic, Scene.py:588,595)
        V3 ce = f4xyz(q0); float r = q1.x;
        V3 oc = ce - o; float c = dot3(oc, oc) - r * r;
        s.pos = o + t * d;
        s.n = normalize3(s.pos - mk3(c, c, c)); s.gn = s.n;
    }
    return s;
}

// texture/Texture.py:36-69
__device__ __forceinline__ V3 env_texel(const WfArgs& a, float fx, float fy) {
    int x = min(max((int)fx, 0), a.env_w - 1), y = min(max((int)fy, 0), a.env_h - 1);
    int c = __ldg(a.env + (size_t)x * a.env_h + y);
    return mk3((float)((c & 0x00FF0000) >> 16) / 255.0f, (float)((c & 0x0000FF00) >> 8) / 255.0f, (float)(c & 0xFF) / 255.0f);
}
__device__ __forceinline__ V3 env_texture2d(const WfArgs& a, float u, float v) {
    float x = clampf(u * (float)a.env_w, 0.0f, (float)a.env_w - 1.0f), y = clampf(v * (float)a.env_h, 0.0f, (float)a.env_h - 1.0f);
    float lx = floorf(x), ly = floorf(y);
    float wbt = y - floorf(y), wlr = x - floorf(x);
    return mix3(mix3(env_texel(a, lx, ly), env_texel(a, lx + 1.0f, ly), wlr),
                mix3(env_texel(a, lx, ly + 1.0f), env_texel(a, lx + 1.0f, ly + 1.0f), wlr), wbt);
}

struct LightSample { V3 pos, normal, dir, emission; float dist, choice_pdf; int prim; int kind; float p0, p1, p2; };   // kind / p*: shading-table kind and shape parameters of the emitter
// Scene.sample_li (Scene.py:477-518) with get_random_light_prim_index (:423-428),
// get_prim_random_point_normal (:381-420) and get_prim_area (:324-350, precomputed per primitive)
__device__ __forceinline__ LightSample sample_li(const WfArgs& a, V3 p, float u_idx, float ua, float ub, bool li_terms = true) {
    LightSample L;
    int index = (int)(u_idx * (float)a.nl); if (index >= a.nl) index = a.nl - 1;
    int pi = __ldg(a.light + index);
    const float4* q = a.shade[pi].q;
    float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
    int kind = __float_as_int(q1.w);
    V3 pos, nor;
    if (kind == 0) {
        float4 q3 = __ldg(q + 3), q4 = __ldg(q + 4), q5 = __ldg(q + 5);
        V3 v1 = f4xyz(q0), v2 = f4xyz(q1), v3 = f4xyz(q2);
        if (ua + ub > 1.0f) { ua = 1.0f - ua; ub = 1.0f - ub; }
        pos = (v1 + (v3 - v1) * ua) + (v2 - v1) * ub;
        nor = normalize3(((1.0f - ua - ub) * f4xyz(q3) + f4xyz(q4) * ua) + f4xyz(q5) * ub);   // roles as in Scene.py:401-402
    } else if (kind == 1) {
        nor = uniform_sample_sphere(ua, ub); pos = f4xyz(q0) + nor * q1.x;
    } else { nor = mk3(q1.x, q1.y, q1.z); pos = f4xyz(q0); }
    nor = normalize3(nor);      // Scene.py:420
    int mid = __float_as_int(q0.w);
    const float* m = a.material + (size_t)mid * 10;
    L.emission = mk3(__ldg(m + 2), __ldg(m + 3), __ldg(m + 4));
    L.choice_pdf = 1.0f / ((float)a.nl * q2.w);
    nor = normalize3(nor);      // Scene.py:486
    V3 dir = p - pos; float dist = length3(dir); dir = dir / dist;
    if (kind >= 2 && li_terms) { // Scene.py:493-516: spot falloff (kind 2) / laser radius cut-off (kind 3) scale the emission
        float visable = 1.0f;
        if (kind == 2) {
            float x = tr_acosf(fabsf(dot3(dir, nor)));
            if (x > q2.y) visable = 0.0f;
            else if (x > q2.x) visable *= 1.0f - (x - q2.x) / (q2.y - q2.x);
        } else if (kind == 3) {
            L.choice_pdf = 1.0f / (float)a.nl;
            float proj = dot3(dir, nor) * dist;
            float r = sqrtf(dist * dist - proj * proj);
            if (r > q2.x) visable = 0.0f;
        }
        L.emission = L.emission * visable;
    }
    L.pos = pos; L.normal = nor; L.dir = dir; L.dist = dist; L.prim = pi;
    L.kind = kind; L.p0 = q2.x; L.p1 = q2.y; L.p2 = q2.z;
    return L;
}

// ------------------------------------------------------------------ shade
// One launch covers the three material-sorted queues back to back: [terminal | disney | glass].
// One path vertex of PathTrace.render (integrator/PT_RGB.py:66-132): emitter / environment terminal terms go
// straight to L[slot]; a surviving path comes back as its next queue record, a NEE sample as a shadow record.
// cl: 0 terminal (miss or emitter), 1 Disney, 2 glass.  Used by the wavefront shade kernel and by the tail kernel.
struct ShadeOut { bool cont, shadow; float4 nA, nB, nC, sA, sB, sC; };

__device__ __forceinline__ void shade_path(const WfArgs& a, const BatchParams& bp, int depth, int cl,
                                           float4 A, float4 B, float4 C, float4 Hh, ShadeOut& out) {
    out.cont = false; out.shadow = false;
    const bool last = depth + 1 >= bp.max_depth;
    V3 o = mk3(A.x, A.y, A.z), d = mk3(A.w, B.x, B.y);
    float brdf_pdf = B.z;
    unsigned sw = __float_as_uint(B.w); unsigned slot = sw & ~SPEC_BIT; bool perfect_spec = (sw & SPEC_BIT) != 0;
    V3 T = mk3(C.x, C.y, C.z); unsigned pix = __float_as_uint(C.w);
    float t = Hh.x; int prim = __float_as_int(Hh.y);
    if (prim < 0) {
        // miss: equirect environment lookup (integrator/PT_RGB.py:127-132)
        if (a.env_w > 0) {
            float dis = sqrtf(d.x * d.x + d.z * d.z);
            float tx = (tr_atan2f(d.z, d.x) + TR_PI_ENV) / TR_PI_ENV / 2.0f;
            float ty = tr_atan2f(d.y, dis) / TR_PI_ENV + 0.5f;
            V3 e = (srgb_to_lrgb(env_texture2d(a, tx, ty)) * T) * a.env_power;
            float4 Lv = a.L[slot]; Lv.x += e.x; Lv.y += e.y; Lv.z += e.z; a.L[slot] = Lv;
        }
        return;
    }
    Surf s = surface_at(a, prim, Hh.z, Hh.w, o, d, t);
    V3 fn = signf_(dot3(-d, s.gn)) * s.n;                 // UF.faceforward(normal, -direction, gnormal)
    const float* m = a.material + (size_t)s.mat * 10;
    V3 mcol = mk3(__ldg(m + 2), __ldg(m + 3), __ldg(m + 4));
    float p0 = __ldg(m + 5), p1 = __ldg(m + 6);
    if (cl == 0) {
        // emitter (integrator/PT_RGB.py:72-81)
        V3 e;
        if (perfect_spec) e = T * mcol;
        else {
            float fCos = fabsf(dot3(d, s.gn));
            float area = s.area * (float)a.nl;
            float light_pdf = (t * t) / (area * fCos);
            e = power_heuristic(brdf_pdf, light_pdf) * T * mcol;
        }
        float4 Lv = a.L[slot]; Lv.x += e.x; Lv.y += e.y; Lv.z += e.z; a.L[slot] = Lv;
        return;
    }
    V3 rc = f4xyz(__ldg(a.matlin + s.mat));            // srgb_to_lrgb(Kd), PT_RGB.py:86, precomputed per material
    unsigned frame = (unsigned)bp.frame_begin + slot / (unsigned)a.npix;
    float4 R0 = rng4(bp.seed, pix, frame, 1u + 2u * (unsigned)depth);
    float4 R1 = rng4(bp.seed, pix, frame, 2u + 2u * (unsigned)depth);
    V3 next_d; float f_or_b = 1.0f, brdf, pdf; bool spec;
    if (cl == 2) {
        spec = true;
        next_d = glass_sample(d, s.n, p0, R0.w, f_or_b);      // brdf/Glass.py:9-34
        brdf = 1.0f; pdf = 1.0f;
    } else {
        spec = false;
        if (a.nl > 0) {
            LightSample ls = sample_li(a, s.pos, R0.x, R0.y, R0.z);
            float NdotL_s = dot3(fn, ls.dir), NdotL_l = dot3(ls.normal, ls.dir);
            if (NdotL_s < 0.0f && NdotL_l > 0.0f) {
                float b2, p2; disney_evaluate_pdf(fn, -d, -ls.dir, p0, p1, b2, p2);
                V3 c = mk3(0.0f, 0.0f, 0.0f);
                if (p2 > 0.0f) {
                    float light_pdf = ls.dist * ls.dist * ls.choice_pdf / NdotL_l;
                    float wgt = power_heuristic(light_pdf, p2) / fmaxf(0.0001f, light_pdf);
                    c = ((((wgt * ls.emission) * T) * rc) * b2) * fabsf(NdotL_s);
                }
                out.shadow = true;
                out.sA = make_float4(ls.pos.x, ls.pos.y, ls.pos.z, ls.dir.x);
                out.sB = make_float4(ls.dir.y, ls.dir.z, __int_as_float(prim), __uint_as_float(slot));
                out.sC = make_float4(c.x, c.y, c.z, 0.0f);
            }
        }
        next_d = disney_sample(d, fn, p0, p1, R0.w, R1.x, R1.y);
        disney_evaluate_pdf(fn, -d, next_d, p0, p1, brdf, pdf);
        brdf *= fabsf(dot3(s.n, next_d));
    }
    V3 next_o = offset_ray(s.pos, signf_(f_or_b) * fn);
    if (pdf > 0.0f) {
        bool alive = true;
        if (f_or_b < 0.0f) { float Rr = tr_expf(-t / p1); if (R1.z >= Rr) alive = false; }   // PT_RGB.py:118-122
        if (alive && !last) {
            T = T * ((brdf / pdf) * rc);
            out.cont = true;
            out.nA = make_float4(next_o.x, next_o.y, next_o.z, next_d.x);
            out.nB = make_float4(next_d.y, next_d.z, pdf, __uint_as_float(slot | (spec ? SPEC_BIT : 0u)));
            out.nC = make_float4(T.x, T.y, T.z, __uint_as_float(pix));
        }
    }
}


// One path vertex of PT_Spec.PathTrace.render (integrator/PT_Spec.py:203-277), four hero wavelengths per path.
// Queue record: B.z = pixel id, C = throughput of the four lanes; the hero wavelength is recomputed from the
// counter-based RNG (block 0, .z) instead of being stored.  RNG blocks: 1+2d = (light index, a, b, lobe / Fresnel
// coin), 2+2d = (r1, r2, hero index of the glass bounce, -).
__device__ __forceinline__ void shade_path_spec(const WfArgs& a, const BatchParams& bp, int depth, int cl,
                                                float4 A, float4 B, float4 C, float4 Hh, ShadeOut& out) {
    out.cont = false; out.shadow = false;
    const SpecDev& sd = a.spec;
    const bool last = depth + 1 >= bp.max_depth;
    V3 o = mk3(A.x, A.y, A.z), d = mk3(A.w, B.x, B.y);
    unsigned pix = __float_as_uint(B.z);
    unsigned slot = __float_as_uint(B.w) & ~SPEC_BIT;
    V4 T = ld4(C);
    unsigned frame = (unsigned)bp.frame_begin + slot / (unsigned)a.npix;
    float Lambda = HERO_LAMBDA_MIN + HERO_STEP * rng4(bp.seed, pix, frame, 0u).z;       // :191
    V4 light_rad = hero_sample(sd.sp[TR_SPEC_D65], Lambda);                                // :217
    float t = Hh.x; int prim = __float_as_int(Hh.y);
    if (prim < 0) {
        // miss: sky dome (:269-276)
        float dis = sqrtf(d.x * d.x + d.z * d.z);
        float beta = tr_atan2f(d.y, dis);
        float gamma = tr_acosf(dot3(d, mk3(__ldg(sd.sky + 110), __ldg(sd.sky + 111), __ldg(sd.sky + 112))));
        float theta = clampf(0.5f * TR_PI_ENV - beta, 0.0f, 0.5f * TR_PI_ENV);
        V4 ibl;
#pragma unroll
        for (int k = 0; k < HERO_N; ++k) ibl.v[k] = sky_radiance(sd, theta, gamma, Lambda + (float)k * HERO_STEP);
        a.L[slot] = st4(ld4(a.L[slot]) + (T * ibl) * light_rad);
        return;
    }
    Surf s = surface_at(a, prim, Hh.z, Hh.w, o, d, t);
    V3 fn = signf_(dot3(-d, s.gn)) * s.n;
    const float* m = a.material + (size_t)s.mat * 10;
    float p0 = __ldg(m + 5), p1 = __ldg(m + 6);
    V4 light_tint = emission_to_rad(sd, s.mat, Lambda);                                   // :218 (the SHADED surface's colour)
    if (cl == 0) {
        // emitter: perfect_spec is reset to 1 every bounce (:219), so the term is never MIS weighted (:222-231)
        float fCos = dot3(d, s.n);
        if (fCos < 0.0f) a.L[slot] = st4(ld4(a.L[slot]) + (T * light_rad) * light_tint);
        return;
    }
    V4 reflect_spec = get_spec_power(sd, s.mat, Lambda);                                  // :236
    float4 R0 = rng4(bp.seed, pix, frame, 1u + 2u * (unsigned)depth);
    float4 R1 = rng4(bp.seed, pix, frame, 2u + 2u * (unsigned)depth);
    V3 next_d; float f_or_b = 1.0f, brdf, pdf;
    if (cl == 2) {
        int index = (int)(R1.z * (float)HERO_N);                                          // Hero.get_rnd_hero (spectrum/HeroSample.py:33-36)
        float rl = Lambda + (float)index * HERO_STEP;
        next_d = glass_sample(d, s.n, get_glass_ior(rl), R0.w, f_or_b);                   // Glass.sample_lambda (brdf/Glass.py:39-65)
        brdf = 1.0f; pdf = 1.0f;
    } else {
        if (a.nl > 0) {
            LightSample ls = sample_li(a, s.pos, R0.x, R0.y, R0.z);
            float NdotL_s = dot3(fn, ls.dir), NdotL_l = dot3(ls.normal, ls.dir);
            if (NdotL_s < 0.0f && NdotL_l > 0.0f) {
                float b2, p2; disney_evaluate_pdf(fn, -d, -ls.dir, p0, p1, b2, p2);
                V4 c = mk4(0.0f);
                if (p2 > 0.0f) {
                    float light_pdf = ls.dist * ls.dist * ls.choice_pdf / NdotL_l;
                    float wgt = power_heuristic(light_pdf, p2) / fmaxf(0.0001f, light_pdf);
                    c = (((((wgt * light_rad) * light_tint) * T) * reflect_spec) * b2) * fabsf(NdotL_s);     // :257
                }
                out.shadow = true;
                out.sA = make_float4(ls.pos.x, ls.pos.y, ls.pos.z, ls.dir.x);
                out.sB = make_float4(ls.dir.y, ls.dir.z, __int_as_float(prim), __uint_as_float(slot));
                out.sC = st4(c);
            }
        }
        next_d = disney_sample(d, fn, p0, p1, R0.w, R1.x, R1.y);
        disney_evaluate_pdf(fn, next_d, -d, p0, p1, brdf, pdf);                           // (N, next_dir, -direction) (:260)
        brdf *= fabsf(dot3(s.n, next_d));
    }
    V3 next_o = offset_ray(s.pos, signf_(f_or_b) * fn);
    float maxc = fmaxf(T.v[2], fmaxf(T.v[0], T.v[1]));                                    // UF.max_component ignores the 4th lane (UtilsFunc.py:487-488)
    if (pdf > 0.0f && maxc > 0.0f && !last) {
        T = T * ((brdf * reflect_spec) / pdf);
        out.cont = true;
        out.nA = make_float4(next_o.x, next_o.y, next_o.z, next_d.x);
        out.nB = make_float4(next_d.y, next_d.z, __uint_as_float(pix), __uint_as_float(slot | SPEC_BIT));
        out.nC = st4(T);
    }
}

template <bool SPEC>
__device__ __forceinline__ void shade_any(const WfArgs& a, const BatchParams& bp, int depth, int cl,
                                          float4 A, float4 B, float4 C, float4 Hh, ShadeOut& out) {
    if (SPEC) shade_path_spec(a, bp, depth, cl, A, B, C, Hh, out); else shade_path(a, bp, depth, cl, A, B, C, Hh, out);
}

#ifndef WF_SHADE_MIN_BLOCKS
#define WF_SHADE_MIN_BLOCKS 4
#endif
template <bool SPEC>
__global__ void __launch_bounds__(WF_THREADS, SPEC ? 3 : WF_SHADE_MIN_BLOCKS) k_shade(WfArgs a, int depth) {
    grid_dep_wait(); grid_dep_launch();
    if (tail_took_over(a, depth)) return;
    const BatchParams bp = *a.bp;
    const int n0 = a.ctr->ncls[depth][0], n1 = a.ctr->ncls[depth][1], n2 = a.ctr->ncls[depth][2];
    const int n = n0 + n1 + n2;
    const int pp = depth & 1, np_ = pp ^ 1;
    const int stride = gridDim.x * blockDim.x;
    const int n_r = (n + WF_THREADS - 1) / WF_THREADS * WF_THREADS;          // block-uniform trip count (block_append)
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_r; w += stride) {
        ShadeOut o; o.cont = false; o.shadow = false;
        if (w < n) {
            int cl = (w < n0) ? 0 : (w < n0 + n1 ? 1 : 2);
            int q = a.cls[(size_t)cl * a.cap + (w - (cl == 0 ? 0 : (cl == 1 ? n0 : n0 + n1)))];
            shade_any<SPEC>(a, bp, depth, cl, a.pa[pp][q], a.pb[pp][q], a.pc[pp][q], a.hit[q], o);
        }
        int qn, qs;
        block_append2(&a.ctr->nq[depth + 1], o.cont, &a.ctr->nshadow[depth], o.shadow, qn, qs);
        if (o.cont) { a.pa[np_][qn] = o.nA; a.pb[np_][qn] = o.nB; a.pc[np_][qn] = o.nC; }
        if (o.shadow) { a.sa[pp][qs] = o.sA; a.sb[pp][qs] = o.sB; a.sc[pp][qs] = o.sC; }
    }
    // last block out decides whether the next depth is small enough for the tail kernel (queue size is final now)
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&a.ctr->shade_done[depth], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        int nn = *((volatile int*)&a.ctr->nq[depth + 1]);
        if (a.tail_max > 0 && depth + 1 < bp.max_depth && nn > 0 && nn <= a.tail_max) a.ctr->tail_from = depth + 1;
    }
}

// ------------------------------------------------------------------ tail (megakernel for the last few paths)
// Deep bounces carry a tiny fraction of the rays (C3: 1.7 % beyond depth 2) but every wavefront stage pays the latency
// of its slowest ray (~0.1-0.2 ms for a 1000-node walk), 3 stages x 13 depths.  Once a chain's live-path count drops to
// tail_max, this kernel takes the remaining paths and runs them to completion in registers.  Every lane is its own state
// machine (closest-hit walk -> shade -> shadow walk -> next bounce ...): a lane never waits for the other lanes' walks, so a
// path costs the sum of ITS OWN walk latencies instead of the per-bounce maximum over a warp -- what matters when a rank of
// an 8-GPU run is left with a few thousand paths and nothing else to overlap them with.  Paths are dealt out a few per warp
// (n / #warps) so that all resident warps work.  Same device functions and the same per-path operation order as the
// wavefront stages (NEE terms are added to Lnee in depth order, after shadow(depth-1): the kernel is enqueued behind it),
// so the film is bit-identical with or without the hand-over.
// Register budget: the RGB instantiation fits 128 registers (2 CTAs per SM).  The spectral one needs ~200; capped at 128 it
// spills, and that spilling build produced garbage tree links once in ~10 small renders on B200 (tools/stress_spec.py: film
// words differ / illegal shared-memory address in leaf_fetch, while the same source built without the cap, or with an extra
// printf, ran 600 renders clean; compute-sanitizer initcheck / racecheck found nothing in the source).  Cause not found
// (ptxas 12.9 spill code is the suspect), so the spectral tail is built without spills: 1 CTA per SM.
template <int MODE, bool SPEC>
__global__ void __launch_bounds__(WF_THREADS, SPEC ? 1 : 2) k_tail(WfArgs a, int depth) {
    grid_dep_wait(); grid_dep_launch();
    if (a.ctr->tail_from != depth) return;
    __shared__ unsigned long long bar;
    TreeView tv; int* stack_base;
    tree_stage_issue<MODE>(a, &bar);
    tree_stage_wait<MODE>(a, &bar, tv, stack_base);
    const BatchParams bp = *a.bp;
    const int n = a.ctr->nq[depth], pp = depth & 1;
    const int lane = threadIdx.x & 31, lane8 = lane & 7;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int chunk = min(32, max(a.tail_chunk, (n + nwarps - 1) / nwarps));
    int* cursor = &a.ctr->wf_trace[TR_MAX_DEPTH_CAP];                  // free slot: stage cursors use [0, max_depth)
    unsigned long long n_closest = 0, n_shadow = 0;
#ifdef TR_COUNTERS
    unsigned long long cnt[4] = {0, 0, 0, 0};
#endif
    int mode = 0, d = depth, q = 0;                                    // 0 idle, 1 closest-hit walk, 2 shadow walk
    float4 A = make_float4(0.f, 0.f, 0.f, 1.f), B = make_float4(1.f, 1.f, 1.f, 0.f), C = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 nA = A, nB = B, nC = C, sC = C; bool st_cont = false; unsigned sslot = 0;
    RayPre r = make_ray(mk3(0.f, 0.f, 0.f), mk3(1.f, 1.f, 1.f)); bool anypar = false;
    int cur = TR_DONE, tleaf = 0; bool visible = false, found = false; float tt = TR_INF;
    SmemStack st; st.init(stack_base);
    HitRec h; hit_reset(h);
    bool more = true;
    while (true) {
        if (__ballot_sync(0xffffffffu, mode != 0) == 0u) {
            if (!more) break;
            int base = 0;
            if (lane == 0) base = atomicAdd(cursor, chunk);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= n) break;
            if (base + chunk >= n) more = false;
            q = base + lane;
            if (lane < chunk && q < n) {
                A = a.pa[pp][q]; B = a.pb[pp][q]; C = a.pc[pp][q]; d = depth; mode = 1;
                r = make_ray(mk3(A.x, A.y, A.z), mk3(A.w, B.x, B.y)); anypar = r.px || r.py || r.pz;
                hit_reset(h); st.reset(); cur = root_enter(a.root, r, anypar);
            }
        }
        // ---- node steps (same code as in k_trace / k_shadow; the mode only selects the pruning bound and the target)
#pragma unroll
        for (int step = 0; step < NodeSteps<MODE>::value; ++step) {
            if (mode != 0 && (unsigned)cur < (unsigned)TR_DONE) {
#ifdef TR_COUNTERS
                ++cnt[mode == 1 ? 0 : 2];
#endif
                if (mode == 1) cur = node_step<MODE, false>(tv, r, anypar, h.t, 0, found, st, cur, lane8);
                else cur = node_step<MODE, true>(tv, r, anypar, tt, -tleaf - 1, found, st, cur, lane8);
            }
        }
#ifdef TR_DEBUG_WALK
        if (mode != 0 && cur != TR_DONE && (cur >= a.nint || cur < -a.nleaves || st.sp < 0 || st.sp > a.stack_cap)) {
            printf("k_tail bad link %d (nint %d nleaves %d) sp %d cap %d mode %d d %d q %d n %d tid %d blk %d tleaf %d\n", cur, a.nint, a.nleaves, st.sp, a.stack_cap, mode, d, q, n, threadIdx.x, blockIdx.x, tleaf);
            cur = TR_DONE; mode = 0;
        }
#endif
        // ---- leaf step
        if (mode != 0 && cur < 0) {
            const int k = -cur - 1;
            float4 la, lb, lc; leaf_fetch<MODE>(tv, k, lane8, la, lb, lc);
            float u, v, t = intersect_leaf(r, la, lb, lc, u, v);
#ifdef TR_COUNTERS
            ++cnt[mode == 1 ? 1 : 3];
#endif
            cur = st.pop();
            if (mode == 1) { if (closer(t, k, h.t, h.leaf)) { h.t = t; h.u = u; h.v = v; h.prim = __float_as_int(la.w); h.mat = __float_as_int(lc.w); h.leaf = k; } }
            else if (blocks_target(t, k, tt, tleaf)) { visible = false; cur = TR_DONE; }
        }
        // ---- a finished walk: shade / add the NEE term, then start the lane's next walk
        if (mode != 0 && cur == TR_DONE) {
            bool advance = true;
            if (mode == 1) {
                ++n_closest;
                if (a.probe) { a.hit[q] = make_float4(h.t, __int_as_float(h.prim), h.u, h.v); mode = 0; continue; }      // test hook: the first walk only
                int cl = 0;
                if (h.prim >= 0) {
                    int mt = (int)__ldg(a.material + (size_t)h.mat * 10);
                    cl = (mt == TR_MAT_LIGHT) ? 0 : (mt == TR_MAT_GLASS ? 2 : 1);
                }
                ShadeOut o; o.cont = false; o.shadow = false;
                shade_any<SPEC>(a, bp, d, cl, A, B, C, make_float4(h.t, __int_as_float(h.prim), h.u, h.v), o);
                st_cont = o.cont; nA = o.nA; nB = o.nB; nC = o.nC;
                if (o.shadow) {
                    ++n_shadow;
                    r = make_ray(mk3(o.sA.x, o.sA.y, o.sA.z), mk3(o.sA.w, o.sB.x, o.sB.y)); anypar = r.px || r.py || r.pz;
                    tleaf = __ldg(a.leaf_of_prim + __float_as_int(o.sB.z)); sslot = __float_as_uint(o.sB.w); sC = o.sC;
                    float4 la, lb, lc; leaf_fetch<MODE>(tv, tleaf, lane8, la, lb, lc);
#ifdef TR_COUNTERS
                    ++cnt[3];
#endif
                    st.reset(); cur = shadow_enter(tv, a.root, r, anypar, tleaf, la, lb, lc, tt, visible, found);
                    mode = 2; advance = false;
                }
            } else if (visible && found) {
                float4 Lv = a.Lnee[sslot]; Lv.x += sC.x; Lv.y += sC.y; Lv.z += sC.z; Lv.w += sC.w; a.Lnee[sslot] = Lv;
            }
            if (advance) {
                if (st_cont) {
                    A = nA; B = nB; C = nC; ++d; mode = 1; st_cont = false;
                    r = make_ray(mk3(A.x, A.y, A.z), mk3(A.w, B.x, B.y)); anypar = r.px || r.py || r.pz;
                    hit_reset(h); st.reset(); cur = root_enter(a.root, r, anypar);
                } else mode = 0;
            }
        }
    }
    if (n_closest) atomicAdd(&a.ctr->tail_rays[0], n_closest);
    if (n_shadow) atomicAdd(&a.ctr->tail_rays[1], n_shadow);
#ifdef TR_COUNTERS
    atomicAdd(a.ctr->visits + 0, 2ull * cnt[0]); atomicAdd(a.ctr->visits + 1, cnt[1]);
    atomicAdd(a.ctr->visits + 2, 2ull * cnt[2]); atomicAdd(a.ctr->visits + 3, cnt[3]);
#endif
}

// ------------------------------------------------------------------ shadow
// Same persistent schedule as k_trace; the walk is bounded by the target's own t and stops at the first
// primitive that would have won the reference's nearest-hit comparison (see trace_shadow_visible).
// QUERY (BDPT connections): instead of adding a contribution, report per queue item (sb.w) the distance to the target when it is
// the nearest hit, -1 otherwise.
#ifndef WF_SHADOW_MIN_BLOCKS
#define WF_SHADOW_MIN_BLOCKS 5      // 48 registers, no spills: 5 CTAs per SM (C2 shadow stages 5.38 -> 5.31 ms)
#endif
template <int MODE, bool QUERY = false>
__global__ void __launch_bounds__(WF_THREADS, WF_SHADOW_MIN_BLOCKS) k_shadow(WfArgs a, int depth) {
    grid_dep_wait(); grid_dep_launch();
    if (tail_took_over(a, depth)) return;
    const int n = a.ctr->nshadow[depth];
    if ((long long)blockIdx.x * WF_THREADS >= (long long)n) return;          // surplus CTAs leave before staging the tree
    __shared__ unsigned long long bar;
    TreeView tv; int* stack_base;
    tree_stage_issue<MODE>(a, &bar);
    tree_stage_wait<MODE>(a, &bar, tv, stack_base);
    int* cursor = &a.ctr->wf_shadow[depth];
    const int lane = threadIdx.x & 31, lane8 = lane & 7;
    WarpFeed feed = make_feed(n);
    unsigned idle = 0xffffffffu;
    bool anypar = false, visible = false, found = false; int q = 0, cur = TR_DONE, tleaf = 0;
    float tt = TR_INF; unsigned slot = 0;
    SmemStack st; st.init(stack_base);
    RayPre r = make_ray(mk3(0.f, 0.f, 0.f), mk3(1.f, 1.f, 1.f));
#ifdef TR_COUNTERS
    unsigned long long cnt_nodes = 0, cnt_leaves = 0;
#endif
    while (true) {
        if ((__popc(idle) >= RefillMin<MODE>::value || idle == 0xffffffffu) && (feed.more || feed.cb < feed.ce)) {
            int nq = feed_lanes(feed, cursor, n, idle, lane);
            if (nq >= 0) {
                float4 A = a.sa[depth & 1][nq], B = a.sb[depth & 1][nq];
                r = make_ray(mk3(A.x, A.y, A.z), mk3(A.w, B.x, B.y));
                anypar = r.px || r.py || r.pz;
                tleaf = __ldg(a.leaf_of_prim + __float_as_int(B.z)); slot = __float_as_uint(B.w);
                float4 la, lb, lc; leaf_fetch<MODE>(tv, tleaf, lane8, la, lb, lc);
                TR_COUNT(cnt_leaves);
                q = nq; st.reset(); cur = shadow_enter(tv, a.root, r, anypar, tleaf, la, lb, lc, tt, visible, found);
            }
        }
        if (idle == 0xffffffffu) { if (!feed.more && feed.cb >= feed.ce) break; else continue; }
        const bool has = !((idle >> lane) & 1u);
        const int tlink = -tleaf - 1;
#pragma unroll
        for (int step = 0; step < NodeSteps<MODE>::value; ++step) {
            if (has && (unsigned)cur < (unsigned)TR_DONE) {
                TR_COUNT(cnt_nodes);
                cur = node_step<MODE, true>(tv, r, anypar, tt, tlink, found, st, cur, lane8);
            }
        }
        if (has && cur < 0) {
            TR_COUNT(cnt_leaves);
            const int k = -cur - 1;
            float4 la, lb, lc; leaf_fetch<MODE>(tv, k, lane8, la, lb, lc);
            float u, v, t = intersect_leaf(r, la, lb, lc, u, v);
            if (blocks_target(t, k, tt, tleaf)) { visible = false; cur = TR_DONE; } else cur = st.pop();
        }
        const unsigned finm = __ballot_sync(0xffffffffu, has && cur == TR_DONE);
        if (QUERY) { if ((finm >> lane) & 1u) a.vis[slot] = (visible && found) ? tt : -1.0f; }
        else if ((finm >> lane) & 1u) {
            if (visible && found) {
                float4 C = a.sc[depth & 1][q];
                float4 Lv = a.Lnee[slot]; Lv.x += C.x; Lv.y += C.y; Lv.z += C.z; Lv.w += C.w; a.Lnee[slot] = Lv;   // .w: 4th hero lane (0 for PT_RGB)
            }
        }
        idle |= finm;
    }
#ifdef TR_COUNTERS
    atomicAdd(a.ctr->visits + 2, 2ull * cnt_nodes); atomicAdd(a.ctr->visits + 3, cnt_leaves);
#endif
}

// ------------------------------------------------------------------ accumulate (PT_RGB.py:134-136)
// SPEC: AddSplat of PT_Spec (integrator/PT_Spec.py:148-165,279-280)
template <bool SPEC>
__global__ void __launch_bounds__(WF_THREADS) k_accumulate(WfArgs a) {
    const BatchParams bp = *a.bp;
    const int stride = gridDim.x * blockDim.x;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < a.npix; p += stride) {
        int x, y;
        if (!slot_to_pixel(a, p, x, y)) continue;
        float* o = a.hdr + ((size_t)x * a.H + y) * 3;
        float r = o[0], g = o[1], b = o[2];
        for (int f = 0; f < bp.n_frames; ++f) {
            // radiance = (NEE terms in depth order) + terminal term: the reference's summation order (PT_RGB.py:76-109,131)
            float4 Lt = a.L[(size_t)f * a.npix + p], Lv = a.Lnee[(size_t)f * a.npix + p];
            Lv.x += Lt.x; Lv.y += Lt.y; Lv.z += Lt.z;
            float coff = 1.0f / ((float)(bp.frame_begin + f) + 1.0f);
            if (SPEC) {
                Lv.w += Lt.w;
                unsigned pix = ((unsigned)x << 16) | (unsigned)y;
                float Lambda = HERO_LAMBDA_MIN + HERO_STEP * rng4(bp.seed, pix, (unsigned)(bp.frame_begin + f), 0u).z;
                add_splat(a.spec, ld4(Lv), Lambda, coff, r, g, b);
                continue;
            }
            r = Lv.x * coff + r * (1.0f - coff);
            g = Lv.y * coff + g * (1.0f - coff);
            b = Lv.z * coff + b * (1.0f - coff);
        }
        o[0] = r; o[1] = g; o[2] = b;
    }
}

// ------------------------------------------------------------------ tone map (UtilsFunc.py:583-586)
__global__ void k_tonemap(const float* __restrict__ hdr, float* __restrict__ rgb, int n, float exposure) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rgb[i] = lrgb_to_srgb1(tone_aces1(hdr[i] * exposure));
}

// ------------------------------------------------------------------ Debug integrator + first hits
// integrator/Debug.py:44-66; one thread per pixel, frame-0 rays (no jitter)
__global__ void k_debug(WfArgs a, float* __restrict__ fh) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = p < a.W * a.H;
    if (!active) p = 0;                       // whole warps stay in the (warp-synchronous) traversal
    int i = p / a.H, j = p - i * a.H;
    V3 o = mk3(a.cam.eye[0], a.cam.eye[1], a.cam.eye[2]), d = camera_dir(a.cam, i, j, 0.0f, 0.0f);
    RayPre r = make_ray(o, d);
    TreeView tv; tv.snodes = tv.sleaves = nullptr; tv.gnodes = a.nodes2; tv.gleaves = a.leaves4; tv.top = 0;
    HitRec h = trace_closest(tv, a.root, r, active, a.ctr->visits);
    if (!active) return;
    float* f = fh + (size_t)p * 16;
    V3 col = mk3(0, 0, 0), pos = mk3(0, 0, 0), gn = mk3(0, 0, 0), nn = mk3(0, 0, 0);
    if (h.prim >= 0) {
        Surf s = surface_at(a, h.prim, h.u, h.v, o, d, h.t);
        const float* m = a.material + (size_t)s.mat * 10;
        col = mk3(m[2], m[3], m[4]); pos = s.pos; gn = s.gn; nn = s.n;
    }
    f[0] = h.t; f[1] = __int_as_float(h.prim); f[2] = h.u; f[3] = h.v;
    f[4] = pos.x; f[5] = pos.y; f[6] = pos.z; f[7] = gn.x; f[8] = gn.y; f[9] = gn.z;
    f[10] = nn.x; f[11] = nn.y; f[12] = nn.z; f[13] = d.x; f[14] = d.y; f[15] = d.z;
    float* hd = a.hdr + (size_t)p * 3; hd[0] = col.x; hd[1] = col.y; hd[2] = col.z;
}

// ------------------------------------------------------------------ shading table
// per-primitive record with precomputed area (Scene.get_prim_area, Scene.py:324-350)
__global__ void k_shade_table(const float* __restrict__ vertex, const int* __restrict__ prim, const float* __restrict__ shape,
                              int n, TrShade* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int type = prim[i * 3], vi = prim[i * 3 + 1], mat = prim[i * 3 + 2];
    TrShade s;
    if (type == TR_PRIM_TRI) {
        const float* p = vertex + (size_t)vi * 9;
        V3 v1 = mk3(p[0], p[1], p[2]), v2 = mk3(p[9], p[10], p[11]), v3 = mk3(p[18], p[19], p[20]);
        float la = length3(v1 - v2), lb = length3(v1 - v3), lc = length3(v3 - v2);
        float sum = ((la + lb) + lc) * 0.5f;
        float area = sqrtf(sum * (sum - la) * (sum - lb) * (sum - lc));
        s.q[0] = make_float4(v1.x, v1.y, v1.z, __int_as_float(mat));
        s.q[1] = make_float4(v2.x, v2.y, v2.z, __int_as_float(0));
        s.q[2] = make_float4(v3.x, v3.y, v3.z, area);
        // geometric normal of Scene.intersect_prim (Scene.py:547-549), same expression as the per-hit evaluation it replaces
        V3 v13 = v3 - v1, v12 = v2 - v1;
        V3 gn = normalize3(cross3(v12, v13));
        s.q[3] = make_float4(p[3], p[4], p[5], gn.x);
        s.q[4] = make_float4(p[12], p[13], p[14], gn.y);
        s.q[5] = make_float4(p[21], p[22], p[23], gn.z);
    } else {
        const float* sp = shape + (size_t)vi * 10;
        int st = (int)sp[0];
        float area = 0.0f;
        if (st == TR_SHAPE_SPHERE || st == TR_SHAPE_SPOT || st == TR_SHAPE_LASER) area = sp[4] * sp[4] * TR_PI_ENV;
        s.q[0] = make_float4(sp[1], sp[2], sp[3], __int_as_float(mat));
        // kind: 1 sphere (q1.x radius), 2 spot (q1 normal, q2 = xita1, xita2, scale), 3 laser (q1 normal, q2.x radius), 4 others
        if (st == TR_SHAPE_SPHERE) s.q[1] = make_float4(sp[4], 0.0f, 0.0f, __int_as_float(1));
        else s.q[1] = make_float4(sp[7], sp[8], sp[9], __int_as_float(st == TR_SHAPE_SPOT ? 2 : (st == TR_SHAPE_LASER ? 3 : 4)));
        s.q[2] = make_float4(st == TR_SHAPE_SPHERE ? 0.0f : sp[4], st == TR_SHAPE_SPOT ? sp[5] : 0.0f, st == TR_SHAPE_SPOT ? sp[6] : 0.0f, area);
        s.q[3] = s.q[4] = s.q[5] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    out[i] = s;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

int tr_build_shade_table(tr_ctx* ctx) {
    if (ctx->shade_ready) return TR_OK;
    int rc; if ((rc = tr_realloc(ctx, &ctx->d_shade, (size_t)ctx->np))) return rc;
    k_shade_table<<<cdiv(ctx->np, 256), 256, 0, ctx->stream>>>(ctx->d_vertex, ctx->d_prim, ctx->d_shape, ctx->np, ctx->d_shade);
    TR_CHECK_LAUNCH(ctx);
    ctx->shade_ready = true;
    return TR_OK;
}

// linear albedo per material: UF.srgb_to_lrgb(material colour) (integrator/PT_RGB.py:86), same device function
__global__ void k_matlin(const float* __restrict__ material, int nm, float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nm) return;
    V3 c = srgb_to_lrgb(mk3(material[i * 10 + 2], material[i * 10 + 3], material[i * 10 + 4]));
    out[i] = make_float4(c.x, c.y, c.z, 0.0f);
}

// the tree as the traversal kernels see it (mode-independent part; launch_cfg picks the mode and the staging sizes)
static void tr_tree_args(tr_ctx* ctx, WfArgs& a) {
    a.nodes2 = (const float4*)ctx->d_nodes2; a.leaves4 = (const float4*)ctx->d_leaves; a.small_img = ctx->d_small_img;
    a.nint = ctx->np - 1; a.nleaves = ctx->np; a.top = 0; a.stack_cap = 0; a.stage_bytes = 0;
    a.root.lo = make_float4(ctx->root_box[0], ctx->root_box[1], ctx->root_box[2], 0.0f);
    a.root.hi = make_float4(ctx->root_box[3], ctx->root_box[4], ctx->root_box[5], 0.0f);
    a.root.link = ctx->np > 1 ? 0 : -1;
}

static int fill_args(tr_ctx* ctx, WfArgs& a, bool spec = false) {
    if (!ctx->bvh_ready) return tr_fail(ctx, TR_ERR_INVALID, "render: BVH not built (call tr_bvh_build)");
    if (!ctx->cam_set) return tr_fail(ctx, TR_ERR_INVALID, "render: camera not set");
    if (!ctx->d_hdr) return tr_fail(ctx, TR_ERR_INVALID, "render: film not created");
    int rc;
    if ((rc = tr_build_shade_table(ctx))) return rc;
    if ((rc = tr_build_tiles(ctx))) return rc;
    if (!ctx->matlin_ready) {
        if ((rc = tr_realloc(ctx, &ctx->d_matlin, (size_t)ctx->nm))) return rc;
        k_matlin<<<cdiv(ctx->nm, 64), 64, 0, ctx->stream>>>(ctx->d_material, ctx->nm, ctx->d_matlin);
        TR_CHECK_LAUNCH(ctx);
        ctx->matlin_ready = true; ctx->gen++;
    }
    memset(&a, 0, sizeof(a));
    tr_tree_args(ctx, a);
    a.shade = ctx->d_shade; a.material = ctx->d_material; a.matlin = ctx->d_matlin; a.light = ctx->d_light; a.nl = ctx->nl; a.leaf_of_prim = ctx->d_leaf_of_prim;
    a.env = ctx->d_env; a.env_w = ctx->d_env ? ctx->env_w : 0; a.env_h = ctx->env_h; a.env_power = ctx->env_power;
    a.cam = ctx->cam; a.W = ctx->W; a.H = ctx->H; a.tiles = ctx->d_tiles; a.npix = ctx->n_local_tiles * TR_TILE * TR_TILE;
    a.hdr = ctx->d_hdr;
    for (int k = 0; k < 2; ++k) { a.pa[k] = ctx->d_path[k][0]; a.pb[k] = ctx->d_path[k][1]; a.pc[k] = ctx->d_path[k][2]; }
    a.hit = ctx->d_hit; a.cls = ctx->d_cls; a.cap = ctx->wf_cap;
    for (int k = 0; k < 2; ++k) { a.sa[k] = ctx->d_shq[k][0]; a.sb[k] = ctx->d_shq[k][1]; a.sc[k] = ctx->d_shq[k][2]; }
    a.L = ctx->d_L; a.Lnee = ctx->d_Lnee;
    a.ctr = ctx->d_ctr; a.bp = (const BatchParams*)ctx->d_batch_params;
    a.frame_off = 0; a.sub_frames = 1 << 20; a.tail_max = ctx->opt_tail_max < 0 ? 4096 : ctx->opt_tail_max; a.tail_chunk = ctx->opt_tail_chunk < 1 ? 1 : ctx->opt_tail_chunk;
    if (spec) { if ((rc = tr_spec_prepare(ctx))) return rc; a.spec = ctx->spec; }
    return TR_OK;
}

static int ensure_wavefront(tr_ctx* ctx, size_t slots) {
    if (slots <= ctx->wf_cap) return TR_OK;
    int rc;
    ctx->wf_cap = 0;                 // a growth that fails half-way (out of memory) must not leave the old capacity standing over freed buffers
    for (int k = 0; k < 2; ++k) for (int j = 0; j < 3; ++j) if ((rc = tr_realloc(ctx, &ctx->d_path[k][j], slots))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_hit, slots))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_cls, slots * 3))) return rc;
    for (int k = 0; k < 2; ++k) for (int j = 0; j < 3; ++j) if ((rc = tr_realloc(ctx, &ctx->d_shq[k][j], slots))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_L, slots))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_Lnee, slots))) return rc;
    ctx->wf_cap = slots; ctx->gen++;
    return TR_OK;
}

struct LaunchCfg { int mode, grid_trace, grid_shadow, grid_simple, grid_tail[2]; size_t smem; };   // grid_tail[SPEC]   // memset before use (compared bytewise)

// instantiate-and-dispatch on the tree mode
#define TR_MODE_SWITCH(mode, CALL) do { switch (mode) { \
    case TM_REP: { constexpr int M = TM_REP; CALL; } break; case TM_SMEM: { constexpr int M = TM_SMEM; CALL; } break; \
    case TM_GTOP: { constexpr int M = TM_GTOP; CALL; } break; default: { constexpr int M = TM_GLOBAL; CALL; } break; } } while (0)

template <typename K> static int kernel_blocks(tr_ctx* ctx, K kern, size_t smem, int& blocks) {
    TR_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, WF_THREADS, smem));
    if (blocks < 1) blocks = 1;
    return TR_OK;
}

// Picks how the traversal kernels read the tree and sizes their dynamic shared memory:
//   TM_REP    whole tree, 8-way replicated conflict-free image      (image <= 40 KB: Cornell box 31 KB)
//   TM_SMEM   whole tree, plain                                     (nodes + leaves <= 64 KB)
//   TM_GTOP   global memory + the breadth-first top `top_nodes` nodes staged (option "top_nodes" > 0)
//   TM_GLOBAL global memory
// plus stack_cap x 1 KB for the traversal stacks (stack_cap = min(entries the tree needs, TR_STACK_SMEM)).
static int launch_cfg(tr_ctx* ctx, WfArgs& a, LaunchCfg& c) {
    const size_t rep_bytes = ((size_t)a.nint * 4 + (size_t)a.nleaves * 3) * 128, plain_bytes = (size_t)a.nint * 64 + (size_t)a.nleaves * 48;
    a.top = 0;
    if (ctx->opt_smem_bvh && ctx->opt_replicas && ctx->d_small_img && rep_bytes <= TR_SMALL_IMG_MAX) { c.mode = TM_REP; a.stage_bytes = (unsigned)rep_bytes; }
    else if (ctx->opt_smem_bvh && plain_bytes <= 64 * 1024) { c.mode = TM_SMEM; a.stage_bytes = (unsigned)((plain_bytes + 127) & ~(size_t)127); }
    else if (ctx->opt_top_nodes > 0 && ctx->top_count > 0) {
        c.mode = TM_GTOP; a.top = ctx->opt_top_nodes < ctx->top_count ? ctx->opt_top_nodes : ctx->top_count; a.stage_bytes = (unsigned)a.top * 64u;
    } else { c.mode = TM_GLOBAL; a.stage_bytes = 0; }
    a.stack_cap = ctx->stack_need < 1 ? 1 : ctx->stack_need;                          // entries per lane: what the tree needs (<= TR_STACK_MAX)
    const size_t stack_bytes = (size_t)a.stack_cap * WF_THREADS * sizeof(int);
    if (c.mode != TM_GLOBAL && (size_t)a.stage_bytes + stack_bytes > 200 * 1024) { c.mode = TM_GLOBAL; a.stage_bytes = 0; a.top = 0; }   // a very deep tree: the stack takes the shared memory
    c.smem = (size_t)a.stage_bytes + stack_bytes;
    int bt = 0, bs = 0, t0 = 0, t1 = 0, rc = TR_OK;
    TR_MODE_SWITCH(c.mode, {
        if (!rc) rc = kernel_blocks(ctx, k_trace<M>, c.smem, bt);
        if (!rc) rc = kernel_blocks(ctx, k_shadow<M, false>, c.smem, bs);
        if (!rc) rc = kernel_blocks(ctx, k_tail<M, false>, c.smem, t0);
        if (!rc) rc = kernel_blocks(ctx, k_tail<M, true>, c.smem, t1);
    });
    if (rc) return rc;
    c.grid_tail[0] = ctx->num_sms * t0; c.grid_tail[1] = ctx->num_sms * t1;        // the tail kernel deals its paths over the warps that are resident at once
    if (ctx->opt_persist_blocks > 0) { bt = bt < ctx->opt_persist_blocks ? bt : ctx->opt_persist_blocks; bs = bs < ctx->opt_persist_blocks ? bs : ctx->opt_persist_blocks; }   // leave SM slots to the other chain's kernels
    c.grid_trace = ctx->num_sms * bt; c.grid_shadow = ctx->num_sms * bs;
    c.grid_simple = ctx->num_sms * 8;
    return TR_OK;
}

// per-chain view of the queues: chain j renders local frames [f0, f0 + nfr) of the batch with its own queue region + counters
static WfArgs chain_args(const WfArgs& a, int j, int f0, int nfr) {
    WfArgs c = a;
    size_t off = (size_t)f0 * a.npix;
    for (int k = 0; k < 2; ++k) { c.pa[k] = a.pa[k] + off; c.pb[k] = a.pb[k] + off; c.pc[k] = a.pc[k] + off; }
    c.hit = a.hit + off; c.cls = a.cls + 3 * off; c.cap = (size_t)nfr * a.npix;
    for (int k = 0; k < 2; ++k) { c.sa[k] = a.sa[k] + off; c.sb[k] = a.sb[k] + off; c.sc[k] = a.sc[k] + off; }
    c.ctr = a.ctr + j; c.frame_off = f0; c.sub_frames = nfr;
    return c;
}

// Launch of a wavefront stage.  pdl: programmatic dependent launch (cudaLaunchAttributeProgrammaticStreamSerialization): the
// CTAs of this kernel may be scheduled while the previous kernel of the stream drains (every stage kernel triggers its dependents
// at its very start and waits for its predecessor with griddepcontrol.wait before it reads anything), which hides the launch
// latency between the ~45 dependent stages of a chain.
template <typename... KArgs, typename... Args>
static inline void wf_launch(bool pdl, void (*kern)(KArgs...), int grid, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(WF_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at; memset(&at, 0, sizeof(at));
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization; at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at; cfg.numAttrs = pdl ? 1u : 0u;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// one chain: generate, then max_depth x (trace, shade) on stream s, with shadow(d) on stream ss:
//   shadow(d) needs shade(d) (its queue) and shadow(d-1) (it sums the NEE terms in depth order);
//   shade(d+2) needs shadow(d) (the shadow queue is a ping-pong pair).
// So the long tail of a trace stage overlaps the previous depth's shadow rays.  ss == s serialises everything.
// ev != nullptr (stage timing, single chain, non-graph mode): 4 events per depth bracket trace / shade / shadow.
template <bool SPEC>
static int enqueue_chain(tr_ctx* ctx, const WfArgs& a, const LaunchCfg& c, int max_depth, cudaStream_t s, cudaStream_t ss,
                         cudaEvent_t* dep, uint64_t* launches, cudaEvent_t* ev) {
    const bool split = (ss != s);
    TR_CUDA(ctx, cudaMemsetAsync(a.ctr, 0, sizeof(TrCounters), s));
    k_generate<SPEC><<<c.grid_simple, WF_THREADS, 0, s>>>(a); ++*launches;
    for (int d = 0; d < max_depth; ++d) {
        if (ev) cudaEventRecord(ev[4 * d + 0], s);
        const bool pdl = ctx->opt_pdl != 0 && !ev;
        TR_MODE_SWITCH(c.mode, wf_launch(pdl && d > 0, k_trace<M>, c.grid_trace, c.smem, s, a, d));      // depth 0 follows k_generate (no trigger there)
        if (ev) cudaEventRecord(ev[4 * d + 1], s);
        if (split && d >= 2) TR_CUDA(ctx, cudaStreamWaitEvent(s, dep[2 * (d - 2) + 1], 0));
        wf_launch(pdl, k_shade<SPEC>, c.grid_simple, 0, s, a, d);
        if (ev) cudaEventRecord(ev[4 * d + 2], s);
        if (a.nl > 0) {
            if (split) { TR_CUDA(ctx, cudaEventRecord(dep[2 * d], s)); TR_CUDA(ctx, cudaStreamWaitEvent(ss, dep[2 * d], 0)); }
            TR_MODE_SWITCH(c.mode, wf_launch(pdl && !split, k_shadow<M, false>, c.grid_shadow, c.smem, ss, a, d));   // on its own stream it follows an event wait
            if (split) TR_CUDA(ctx, cudaEventRecord(dep[2 * d + 1], ss));
            ++*launches;
        }
        if (a.tail_max > 0 && d + 1 < max_depth) {
            // behind shadow(d) on the same stream: NEE terms stay in depth order; shade(d) (the hand-over decision) is done
            if (split && a.nl == 0) { TR_CUDA(ctx, cudaEventRecord(dep[2 * d], s)); TR_CUDA(ctx, cudaStreamWaitEvent(ss, dep[2 * d], 0)); }
            TR_MODE_SWITCH(c.mode, wf_launch(pdl && a.nl > 0, k_tail<M, SPEC>, c.grid_tail[SPEC], c.smem, ss, a, d + 1));   // follows k_shadow(d) on ss
            ++*launches;
        }
        if (ev) cudaEventRecord(ev[4 * d + 3], s);
        *launches += 2;
    }
    if (split) {                                                                           // join: everything enqueued on ss
        TR_CUDA(ctx, cudaEventRecord(dep[2 * max_depth], ss));
        TR_CUDA(ctx, cudaStreamWaitEvent(s, dep[2 * max_depth], 0));
    }
    TR_CHECK_LAUNCH(ctx);
    return TR_OK;
}

// The launch sequence of one batch (captured into a CUDA graph when opt_graph is on): K independent chains over
// disjoint frame ranges run as parallel branches, so the tail of one chain's stage (a few slow rays) overlaps the
// next stage of another chain; the frame-ordered running mean (k_accumulate) joins them.
template <bool SPEC>
static int enqueue_batch(tr_ctx* ctx, const WfArgs& a, const LaunchCfg& c, int max_depth, int K, int fs, cudaStream_t s,
                         uint64_t* launches, cudaEvent_t* ev) {
    int rc;
    if (K > 1) {
        TR_CUDA(ctx, cudaEventRecord(ctx->ev_fork, s));
        for (int j = 1; j < K; ++j) TR_CUDA(ctx, cudaStreamWaitEvent(ctx->sub_stream[j], ctx->ev_fork, 0));
    }
    // frames per chain: even split, or (two chains, option "chain_skew" = percent of the frames for chain 0) an uneven one, so that
    // the chains do not reach their thin, latency-bound deep stages at the same time
    int f0 = 0;
    for (int j = 0; j < K; ++j) {
        int nfr = fs;
        if (K == 2 && ctx->opt_chain_skew > 0) { const int n0 = max(1, min(2 * fs - 1, (2 * fs * ctx->opt_chain_skew + 50) / 100)); nfr = j == 0 ? n0 : 2 * fs - n0; }
        WfArgs cj = chain_args(a, j, f0, nfr);
        f0 += nfr;
        cudaStream_t sj = (j == 0) ? s : ctx->sub_stream[j];
        cudaStream_t ssj = (ev || !ctx->opt_shadow_overlap) ? sj : ctx->shadow_stream[j];
        if ((rc = enqueue_chain<SPEC>(ctx, cj, c, max_depth, sj, ssj, ctx->dep_ev.data() + (size_t)j * 2 * (TR_MAX_DEPTH_CAP + 1), launches, K == 1 ? ev : nullptr))) return rc;
    }
    for (int j = 1; j < K; ++j) {
        TR_CUDA(ctx, cudaEventRecord(ctx->ev_join[j], ctx->sub_stream[j]));
        TR_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_join[j], 0));
    }
    k_accumulate<SPEC><<<c.grid_simple, WF_THREADS, 0, s>>>(a); ++*launches;
    TR_CHECK_LAUNCH(ctx);
    return TR_OK;
}

// the wavefront driver shared by PT_RGB (SPEC = false) and PT_Spec (SPEC = true)
template <bool SPEC>
static int render_wavefront(tr_ctx* ctx, int frame_begin, int n_frames, int max_depth, uint64_t seed) {
    if (!ctx || n_frames <= 0 || frame_begin < 0 || max_depth <= 0 || max_depth > TR_MAX_DEPTH_CAP)
        return tr_fail(ctx, TR_ERR_INVALID, "%s: bad arguments (frames %d+%d, depth %d)", SPEC ? "tr_render_pt_spec" : "tr_render_pt_rgb", frame_begin, n_frames, max_depth);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    WfArgs a; int rc;
    if ((rc = tr_stats_resolve(ctx))) return rc;                        // an earlier asynchronous render still owns ev0 / ev1 / the ring
    ctx->present_sum = false;                                           // the partial film changes: a reduced copy is stale
    if (!ctx->h_ring) TR_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_ring, sizeof(TrCounters) * TR_MAX_CHAINS * TR_RING_BATCHES, cudaHostAllocDefault));
    if ((rc = fill_args(ctx, a, SPEC))) return rc;
    if (a.npix == 0) { memset(&ctx->stats, 0, sizeof(ctx->stats)); ctx->stats.frames = n_frames; return TR_OK; }    // this rank owns no tile
    // frames per batch: enough paths in flight to fill the chip a few times, bounded by max_paths
    int F = ctx->opt_batch_frames;
    if (F <= 0) { F = (int)(ctx->opt_max_paths / (size_t)a.npix); if (F < 1) F = 1; }     // auto: as many frames per batch as the path budget allows
    while (F > 1 && (size_t)F * a.npix > ctx->opt_max_paths) --F;
    if (F > n_frames) F = n_frames;
    if (ctx->opt_batch_frames <= 0) {                                  // auto: equal batches instead of full ones plus a small rest (64 spp of C3: 4 x 16, not 3 x 20 + 4)
        const int nb = (n_frames + F - 1) / F;
        F = (n_frames + nb - 1) / nb;
    }
    const bool timing = ctx->opt_stage_timing != 0;
    int K = timing ? 1 : ctx->opt_chains;
    if (K < 1) K = 1; if (K > TR_MAX_CHAINS) K = TR_MAX_CHAINS; if (K > F) K = F;
    const int fs = (F + K - 1) / K;                                   // frames per chain
    if ((rc = ensure_wavefront(ctx, (size_t)fs * K * a.npix))) return rc;
    if ((rc = fill_args(ctx, a, SPEC))) return rc;
    LaunchCfg cfg; memset(&cfg, 0, sizeof(cfg)); if ((rc = launch_cfg(ctx, a, cfg))) return rc;
    cudaStream_t s = ctx->stream;
    if (ctx->dep_ev.empty()) {
        ctx->dep_ev.resize((size_t)TR_MAX_CHAINS * 2 * (TR_MAX_DEPTH_CAP + 1));
        for (auto& e : ctx->dep_ev) TR_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (int j = 0; j < TR_MAX_CHAINS; ++j) TR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->shadow_stream[j], cudaStreamNonBlocking));
    }
    if (K > 1 && !ctx->sub_stream[1]) {
        for (int j = 1; j < TR_MAX_CHAINS; ++j) {
            TR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->sub_stream[j], cudaStreamNonBlocking));
            TR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join[j], cudaEventDisableTiming));
        }
        TR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    }
    if (timing && ctx->stage_ev.empty()) {
        ctx->stage_ev.resize(4 * TR_MAX_DEPTH_CAP);
        for (auto& e : ctx->stage_ev) TR_CUDA(ctx, cudaEventCreate(&e));
    }
    uint64_t launches = 0, rays_c = 0, rays_s = 0, vis[4] = {0, 0, 0, 0}, n_term = 0;
    float ms_trace = 0.0f, ms_shade = 0.0f, ms_shadow = 0.0f;
    TR_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    for (int f0 = 0; f0 < n_frames; f0 += F) {
        int nf = (n_frames - f0 < F) ? n_frames - f0 : F;
        BatchParams bp; bp.frame_begin = frame_begin + f0; bp.n_frames = nf; bp.seed = seed; bp.max_depth = max_depth; bp.pad = 0;
        TR_CUDA(ctx, cudaMemcpyAsync(ctx->d_batch_params, &bp, sizeof(bp), cudaMemcpyHostToDevice, s));
        if (ctx->opt_graph && !timing) {
            // a captured graph stays valid as long as the kernel arguments it baked in are unchanged
            if (!ctx->graph_exec || ctx->graph_spec != (int)SPEC || ctx->graph_depth != max_depth || ctx->graph_chains != K || ctx->graph_fs != fs ||
                ctx->graph_args.size() != sizeof(WfArgs) + sizeof(LaunchCfg) || memcmp(ctx->graph_args.data(), &a, sizeof(WfArgs)) != 0 ||
                memcmp(ctx->graph_args.data() + sizeof(WfArgs), &cfg, sizeof(LaunchCfg)) != 0) {
                if (ctx->graph_exec) { cudaStreamSynchronize(s); cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }   // the previous (asynchronous) batch may still replay it
                cudaGraph_t g; uint64_t l2 = 0;
                TR_CUDA(ctx, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
                rc = enqueue_batch<SPEC>(ctx, a, cfg, max_depth, K, fs, s, &l2, nullptr);
                cudaError_t e = cudaStreamEndCapture(s, &g);
                if (rc) return rc;
                if (e != cudaSuccess) return tr_fail(ctx, TR_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
                TR_CUDA(ctx, cudaGraphInstantiate(&ctx->graph_exec, g, 0));
                cudaGraphDestroy(g);
                ctx->graph_depth = max_depth; ctx->graph_launches = (int)l2; ctx->graph_gen = ctx->gen; ctx->graph_chains = K; ctx->graph_fs = fs; ctx->graph_spec = (int)SPEC;
                ctx->graph_args.resize(sizeof(WfArgs) + sizeof(LaunchCfg));
                memcpy(ctx->graph_args.data(), &a, sizeof(WfArgs)); memcpy(ctx->graph_args.data() + sizeof(WfArgs), &cfg, sizeof(LaunchCfg));
            }
            TR_CUDA(ctx, cudaGraphLaunch(ctx->graph_exec, s));
            launches += (uint64_t)ctx->graph_launches;
        } else {
            if ((rc = enqueue_batch<SPEC>(ctx, a, cfg, max_depth, K, fs, s, &launches, timing ? ctx->stage_ev.data() : nullptr))) return rc;
        }
        // ray counters of this batch = queue sizes (device counters).  Asynchronous mode: snapshot into the pinned ring and go
        // on (no host round trip between batches or before whatever the caller enqueues next, e.g. the NCCL film reduce);
        // tr_stats_get folds the snapshots.  Stage timing needs the events of this batch, so it stays synchronous.
        const int bi = f0 / F;
        if (!timing && bi < TR_RING_BATCHES) {
            TR_CUDA(ctx, cudaMemcpyAsync(ctx->h_ring + (size_t)bi * TR_MAX_CHAINS, ctx->d_ctr, sizeof(TrCounters) * K, cudaMemcpyDeviceToHost, s));
            ctx->ring_batches = bi + 1;
            if (bi + 1 == TR_RING_BATCHES && f0 + F < n_frames) TR_CUDA(ctx, cudaStreamSynchronize(s));     // ring full: later batches fall back to the synchronous path
            continue;
        }
        TR_CUDA(ctx, cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(TrCounters) * K, cudaMemcpyDeviceToHost, s));
        TR_CUDA(ctx, cudaStreamSynchronize(s));
        for (int j = 0; j < K; ++j) {
            const int tf = ctx->h_ctr[j].tail_from;
            for (int d = 0; d < max_depth; ++d) { if (tf == 0 || d != tf) rays_c += (uint64_t)ctx->h_ctr[j].nq[d]; rays_s += (uint64_t)ctx->h_ctr[j].nshadow[d]; n_term += (uint64_t)ctx->h_ctr[j].ncls[d][0]; }
            rays_c += ctx->h_ctr[j].tail_rays[0]; rays_s += ctx->h_ctr[j].tail_rays[1];
            for (int k = 0; k < 4; ++k) vis[k] += ctx->h_ctr[j].visits[k];
        }
        if (timing) for (int d = 0; d < max_depth; ++d) {
            float t0 = 0, t1 = 0, t2 = 0; cudaEvent_t* e = ctx->stage_ev.data() + 4 * d;
            cudaEventElapsedTime(&t0, e[0], e[1]); cudaEventElapsedTime(&t1, e[1], e[2]); cudaEventElapsedTime(&t2, e[2], e[3]);
            ms_trace += t0; ms_shade += t1; ms_shadow += t2;
        }
    }
    TR_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    ctx->stats.rays_closest = rays_c; ctx->stats.rays_shadow = rays_s; ctx->stats.shade_terminal = n_term;
    ctx->stats.node_visits = vis[0]; ctx->stats.leaf_tests = vis[1]; ctx->stats.node_visits_shadow = vis[2]; ctx->stats.leaf_tests_shadow = vis[3];
    ctx->stats.kernel_launches = launches; ctx->stats.ms_total = 0.0f; ctx->stats.frames = n_frames; ctx->stats.paths_in_flight = (int)((size_t)F * a.npix); ctx->stats.chains = K;
    ctx->stats.ms_trace = ms_trace; ctx->stats.ms_shade = ms_shade; ctx->stats.ms_shadow = ms_shadow;
    ctx->ring_K = K; ctx->ring_depth = max_depth; ctx->ring_mode = 0; ctx->stats_pending = true;
    if (timing) return tr_stats_resolve(ctx);
    return TR_OK;
}

// Wait for the last render (its ev1) and fold the per-batch counter snapshots into ctx->stats.
int tr_stats_resolve(tr_ctx* ctx) {
    if (!ctx->stats_pending) return TR_OK;
    TR_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    ctx->stats_pending = false;
    uint64_t rays_c = 0, rays_s = 0, vis[4] = {0, 0, 0, 0}, n_term = 0;
    if (ctx->ring_mode != 0) {
        // BDPT_RGB batches: slot 0 = wavefront counters (ring_depth sub-path stages, one connection-query stage), slot 1 = the
        // BDPT counters (lock-step pipeline: closest / shadow traversals)
        for (int b = 0; b < ctx->ring_batches; ++b) {
            const TrCounters& c = ctx->h_ring[(size_t)b * TR_MAX_CHAINS];
            const unsigned long long* hc = (const unsigned long long*)(ctx->h_ring + (size_t)b * TR_MAX_CHAINS + 1);
            if (ctx->ring_mode == 1) { for (int d = 0; d < ctx->ring_depth; ++d) rays_c += (uint64_t)c.nq[d]; rays_s += (uint64_t)c.nshadow[0]; }
            else { rays_c += hc[0]; rays_s += hc[1]; }
            for (int k = 0; k < 4; ++k) vis[k] += c.visits[k];
        }
        ctx->ring_batches = 0; ctx->ring_K = 0;
    }
    for (int b = 0; b < ctx->ring_batches; ++b) for (int j = 0; j < ctx->ring_K; ++j) {
        const TrCounters& c = ctx->h_ring[(size_t)b * TR_MAX_CHAINS + j];
        const int tf = c.tail_from;
        for (int d = 0; d < ctx->ring_depth; ++d) { if (tf == 0 || d != tf) rays_c += (uint64_t)c.nq[d]; rays_s += (uint64_t)c.nshadow[d]; n_term += (uint64_t)c.ncls[d][0]; }
        rays_c += c.tail_rays[0]; rays_s += c.tail_rays[1];
        for (int k = 0; k < 4; ++k) vis[k] += c.visits[k];
    }
    ctx->ring_batches = 0;
    ctx->stats.rays_closest += rays_c; ctx->stats.rays_shadow += rays_s; ctx->stats.shade_terminal += n_term;
    ctx->stats.node_visits += vis[0]; ctx->stats.leaf_tests += vis[1]; ctx->stats.node_visits_shadow += vis[2]; ctx->stats.leaf_tests_shadow += vis[3];
    float ms = 0.0f; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->stats.ms_total = ms;
    return TR_OK;
}

extern "C" int tr_render_pt_rgb(tr_ctx* ctx, int frame_begin, int n_frames, int max_depth, uint64_t seed) {
    return render_wavefront<false>(ctx, frame_begin, n_frames, max_depth, seed);
}
extern "C" int tr_render_pt_spec(tr_ctx* ctx, int frame_begin, int n_frames, int max_depth, uint64_t seed) {
    return render_wavefront<true>(ctx, frame_begin, n_frames, max_depth, seed);
}

extern "C" int tr_render_debug(tr_ctx* ctx) {
    if (!ctx) return TR_ERR_INVALID;
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    WfArgs a; int rc;
    if ((rc = tr_stats_resolve(ctx))) return rc;
    if ((rc = fill_args(ctx, a))) return rc;
    TR_CUDA(ctx, cudaMemsetAsync(ctx->d_ctr, 0, sizeof(TrCounters), ctx->stream));
    k_debug<<<cdiv(ctx->W * ctx->H, 128), 128, 0, ctx->stream>>>(a, ctx->d_fh);
    TR_CHECK_LAUNCH(ctx);
    TR_CUDA(ctx, cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(TrCounters), cudaMemcpyDeviceToHost, ctx->stream));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.rays_closest = (uint64_t)ctx->W * ctx->H; ctx->stats.rays_shadow = 0;
    ctx->stats.node_visits = ctx->h_ctr[0].visits[0]; ctx->stats.leaf_tests = ctx->h_ctr[0].visits[1];
    ctx->fh_ready = true;
    return TR_OK;
}

extern "C" int tr_first_hit_download(tr_ctx* ctx, float* t, int32_t* prim, float* uv, float* pos, float* gnormal, float* normal, float* dir) {
    if (!ctx || !ctx->fh_ready) return tr_fail(ctx, TR_ERR_INVALID, "tr_first_hit_download: call tr_render_debug first");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n = (size_t)ctx->W * ctx->H;
    float* h = (float*)malloc(n * 16 * 4);
    if (!h) return tr_fail(ctx, TR_ERR_INVALID, "out of host memory");
    cudaError_t e = cudaMemcpyAsync(h, ctx->d_fh, n * 64, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { free(h); return tr_fail(ctx, TR_ERR_CUDA, "first-hit download: %s", cudaGetErrorString(e)); }
    for (size_t p = 0; p < n; ++p) {
        const float* f = h + p * 16;
        if (t) t[p] = f[0];
        if (prim) memcpy(&prim[p], &f[1], 4);
        if (uv) { uv[p * 2] = f[2]; uv[p * 2 + 1] = f[3]; }
        if (pos) memcpy(pos + p * 3, f + 4, 12);
        if (gnormal) memcpy(gnormal + p * 3, f + 7, 12);
        if (normal) memcpy(normal + p * 3, f + 10, 12);
        if (dir) memcpy(dir + p * 3, f + 13, 12);
    }
    free(h);
    return TR_OK;
}

extern "C" int tr_tonemap(tr_ctx* ctx, float exposure) {
    if (!ctx || !ctx->d_hdr) return tr_fail(ctx, TR_ERR_INVALID, "tr_tonemap: no film");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    int n = ctx->W * ctx->H * 3;
    k_tonemap<<<cdiv(n, 256), 256, 0, ctx->stream>>>(tr_present_hdr(ctx), ctx->d_rgb, n, exposure);
    TR_CHECK_LAUNCH(ctx);
    return TR_OK;
}

// ------------------------------------------------------------------ arbitrary-ray test hooks
// kernel 0: the simple one-lane-per-ray walk (trace_closest / trace_shadow_visible, also used by the Debug integrator)
__global__ void k_test_trace(WfArgs a, int n, const float* __restrict__ o, const float* __restrict__ d, int shadow,
                             float* __restrict__ t, int* __restrict__ prim, float* __restrict__ uv) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = k < n;
    if (!active) k = 0;
    RayPre r = make_ray(mk3(o[k * 3], o[k * 3 + 1], o[k * 3 + 2]), mk3(d[k * 3], d[k * 3 + 1], d[k * 3 + 2]));
    TreeView tv; tv.snodes = tv.sleaves = nullptr; tv.gnodes = a.nodes2; tv.gleaves = a.leaves4; tv.top = 0;
    HitRec h = trace_closest(tv, a.root, r, active, nullptr);
    if (shadow) {
        // cross-check the early-exit shadow query against the closest-hit answer it must reproduce
        bool has = active && h.prim >= 0;
        bool vis = trace_shadow_visible(tv, a.root, r, has, has ? a.leaf_of_prim[h.prim] : 0, nullptr);
        if (has && !vis) h.prim = -2;
    }
    if (!active) return;
    t[k] = h.t; prim[k] = h.prim;
    if (uv) { uv[k * 2] = h.u; uv[k * 2 + 1] = h.v; }
}

// the rays as queue records of the PRODUCTION kernels: path queue (k_trace, k_tail) or shadow queue with a per-ray target
// primitive and unit contribution (k_shadow)
__global__ void k_test_fill(WfArgs a, int n, const float* __restrict__ o, const float* __restrict__ d, const int* __restrict__ target, int pp) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float4 A = make_float4(o[k * 3], o[k * 3 + 1], o[k * 3 + 2], d[k * 3]);
    if (target) {
        a.sa[0][k] = A; a.sb[0][k] = make_float4(d[k * 3 + 1], d[k * 3 + 2], __int_as_float(target[k]), __uint_as_float((unsigned)k));
        a.sc[0][k] = make_float4(1.0f, 0.0f, 0.0f, 0.0f); a.Lnee[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    } else {
        a.pa[pp][k] = A; a.pb[pp][k] = make_float4(d[k * 3 + 1], d[k * 3 + 2], 1.0f, __uint_as_float((unsigned)k | SPEC_BIT));
        a.pc[pp][k] = make_float4(1.0f, 1.0f, 1.0f, 0.0f); a.hit[k] = make_float4(-1.0f, __int_as_float(-7), 0.0f, 0.0f);
    }
}
__global__ void k_test_ctr(TrCounters* c, int n, int which) {
    if (which == 2) c->nshadow[0] = n; else if (which == 3) { c->nq[1] = n; c->tail_from = 1; } else c->nq[0] = n;
}

// kernel: 0 simple walk (shadow != 0: plus the shadow-query cross-check); 1 production k_trace; 2 production k_shadow (target[] =
// primitive each ray must see: prim out = target if visible else -2, t = 1 / 0); 3 production k_tail (first walk of each path).
// Kernels 1..3 run in the tree mode the renderer would use (options smem_bvh / replicas / top_nodes apply).
extern "C" int tr_test_trace_kernel(tr_ctx* ctx, int kernel, int n, const float* o, const float* d, const int32_t* target, int shadow,
                                    float* t, int32_t* prim, float* uv) {
    if (!ctx || n <= 0 || !o || !d || !t || !prim || kernel < 0 || kernel > 3 || (kernel == 2 && !target))
        return tr_fail(ctx, TR_ERR_INVALID, "tr_test_trace_kernel: bad arguments");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->bvh_ready) return tr_fail(ctx, TR_ERR_INVALID, "tr_test_trace_kernel: BVH not built");
    int rc; if ((rc = tr_stats_resolve(ctx))) return rc;
    if ((rc = tr_build_shade_table(ctx))) return rc;
    WfArgs a; memset(&a, 0, sizeof(a));
    tr_tree_args(ctx, a);
    a.leaf_of_prim = ctx->d_leaf_of_prim; a.material = ctx->d_material; a.ctr = ctx->d_ctr; a.bp = (const BatchParams*)ctx->d_batch_params;
    a.tail_chunk = 8; a.probe = 1;
    cudaStream_t s = ctx->stream;
    if (kernel > 0) {
        if ((rc = ensure_wavefront(ctx, (size_t)n))) return rc;
        for (int k = 0; k < 2; ++k) { a.pa[k] = ctx->d_path[k][0]; a.pb[k] = ctx->d_path[k][1]; a.pc[k] = ctx->d_path[k][2]; }
        a.hit = ctx->d_hit; a.cls = ctx->d_cls; a.cap = ctx->wf_cap; a.Lnee = ctx->d_Lnee; a.L = ctx->d_L;
        for (int k = 0; k < 2; ++k) { a.sa[k] = ctx->d_shq[k][0]; a.sb[k] = ctx->d_shq[k][1]; a.sc[k] = ctx->d_shq[k][2]; }
    }
    float *d_o, *d_d, *d_t, *d_uv; int *d_p, *d_tg = nullptr;
    TR_CUDA(ctx, cudaMalloc((void**)&d_o, (size_t)n * 12)); TR_CUDA(ctx, cudaMalloc((void**)&d_d, (size_t)n * 12));
    TR_CUDA(ctx, cudaMalloc((void**)&d_t, (size_t)n * 4)); TR_CUDA(ctx, cudaMalloc((void**)&d_p, (size_t)n * 4));
    TR_CUDA(ctx, cudaMalloc((void**)&d_uv, (size_t)n * 8));
    cudaMemcpyAsync(d_o, o, (size_t)n * 12, cudaMemcpyHostToDevice, s); cudaMemcpyAsync(d_d, d, (size_t)n * 12, cudaMemcpyHostToDevice, s);
    if (kernel == 2) { TR_CUDA(ctx, cudaMalloc((void**)&d_tg, (size_t)n * 4)); cudaMemcpyAsync(d_tg, target, (size_t)n * 4, cudaMemcpyHostToDevice, s); }
    std::vector<float4> hbuf;
    if (kernel == 0) {
        k_test_trace<<<cdiv(n, 128), 128, 0, s>>>(a, n, d_o, d_d, shadow, d_t, d_p, d_uv);
        cudaMemcpyAsync(t, d_t, (size_t)n * 4, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(prim, d_p, (size_t)n * 4, cudaMemcpyDeviceToHost, s);
        if (uv) cudaMemcpyAsync(uv, d_uv, (size_t)n * 8, cudaMemcpyDeviceToHost, s);
    } else {
        LaunchCfg c; memset(&c, 0, sizeof(c));
        if ((rc = launch_cfg(ctx, a, c))) return rc;
        BatchParams bp; bp.frame_begin = 0; bp.n_frames = 1; bp.seed = 0; bp.max_depth = 4; bp.pad = 0;
        cudaMemcpyAsync(ctx->d_batch_params, &bp, sizeof(bp), cudaMemcpyHostToDevice, s);
        cudaMemsetAsync(ctx->d_ctr, 0, sizeof(TrCounters), s);
        k_test_fill<<<cdiv(n, 256), 256, 0, s>>>(a, n, d_o, d_d, d_tg, kernel == 3 ? 1 : 0);
        k_test_ctr<<<1, 1, 0, s>>>(ctx->d_ctr, n, kernel);
        if (kernel == 1) TR_MODE_SWITCH(c.mode, (k_trace<M><<<c.grid_trace, WF_THREADS, c.smem, s>>>(a, 0)));
        else if (kernel == 2) TR_MODE_SWITCH(c.mode, (k_shadow<M, false><<<c.grid_shadow, WF_THREADS, c.smem, s>>>(a, 0)));
        else TR_MODE_SWITCH(c.mode, (k_tail<M, false><<<c.grid_tail[0], WF_THREADS, c.smem, s>>>(a, 1)));
        hbuf.resize((size_t)n);
        cudaMemcpyAsync(hbuf.data(), kernel == 2 ? a.Lnee : a.hit, (size_t)n * 16, cudaMemcpyDeviceToHost, s);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d_o); cudaFree(d_d); cudaFree(d_t); cudaFree(d_p); cudaFree(d_uv); if (d_tg) cudaFree(d_tg);
    if (e != cudaSuccess) return tr_fail(ctx, TR_ERR_CUDA, "tr_test_trace_kernel: %s", cudaGetErrorString(e));
    if (kernel == 2) for (int k = 0; k < n; ++k) { const bool vis = hbuf[k].x == 1.0f; t[k] = vis ? 1.0f : 0.0f; prim[k] = vis ? target[k] : -2; }
    else if (kernel > 0) for (int k = 0; k < n; ++k) {
        t[k] = hbuf[k].x; memcpy(&prim[k], &hbuf[k].y, 4);
        if (uv) { uv[k * 2] = hbuf[k].z; uv[k * 2 + 1] = hbuf[k].w; }
    }
    ctx->gen++;                                 // the queues were overwritten
    return TR_OK;
}

extern "C" int tr_test_trace(tr_ctx* ctx, int n, const float* o, const float* d, int shadow, float* t, int32_t* prim, float* uv) {
    return tr_test_trace_kernel(ctx, 0, n, o, d, nullptr, shadow, t, prim, uv);
}

#include "bdpt.cuh"
