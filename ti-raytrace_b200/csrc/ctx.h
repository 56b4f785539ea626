// ctx.h — the per-GPU context behind the C-ABI of include/tiray.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <unordered_map>
#include "../../include/tiray.h"
#include "spectral.cuh"

#define TR_MAX_DEPTH_CAP 64       // counters are sized for this many wavefront stages
#define TR_MAX_CHAINS 8           // independent wavefront chains per batch (parallel graph branches)
#define TR_TILE 32                // framebuffer tile edge used for sharding and ray coherence

// Pre-order node of the reference's flattened tree, 32 B (left child = idx+1): lo = (min.xyz, as_float(escape = idx + subtree size)),
// hi = (max.xyz, as_float(link)), link >= 0: pre-order index of the right child, link < 0: leaf with sorted position -link-1.
// Walked by process_normal's point queries (normals.cu), which must visit candidates in the reference's order.
struct TrNode { float4 lo, hi; };
// Traversal node of the render kernels, 64 B, one per INTERNAL node: the boxes of both children, so one dependent load serves two
// slab tests.  a = (lo0.xyz, as_float(link0)), b = (hi0.xyz, as_float(link1)), c = (lo1.xyz, 0), d = (hi1.xyz, 0); child 0 = left.
//   link >= 0: index of an internal node in this array; link < 0: leaf, sorted position k = -link-1 (its box carries the guard band).
// Order: the first top_count nodes are the breadth-first top of the tree (the part TM_GTOP stages into shared memory), the rest
// follows in pre-order.
struct TrNode2 { float4 a, b, c, d; };
#define TR_STACK_MAX 128                     // traversal-stack entries per lane a tree may need (1 KB of shared memory per entry and CTA; else TR_ERR_STACK at build)
#define TR_TOP_MAX 1024                      // breadth-first top nodes kept contiguous at the front of the TrNode2 array
#define TR_SMALL_IMG_MAX (40 * 1024)         // largest replicated shared-memory image (see trace.cuh)
// Leaf record, 48 B, in sorted (Morton) order:
//   a = (v0.xyz, as_float(prim id)), b = (E1.xyz, as_float(kind)), c = (E2.xyz, as_float(material id))   kind 0: triangle
//   a = (centre.xyz, prim id),       b = (radius, 0, 0, kind)                         kind 1: sphere
//   kind 2: shape that never intersects (Scene.py:597-598)
struct TrLeaf { float4 a, b, c; };
// Shading record per primitive, 96 B (built lazily: process_normal may rewrite normals)
//   q0 = (v1, as_float(mat)), q1 = (v2, as_float(kind)), q2 = (v3, area), q3..q5 = (n1, gn.x), (n2, gn.y), (n3, gn.z); gn = geometric normal
//   sphere: q0 = (centre, mat), q1 = (radius,0,0,kind=1), q2 = (0,0,0,area)
struct TrShade { float4 q[6]; };

struct TrCamera { float view_inv[16]; float eye[3]; float fx, fy, cx, cy; };

// device-side counters of one wavefront batch
struct TrCounters {
    int nq[TR_MAX_DEPTH_CAP + 1];        // live paths entering stage d
    int nshadow[TR_MAX_DEPTH_CAP + 1];   // shadow rays emitted by stage d
    int ncls[TR_MAX_DEPTH_CAP + 1][4];   // material-sorted shade queue sizes: terminal, disney, glass
    int wf_trace[TR_MAX_DEPTH_CAP + 1];  // work-fetch cursors of the persistent trace kernels
    int wf_shadow[TR_MAX_DEPTH_CAP + 1];
    int shade_done[TR_MAX_DEPTH_CAP + 1]; // blocks of k_shade(d) that finished (last one decides the hand-over)
    int tail_from;                        // depth (>= 1) from which k_tail owns the chain's paths; 0: wavefront all the way
    int pad_;
    unsigned long long tail_rays[2];      // closest / shadow traversals executed by k_tail
    unsigned long long visits[4];        // closest: nodes, leaves; shadow: nodes, leaves (only with -DTR_COUNTERS)
};

struct tr_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    std::string err;

    // scene tables (reference layouts, device copies)
    float* d_vertex = nullptr; int nv = 0;
    int*   d_prim = nullptr;   int np = 0;
    float* d_material = nullptr; int nm = 0;
    float* d_shape = nullptr;  int ns = 0;
    int*   d_light = nullptr;  int nl = 0;
    float bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
    int*   d_env = nullptr; int env_w = 0, env_h = 0; float env_power = 0.0f;

    // LBVH build products
    bool bvh_ready = false;
    int*   d_morton_unsorted = nullptr;  // n x 2
    int*   d_keys[2] = {nullptr, nullptr};
    int*   d_vals[2] = {nullptr, nullptr};
    int    sorted_buf = 0;
    int*   d_left = nullptr; int* d_right = nullptr; int* d_parent = nullptr;   // 2n-1 (build-order ids)
    float* d_boxes = nullptr;            // (2n-1) x 6
    int*   d_leafcount = nullptr; int* d_flag = nullptr; int* d_pre = nullptr;
    int*   d_build_status = nullptr;     // [0] = refit-completed internal nodes
    TrNode* d_nodes = nullptr; TrLeaf* d_leaves = nullptr; int* d_leaf_of_prim = nullptr;
    TrNode2* d_nodes2 = nullptr; float4* d_small_img = nullptr;         // render-kernel nodes; 8-way replicated image of a small tree
    int* d_first = nullptr; int* d_sneed = nullptr; int* d_irank = nullptr; int* d_top = nullptr;   // build scratch: first leaf, stack need, node order
    float root_box[6] = {0, 0, 0, 0, 0, 0}; int stack_need = 1, top_count = 0;
    TrShade* d_shade = nullptr; bool shade_ready = false;
    int*   d_hist = nullptr; size_t hist_cap = 0;

    // camera + film
    TrCamera cam; bool cam_set = false;
    int W = 0, H = 0;
    float* d_hdr = nullptr; float* d_rgb = nullptr;
    // multi-GPU: tr_film_reduce sums the per-rank partial films (d_hdr stays a pure partial) into d_hdr_sum on the root; while
    // present_sum is set, tone map and download read the sum.  Any render / clear / upload of the film resets it.
    float* d_hdr_sum = nullptr; bool present_sum = false;
    void* nccl_comm = nullptr; int comm_rank = 0, comm_nranks = 1;
    // pinned host memory: upload staging (bump allocator), film download target, LBVH build status
    char* h_stage = nullptr; size_t stage_cap = 0, stage_used = 0; bool stage_busy = false, stage_direct = false; cudaEvent_t ev_stage = nullptr;
    float* h_film = nullptr; size_t film_host_cap = 0;
    int* h_build_status = nullptr;
    // first-hit buffers (Debug integrator)
    float* d_fh = nullptr;      // W*H*16 floats: t, prim, u, v, pos3, gn3, n3, dir3
    bool fh_ready = false;

    // sharding
    int rank = 0, nranks = 1;
    int* d_tiles = nullptr; int n_local_tiles = 0; bool tiles_ready = false;

    // wavefront buffers
    size_t wf_cap = 0;              // path slots allocated
    float4* d_path[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    float4* d_hit = nullptr;
    int*    d_cls = nullptr;        // 3 x cap indices (material-sorted shade queues)
    float4* d_shq[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    float4* d_L = nullptr; float4* d_Lnee = nullptr;          // per-sample radiance: terminal term, NEE sum
    TrCounters* d_ctr = nullptr;          // TR_MAX_CHAINS entries
    TrCounters h_ctr[TR_MAX_CHAINS];
    // deferred statistics of an asynchronous render: per-batch counter snapshots land in a pinned ring, tr_stats_get folds them
    TrCounters* h_ring = nullptr; int ring_batches = 0, ring_K = 0, ring_depth = 0, ring_mode = 0; bool stats_pending = false;   // ring_mode 0 PT, 1 BDPT wavefront, 2 BDPT lock-step
    float4* d_matlin = nullptr; bool matlin_ready = false;
    cudaStream_t sub_stream[TR_MAX_CHAINS] = {}; cudaEvent_t ev_join[TR_MAX_CHAINS] = {}; cudaEvent_t ev_fork = nullptr;

    // spectral integrator tables (PT_Spec): device copies + the view handed to the kernels
    SpecDev spec = {};
    float4* d_sensor = nullptr; float* d_spectrum[4] = {nullptr, nullptr, nullptr, nullptr};
    float* d_rs_scale = nullptr; float* d_rs_data = nullptr; int rs_res = 0;
    float* d_sky = nullptr; float4* d_matspec = nullptr; bool matspec_ready = false;
    float* d_white_point = nullptr;

    // bidirectional integrator (BDPT_RGB): vertex records, per-strategy contributions, splat film, connection queue
    float view[16] = {0}; bool view_set = false;   // Camera.view (get_image_point / get_optical_axis)
    float4* d_bd_vb = nullptr; int* d_bd_depths = nullptr; float4* d_bd_contrib = nullptr; float* d_bd_splat = nullptr;
    unsigned* d_bd_items = nullptr; int* d_bd_tile_slot = nullptr; unsigned long long* d_bd_ctr = nullptr;
    float4* d_bd_sq[2] = {nullptr, nullptr}; float* d_bd_vis = nullptr;      // wavefront pipeline: connection shadow queue (sa, sb) + query results
    size_t bd_cap = 0; int bd_splat_frames = 0; float bd_ms_ktrace = 0.0f, bd_ms_kshadow = 0.0f;

    // options
    int opt_batch_frames = 0;       // 0 = auto
    int opt_stage_timing = 0;
    int opt_graph = 1;
    int opt_smem_bvh = 1;
    int opt_chains = 2;
    int opt_shadow_overlap = 1;
    int opt_tail_max = -1;          // hand-over threshold of k_tail (live paths of a chain): -1 = default (4096, measured on 8-way shards of C2 / C3), 0 = never
    int opt_tail_chunk = 8;
    int opt_bdpt_wavefront = 1;     // 0: lock-step BDPT pipeline (cross-check)
    int opt_top_nodes = 0;          // large trees: this many breadth-first top nodes are staged into shared memory per CTA (0 = off)
    int opt_chain_skew = 65;        // two chains: percent of a batch's frames given to chain 0 (0 = even split); 65: their thin deep stages do not coincide
    int opt_persist_blocks = 0;     // cap on the resident CTAs per SM of the persistent trace / shadow kernels (0 = what fits)
    int opt_pdl = 0;                // programmatic dependent launch between the stages of a chain
    int opt_replicas = 1;           // small trees: bank-conflict-free 8-replica shared-memory image
    size_t opt_max_paths = (size_t)64 << 20;   // path slots per batch (188 B each = 12.6 GB of the 180 GB): a 64-spp step of a 1024^2 image is ONE batch
                                               // -- every batch ends with an exposed tail and thin deep stages (C3: 4 batches 29.4 ms, 2: 26.7, 1: 25.1 ms/step)

    // cuda graph cache for the batch pipeline
    cudaGraphExec_t graph_exec = nullptr; int graph_launches = 0, graph_depth = 0, graph_chains = 0, graph_fs = 0, graph_spec = 0;
    std::vector<char> graph_args;
    unsigned long long gen = 0, graph_gen = 0;   // any state change bumps gen; a captured graph is valid for one gen
    void* d_batch_params = nullptr;

    tr_stats stats;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> stage_ev;
    std::vector<cudaEvent_t> dep_ev;               // [chain][depth][shade done, shadow done]
    cudaStream_t shadow_stream[TR_MAX_CHAINS] = {};
    std::unordered_map<void*, size_t> capacity;   // bytes behind every buffer handed out by tr_realloc (re-used when large enough)
    float* d_smooth = nullptr;                    // process_normal scratch
    cudaStream_t own_stream = nullptr;   // ctx->stream may be redirected to a caller's stream (tr_stream_set)
};

int tr_fail(tr_ctx* ctx, int code, const char* fmt, ...);
static inline float* tr_present_hdr(tr_ctx* ctx) { return (ctx->present_sum && ctx->d_hdr_sum) ? ctx->d_hdr_sum : ctx->d_hdr; }
void tr_comm_release(tr_ctx* ctx);                                              // comm.cu
int tr_stage_h2d(tr_ctx* ctx, void* dst, const void* src, size_t bytes);        // api.cu: asynchronous upload through pinned staging
#define TR_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return tr_fail(ctx, TR_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)
#define TR_CHECK_LAUNCH(ctx) TR_CUDA(ctx, cudaGetLastError())

// (Re)allocate a context-owned device buffer. An existing buffer is kept when it is large enough: cudaFree /
// cudaMalloc synchronise the device and cost milliseconds, which would dominate a scene re-upload.
template <typename T> static inline int tr_realloc(tr_ctx* ctx, T** p, size_t count) {
    if (count == 0) count = 1;
    const size_t bytes = count * sizeof(T);
    if (*p) {
        auto it = ctx->capacity.find((void*)*p);
        if (it != ctx->capacity.end() && it->second >= bytes) return TR_OK;
        if (it != ctx->capacity.end()) ctx->capacity.erase(it);
        cudaFree(*p); *p = nullptr;
    }
    TR_CUDA(ctx, cudaMalloc((void**)p, bytes));
    ctx->capacity[(void*)*p] = bytes;
    return TR_OK;
}

// implemented across the .cu files
int tr_build_shade_table(tr_ctx* ctx);
int tr_build_tiles(tr_ctx* ctx);
int tr_spec_prepare(tr_ctx* ctx);
int tr_stats_resolve(tr_ctx* ctx);     // wavefront.cu: wait for the last asynchronous render and fold its counters into ctx->stats
#define TR_RING_BATCHES 64      // spectral.cu: checks the PT_Spec tables and builds the per-material coefficient table
