// api.cu — context, uploads/downloads, camera, film and sharding entry points of include/tiray.h
#include <stdarg.h>
#include <string.h>
#include <vector>
#include "ctx.h"
#include "common.cuh"

static thread_local std::string g_last_error;     // per calling thread (ctx == NULL errors: tr_ctx_create, tr_obj_*)

int tr_fail(tr_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (ctx) ctx->err = buf;
    g_last_error = buf;
    return code;
}

extern "C" {

const char* tr_last_error(tr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int tr_host_register(void* p, size_t bytes) {
    if (!p || bytes == 0) return tr_fail(nullptr, TR_ERR_INVALID, "tr_host_register: NULL or empty range");
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return TR_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); return tr_fail(nullptr, TR_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e)); }
    return TR_OK;
}
int tr_host_unregister(void* p) {
    if (!p) return TR_ERR_INVALID;
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { cudaGetLastError(); return tr_fail(nullptr, TR_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e)); }
    return TR_OK;
}

int tr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int ctx_init(tr_ctx* ctx, int device) {
    ctx->device = device;
    TR_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    TR_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    if (prop.major < 10) return tr_fail(ctx, TR_ERR_NO_DEVICE, "device %d is sm_%d%d; libtiray is built for sm_100a only", device, prop.major, prop.minor);
    TR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = ctx->stream;
    TR_CUDA(ctx, cudaEventCreate(&ctx->ev0));
    TR_CUDA(ctx, cudaEventCreate(&ctx->ev1));
    TR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_stage, cudaEventDisableTiming));
    TR_CUDA(ctx, cudaMalloc((void**)&ctx->d_ctr, sizeof(TrCounters) * TR_MAX_CHAINS));
    TR_CUDA(ctx, cudaMalloc((void**)&ctx->d_build_status, 16 * sizeof(int)));
    TR_CUDA(ctx, cudaMalloc((void**)&ctx->d_batch_params, 64));
    TR_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_build_status, 16 * sizeof(int), cudaHostAllocDefault));
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    memset(&ctx->cam, 0, sizeof(ctx->cam));
    return TR_OK;
}

int tr_ctx_create(int device, tr_ctx** out) {
    if (!out) return tr_fail(nullptr, TR_ERR_INVALID, "tr_ctx_create: out is NULL");
    *out = nullptr;
    int n = tr_device_count();
    if (n <= 0) return tr_fail(nullptr, TR_ERR_NO_DEVICE, "no CUDA device visible: libtiray has no CPU fallback");
    if (device < 0 || device >= n) return tr_fail(nullptr, TR_ERR_INVALID, "device %d out of range (%d devices)", device, n);
    tr_ctx* ctx = new tr_ctx();
    int rc = ctx_init(ctx, device);
    if (rc != TR_OK) {                       // nothing leaks on a failed create: the partially built context goes through the destructor
        std::string why = ctx->err;
        tr_ctx_destroy(ctx);
        return tr_fail(nullptr, rc, "%s", why.c_str());
    }
    *out = ctx;
    return TR_OK;
}

void tr_ctx_destroy(tr_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    tr_comm_release(ctx);
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_film) cudaFreeHost(ctx->h_film);
    if (ctx->h_build_status) cudaFreeHost(ctx->h_build_status);
    if (ctx->ev_stage) cudaEventDestroy(ctx->ev_stage);
    void* ptrs[] = {ctx->d_vertex, ctx->d_prim, ctx->d_material, ctx->d_shape, ctx->d_light, ctx->d_env,
                    ctx->d_morton_unsorted, ctx->d_keys[0], ctx->d_keys[1], ctx->d_vals[0], ctx->d_vals[1],
                    ctx->d_left, ctx->d_right, ctx->d_parent, ctx->d_boxes, ctx->d_leafcount, ctx->d_flag, ctx->d_pre,
                    ctx->d_build_status, ctx->d_nodes, ctx->d_leaves, ctx->d_leaf_of_prim, ctx->d_shade, ctx->d_hist, ctx->d_hdr, ctx->d_rgb,
                    ctx->d_fh, ctx->d_tiles, ctx->d_path[0][0], ctx->d_path[0][1], ctx->d_path[0][2], ctx->d_path[1][0],
                    ctx->d_path[1][1], ctx->d_path[1][2], ctx->d_hit, ctx->d_cls, ctx->d_shq[0][0], ctx->d_shq[0][1],
                    ctx->d_shq[0][2], ctx->d_shq[1][0], ctx->d_shq[1][1], ctx->d_shq[1][2], ctx->d_Lnee, ctx->d_L, ctx->d_ctr, ctx->d_batch_params, ctx->d_matlin, ctx->d_smooth,
                    ctx->d_sensor, ctx->d_spectrum[0], ctx->d_spectrum[1], ctx->d_spectrum[2], ctx->d_spectrum[3], ctx->d_rs_scale, ctx->d_rs_data,
                    ctx->d_sky, ctx->d_matspec, ctx->d_white_point,
                    ctx->d_bd_vb, ctx->d_bd_depths, ctx->d_bd_contrib, ctx->d_bd_splat, ctx->d_bd_items, ctx->d_bd_tile_slot, ctx->d_bd_ctr,
                    ctx->d_bd_sq[0], ctx->d_bd_sq[1], ctx->d_bd_vis, ctx->d_hdr_sum, ctx->d_nodes2, ctx->d_small_img, ctx->d_first, ctx->d_sneed, ctx->d_irank, ctx->d_top};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (auto e : ctx->stage_ev) cudaEventDestroy(e);
    for (auto e : ctx->dep_ev) cudaEventDestroy(e);
    for (int j = 0; j < TR_MAX_CHAINS; ++j) if (ctx->shadow_stream[j]) cudaStreamDestroy(ctx->shadow_stream[j]);
    for (int j = 1; j < TR_MAX_CHAINS; ++j) { if (ctx->sub_stream[j]) cudaStreamDestroy(ctx->sub_stream[j]); if (ctx->ev_join[j]) cudaEventDestroy(ctx->ev_join[j]); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int tr_stream_set(tr_ctx* ctx, void* cuda_stream) {
    if (!ctx) return TR_ERR_INVALID;
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    ctx->gen++;
    return TR_OK;
}

int tr_synchronize(tr_ctx* ctx) {
    if (!ctx) return TR_ERR_INVALID;
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

// ---- pinned staging of host -> device uploads
// The caller's arrays are borrowed for the duration of the call only.  They are copied once into a context-owned pinned buffer
// (host memcpy) and DMA'd from there asynchronously, so no upload synchronises the stream; a pageable cudaMemcpyAsync would
// stage through the driver's bounce buffers AND block.  The buffer is a bump allocator: when it runs out, the event of its last
// use is awaited and it starts over (or grows).
static int stage_reserve(tr_ctx* ctx, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (ctx->stage_used + bytes <= ctx->stage_cap) return TR_OK;
    if (ctx->stage_busy) { TR_CUDA(ctx, cudaEventSynchronize(ctx->ev_stage)); ctx->stage_busy = false; }
    ctx->stage_used = 0;
    if (bytes > ctx->stage_cap) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr; ctx->stage_cap = 0;
        size_t cap = bytes * 2 < ((size_t)4 << 20) ? ((size_t)4 << 20) : bytes * 2;
        TR_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_stage, cap, cudaHostAllocDefault));
        ctx->stage_cap = cap;
    }
    return TR_OK;
}
// true when the caller's array already is page-locked (tr_host_register, or any cudaHostAlloc / cudaHostRegister memory)
static bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}
// copy `bytes` from the caller's array into the staging buffer and enqueue the DMA to dst (reserve the total first).
// Page-locked source arrays (>= 64 KB) skip the staging copy: the DMA reads them directly and stage_commit waits for it, so the
// arrays are still only borrowed for the duration of the call.
static int stage_upload(tr_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return TR_OK;
    if (bytes >= (64u << 10) && is_pinned(src)) {
        TR_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        ctx->stage_direct = true;
        return TR_OK;
    }
    char* h = ctx->h_stage + ctx->stage_used;
    memcpy(h, src, bytes);
    ctx->stage_used += (bytes + 255) & ~(size_t)255;
    TR_CUDA(ctx, cudaMemcpyAsync(dst, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return TR_OK;
}
static int stage_commit(tr_ctx* ctx) {
    TR_CUDA(ctx, cudaEventRecord(ctx->ev_stage, ctx->stream));
    ctx->stage_busy = true;
    if (ctx->stage_direct) {                      // a DMA is reading the caller's own (page-locked) array: wait until it has
        ctx->stage_direct = false;
        TR_CUDA(ctx, cudaEventSynchronize(ctx->ev_stage));
        ctx->stage_busy = false;
    }
    return TR_OK;
}
}  // extern "C"
int tr_stage_h2d(tr_ctx* ctx, void* dst, const void* src, size_t bytes) {      // for the other .cu files (spectral tables, tiles)
    int rc;
    if ((rc = stage_reserve(ctx, bytes))) return rc;
    if ((rc = stage_upload(ctx, dst, src, bytes))) return rc;
    return stage_commit(ctx);
}
extern "C" {

// O(n) host validation of the packed tables (Scene.py:225-273 layouts): an index out of range would otherwise become an
// out-of-bounds device read in k_morton / k_refit / k_shade_table
static int validate_scene(tr_ctx* ctx, const float* vertex, int nv, const int32_t* prim, int np, const float* material, int nm,
                          const float* shape, int ns, const int32_t* light, int nl) {
    (void)vertex;
    for (int i = 0; i < np; ++i) {
        const int type = prim[i * 3], idx = prim[i * 3 + 1], mat = prim[i * 3 + 2];
        if (mat < 0 || mat >= nm) return tr_fail(ctx, TR_ERR_INVALID, "tr_scene_upload: primitive %d: material index %d out of range (%d materials)", i, mat, nm);
        if (type == TR_PRIM_TRI) {
            if (idx < 0 || (long long)idx + 2 >= (long long)nv) return tr_fail(ctx, TR_ERR_INVALID, "tr_scene_upload: primitive %d: vertex index %d + 2 out of range (%d vertices)", i, idx, nv);
        } else if (type == 2) {
            if (!shape || idx < 0 || idx >= ns) return tr_fail(ctx, TR_ERR_INVALID, "tr_scene_upload: primitive %d: shape index %d out of range (%d shapes)", i, idx, ns);
            // emitters: triangle, sphere, spot and laser lights (Scene.sample_li / sample_light, Scene.py:430-518).  A QUAD emitter
            // has no area in the reference (get_prim_area returns 0 -> infinite choice pdf): rejected instead of rendered wrong.
            const int st = (int)shape[(size_t)idx * 10];
            if ((int)material[(size_t)mat * 10] == TR_MAT_LIGHT && st != TR_SHAPE_SPHERE && st != TR_SHAPE_SPOT && st != TR_SHAPE_LASER)
                return tr_fail(ctx, TR_ERR_INVALID, "tr_scene_upload: primitive %d: emitter shape type %d is not supported (triangle, sphere, spot and laser emitters only)", i, st);
        } else return tr_fail(ctx, TR_ERR_INVALID, "tr_scene_upload: primitive %d: unknown primitive type %d", i, type);
    }
    for (int i = 0; i < nl; ++i)
        if (light[i] < 0 || light[i] >= np) return tr_fail(ctx, TR_ERR_INVALID, "tr_scene_upload: light %d: primitive index %d out of range (%d primitives)", i, light[i], np);
    return TR_OK;
}

int tr_scene_upload(tr_ctx* ctx, const float* vertex, int nv, const int32_t* prim, int np,
                    const float* material, int nm, const float* shape, int ns,
                    const int32_t* light, int nl, const float bmin[3], const float bmax[3]) {
    if (!ctx || !vertex || !prim || !material || np <= 0 || nm <= 0 || nv < 0 || ns < 0 || nl < 0 || !bmin || !bmax)
        return tr_fail(ctx, TR_ERR_INVALID, "tr_scene_upload: NULL table or empty scene (np=%d nm=%d)", np, nm);
    if (np >= (1 << 24)) return tr_fail(ctx, TR_ERR_INVALID, "tr_scene_upload: %d primitives (node indices are f32-exact below 2^24, UtilsFunc.py:295-308)", np);
    if (!light) nl = 0;
    if (!shape) ns = 0;
    int rc;
    if ((rc = validate_scene(ctx, vertex, nv, prim, np, material, nm, shape, ns, light, nl))) return rc;
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->nv = nv; ctx->np = np; ctx->nm = nm; ctx->ns = ns; ctx->nl = nl;
    for (int k = 0; k < 3; ++k) { ctx->bmin[k] = bmin[k]; ctx->bmax[k] = bmax[k]; }
    if ((rc = tr_realloc(ctx, &ctx->d_vertex, (size_t)nv * 9))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_prim, (size_t)np * 3))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_material, (size_t)nm * 10))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_shape, (size_t)(ns > 0 ? ns : 1) * 10))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_light, (size_t)(nl > 0 ? nl : 1)))) return rc;
    cudaStream_t s = ctx->stream;
    const size_t bv = (size_t)nv * 36, bp = (size_t)np * 12, bm = (size_t)nm * 40, bs = (size_t)ns * 40, bl = (size_t)nl * 4;
    if ((rc = stage_reserve(ctx, bv + bp + bm + bs + bl + 5 * 256))) return rc;
    if ((rc = stage_upload(ctx, ctx->d_vertex, vertex, bv))) return rc;
    if ((rc = stage_upload(ctx, ctx->d_prim, prim, bp))) return rc;
    if ((rc = stage_upload(ctx, ctx->d_material, material, bm))) return rc;
    if (ns > 0) { if ((rc = stage_upload(ctx, ctx->d_shape, shape, bs))) return rc; }
    else TR_CUDA(ctx, cudaMemsetAsync(ctx->d_shape, 0, 10 * 4, s));
    if (nl > 0 && (rc = stage_upload(ctx, ctx->d_light, light, bl))) return rc;
    if ((rc = stage_commit(ctx))) return rc;
    ctx->bvh_ready = false; ctx->shade_ready = false; ctx->fh_ready = false; ctx->matlin_ready = false; ctx->matspec_ready = false; ctx->gen++;
    return TR_OK;
}

int tr_material_upload(tr_ctx* ctx, const float* material, int nm) {
    if (!ctx || !material || nm != ctx->nm) return tr_fail(ctx, TR_ERR_INVALID, "tr_material_upload: nm=%d does not match scene (%d)", nm, ctx ? ctx->nm : -1);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc; if ((rc = tr_stage_h2d(ctx, ctx->d_material, material, (size_t)nm * 10 * 4))) return rc;
    ctx->matlin_ready = false; ctx->matspec_ready = false;
    return TR_OK;
}

int tr_env_upload(tr_ctx* ctx, const int32_t* rgb, int w, int h, float power) {
    if (!ctx || !rgb || w <= 0 || h <= 0) return tr_fail(ctx, TR_ERR_INVALID, "tr_env_upload: bad image %dx%d", w, h);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc; if ((rc = tr_realloc(ctx, &ctx->d_env, (size_t)w * h))) return rc;
    if ((rc = tr_stage_h2d(ctx, ctx->d_env, rgb, (size_t)w * h * 4))) return rc;
    ctx->env_w = w; ctx->env_h = h; ctx->env_power = power; ctx->gen++;
    return TR_OK;
}

int tr_camera_set(tr_ctx* ctx, const float view[16], const float view_inv[16], const float eye[3],
                  float fx, float fy, float cx, float cy) {
    if (!ctx || !view_inv || !eye) return tr_fail(ctx, TR_ERR_INVALID, "tr_camera_set: NULL argument");
    if (view) memcpy(ctx->view, view, 64);
    ctx->view_set = view != nullptr;
    memcpy(ctx->cam.view_inv, view_inv, 64); memcpy(ctx->cam.eye, eye, 12);
    ctx->cam.fx = fx; ctx->cam.fy = fy; ctx->cam.cx = cx; ctx->cam.cy = cy;
    ctx->cam_set = true; ctx->fh_ready = false; ctx->gen++;
    return TR_OK;
}

int tr_film_create(tr_ctx* ctx, int W, int H) {
    if (!ctx || W <= 0 || H <= 0 || W > 65535 || H > 65535) return tr_fail(ctx, TR_ERR_INVALID, "tr_film_create: bad size %dx%d", W, H);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->W = W; ctx->H = H;
    int rc;
    if ((rc = tr_realloc(ctx, &ctx->d_hdr, (size_t)W * H * 3))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_rgb, (size_t)W * H * 3))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_fh, (size_t)W * H * 16))) return rc;
    TR_CUDA(ctx, cudaMemsetAsync(ctx->d_hdr, 0, (size_t)W * H * 12, ctx->stream));
    TR_CUDA(ctx, cudaMemsetAsync(ctx->d_rgb, 0, (size_t)W * H * 12, ctx->stream));
    ctx->tiles_ready = false; ctx->fh_ready = false; ctx->present_sum = false; ctx->gen++;
    return TR_OK;
}

int tr_film_clear(tr_ctx* ctx) {
    if (!ctx || !ctx->d_hdr) return tr_fail(ctx, TR_ERR_INVALID, "tr_film_clear: no film");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    TR_CUDA(ctx, cudaMemsetAsync(ctx->d_hdr, 0, (size_t)ctx->W * ctx->H * 12, ctx->stream));
    TR_CUDA(ctx, cudaMemsetAsync(ctx->d_rgb, 0, (size_t)ctx->W * ctx->H * 12, ctx->stream));
    ctx->present_sum = false;
    return TR_OK;
}

// DMA the presented film(s) into the context's pinned host buffers and wait: (hdr | rgb), W*H*3 f32 each
static int film_to_pinned(tr_ctx* ctx, bool hdr, bool rgb) {
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)ctx->W * ctx->H * 12;
    if (ctx->film_host_cap < 2 * bytes) {
        if (ctx->h_film) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->h_film); ctx->h_film = nullptr; ctx->film_host_cap = 0; }
        TR_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_film, 2 * bytes, cudaHostAllocDefault));
        ctx->film_host_cap = 2 * bytes;
    }
    if (hdr) TR_CUDA(ctx, cudaMemcpyAsync(ctx->h_film, tr_present_hdr(ctx), bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (rgb) TR_CUDA(ctx, cudaMemcpyAsync(ctx->h_film + bytes / 4, ctx->d_rgb, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

int tr_film_download(tr_ctx* ctx, float* hdr, float* rgb) {
    if (!ctx || !ctx->d_hdr) return tr_fail(ctx, TR_ERR_INVALID, "tr_film_download: no film");
    int rc; if ((rc = film_to_pinned(ctx, hdr != nullptr, rgb != nullptr))) return rc;
    const size_t bytes = (size_t)ctx->W * ctx->H * 12;
    if (hdr) memcpy(hdr, ctx->h_film, bytes);
    if (rgb) memcpy(rgb, ctx->h_film + bytes / 4, bytes);
    return TR_OK;
}

int tr_film_download_pinned(tr_ctx* ctx, int want_hdr, int want_rgb, const float** hdr, const float** rgb) {
    if (!ctx || !ctx->d_hdr) return tr_fail(ctx, TR_ERR_INVALID, "tr_film_download_pinned: no film");
    int rc; if ((rc = film_to_pinned(ctx, want_hdr != 0, want_rgb != 0))) return rc;
    if (hdr) *hdr = want_hdr ? ctx->h_film : nullptr;
    if (rgb) *rgb = want_rgb ? ctx->h_film + (size_t)ctx->W * ctx->H * 3 : nullptr;
    return TR_OK;
}

int tr_film_upload(tr_ctx* ctx, const float* hdr) {
    if (!ctx || !ctx->d_hdr || !hdr) return tr_fail(ctx, TR_ERR_INVALID, "tr_film_upload: no film");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->present_sum = false;
    return tr_stage_h2d(ctx, ctx->d_hdr, hdr, (size_t)ctx->W * ctx->H * 12);
}

int tr_film_device_ptr(tr_ctx* ctx, void** hdr_dev, void** rgb_dev) {
    if (!ctx || !ctx->d_hdr) return tr_fail(ctx, TR_ERR_INVALID, "tr_film_device_ptr: no film");
    if (hdr_dev) *hdr_dev = ctx->d_hdr;
    if (rgb_dev) *rgb_dev = ctx->d_rgb;
    return TR_OK;
}

int tr_set_shard(tr_ctx* ctx, int rank, int nranks) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return tr_fail(ctx, TR_ERR_INVALID, "tr_set_shard: bad rank %d/%d", rank, nranks);
    ctx->rank = rank; ctx->nranks = nranks; ctx->tiles_ready = false; ctx->gen++;
    return TR_OK;
}

int tr_set_option(tr_ctx* ctx, const char* name, int value) {
    if (!ctx || !name) return TR_ERR_INVALID;
    struct Opt { const char* name; int* dst; int lo, hi; };
    const Opt opts[] = {
        {"batch_frames", &ctx->opt_batch_frames, 0, 1 << 20}, {"stage_timing", &ctx->opt_stage_timing, 0, 1}, {"graph", &ctx->opt_graph, 0, 1},
        {"smem_bvh", &ctx->opt_smem_bvh, 0, 1}, {"chains", &ctx->opt_chains, 1, TR_MAX_CHAINS}, {"shadow_overlap", &ctx->opt_shadow_overlap, 0, 1},
        {"tail_max", &ctx->opt_tail_max, -1, 1 << 30}, {"tail_chunk", &ctx->opt_tail_chunk, 1, 32}, {"bdpt_wavefront", &ctx->opt_bdpt_wavefront, 0, 1},
        {"top_nodes", &ctx->opt_top_nodes, 0, 1 << 16}, {"pdl", &ctx->opt_pdl, 0, 1}, {"replicas", &ctx->opt_replicas, 0, 1}, {"chain_skew", &ctx->opt_chain_skew, 0, 95}, {"persist_blocks", &ctx->opt_persist_blocks, 0, 8},
    };
    bool found = false;
    for (const Opt& o : opts) if (!strcmp(name, o.name)) {
        if (value < o.lo || value > o.hi) return tr_fail(ctx, TR_ERR_INVALID, "tr_set_option: %s = %d out of range [%d, %d]", name, value, o.lo, o.hi);
        *o.dst = value; found = true;
    }
    if (!found && !strcmp(name, "max_paths")) {
        // path slots per batch: at least one frame of a 1-tile film, below 2^30 (queue indices carry two class bits)
        if (value < TR_TILE * TR_TILE || value >= (1 << 30)) return tr_fail(ctx, TR_ERR_INVALID, "tr_set_option: max_paths = %d out of range [1024, 2^30)", value);
        ctx->opt_max_paths = (size_t)value; found = true;
    }
    if (!found) return tr_fail(ctx, TR_ERR_INVALID, "tr_set_option: unknown option '%s'", name);
    ctx->gen++;
    if (ctx->graph_exec) {                                                                       // options are baked into the captured graph
        cudaStreamSynchronize(ctx->stream);                                                      // an asynchronous render may still be replaying it
        cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr;
    }
    return TR_OK;
}

int tr_stats_get(tr_ctx* ctx, tr_stats* out) {
    if (!ctx || !out) return TR_ERR_INVALID;
    int rc = tr_stats_resolve(ctx); if (rc) return rc;
    *out = ctx->stats;
    return TR_OK;
}

}  // extern "C"

// Host-side tile list of this rank: tiles in row-major order whose (tx + 3*ty) % nranks == rank.
int tr_build_tiles(tr_ctx* ctx) {
    if (ctx->tiles_ready) return TR_OK;
    if (ctx->W <= 0) return tr_fail(ctx, TR_ERR_INVALID, "film not created");
    int ntx = (ctx->W + TR_TILE - 1) / TR_TILE, nty = (ctx->H + TR_TILE - 1) / TR_TILE;
    std::vector<int> tiles, slot_of((size_t)ntx * nty, -1);      // slot_of: tile -> ordinal in this rank's list (BDPT film pass)
    for (int ty = 0; ty < nty; ++ty) for (int tx = 0; tx < ntx; ++tx)
        if ((tx + 3 * ty) % ctx->nranks == ctx->rank) { slot_of[(size_t)ty * ntx + tx] = (int)tiles.size(); tiles.push_back(ty * ntx + tx); }
    ctx->n_local_tiles = (int)tiles.size();
    int rc; if ((rc = tr_realloc(ctx, &ctx->d_tiles, tiles.size()))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_bd_tile_slot, slot_of.size()))) return rc;
    if ((rc = tr_stage_h2d(ctx, ctx->d_bd_tile_slot, slot_of.data(), slot_of.size() * 4))) return rc;
    if (!tiles.empty() && (rc = tr_stage_h2d(ctx, ctx->d_tiles, tiles.data(), tiles.size() * 4))) return rc;
    ctx->tiles_ready = true;
    return TR_OK;
}
