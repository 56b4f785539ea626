// api.cu — context, uploads/downloads, camera, film and sharding entry points of include/tiray.h
#include <stdarg.h>
#include <string.h>
#include <vector>
#include "ctx.h"
#include "common.cuh"

static std::string g_last_error;

int tr_fail(tr_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (ctx) ctx->err = buf;
    g_last_error = buf;
    return code;
}

extern "C" {

const char* tr_last_error(tr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int tr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int tr_ctx_create(int device, tr_ctx** out) {
    if (!out) return tr_fail(nullptr, TR_ERR_INVALID, "tr_ctx_create: out is NULL");
    *out = nullptr;
    int n = tr_device_count();
    if (n <= 0) return tr_fail(nullptr, TR_ERR_NO_DEVICE, "no CUDA device visible: libtiray has no CPU fallback");
    if (device < 0 || device >= n) return tr_fail(nullptr, TR_ERR_INVALID, "device %d out of range (%d devices)", device, n);
    tr_ctx* ctx = new tr_ctx();
    ctx->device = device;
    TR_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    TR_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    if (prop.major < 10) {
        int rc = tr_fail(nullptr, TR_ERR_NO_DEVICE, "device %d is sm_%d%d; libtiray is built for sm_100a only", device, prop.major, prop.minor);
        delete ctx; return rc;
    }
    TR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = ctx->stream;
    TR_CUDA(ctx, cudaEventCreate(&ctx->ev0));
    TR_CUDA(ctx, cudaEventCreate(&ctx->ev1));
    TR_CUDA(ctx, cudaMalloc((void**)&ctx->d_ctr, sizeof(TrCounters) * TR_MAX_CHAINS));
    TR_CUDA(ctx, cudaMalloc((void**)&ctx->d_build_status, 16 * sizeof(int)));
    TR_CUDA(ctx, cudaMalloc((void**)&ctx->d_batch_params, 64));
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    memset(&ctx->cam, 0, sizeof(ctx->cam));
    *out = ctx;
    return TR_OK;
}

void tr_ctx_destroy(tr_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
    void* ptrs[] = {ctx->d_vertex, ctx->d_prim, ctx->d_material, ctx->d_shape, ctx->d_light, ctx->d_env,
                    ctx->d_morton_unsorted, ctx->d_keys[0], ctx->d_keys[1], ctx->d_vals[0], ctx->d_vals[1],
                    ctx->d_left, ctx->d_right, ctx->d_parent, ctx->d_boxes, ctx->d_leafcount, ctx->d_flag, ctx->d_pre,
                    ctx->d_build_status, ctx->d_nodes, ctx->d_leaves, ctx->d_leaf_of_prim, ctx->d_shade, ctx->d_hist, ctx->d_hdr, ctx->d_rgb,
                    ctx->d_fh, ctx->d_tiles, ctx->d_path[0][0], ctx->d_path[0][1], ctx->d_path[0][2], ctx->d_path[1][0],
                    ctx->d_path[1][1], ctx->d_path[1][2], ctx->d_hit, ctx->d_cls, ctx->d_shq[0][0], ctx->d_shq[0][1],
                    ctx->d_shq[0][2], ctx->d_shq[1][0], ctx->d_shq[1][1], ctx->d_shq[1][2], ctx->d_Lnee, ctx->d_L, ctx->d_ctr, ctx->d_batch_params, ctx->d_matlin, ctx->d_axis, ctx->d_nodesx, ctx->d_smooth,
                    ctx->d_sensor, ctx->d_spectrum[0], ctx->d_spectrum[1], ctx->d_spectrum[2], ctx->d_spectrum[3], ctx->d_rs_scale, ctx->d_rs_data,
                    ctx->d_sky, ctx->d_matspec, ctx->d_white_point,
                    ctx->d_bd_vb, ctx->d_bd_depths, ctx->d_bd_contrib, ctx->d_bd_splat, ctx->d_bd_items, ctx->d_bd_tile_slot, ctx->d_bd_ctr,
                    ctx->d_bd_sq[0], ctx->d_bd_sq[1], ctx->d_bd_vis};
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

int tr_scene_upload(tr_ctx* ctx, const float* vertex, int nv, const int32_t* prim, int np,
                    const float* material, int nm, const float* shape, int ns,
                    const int32_t* light, int nl, const float bmin[3], const float bmax[3]) {
    if (!ctx || !vertex || !prim || !material || np <= 0 || nm <= 0 || !bmin || !bmax)
        return tr_fail(ctx, TR_ERR_INVALID, "tr_scene_upload: NULL table or empty scene (np=%d nm=%d)", np, nm);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->nv = nv; ctx->np = np; ctx->nm = nm; ctx->ns = ns; ctx->nl = nl;
    for (int k = 0; k < 3; ++k) { ctx->bmin[k] = bmin[k]; ctx->bmax[k] = bmax[k]; }
    int rc;
    if ((rc = tr_realloc(ctx, &ctx->d_vertex, (size_t)nv * 9))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_prim, (size_t)np * 3))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_material, (size_t)nm * 10))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_shape, (size_t)(ns > 0 ? ns : 1) * 10))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_light, (size_t)(nl > 0 ? nl : 1)))) return rc;
    cudaStream_t s = ctx->stream;
    TR_CUDA(ctx, cudaMemcpyAsync(ctx->d_vertex, vertex, (size_t)nv * 9 * 4, cudaMemcpyHostToDevice, s));
    TR_CUDA(ctx, cudaMemcpyAsync(ctx->d_prim, prim, (size_t)np * 3 * 4, cudaMemcpyHostToDevice, s));
    TR_CUDA(ctx, cudaMemcpyAsync(ctx->d_material, material, (size_t)nm * 10 * 4, cudaMemcpyHostToDevice, s));
    if (ns > 0 && shape) TR_CUDA(ctx, cudaMemcpyAsync(ctx->d_shape, shape, (size_t)ns * 10 * 4, cudaMemcpyHostToDevice, s));
    else TR_CUDA(ctx, cudaMemsetAsync(ctx->d_shape, 0, 10 * 4, s));
    if (nl > 0 && light) TR_CUDA(ctx, cudaMemcpyAsync(ctx->d_light, light, (size_t)nl * 4, cudaMemcpyHostToDevice, s));
    TR_CUDA(ctx, cudaStreamSynchronize(s));   // host arrays are only borrowed for the call
    ctx->bvh_ready = false; ctx->shade_ready = false; ctx->fh_ready = false; ctx->matlin_ready = false; ctx->matspec_ready = false; ctx->gen++;
    return TR_OK;
}

int tr_material_upload(tr_ctx* ctx, const float* material, int nm) {
    if (!ctx || !material || nm != ctx->nm) return tr_fail(ctx, TR_ERR_INVALID, "tr_material_upload: nm=%d does not match scene (%d)", nm, ctx ? ctx->nm : -1);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    TR_CUDA(ctx, cudaMemcpyAsync(ctx->d_material, material, (size_t)nm * 10 * 4, cudaMemcpyHostToDevice, ctx->stream));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->matlin_ready = false; ctx->matspec_ready = false;
    return TR_OK;
}

int tr_env_upload(tr_ctx* ctx, const int32_t* rgb, int w, int h, float power) {
    if (!ctx || !rgb || w <= 0 || h <= 0) return tr_fail(ctx, TR_ERR_INVALID, "tr_env_upload: bad image %dx%d", w, h);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc; if ((rc = tr_realloc(ctx, &ctx->d_env, (size_t)w * h))) return rc;
    TR_CUDA(ctx, cudaMemcpyAsync(ctx->d_env, rgb, (size_t)w * h * 4, cudaMemcpyHostToDevice, ctx->stream));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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
    ctx->tiles_ready = false; ctx->fh_ready = false; ctx->gen++;
    return TR_OK;
}

int tr_film_clear(tr_ctx* ctx) {
    if (!ctx || !ctx->d_hdr) return tr_fail(ctx, TR_ERR_INVALID, "tr_film_clear: no film");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    TR_CUDA(ctx, cudaMemsetAsync(ctx->d_hdr, 0, (size_t)ctx->W * ctx->H * 12, ctx->stream));
    TR_CUDA(ctx, cudaMemsetAsync(ctx->d_rgb, 0, (size_t)ctx->W * ctx->H * 12, ctx->stream));
    return TR_OK;
}

int tr_film_download(tr_ctx* ctx, float* hdr, float* rgb) {
    if (!ctx || !ctx->d_hdr) return tr_fail(ctx, TR_ERR_INVALID, "tr_film_download: no film");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t bytes = (size_t)ctx->W * ctx->H * 12;
    if (hdr) TR_CUDA(ctx, cudaMemcpyAsync(hdr, ctx->d_hdr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (rgb) TR_CUDA(ctx, cudaMemcpyAsync(rgb, ctx->d_rgb, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

int tr_film_upload(tr_ctx* ctx, const float* hdr) {
    if (!ctx || !ctx->d_hdr || !hdr) return tr_fail(ctx, TR_ERR_INVALID, "tr_film_upload: no film");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    TR_CUDA(ctx, cudaMemcpyAsync(ctx->d_hdr, hdr, (size_t)ctx->W * ctx->H * 12, cudaMemcpyHostToDevice, ctx->stream));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
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
    if (!strcmp(name, "batch_frames")) ctx->opt_batch_frames = value;
    else if (!strcmp(name, "stage_timing")) ctx->opt_stage_timing = value;
    else if (!strcmp(name, "graph")) ctx->opt_graph = value;
    else if (!strcmp(name, "smem_bvh")) ctx->opt_smem_bvh = value;
    else if (!strcmp(name, "chains")) ctx->opt_chains = value;
    else if (!strcmp(name, "shadow_overlap")) ctx->opt_shadow_overlap = value;
    else if (!strcmp(name, "tail_max")) ctx->opt_tail_max = value;
    else if (!strcmp(name, "tail_chunk")) ctx->opt_tail_chunk = value;
    else if (!strcmp(name, "bdpt_wavefront")) ctx->opt_bdpt_wavefront = value;
    else if (!strcmp(name, "max_paths")) ctx->opt_max_paths = (size_t)value;
    else return tr_fail(ctx, TR_ERR_INVALID, "tr_set_option: unknown option '%s'", name);
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
    TR_CUDA(ctx, cudaMemcpyAsync(ctx->d_bd_tile_slot, slot_of.data(), slot_of.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (!tiles.empty()) TR_CUDA(ctx, cudaMemcpyAsync(ctx->d_tiles, tiles.data(), tiles.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->tiles_ready = true;
    return TR_OK;
}
