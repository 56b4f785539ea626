// comm.cu — the multi-GPU entry points of include/tiray.h: tr_comm_unique_id / tr_comm_init / tr_film_reduce / tr_comm_destroy.
//
// The path shards by framebuffer tiles: scene and BVH are replicated, every rank renders the 32x32 tiles it owns
// into a full-size partial film whose foreign pixels stay 0 (BDPT: plus its light-tracing splats anywhere), and ONE sum-reduce
// of the film per sample batch reconstructs the image on rank 0.  The reduce is an ncclReduce enqueued on the context's own
// stream right behind the last render kernel (no host round trip), out of place: d_hdr stays the rank's pure partial, the sum
// lands in d_hdr_sum on the root, which tone map / download then present.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 — the copy a host process such as torch already mapped, else the system
// one), so libtiray.so has no link-time dependency on it and single-GPU hosts never load it.
#include <dlfcn.h>
#include <string.h>
#include <nccl.h>
#include "ctx.h"

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;

int nccl_load(tr_ctx* ctx) {
    if (g_nccl.handle) return TR_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // already mapped by the host process (e.g. torch's bundled copy)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return tr_fail(ctx, TR_ERR_COMM, "tr_comm: cannot load libnccl.so.2: %s", dlerror());
    NcclApi a; a.handle = h;
    *(void**)&a.GetUniqueId = dlsym(h, "ncclGetUniqueId");
    *(void**)&a.CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void**)&a.CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void**)&a.Reduce = dlsym(h, "ncclReduce");
    *(void**)&a.AllReduce = dlsym(h, "ncclAllReduce");
    *(void**)&a.GetErrorString = dlsym(h, "ncclGetErrorString");
    *(void**)&a.GetVersion = dlsym(h, "ncclGetVersion");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Reduce || !a.AllReduce || !a.GetErrorString)
        return tr_fail(ctx, TR_ERR_COMM, "tr_comm: libnccl.so.2 lacks a required symbol");
    g_nccl = a;
    return TR_OK;
}
}  // namespace

#define TR_NCCL(ctx, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) \
    return tr_fail(ctx, TR_ERR_COMM, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); } while (0)

void tr_comm_release(tr_ctx* ctx) {
    if (ctx && ctx->nccl_comm && g_nccl.CommDestroy) { g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
}

extern "C" {

int tr_comm_unique_id(void* id_out) {
    if (!id_out) return tr_fail(nullptr, TR_ERR_INVALID, "tr_comm_unique_id: NULL output");
    static_assert(sizeof(ncclUniqueId) == TR_COMM_ID_BYTES, "TR_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
    int rc; if ((rc = nccl_load(nullptr))) return rc;
    ncclUniqueId id;
    TR_NCCL(nullptr, g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return TR_OK;
}

int tr_comm_init(tr_ctx* ctx, int rank, int nranks, const void* unique_id) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !unique_id))
        return tr_fail(ctx, TR_ERR_INVALID, "tr_comm_init: bad rank %d / %d or NULL id", rank, nranks);
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    tr_comm_release(ctx);
    int rc;
    if (nranks > 1) {
        if ((rc = nccl_load(ctx))) return rc;
        ncclUniqueId id; memcpy(&id, unique_id, sizeof(id));
        ncclComm_t comm = nullptr;
        TR_NCCL(ctx, g_nccl.CommInitRank(&comm, nranks, id, rank));
        ctx->nccl_comm = comm;
    }
    ctx->comm_rank = rank; ctx->comm_nranks = nranks;
    return tr_set_shard(ctx, rank, nranks);            // the communicator's rank owns the tiles (tx + 3 ty) % nranks == rank
}

int tr_comm_destroy(tr_ctx* ctx) {
    if (!ctx) return TR_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    tr_comm_release(ctx);
    ctx->comm_rank = 0; ctx->comm_nranks = 1; ctx->present_sum = false;
    return TR_OK;
}

int tr_film_reduce(tr_ctx* ctx, int all_ranks) {
    if (!ctx || !ctx->d_hdr) return tr_fail(ctx, TR_ERR_INVALID, "tr_film_reduce: no film");
    if (ctx->comm_nranks <= 1) return TR_OK;             // single rank: the partial film is the image
    if (!ctx->nccl_comm) return tr_fail(ctx, TR_ERR_INVALID, "tr_film_reduce: tr_comm_init has not been called");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t count = (size_t)ctx->W * ctx->H * 3;
    const bool recv = all_ranks || ctx->comm_rank == 0;
    int rc;
    if (recv && (rc = tr_realloc(ctx, &ctx->d_hdr_sum, count))) return rc;
    if (all_ranks) TR_NCCL(ctx, g_nccl.AllReduce(ctx->d_hdr, ctx->d_hdr_sum, count, ncclFloat, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    else TR_NCCL(ctx, g_nccl.Reduce(ctx->d_hdr, ctx->d_hdr_sum, count, ncclFloat, ncclSum, 0, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    ctx->present_sum = recv;
    return TR_OK;
}

}  // extern "C"
