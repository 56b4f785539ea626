"""Multi-GPU plumbing: one process per GPU, framebuffer tiles sharded across ranks, one sum-reduce of
the accumulated image per sample batch (the reference itself is single-device).

Sharding rule (same as csrc/api.cu tr_build_tiles): the image is cut into 32x32 tiles; tile (tx, ty)
belongs to rank (tx + 3*ty) % nranks.  Every rank holds a full-size hdr film whose foreign pixels
stay 0, so a SUM over ranks reconstructs the image; per-pixel RNG is keyed by the global pixel, so the
result does not depend on nranks.  torch.distributed is used only as plumbing (NCCL over NVLink on
GPUs, gloo in the CPU tests)."""
import os
import numpy as np

TILE = 32


def tile_owner(W, H, nranks):
    """(W,H) int array: owning rank of every pixel, index [x][y]"""
    tx = np.arange(W) // TILE
    ty = np.arange(H) // TILE
    return ((tx[:, None] + 3 * ty[None, :]) % nranks).astype(np.int32)


def tile_mask(W, H, rank, nranks):
    return tile_owner(W, H, nranks) == rank


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


class _DevArray:
    """__cuda_array_interface__ view of a raw device pointer (zero-copy hand-off to torch)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


def film_tensor(ctx):
    """torch tensor aliasing the context's device hdr film (W,H,3 f32)."""
    import torch
    hdr_ptr, _ = ctx.film_device_ptr()
    return torch.as_tensor(_DevArray(hdr_ptr, (ctx.W, ctx.H, 3)), device="cuda:%d" % ctx.device)


def reduce_film(film, dst=0, all_ranks=False):
    """Sum per-rank partial films held in torch tensors (CPU tensors under gloo: the host-logic tests; the GPU path is
    comm_init() + ctx.film_reduce() below).  OUT OF PLACE: `film` stays this rank's pure partial, so rendering more frames and
    reducing again never double-counts history.  Returns the summed image on rank `dst` (on every rank with all_ranks), None
    elsewhere."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return film
    out = film.clone()
    if all_ranks:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
        return out
    dist.reduce(out, dst=dst, op=dist.ReduceOp.SUM)
    return out if dist.get_rank() == dst else None


def comm_init(ctx):
    """Create the library-owned NCCL communicator of this rank's context (tr_comm_init): rank 0 makes the unique id, the 128
    bytes travel over the torch.distributed process group (plumbing only), every rank joins.  Also shards the film."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank()
    if world == 1 or not dist.is_initialized():
        ctx.comm_init(0, 1, None)
        return
    uid = ctx.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)
    t = torch.from_numpy(uid)
    if dist.get_backend() == "nccl":
        t = t.cuda(local)
    dist.broadcast(t, src=0)
    ctx.comm_init(rank, world, t.cpu().numpy())


def reduce_device_film(ctx, all_ranks=False):
    """One ncclReduce of the device film over NVLink, enqueued on the context's render stream (tr_film_reduce): no host round
    trip between the last render kernel and the reduce.  Rank 0 (every rank with all_ranks) then presents the summed image
    through tone_map / film_download."""
    ctx.film_reduce(all_ranks)
