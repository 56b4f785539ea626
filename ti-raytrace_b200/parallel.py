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
    """Sum the per-rank partial films. `film` is a torch tensor (device film alias on GPUs, CPU tensor
    under gloo). In place; after the call rank `dst` (or every rank) holds the full image."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return film
    if all_ranks:
        dist.all_reduce(film, op=dist.ReduceOp.SUM)
    else:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


def reduce_device_film(ctx, dst=0, all_ranks=False):
    """One NCCL reduce of the device film over NVLink, ordered after the context's render stream."""
    import torch
    ctx.synchronize()
    t = film_tensor(ctx)
    reduce_film(t, dst, all_ranks)
    torch.cuda.synchronize(t.device)
    return t
