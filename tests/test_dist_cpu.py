"""CPU, world_size 2 over gloo: the multi-rank host logic (tile sharding + one sum-reduce of the film).
Each rank renders only the pixels of its tiles (here with the CPU oracle standing in for the device
film), the partial films are summed with the product's parallel.reduce_film(), and rank 0 must hold
exactly the single-rank image."""
import os
import socket
import sys
import numpy as np
import pytest

from conftest import ROOT, PKG, model


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, all_ranks, bdpt=False):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    import parallel
    from oracle import oracle, objload
    r, w, _ = parallel.init_process_group("gloo")
    assert (r, w) == (rank, world)
    W, H = 96, 64
    t = objload.load_scene([model("cornell_box.obj")])
    s = oracle.OracleScene(t).build()
    cam = oracle.fit_camera(t, W, H)
    s.set_camera(cam[1], cam[2], *cam[3:])
    if bdpt:
        # BDPT: a rank traces only its tiles' paths but its light-tracing splats land on ANY pixel, so every rank's film is
        # dense and the exchange has to be a sum (SURVEY 8e "Exception"); float summation order differs -> allclose
        s.set_camera_view(cam[0], W, H)
        part, _ = s.render_bdpt_rgb(W, H, 0, 2, seed=3, mask=parallel.tile_mask(W, H, rank, world))
        foreign = part[~parallel.tile_mask(W, H, rank, world)]
        assert foreign.any(), "no splat crossed a tile boundary: the test would not exercise the sum"
    else:
        part, _ = s.render_pt_rgb(W, H, 0, 2, seed=3, mask=parallel.tile_mask(W, H, rank, world))
    film = torch.from_numpy(part.copy())
    out = parallel.reduce_film(film, dst=0, all_ranks=all_ranks)
    assert np.array_equal(film.numpy(), part), "the reduce must leave the rank's partial film untouched"
    assert (out is not None) == (rank == 0 or all_ranks)
    if out is not None:
        if bdpt:
            full, _ = s.render_bdpt_rgb(W, H, 0, 2, seed=3)
            assert np.allclose(out.numpy(), full, rtol=1e-4, atol=1e-6)
        else:
            full, _ = s.render_pt_rgb(W, H, 0, 2, seed=3)
            assert np.array_equal(out.numpy(), full)
    if not bdpt:
        # progressive use: two more frames accumulated into the SAME partial film, reduced again -> no double counting
        part2, _ = s.render_pt_rgb(W, H, 2, 2, seed=3, hdr=part.copy(), mask=parallel.tile_mask(W, H, rank, world))
        out2 = parallel.reduce_film(torch.from_numpy(part2), dst=0, all_ranks=all_ranks)
        if out2 is not None:
            full2, _ = s.render_pt_rgb(W, H, 0, 4, seed=3)
            assert np.array_equal(out2.numpy(), full2)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("all_ranks", [False, True])
def test_two_rank_tile_shard_and_reduce(all_ranks):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), all_ranks), nprocs=2, join=True)


def test_two_rank_bdpt_splats_need_a_sum():
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), False, True), nprocs=2, join=True)
