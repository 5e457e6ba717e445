"""Multi-GPU host logic on CPU: the pixel-tile sharding of bench.py and the per-pass reduce, world_size 2 over gloo.

Each rank traces ITS samples (here with the CPU oracle standing in for the device tracer — this test is about the
partition and the reduction, not the kernels), accumulates per-pixel float4 sums like k_accumulate /
lum_image_s_push (reference src/scene.c:804-813) and all-reduces them; the result must equal the single-process
accumulation of all samples, and the tiles must be a partition.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_tiles_partition_the_passes():
    W, H = 40, 24
    for n in (1, 2, 4, 8):
        parts = [bench.rank_samples(W, H, n, r) for r in range(n)]
        allxy = np.concatenate(parts)
        assert len(allxy) == W * H * n                          # every sample of every pass exactly once
        px = allxy[:, 1].astype(int) * W + allxy[:, 0].astype(int)
        assert (np.bincount(px, minlength=W * H) == n).all()    # n passes -> n samples in every pixel
        for r, p in enumerate(parts):                           # all samples of a pixel stay on one rank
            tx, ty = p[:, 0].astype(int) // bench.TILE, p[:, 1].astype(int) // bench.TILE
            assert ((tx + ty) % n == r).all()
    for n in (2, 4, 8):                                         # weak scaling: near-equal work at the benchmark's image size
        tx, ty = np.meshgrid(np.arange(400 // bench.TILE), np.arange(400 // bench.TILE))
        sizes = np.bincount(((tx + ty) % n).ravel(), minlength=n)
        assert sizes.max() <= 1.05 * sizes.min()


def accumulate(xy, rgb, W, H):
    acc = np.zeros((H, W, 4), dtype=np.float64)
    x, y = xy[:, 0].astype(int), xy[:, 1].astype(int)
    np.add.at(acc, (y, x, 0), rgb[:, 0]); np.add.at(acc, (y, x, 1), rgb[:, 1]); np.add.at(acc, (y, x, 2), rgb[:, 2])
    np.add.at(acc, (y, x, 3), 1.0)
    return acc


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import actinon_b200 as acn
    from tests.oracle_lib import Oracle
    sc = acn.scenes.primitives(32, 24, 4, 0)
    flat = sc.flatten()
    xy = bench.rank_samples(32, 24, world, rank)
    rgb, _ = Oracle().render(flat, xy, seed_mode=acn.SEED_POSITION_HASH, threads=1)   # seeds depend on geometry only: order-free
    acc = torch.from_numpy(accumulate(xy, rgb, 32, 24))
    dist.all_reduce(acc)                                        # the one exchange step of the path
    if rank == 0:
        np.save(out, acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reduce_to_the_single_process_image(tmp_path):
    import actinon_b200 as acn
    from tests.oracle_lib import Oracle
    world, port = 2, 29631 + os.getpid() % 200
    out = str(tmp_path / "acc.npy")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = np.load(out)
    flat = acn.scenes.primitives(32, 24, 4, 0).flatten()
    xy = np.concatenate([bench.pass_positions(32, 24, p) for p in range(world)])
    rgb, _ = Oracle().render(flat, xy, seed_mode=acn.SEED_POSITION_HASH, threads=1)   # seeds depend on geometry only: order-free
    ref = accumulate(xy, rgb, 32, 24)
    assert np.allclose(got[..., 3], world)                      # N passes -> weight N per pixel
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)
