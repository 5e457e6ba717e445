"""Multi-GPU host logic on CPU: the pixel-tile dealing of the device image (acn_dimage_set_shard / acn_pixel_owner) and the
per-pass exchange, world_size 2 over gloo.

Each rank traces the samples of ITS pixels (here with the CPU oracle standing in for the device tracer — this test is about
the partition and the reduction, not the kernels), accumulates them per pixel as 64-bit fixed-point sums like
k_img_accumulate / lum_image_s_push (reference src/scene.c:804-813) and all-reduces the sums; the result must equal the
single-process accumulation of all samples TO THE BIT (integer sums, disjoint support), and the tiles must be a balanced
partition.
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

import actinon_b200 as acn  # noqa: E402
import bench  # noqa: E402

TILE = 4


def owners(W, H, n):
    return np.array([[acn.pixel_owner(x, y, TILE, n) for x in range(W)] for y in range(H)])


def test_tiles_are_a_balanced_partition():
    W, H = 64, 48
    for n in (1, 2, 3, 4, 8):
        o = owners(W, H, n)
        assert o.min() == 0 and o.max() == n - 1
        # whole tiles: every pixel of a 4x4 tile has one owner, so all samples of a pixel stay on one rank
        t = o.reshape(H // TILE, TILE, W // TILE, TILE)
        assert (t == t[:, :1, :, :1]).all()
        sizes = np.bincount(o.ravel(), minlength=n)
        assert sizes.max() <= 1.05 * sizes.min()
    # powers of two: every aligned group of n tiles along the Morton curve holds each rank once, so any 16x8 (n = 8)
    # block of pixels — a caustic, a silhouette — is spread over all ranks
    o = owners(64, 64, 8)[::TILE, ::TILE]
    for by in range(0, 16, 2):
        for bx in range(0, 16, 4):
            assert sorted(o[by:by + 2, bx:bx + 4].ravel()) == list(range(8))
    assert acn.pixel_owner(-1, 0, 4, 2) == -1


def test_weak_scaling_sizes():
    for n in (1, 2, 4, 8):
        w, h = bench.scaled_size(400, 400, n)
        assert abs(w * h / (160000 * n) - 1) < 0.01


def fixed_sums(xy, rgb, W, H):
    """Per-pixel Q20.44 sums of float32 sample colours + weight, like k_img_accumulate."""
    acc = np.zeros((H, W, 4), dtype=np.int64)
    x, y = xy[:, 0].astype(int), xy[:, 1].astype(int)
    q = np.rint(rgb.astype(np.float32).astype(np.float64) * acn.PIX_SCALE).astype(np.int64)
    for c in range(3):
        np.add.at(acc, (y, x, c), q[:, c])
    np.add.at(acc, (y, x, 3), 1)
    return acc


def samples(W, H, spp, seed):
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:H, 0:W]
    return np.concatenate([np.stack([(xs + rng.random((H, W))).ravel(), (ys + rng.random((H, W))).ravel()], axis=1) for _ in range(spp)])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.oracle_lib import Oracle
    flat = acn.scenes.primitives(32, 24, 4, 0).flatten()
    xy = samples(32, 24, 3, 7)
    own = np.array([acn.pixel_owner(int(x), int(y), TILE, world) for x, y in xy])
    mine = np.ascontiguousarray(xy[own == rank])
    rgb, _ = Oracle().render(flat, mine, seed_mode=acn.SEED_POSITION_HASH, threads=1)   # seeds depend on geometry only: order-free
    acc = torch.from_numpy(fixed_sums(mine, rgb, 32, 24))
    dist.all_reduce(acc)                                        # the one exchange step of the path
    if rank == 0:
        np.save(out, acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reduce_to_the_single_process_image_bit_for_bit(tmp_path):
    from tests.oracle_lib import Oracle
    world, port = 2, 29631 + os.getpid() % 200
    out = str(tmp_path / "acc.npy")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = np.load(out)
    flat = acn.scenes.primitives(32, 24, 4, 0).flatten()
    xy = samples(32, 24, 3, 7)
    rgb, _ = Oracle().render(flat, xy, seed_mode=acn.SEED_POSITION_HASH, threads=1)
    ref = fixed_sums(xy, rgb, 32, 24)
    assert (got[..., 3] == 3).all()
    assert np.array_equal(got, ref)
