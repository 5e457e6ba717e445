"""The pass controller on the device (acn_dimage_*) against the host controller (acn_image_*, the restatement
of reference scene.c:804-862,1103-1159).  -m gpu.

The device builds each pass's sample list with a stencil kernel, a prefix sum and LCG skip-ahead; the host loops over the
pixels in raster order and draws the jitter sequentially.  The lists must be identical to the bit, and — because the samples'
colours are accumulated in fixed point on the device — so must the images (pnm hash), however the pixels are sharded."""
import numpy as np
import pytest

import actinon_b200 as acn

pytestmark = pytest.mark.gpu


def host_controller(flat, tracer, passes=None):
    prm = flat.params
    img = acn.Image(prm.image_width, prm.image_height)
    lists = []
    while passes is None or len(lists) < passes:
        xy = img.next_pass(prm)
        if xy.shape[0] == 0:
            break
        lists.append(xy)
        img.push(xy, tracer.render_samples(xy))
    return img, lists


@pytest.mark.parametrize("name,ov", [
    ("primitives", dict(image_width=160, image_height=120, direct_samples=6, path_samples=3, gradient_cycles=6)),
    ("wine_glass", dict(image_width=100, image_height=100, direct_samples=12, path_samples=8, gradient_cycles=5)),
])
def test_device_controller_draws_the_same_samples_and_gives_the_same_image(name, ov):
    flat = acn.scenes.load(name, **ov)
    prm = flat.params
    t = acn.Tracer(flat, acn.Options())
    img_h, lists = host_controller(flat, t)
    d = acn.DeviceImage(prm.image_width, prm.image_height)
    k = 0
    while True:
        nl, nt = d.render_pass(t, prm)
        if nt == 0:
            break
        assert nl == nt == lists[k].shape[0]
        assert np.array_equal(d.pass_xy(nl), lists[k])            # bit-identical sample positions, same order
        k += 1
    assert k == len(lists) == prm.gradient_cycles + 1
    img_d = d.download()
    assert img_d.cycle == img_h.cycle and img_d.rval == img_h.rval
    # the host sums doubles, the device Q20.44 integers: a sample value below 2^-20 (or a jittered position, whose 53-bit
    # mantissa reaches below 2^-44) is rounded to 2^-45 before it is added; everything else is exact
    sd, sh = img_d.sums(), img_h.sums()
    assert np.array_equal(sd[..., 5], sh[..., 5])
    assert np.abs(sd - sh).max() < 1e-9
    assert np.array_equal(img_d.average(), img_h.average())
    assert img_d.write_pnm(None) == img_h.write_pnm(None)
    d.close(); t.close()


def test_sharded_passes_sum_to_the_unsharded_image():
    """Three 'ranks' on one GPU: each traces its own Morton-dealt 4x4 pixel tiles, the pass deltas are summed (here through
    torch tensors, as the multi-process path all-reduces them over NCCL) — the image must be bit-identical to one rank's."""
    import torch
    flat = acn.scenes.load("diamond", image_width=120, image_height=120, direct_samples=8, path_samples=6, gradient_cycles=4)
    prm = flat.params
    t = acn.Tracer(flat, acn.Options())
    one = acn.DeviceImage(120, 120)
    while one.render_pass(t, prm)[1]:
        pass
    ref = one.download()
    R = 3
    imgs = [acn.DeviceImage(120, 120) for _ in range(R)]
    for r, d in enumerate(imgs):
        d.set_shard(R, r, 4)
    bufs = [torch.zeros(imgs[0].words, dtype=torch.int64, device="cuda") for _ in range(R)]
    while True:
        counts = [d.render_pass(t, prm) for d in imgs]
        if counts[0][1] == 0:
            break
        assert len({c[1] for c in counts}) == 1 and sum(c[0] for c in counts) == counts[0][1]     # same pass everywhere, disjoint shares
        for d, b in zip(imgs, bufs):
            d.copy_delta(b)
        torch.cuda.synchronize()
        total = bufs[0] + bufs[1] + bufs[2]
        for d in imgs:
            d.set_delta(total); d.end_pass()
    for d in imgs:
        got = d.download()
        assert np.array_equal(got.sums(), ref.sums()) and got.write_pnm(None) == ref.write_pnm(None)
        d.close()
    one.close(); t.close()


def test_resume_from_a_host_checkpoint(tmp_path):
    flat = acn.scenes.load("primitives", image_width=96, image_height=72, direct_samples=4, path_samples=0, gradient_cycles=4)
    prm = flat.params
    t = acn.Tracer(flat, acn.Options())
    a = acn.DeviceImage(96, 72)
    while a.render_pass(t, prm)[1]:
        pass
    b = acn.DeviceImage(96, 72)
    for _ in range(2):
        b.render_pass(t, prm)
    ck = b.download(); ck.save(str(tmp_path / "ck.lum")); b.close()
    c = acn.DeviceImage(96, 72)
    c.upload(acn.Image.load(str(tmp_path / "ck.lum")))
    assert c.cycle == 2
    while c.render_pass(t, prm)[1]:
        pass
    assert c.download().write_pnm(None) == a.download().write_pnm(None)
    a.close(); c.close(); t.close()


def _group_image(flat, devices, options=None):
    g = acn.Group(flat, devices, options)
    n_samples, n_pass, rays = g.render()
    img = g.download()
    peer = g.uses_peer_access
    g.close()
    return img, n_samples, n_pass, rays, peer


def test_group_of_three_ranks_on_one_device_is_bit_identical_to_one_rank():
    """acn_group_* (one process, a worker thread + tracer + device image per rank, deltas exchanged through the owners'
    memory): three ranks sharing GPU 0 must produce the image of a single rank, to the bit."""
    flat = acn.scenes.load("wine_glass", image_width=96, image_height=96, direct_samples=10, path_samples=6, gradient_cycles=4)
    one, n1, p1, r1, _ = _group_image(flat, [0])
    three, n3, p3, r3, _ = _group_image(flat, [0, 0, 0])
    assert (n1, p1) == (n3, p3) and p1 == 5
    assert r1 == r3                                       # the same rays were traced, only by other ranks
    assert np.array_equal(one.sums(), three.sums())
    assert one.write_pnm(None) == three.write_pnm(None)


def test_group_over_all_gpus_of_the_box_is_bit_identical():
    n = acn.device_count()
    if n < 2:
        pytest.skip("one GPU")
    flat = acn.scenes.load("diamond", image_width=128, image_height=128, direct_samples=8, path_samples=6, gradient_cycles=3)
    one, n1, p1, r1, _ = _group_image(flat, [0])
    many, nn, pn, rn, peer = _group_image(flat, list(range(n)))
    print(f"{n} GPUs, peer access {peer}")
    assert (n1, p1, r1) == (nn, pn, rn)
    assert np.array_equal(one.sums(), many.sums())
