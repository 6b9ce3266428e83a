"""Ray setup, sample points for the radiance-field callback and depth re-projection
 against fixtures produced by the reference's own utils module."""
import pytest
import torch

import avr_oracle as O
from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu


def test_world_rays_golden(dev):
    from avr_b200 import ops
    g = load_golden("geometry")
    ros, rds = ops.world_rays(g["x_pix"].to(dev), g["intrinsics"].to(dev), g["cam2world"].to(dev))
    assert torch.equal(ros.cpu(), g["ref_ros"])                      # a copy of pose[:3, 3]
    assert_close(rds, g["ref_rds"], rtol=1e-5, atol=1e-6, what="rds")
    assert_close(rds.norm(dim=-1), torch.ones(2, 75), rtol=1e-6, atol=1e-6, what="unit directions")


@pytest.mark.parametrize("k", [12, 7, 64, 1])
def test_ray_points_bit_exact_and_grad(k, dev):
    from avr_b200 import ops
    g = load_golden("geometry")
    ros, rds = g["ref_ros"], g["ref_rds"]
    gen = torch.Generator().manual_seed(k)
    z = g["z"] if k == 12 else torch.sort(0.8 + torch.rand(2, 75, k, generator=gen), -1).values
    zd = z.to(dev).requires_grad_(True)
    pts, vd = ops.ray_points(ros.to(dev), rds.to(dev), zd)
    want = ros.unsqueeze(-2) + rds.unsqueeze(-2) * z.unsqueeze(-1)     # renderers.py:171: mul, then add
    assert torch.equal(pts.cpu(), want)
    if k == 12:
        assert torch.equal(pts.cpu(), g["ref_pts"])
        assert torch.equal(vd.reshape(2, -1, 3).cpu(), g["ref_viewdirs"])
    assert torch.equal(vd.cpu(), rds.unsqueeze(-2).expand(2, 75, k, 3))
    gp = torch.randn(2, 75, k, 3, generator=gen)
    pts.backward(gp.to(dev))
    assert_close(zd.grad, (gp * rds.unsqueeze(-2)).sum(-1), rtol=1e-5, atol=1e-6, what="d_z")


def test_ray_points_grad_to_rays(dev):
    """Nobody on the reference's path differentiates w.r.t. the rays, but autograd must stay right."""
    from avr_b200 import ops
    gen = torch.Generator().manual_seed(3)
    ros = torch.randn(1, 9, 3, generator=gen).to(dev).requires_grad_(True)
    rds = torch.randn(1, 9, 3, generator=gen).to(dev).requires_grad_(True)
    z = torch.rand(1, 9, 5, generator=gen).to(dev).requires_grad_(True)
    gp, gv = torch.randn(1, 9, 5, 3, generator=gen).to(dev), torch.randn(1, 9, 5, 3, generator=gen).to(dev)
    pts, vd = ops.ray_points(ros, rds, z)
    torch.autograd.backward([pts, vd], [gp, gv])
    o2, d2, z2 = (t.detach().clone().requires_grad_(True) for t in (ros, rds, z))
    p2 = o2.unsqueeze(-2) + d2.unsqueeze(-2) * z2.unsqueeze(-1)
    v2 = d2.unsqueeze(-2).expand(1, 9, 5, 3)
    torch.autograd.backward([p2, v2], [gp, gv])
    for a, b, what in ((ros, o2, "d_ros"), (rds, d2, "d_rds"), (z, z2, "d_z")):
        assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-6, what=what)


@pytest.mark.parametrize("k,per_ray", [(64, False), (20, True), (7, True), (96, False)])
def test_coarse_sample_points_matches_separate_kernels(k, per_ray, dev):
    from avr_b200 import ops
    g = load_golden("geometry")
    ros, rds = g["ref_ros"].to(dev), g["ref_rds"].to(dev)
    gen = torch.Generator().manual_seed(k)
    u = torch.rand(2, 75, k, generator=gen)
    if per_ray:
        d = 0.9 + 0.8 * torch.rand(2, 75, generator=gen)
        near, far, stride = (d - 0.15).reshape(-1), (d + 0.15).reshape(-1), 1
        want_z = O.coarse_z(d - 0.15, d + 0.15, k, u)
    else:
        near, far, stride = torch.tensor([0.8]), torch.tensor([1.8]), 0
        want_z = O.coarse_z(near.expand(2, 75), far.expand(2, 75), k, u)
    z, pts, vd = ops.coarse_sample_points(near.to(dev), far.to(dev), stride, u.to(dev), ros, rds)
    assert torch.equal(z.cpu(), want_z)                               # same bits as sample_coarse
    assert torch.equal(pts.cpu(), g["ref_ros"].unsqueeze(-2) + g["ref_rds"].unsqueeze(-2) * want_z.unsqueeze(-1))
    assert torch.equal(vd.cpu(), g["ref_rds"].unsqueeze(-2).expand(2, 75, k, 3))


def test_depth_from_world_golden_and_grad(dev):
    from avr_b200 import ops
    g = load_golden("geometry")
    ros, rds, c2w = g["ref_ros"].to(dev), g["ref_rds"].to(dev), g["cam2world"].to(dev)
    dist = g["dist"].to(dev).requires_grad_(True)
    depth = ops.depth_from_world(ros, rds, dist, c2w)
    assert depth.shape == (2, 75)
    assert_close(depth, g["ref_depth"], rtol=1e-5, atol=2e-6, what="depth")
    depth.backward(g["g_depth"].to(dev))
    want = (g["ref_d_world"] * g["ref_rds"]).sum(-1)                  # chain rule through p = o + d*dist
    assert_close(dist.grad, want, rtol=1e-5, atol=2e-6, what="d_dist")
    # points given directly (the adaptive renderer's coarse depth, renderers.py:486)
    world = (g["ref_ros"] + g["ref_rds"] * g["dist"].unsqueeze(-1)).to(dev).requires_grad_(True)
    depth2 = ops.depth_from_world(world, None, None, c2w)
    assert_close(depth2, g["ref_depth"], rtol=1e-5, atol=2e-6, what="depth (points)")
    depth2.backward(g["g_depth"].to(dev))
    assert_close(world.grad, g["ref_d_world"], rtol=1e-5, atol=2e-6, what="d_world")


def test_geometry_full_size_properties(dev):
    """2^20 rays x 96: points lie on their rays, view directions are the rays' directions."""
    from avr_b200 import ops
    r, k = 1 << 20, 96
    gen = torch.Generator(device=dev).manual_seed(0)
    ros = torch.randn(1, r, 3, device=dev, generator=gen)
    rds = torch.nn.functional.normalize(torch.randn(1, r, 3, device=dev, generator=gen), dim=-1)
    u = torch.rand(1, r, k, device=dev, generator=gen)
    near, far = torch.tensor([0.8], device=dev), torch.tensor([1.8], device=dev)
    z, pts, vd = ops.coarse_sample_points(near, far, 0, u, ros, rds)
    assert torch.equal(z, ops.coarse_sample_raw(near, far, 0, u))
    assert torch.equal(pts, ros.unsqueeze(-2) + rds.unsqueeze(-2) * z.unsqueeze(-1))
    assert torch.equal(vd, rds.unsqueeze(-2).expand(1, r, k, 3))
    p2, v2 = ops.ray_points(ros, rds, z)
    assert torch.equal(p2, pts) and torch.equal(v2, vd)


def test_ray_points_packed(dev):
    """Packed rays (counts 0..300): points lie on their rays bit for bit, d_z = g_pts . rds."""
    from avr_b200 import ops
    gen = torch.Generator().manual_seed(9)
    r = 700
    counts = torch.randint(0, 301, (r,), generator=gen)
    counts[::50] = 0
    offsets = torch.zeros(r + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(counts, 0)
    s = int(offsets[-1])
    ros, rds = torch.randn(r, 3, generator=gen), torch.nn.functional.normalize(torch.randn(r, 3, generator=gen), dim=-1)
    z = torch.rand(s, generator=gen) + 0.5
    seg = torch.repeat_interleave(torch.arange(r), counts)
    zd = z.to(dev).requires_grad_(True)
    pts, vd = ops.ray_points_packed(ros.to(dev), rds.to(dev), zd, offsets.to(dev))
    assert torch.equal(pts.cpu(), ros[seg] + rds[seg] * z.unsqueeze(-1))
    assert torch.equal(vd.cpu(), rds[seg])
    gp = torch.randn(s, 3, generator=gen)
    pts.backward(gp.to(dev))
    assert_close(zd.grad, (gp * rds[seg]).sum(-1), rtol=1e-5, atol=1e-6, what="d_z")
