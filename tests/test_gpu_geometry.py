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


@pytest.mark.parametrize("k", [64, 20, 7])
def test_rays_coarse_sample_points_is_the_three_calls_in_one(k, dev):
    """VolumeRenderer's first launch (renderers.py:166-175): rays, coarse depths, points and view
    directions from one kernel, each bit-identical to the separate kernels (and so to the goldens
    those are held to), plus the camera-depth coefficients."""
    from avr_b200 import ops
    g = load_golden("geometry")
    x_pix, intr, c2w = g["x_pix"].to(dev), g["intrinsics"].to(dev), g["cam2world"].to(dev)
    u = torch.rand(2, 75, k, generator=torch.Generator().manual_seed(k)).to(dev)
    near, far = torch.tensor([0.8], device=dev), torch.tensor([1.8], device=dev)
    ros0, rds0, aff0 = ops.world_rays(x_pix, intr, c2w, want_affine=True)
    z0, pts0, vd0 = ops.coarse_sample_points(near, far, 0, u, ros0, rds0)
    ros, rds, aff, z, pts, vd = ops.rays_coarse_sample_points(x_pix, intr, c2w, near, far, 0, u)
    for a, b, what in ((ros, ros0, "ros"), (rds, rds0, "rds"), (aff, aff0, "affine"), (z, z0, "z"), (pts, pts0, "pts"), (vd, vd0, "viewdirs")):
        assert torch.equal(a, b), what
    assert torch.equal(ros.cpu(), g["ref_ros"])
    assert_close(rds, g["ref_rds"], rtol=1e-5, atol=1e-6, what="rds")
    # depth_from_world(ros + rds * t) == A t + B   (utils.py:358-361)
    depth = aff[..., 0] * g["dist"].to(dev) + aff[..., 1]
    assert_close(depth, g["ref_depth"], rtol=1e-5, atol=2e-6, what="camera depth from the affine form")


@pytest.mark.parametrize("r,k,want_dz", [(3000, 96, False), (1200, 20, True), (50, 64, False), (5, 7, True)])
def test_composite_returns_the_camera_depth(r, k, want_dz, dev):
    """utils.depth_from_world folded into the compositing kernels (renderers.py:270-275, 505-509):
    the span kernels (3000 x 96, 1200 x 20 with d_z), the warp-per-ray kernels (tails / tiny batches)
    and the generic kernels return A * dist + B and differentiate through it, against composite +
    depth_from_world as separate steps."""
    import avr_b200
    from avr_b200 import ops
    lib = avr_b200.load_library()
    gen = torch.Generator().manual_seed(r + k)
    z = torch.sort(0.8 + torch.rand(1, r, k, generator=gen), -1).values.to(dev)
    x = torch.cat([torch.sigmoid(torch.randn(1, r, k, 3, generator=gen)), torch.relu(torch.randn(1, r, k, 1, generator=gen)) * 30], -1).to(dev)
    x_pix = torch.rand(1, r, 2, generator=gen).to(dev)
    from fields import camera_setup
    c2w, intr, _ = (t.to(dev) for t in camera_setup(1, r, seed=4))
    ros, rds, aff = ops.world_rays(x_pix, intr, c2w, want_affine=True)
    g_rgb, g_d = torch.randn(1, r, 3, generator=gen).to(dev), torch.randn(1, r, generator=gen).to(dev)
    for generic in (0, 1):
        lib.avr_set_force_generic(generic)
        try:
            xa, za = x.clone().requires_grad_(True), z.clone().requires_grad_(want_dz)
            rgb_a, depth_a, _ = ops.composite(xa, za, True, 1.8, want_w=False, depth_affine=aff)
            torch.autograd.backward([rgb_a, depth_a], [g_rgb, g_d])
            xb, zb = x.clone().requires_grad_(True), z.clone().requires_grad_(want_dz)
            rgb_b, dist_b, _ = ops.composite(xb, zb, True, 1.8, want_w=False)
            depth_b = ops.depth_from_world(ros, rds, dist_b, c2w)
            torch.autograd.backward([rgb_b, depth_b], [g_rgb, g_d])
        finally:
            lib.avr_set_force_generic(0)
        assert torch.equal(rgb_a, rgb_b)
        assert_close(depth_a, depth_b, rtol=1e-5, atol=2e-6, what=f"camera depth (generic={generic})")
        assert_close(xa.grad[..., :3], xb.grad[..., :3], rtol=1e-5, atol=2e-6, what="d_rgb")
        assert_close(xa.grad[..., :-1, 3], xb.grad[..., :-1, 3], rtol=1e-5, atol=2e-6, what="d_sigma[:-1]")
        assert_close(xa.grad[..., -1, 3] / 1e10, xb.grad[..., -1, 3] / 1e10, rtol=1e-5, atol=2e-6, what="d_sigma[-1] / 1e10")
        if want_dz:
            scale = zb.grad.abs().amax(-1, keepdim=True).clamp_min(1.0)
            assert_close(za.grad / scale, zb.grad / scale, rtol=1e-5, atol=2e-5, what="d_z")
