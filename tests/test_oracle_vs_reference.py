"""Pin the oracle: bit-for-bit against the UNMODIFIED reference, imported from
/root/reference (present in the build container only; skipped elsewhere)."""
import pytest
import torch

import avr_oracle as O
import ref_shim
from fields import TinyField, camera_setup

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


def _bounds(sb, r):
    return torch.tensor([0.8]).expand(sb, r), torch.tensor([1.8]).expand(sb, r)


@pytest.mark.parametrize("k", [1, 20, 64])
def test_coarse_bit_exact(ref, k):
    near, far = _bounds(2, 33)
    torch.manual_seed(k)
    want = ref.sample_coarse(near, far, k, device="cpu")
    torch.manual_seed(k)
    u = torch.rand(2, 33, k)
    assert torch.equal(O.coarse_z(near, far, k, u), want)


@pytest.mark.parametrize("k,wb", [(96, True), (64, False), (1, True), (7, True)])
def test_composite_bit_exact_and_grads(ref, k, wb):
    g = torch.Generator().manual_seed(k)
    z = torch.sort(0.8 + torch.rand(1, 50, k, generator=g), -1).values
    x = torch.cat([torch.sigmoid(torch.randn(1, 50, k, 3, generator=g)),
                   torch.relu(torch.randn(1, 50, k, 1, generator=g)) * 30], -1)
    want = ref.volume_integral(z, x[..., 3:4], x[..., :3], white_back=wb)
    got = O.composite_rgbs(z, x, wb)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    g_rgb, g_d = torch.randn(1, 50, 3, generator=g), torch.randn(1, 50, 1, generator=g)
    xx = x.clone().requires_grad_(True)
    zz = z.clone().requires_grad_(True)
    rgb, depth, _ = ref.volume_integral(zz, xx[..., 3:4], xx[..., :3], white_back=wb)
    torch.autograd.backward([rgb, depth], [g_rgb, g_d])
    dx, dz = O.composite_grads(z, x, g_rgb, g_d, None, wb, want_dz=True)
    assert torch.equal(dx, xx.grad) and torch.equal(dz, zz.grad)


def test_fine_bit_exact(ref):
    near, far = _bounds(1, 100)
    g = torch.Generator().manual_seed(3)
    w = torch.rand(1, 100, 64, 1, generator=g) ** 6
    torch.manual_seed(9)
    want = ref.sample_fine(near, far, 128, w, device="cpu")
    torch.manual_seed(9)
    u, u2 = torch.rand(1, 100, 128), torch.rand(1, 100, 128)
    assert torch.equal(O.fine_z(near, far, w, u, u2), want)


def test_volume_renderer_bit_exact(ref):
    sb, r = 2, 40
    cam2world, intrinsics, x_pix = camera_setup(sb, r, seed=5)
    field = TinyField(seed=6)
    ren = ref.VolumeRenderer(0.8, 1.8, 64, 32, 16, 0.01, white_back=True)
    ren.near, ren.far = ren.near.cpu(), ren.far.cpu()
    torch.manual_seed(12)
    with torch.no_grad():
        want = ren(cam2world, intrinsics, x_pix, field)
    torch.manual_seed(12)
    draws = O.draw_volume_randoms(sb, r, 64, 32, 16)
    with torch.no_grad():
        got = O.render_volume(cam2world, intrinsics, x_pix, field, 0.8, 1.8, 64, 32, 16, 0.01, True, draws)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
