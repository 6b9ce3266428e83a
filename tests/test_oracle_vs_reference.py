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


class _KeepInput(torch.nn.Module):
    """Stands in for the reference net's MLP: records the tensor the front end built."""

    def forward(self, x, combine_inner_dims=(1,), **kw):
        self.seen = x
        return x.new_zeros(x.shape[0] // combine_inner_dims[0], 4) + 0 * x.sum()


def test_field_front_end_bit_exact_and_config(ref):
    """oracle/field_oracle.py against NewPixelNeRFNet.forward itself (models.py:754-826) on fresh
    inputs: the MLP's input, the return_features path and the autograd gradients, bit for bit;
    and the host-side logic that reads the launch constants off the reference module."""
    import models
    import field_oracle as FO
    from avr_b200 import field
    from ref_shim import Conf

    torch.manual_seed(4)
    conf = Conf(use_encoder=True, use_global_encoder=False, use_xyz=True, canon_xyz=False, use_code=True,
                code=dict(num_freqs=6, freq_factor=1.5, include_input=True), use_viewdirs=True, use_code_viewdirs=False,
                mlp_coarse=dict(type="resnet", n_blocks=1, d_hidden=32), mlp_fine=dict(type="resnet", n_blocks=1, d_hidden=32),
                encoder=dict(backbone="resnet34", pretrained=False, num_layers=4))
    net = models.make_new_model(conf)
    sb, ns, b, sl = 2, 2, 37, 16
    g = torch.Generator().manual_seed(8)
    c2w = camera_setup(sb * ns, 1, seed=2)[0][:, 0].reshape(sb, ns, 4, 4)
    with torch.no_grad():
        net.encode(torch.rand(sb, ns, 3, sl, sl, generator=g) * 2 - 1, c2w, torch.tensor(131.25 / 128 * sl))
    field._check_supported(net)
    cfg = field._config_of(net)
    assert cfg.ns == ns and len(cfg.freqs) == 12 and cfg.include_input and cfg.normalize_z and cfg.use_viewdirs
    assert cfg.scale == tuple((net.encoder.latent_scaling / net.image_shape).tolist())
    assert cfg.code_width() == net.d_in == 42
    net.mlp_coarse = _KeepInput()
    xyz = (torch.randn(sb, b, 3, generator=g) * 0.3).requires_grad_(True)
    vd = torch.nn.functional.normalize(torch.randn(sb, b, 3, generator=g), dim=-1).requires_grad_(True)
    lat = net.encoder.latent.detach().clone().requires_grad_(True)
    net.encoder.latent = lat
    net(xyz, coarse=True, viewdirs=vd)
    want = net.mlp_coarse.seen
    g_out = torch.randn(want.shape, generator=g)
    want.backward(g_out)
    grads = [t.grad.clone() for t in (xyz, vd, lat)]
    for t in (xyz, vd, lat):
        t.grad = None
    args = (xyz, vd, net.poses, net.focal, net.c, net.image_shape, lat, net.encoder.latent_scaling, net.code._freqs, net.code._phases)
    got = FO.field_inputs(*args, ns=ns)
    assert torch.equal(got, want)
    got.backward(g_out)
    assert all(torch.equal(t.grad, w) for t, w in zip((xyz, vd, lat), grads))
    with torch.no_grad():
        assert torch.equal(FO.field_inputs(*args, ns=ns, features_only=True), net(xyz, viewdirs=vd, return_features=True))
    # configurations outside the kernels' family are refused up front
    net.use_global_encoder = True
    with pytest.raises(field.AvrError):
        field.fuse_field_inputs(net)


def test_lstm_march_bit_exact(ref):
    """oracle.lstm_march against the reference's own loop: Raymarcher.forward (renderers.py:313-351)
    is the loop of AdaptiveVolumeRenderer (:411-435) followed by phi at the last point and
    depth_from_world, so its outputs and its gradients pin the march forward and backward."""
    from fields import TinyFeatureField
    sb, r, ch, steps = 2, 30, 32, 4
    cam2world, intrinsics, x_pix = camera_setup(sb, r, seed=8)
    phi = TinyFeatureField(ch, seed=9)
    torch.manual_seed(21)
    rm = ref.Raymarcher(ch, steps)
    torch.manual_seed(33)
    rgb, _, depth, _ = rm(cam2world, intrinsics, x_pix, phi)
    (rgb.sum() * 30 + depth.sum()).backward()
    want = {k: p.grad.clone() for k, p in rm.named_parameters()}
    want_phi = [p.grad.clone() for p in phi.parameters()]
    for p in list(rm.parameters()) + list(phi.parameters()):
        p.grad = None
    torch.manual_seed(33)
    init = torch.zeros((sb, r, 1)).normal_(mean=0.8, std=5e-2)          # renderers.py:320
    ros, rds = O.world_rays(x_pix, intrinsics, cam2world)
    world = O.lstm_march(ros, rds, init, phi, rm.lstm, rm.out_layer, steps)
    out = phi(world.reshape(sb, -1, 3), viewdirs=rds.reshape(sb, -1, 3), coarse=True, return_features=False)
    got_rgb = out[..., :3].reshape(sb, r, 3)
    got_depth = O.camera_depth(world, cam2world).reshape(sb, r, -1)
    assert torch.equal(got_rgb, rgb) and torch.equal(got_depth, depth)
    (got_rgb.sum() * 30 + got_depth.sum()).backward()
    for k, p in rm.named_parameters():
        assert torch.equal(p.grad, want[k]), k
    for p, w in zip(phi.parameters(), want_phi):
        assert torch.equal(p.grad, w)
