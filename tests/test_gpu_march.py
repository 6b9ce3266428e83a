"""The LSTM ray march kernels (csrc/lstm_march.cu; renderers.py:411-435) against
the oracle's restatement of the loop (pinned bit for bit to the reference's Raymarcher in
tests/test_oracle_vs_reference.py) and against a fixture the reference's AdaptiveVolumeRenderer made."""
import pytest
import torch

import avr_oracle as O
from conftest import assert_close, load_golden
from field_stub import StubNet
from fields import camera_setup

pytestmark = pytest.mark.gpu


def _scene(sb, r, seed, size=(20, 16)):
    cam2world, intrinsics, x_pix = camera_setup(sb, r, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    images = torch.rand(sb, 1, 3, *size, generator=g) * 2 - 1
    ros, rds = O.world_rays(x_pix, intrinsics, cam2world)
    init = 0.8 + 0.05 * torch.randn(sb, r, 1, generator=g)
    return cam2world, images, ros.contiguous(), rds.contiguous(), init


def _head(ch, seed, gain=3.0):
    torch.manual_seed(seed)
    lstm = torch.nn.LSTMCell(ch, 16)
    out_layer = torch.nn.Linear(16, 1)
    with torch.no_grad():
        out_layer.weight.mul_(gain)
    return lstm, out_layer


@pytest.mark.parametrize("sb,r,steps,g_scale", [(2, 37, 4, 1.0),       # groups of 4 rays with a ragged tail
                                                (1, 256, 10, 1.0),     # conf/default.conf's raymarch_steps
                                                (2, 64, 6, 3e3)])      # upstream gradient large enough for the clamp hook
def test_march_matches_oracle(sb, r, steps, g_scale, dev):
    import copy

    import avr_b200
    torch.manual_seed(3)
    phi = StubNet()
    cam2world, images, ros, rds, init = _scene(sb, r, seed=20 + steps)
    lstm, out_layer = _head(128, seed=4)
    phi_d = copy.deepcopy(phi).to(dev)            # before encode(): the feature map is a non-leaf
    phi.encode(images, cam2world[:, :1], 22.0)
    g_out = torch.randn(sb, r, 3, generator=torch.Generator().manual_seed(5)) * g_scale
    # oracle: the reference's loop on the CPU
    world = O.lstm_march(ros, rds, init, phi, lstm, out_layer, steps)
    (world * g_out).sum().backward()
    want = {k: p.grad.clone() for k, p in list(lstm.named_parameters()) + list(out_layer.named_parameters())}
    want_conv = phi.encoder.conv.weight.grad.clone()
    # the march feeds its own output back `steps` times: rounding differences grow from step to step
    # (the reference's fp32 run itself drifts from an fp64 run of the same code).  Yardstick, as for
    # the other ill-conditioned quantities of this suite: the same loop in fp64.
    phi64, lstm64, out64 = copy.deepcopy(phi_d).cpu().double(), copy.deepcopy(lstm).double(), copy.deepcopy(out_layer).double()
    for p in list(phi64.parameters()) + list(lstm64.parameters()) + list(out64.parameters()):
        p.grad = None
    phi64.encode(images.double(), cam2world[:, :1].double(), 22.0)
    world64 = O.lstm_march(ros.double(), rds.double(), init.double(), phi64, lstm64, out64, steps)
    (world64 * g_out.double()).sum().backward()
    want64 = {k: p.grad.clone() for k, p in list(lstm64.named_parameters()) + list(out64.named_parameters())}
    want64_conv = phi64.encoder.conv.weight.grad.clone()

    def held_to_fp64(got_t, ref32, ref64, what, floor):
        """within the 1e-5 / 1e-6 bar of the fp32 oracle, or no further from fp64 than 4x the oracle's
        own fp32 error (plus a floor for quantities the oracle happens to hit exactly)"""
        got_t, ref32 = got_t.detach().cpu().double(), ref32.double()
        scale = max(ref64.abs().max().item(), 1e-12)
        ok = (got_t - ref32).abs() <= 1e-6 * scale + 1e-5 * ref32.abs()
        own = ((ref32 - ref64).abs().max() / scale).item()
        err = ((got_t - ref64).abs().max() / scale).item()
        assert bool(ok.all()) or err <= 4 * own + floor, f"{what}: {err:.3g} of scale from fp64, the fp32 oracle {own:.3g}"

    lstm_d, out_d = copy.deepcopy(lstm).to(dev), copy.deepcopy(out_layer).to(dev)
    for p in list(lstm_d.parameters()) + list(out_d.parameters()):
        p.grad = None
    phi_d.encode(images.to(dev), cam2world[:, :1].to(dev), 22.0)
    avr_b200.fuse_field_inputs(phi_d)
    assert avr_b200.march.march_supported(phi_d, lstm_d, out_d)
    got = avr_b200.lstm_march(ros.to(dev), rds.to(dev), init.to(dev), phi_d, lstm_d, out_d, steps)
    held_to_fp64(got, world, world64, "world_coords[-1]", 2e-6)
    (got * g_out.to(dev)).sum().backward()
    for k, p in list(lstm_d.named_parameters()) + list(out_d.named_parameters()):
        held_to_fp64(p.grad, want[k], want64[k], f"grad {k}", 2e-5)
    held_to_fp64(phi_d.encoder.conv.weight.grad, want_conv, want64_conv, "grad encoder (d_latent)", 2e-5)
    if g_scale > 1:
        # the case exists to exercise the hook: without the clamp the gradients are different
        lstm2, out2 = copy.deepcopy(lstm), copy.deepcopy(out_layer)
        for p in list(lstm2.parameters()) + list(out2.parameters()) + list(phi.parameters()):
            p.grad = None
        phi.encode(images, cam2world[:, :1], 22.0)          # a fresh graph for the second backward
        w = ros + rds * init
        state = None
        for _ in range(steps):
            v = phi(w.reshape(sb, -1, 3), viewdirs=rds.reshape(sb, -1, 3), return_features=True)
            state = lstm2(v.reshape(-1, 128), state)
            w = w + rds * out2(state[0]).view(sb, r, 1)
        (w * g_out).sum().backward()
        assert not torch.allclose(lstm2.weight_ih.grad, want["weight_ih"], rtol=1e-3, atol=0)
    # inference: nothing saved, same result
    with torch.no_grad():
        again = avr_b200.lstm_march(ros.to(dev), rds.to(dev), init.to(dev), phi_d, lstm_d, out_d, steps)
    assert torch.equal(again, got)


def test_adaptive_renderer_with_fused_march_golden(dev):
    """The reference's AdaptiveVolumeRenderer around a field with a feature map (oracle/make_golden.py
    case_adaptive_march): the drop-in takes the fused march (and the fused front end) and must
    reproduce outputs and gradients; with fused_march = False it takes the reference's loop."""
    import avr_b200
    g = load_golden("adaptive_march")
    steps = int(g["steps"])
    phi = StubNet()
    phi.load_state_dict({k[len("phi_"):].replace("__", "."): v for k, v in g.items() if k.startswith("phi_")})
    phi = phi.to(dev)
    phi.encode(g["images"].to(dev), g["src_pose"].to(dev), float(g["focal"]))
    avr_b200.fuse_field_inputs(phi)
    ren = avr_b200.AdaptiveVolumeRenderer(128, raymarch_steps=steps, epsilon=0.15, n_coarse=20, white_back=True)
    ren.load_state_dict({k[len("state_"):].replace("__", "."): v for k, v in g.items() if k.startswith("state_")})
    ren = ren.to(dev)
    args = (g["cam2world"].to(dev), g["intrinsics"].to(dev), g["x_pix"].to(dev), phi)
    draws = (g["init_distance"].to(dev), g["u_coarse"].to(dev))
    results = {}
    for fused in (True, False):
        ren.fused_march = fused
        for p in list(ren.parameters()) + list(phi.parameters()):
            p.grad = None
        phi.encode(g["images"].to(dev), g["src_pose"].to(dev), float(g["focal"]))      # a fresh graph per pass
        rc, rgb, dc, depth = ren(*args, draws=draws)
        assert_close(rc, g["ref_rgb_coarse"], rtol=1e-4, atol=1e-5, what="rgb_coarse")
        assert_close(dc, g["ref_depth_coarse"], rtol=1e-4, atol=1e-5, what="depth_coarse")
        assert_close(rgb, g["ref_rgb"], rtol=1e-4, atol=2e-5, what="rgb")
        assert_close(depth, g["ref_depth"], rtol=1e-4, atol=2e-5, what="depth")
        loss = ((rgb - 0.3) ** 2).mean() + 0.1 * depth.mean() + ((rc - 0.2) ** 2).mean() + 0.05 * dc.mean()
        assert abs(loss.item() - g["ref_loss"].item()) < 1e-5
        loss.backward()
        for k, p in ren.named_parameters():
            ref = g["ref_grad_" + k.replace(".", "__")]
            scale = max(ref.abs().max().item(), 1e-9)
            assert_close(p.grad.cpu() / scale, ref / scale, rtol=1e-3, atol=5e-3, what=f"grad {k} (fused={fused})")
        for k, p in phi.named_parameters():
            ref = g["ref_phi_grad_" + k.replace(".", "__")]
            scale = max(ref.abs().max().item(), 1e-9)
            assert_close(p.grad.cpu() / scale, ref / scale, rtol=1e-3, atol=5e-3, what=f"phi grad {k} (fused={fused})")
        results[fused] = rgb.detach()
    assert_close(results[True], results[False], rtol=1e-4, atol=1e-5, what="fused vs loop")


def test_march_falls_back_for_an_arbitrary_field(dev):
    """A phi without a feature map the kernels know (any callable honouring the callback contract)
    keeps the reference's loop: the golden of the round-1 adaptive test still holds (that test runs
    it); here only the dispatch is checked."""
    import avr_b200
    from fields import TinyFeatureField
    phi = TinyFeatureField(32, seed=4).to(dev)
    ren = avr_b200.AdaptiveVolumeRenderer(32, raymarch_steps=2, epsilon=0.15, n_coarse=20, white_back=True).to(dev)
    assert not avr_b200.march.march_supported(phi, ren.lstm, ren.out_layer)


@pytest.mark.parametrize("ch", [256, 512])
def test_march_at_pixelnerf_channel_counts(ch, dev):
    """conf/default.conf's 512 feature channels (and 256): the kernel against the SAME loop run by torch
    on the GPU (fused front-end kernel for the features, torch's LSTMCell / Linear / autograd for the
    rest — an independent implementation of everything the march kernels add), forward and every
    gradient, with the fp64 CPU oracle as the yardstick for how far two fp32 runs may drift."""
    import copy

    import avr_b200
    from field_stub import MapField
    sb, r, steps = 2, 96, 5
    cam2world, images, ros, rds, init = _scene(sb, r, seed=40)
    phi = MapField(sb, ch, seed=41)
    phi64 = copy.deepcopy(phi).double()
    lstm, out_layer = _head(ch, seed=42, gain=2.0)
    lstm64, out64 = copy.deepcopy(lstm).double(), copy.deepcopy(out_layer).double()
    g_out = torch.randn(sb, r, 3, generator=torch.Generator().manual_seed(43))
    # fp64 and fp32 oracle runs on the CPU
    phi.place(cam2world[:, 0], 36.0)
    world32 = O.lstm_march(ros, rds, init, phi, lstm, out_layer, steps)
    (world32 * g_out).sum().backward()
    phi64.place(cam2world[:, 0].double(), 36.0)
    phi64.encoder.latent_scaling = phi64.encoder.latent_scaling.double()
    phi64.image_shape, phi64.focal, phi64.c = phi64.image_shape.double(), phi64.focal.double(), phi64.c.double()
    world64 = O.lstm_march(ros.double(), rds.double(), init.double(), phi64, lstm64, out64, steps)
    (world64 * g_out.double()).sum().backward()
    # the kernels
    phi_d, lstm_d, out_d = MapField(sb, ch, seed=41).to(dev), copy.deepcopy(lstm).to(dev), copy.deepcopy(out_layer).to(dev)
    for p in list(lstm_d.parameters()) + list(out_d.parameters()):
        p.grad = None
    phi_d.place(cam2world[:, 0].to(dev), 36.0)
    avr_b200.fuse_field_inputs(phi_d)
    assert avr_b200.march.march_supported(phi_d, lstm_d, out_d)
    got = avr_b200.lstm_march(ros.to(dev), rds.to(dev), init.to(dev), phi_d, lstm_d, out_d, steps)
    (got * g_out.to(dev)).sum().backward()

    def held(got_t, ref32, ref64, what):
        scale = max(ref64.abs().max().item(), 1e-12)
        own = ((ref32.double() - ref64).abs().max() / scale).item()
        err = ((got_t.detach().cpu().double() - ref64).abs().max() / scale).item()
        assert err <= 4 * own + 2e-5, f"{what}: {err:.3g} of scale from fp64, the fp32 oracle {own:.3g}"

    held(got, world32, world64, "world_coords[-1]")
    for (k, p), (_, p32), (_, p64) in zip(list(lstm_d.named_parameters()) + list(out_d.named_parameters()),
                                           list(lstm.named_parameters()) + list(out_layer.named_parameters()),
                                           list(lstm64.named_parameters()) + list(out64.named_parameters())):
        held(p.grad, p32.grad, p64.grad, f"grad {k}")
    held(phi_d.map.grad, phi.map.grad, phi64.map.grad, "grad feature map (d_latent)")


def test_raymarcher_golden(dev):
    """The reference's third renderer, Raymarcher (renderers.py:292-358), as a drop-in on the march
    kernels: outputs, the 4-tuple's shape conventions, parameter names and gradients against a fixture
    the reference's own class produced (oracle/make_golden.py case_raymarcher)."""
    import avr_b200
    from ref_shim import Conf
    g = load_golden("raymarcher")
    steps = int(g["steps"])
    phi = StubNet()
    phi.load_state_dict({k[len("phi_"):].replace("__", "."): v for k, v in g.items() if k.startswith("phi_")})
    phi = phi.to(dev)
    avr_b200.fuse_field_inputs(phi)
    ren = avr_b200.Raymarcher.from_conf(Conf(num_feature_channels=128), steps)
    state = {k[len("state_"):].replace("__", "."): v for k, v in g.items() if k.startswith("state_")}
    assert sorted(state) == sorted(ren.state_dict().keys())
    ren.load_state_dict(state)
    ren = ren.to(dev)
    for fused in (True, False):
        ren.fused_march = fused
        for p in list(ren.parameters()) + list(phi.parameters()):
            p.grad = None
        phi.encode(g["images"].to(dev), g["src_pose"].to(dev), float(g["focal"]))
        rgb, none, depth, depth2 = ren(g["cam2world"].to(dev), g["intrinsics"].to(dev), g["x_pix"].to(dev), phi,
                                       draws=(g["init_distance"].to(dev),))
        assert none is None and depth is depth2 and depth.shape == g["ref_depth"].shape
        # five steps with out_layer x 3: the march amplifies rounding from step to step (test_march_matches_oracle
        # measures that drift against fp64), and the colour is a sigmoid of an MLP at the point it reaches
        assert_close(rgb, g["ref_rgb"], rtol=1e-3, atol=5e-4, what="rgb")
        assert_close(depth, g["ref_depth"], rtol=1e-3, atol=1e-4, what="depth")
        loss = ((rgb - 0.3) ** 2).mean() + 0.1 * depth.mean()
        assert abs(loss.item() - g["ref_loss"].item()) < 1e-4
        loss.backward()
        for k, p in ren.named_parameters():
            ref = g["ref_grad_" + k.replace(".", "__")]
            scale = max(ref.abs().max().item(), 1e-9)
            assert_close(p.grad.cpu() / scale, ref / scale, rtol=1e-3, atol=5e-3, what=f"grad {k} (fused={fused})")
        for k, p in phi.named_parameters():
            key = "ref_phi_grad_" + k.replace(".", "__")
            if key in g:
                scale = max(g[key].abs().max().item(), 1e-9)
                assert_close(p.grad.cpu() / scale, g[key] / scale, rtol=1e-3, atol=5e-3, what=f"phi grad {k} (fused={fused})")
