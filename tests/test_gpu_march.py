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
    phi.encode(images, cam2world[:, :1], 22.0)
    g_out = torch.randn(sb, r, 3, generator=torch.Generator().manual_seed(5)) * g_scale
    # oracle: the reference's loop on the CPU
    world = O.lstm_march(ros, rds, init, phi, lstm, out_layer, steps)
    (world * g_out).sum().backward()
    want = {k: p.grad.clone() for k, p in list(lstm.named_parameters()) + list(out_layer.named_parameters())}
    want_conv = phi.encoder.conv.weight.grad.clone()

    phi_d = copy.deepcopy(phi).to(dev)
    for p in phi_d.parameters():
        p.grad = None
    lstm_d, out_d = copy.deepcopy(lstm).to(dev), copy.deepcopy(out_layer).to(dev)
    for p in list(lstm_d.parameters()) + list(out_d.parameters()):
        p.grad = None
    phi_d.encode(images.to(dev), cam2world[:, :1].to(dev), 22.0)
    avr_b200.fuse_field_inputs(phi_d)
    assert avr_b200.march.march_supported(phi_d, lstm_d, out_d)
    got = avr_b200.lstm_march(ros.to(dev), rds.to(dev), init.to(dev), phi_d, lstm_d, out_d, steps)
    assert_close(got, world, rtol=2e-5, atol=2e-6, what="world_coords[-1]")
    (got * g_out.to(dev)).sum().backward()
    for k, p in list(lstm_d.named_parameters()) + list(out_d.named_parameters()):
        scale = max(want[k].abs().max().item(), 1e-12)
        assert_close(p.grad / scale, want[k] / scale, rtol=1e-4, atol=2e-5, what=f"grad {k}")
    scale = want_conv.abs().max().item()
    assert_close(phi_d.encoder.conv.weight.grad / scale, want_conv / scale, rtol=1e-4, atol=2e-5, what="grad encoder (d_latent)")
    if g_scale > 1:
        # the case exists to exercise the hook: without the clamp the gradients are different
        lstm2, out2 = copy.deepcopy(lstm), copy.deepcopy(out_layer)
        for p in list(lstm2.parameters()) + list(out2.parameters()) + list(phi.parameters()):
            p.grad = None
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
