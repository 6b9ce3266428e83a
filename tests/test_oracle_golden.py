"""The oracle against the committed fixtures, which hold outputs of the reference itself
(oracle/make_golden.py).  Runs anywhere (CPU).  Tolerance is a few ulp rather than exact
because torch-CPU's exp/sum kernels are ISA-dependent across machines."""
import pytest
import torch

import avr_oracle as O
from conftest import assert_close, load_golden

TIGHT = dict(rtol=2e-6, atol=2e-7)


def test_coarse_golden():
    g = load_golden("coarse")
    near = torch.tensor([0.8]).expand(2, 37)
    far = torch.tensor([1.8]).expand(2, 37)
    for k in (64, 20, 1):
        assert torch.equal(O.coarse_z(near, far, k, g[f"u_k{k}"]), g[f"ref_z_k{k}"])
    d = g["avr_d"]
    assert torch.equal(O.coarse_z(d - 0.15, d + 0.15, 20, g["avr_u"]), g["ref_avr_z"])


COMPOSITE_CASES = ["k96", "k64", "k192", "k20", "k20_noback", "k1", "k7", "dense_pos"]


@pytest.mark.parametrize("name", COMPOSITE_CASES)
def test_composite_golden(name):
    g = load_golden("composite")
    wb = bool(g[f"{name}_white_back"])
    z, x = g[f"{name}_z"], g[f"{name}_rgbs"]
    rgb, depth, w = O.composite_rgbs(z, x, wb)
    assert_close(rgb, g[f"{name}_ref_rgb"], what="rgb", **TIGHT)
    assert_close(depth, g[f"{name}_ref_depth"], what="depth", **TIGHT)
    assert_close(w, g[f"{name}_ref_w"], what="w", **TIGHT)
    # fp64 run of the reference is the accuracy yardstick: fp32 sits well inside the parity bar
    assert_close(rgb, g[f"{name}_ref64_rgb"].float(), what="rgb vs fp64")
    dx, dz = O.composite_grads(z, x, g[f"{name}_g_rgb"], g[f"{name}_g_depth"], None, wb, want_dz=True)
    ref_dx = g[f"{name}_ref_d_rgbs"]
    scale = ref_dx.abs().max().item()
    assert_close(dx / scale, ref_dx / scale, what="d_rgbs", **TIGHT)
    ref_dz = g[f"{name}_ref_d_z"]
    assert_close(dz / ref_dz.abs().max().clamp_min(1.0), ref_dz / ref_dz.abs().max().clamp_min(1.0), what="d_z", **TIGHT)


def test_composite_known_answers():
    g = load_golden("composite")
    rgb, depth, w = O.composite_rgbs(g["kat_z"], g["kat_rgbs"])
    assert_close(rgb, g["kat_ref_rgb"], what="rgb", **TIGHT)
    # empty ray: pure white background, depth 0
    assert torch.equal(rgb[0, 0], torch.ones(3)) and depth[0, 0, 0] == 0 and w[0, 0].abs().sum() == 0
    # very opaque ray: weights fall by 1e-10 per sample through the denormals
    assert w[0, 1, 0, 0] == 1.0 and 0 < w[0, 1, 3, 0] < 1e-29
    # one opaque sample: one-hot weights
    assert w[0, 2, 5, 0] == 1.0 and w[0, 2].sum() == 1.0


def test_fine_golden():
    g = load_golden("fine")
    r = g["w"].shape[1]
    near = torch.tensor([0.8]).expand(1, r)
    far = torch.tensor([1.8]).expand(1, r)
    z, cdf, idx = O.fine_z(near, far, g["w"], g["u"], g["u2"], return_aux=True)
    assert_close(z, g["ref_z"], what="z_fine", **TIGHT)
    assert idx.min() >= 0 and idx.max() <= 64
    # uniform weights: cdf_j = j/K and the bin is floor(u*K) (up to rounding of the cdf)
    assert (idx[0, 1] - torch.floor(g["u"][0, 1] * 64).long()).abs().max() <= 1
    z_adv, _, idx_adv = O.fine_z(near, far, g["w"], g["u_adv"], g["u2_adv"], return_aux=True)
    assert_close(z_adv, g["ref_z_adv"], what="z_fine adversarial", **TIGHT)
    assert (idx_adv[..., 0] == 0).all()                 # u = 0 -> first bin
    assert (idx_adv[..., 2] >= 63).all()                # u = 1 - 2^-24 -> last bin, or one past it
    d = g["pr_d"]
    assert_close(O.fine_z(d - 0.15, d + 0.15, g["w"], g["pr_u"], g["pr_u2"]), g["ref_pr_z"], what="per-ray bounds", **TIGHT)
    zc = O.coarse_z(near, far, 64, g["merge_uc"])
    zd = O.depth_z(g["merge_normals"], 0.01, torch.tensor([0.8]), torch.tensor([1.8]))
    assert torch.equal(zc, g["ref_merge_zc"]) and torch.equal(zd, g["ref_merge_zd"])
    assert (zd == 0.8).all()                            # the sample_depth quirk: all clamp to near
    assert_close(O.merge_sorted(zc, z, zd), g["ref_merge_sorted"], what="merged", **TIGHT)


@pytest.mark.parametrize("name", ["default", "small"])
def test_volume_renderer_golden(name):
    from fields import TinyField

    g = load_golden(f"volume_renderer_{name}")
    kc, nf, nd, wb = [int(v) for v in g["cfg"]]
    field = TinyField(seed=2)
    draws = (g["u_coarse"], g["u_cdf"], g["u_bin"], g["normals"])
    with torch.no_grad():
        rc, rf, d, _ = O.render_volume(g["cam2world"], g["intrinsics"], g["x_pix"], field, 0.8, 1.8, kc, nf, nd, 0.01,
                                       bool(wb), draws)
    assert_close(rc, g["ref_rgb_coarse"], what="rgb_coarse", rtol=1e-5, atol=1e-6)
    assert_close(rf, g["ref_rgb_fine"], what="rgb_fine", rtol=1e-4, atol=1e-5)
    assert_close(d, g["ref_depth"], what="depth", rtol=1e-4, atol=1e-5)


def test_geometry_golden():
    """utils.get_world_rays / depth_from_world and the sample-point generation."""
    g = load_golden("geometry")
    ros, rds = O.world_rays(g["x_pix"], g["intrinsics"], g["cam2world"])
    assert_close(ros, g["ref_ros"], what="ros", **TIGHT)
    assert_close(rds, g["ref_rds"], what="rds", **TIGHT)
    pts = g["ref_ros"].unsqueeze(-2) + g["ref_rds"].unsqueeze(-2) * g["z"].unsqueeze(-1)
    assert torch.equal(pts, g["ref_pts"])
    world = (g["ref_ros"] + g["ref_rds"] * g["dist"].unsqueeze(-1)).requires_grad_(True)
    depth = O.camera_depth(world, g["cam2world"])
    assert_close(depth, g["ref_depth"], what="depth", **TIGHT)
    depth.backward(g["g_depth"])
    assert_close(world.grad, g["ref_d_world"], what="d_world", **TIGHT)


def test_pixelnerf_replay_golden():
    """BASELINE.json config 1: the oracle's renderer against the reference VolumeRenderer run
    around the reference's own PixelNeRF (the field side is replayed from the fixture)."""
    g = load_golden("pixelnerf_replay")
    asked = {}

    def field(xyz, viewdirs=None, coarse=True):
        key = "coarse" if coarse else "fine"
        asked[key] = xyz
        return g[f"field_out_{key}"]

    draws = tuple(g[k] for k in ("u_coarse", "u_cdf", "u_bin", "normals"))
    rc, rf, depth, _ = O.render_volume(g["cam2world"], g["intrinsics"], g["x_pix"], field, 0.8, 1.8, 64, 32, 16, 0.01,
                                       True, draws)
    assert_close(asked["coarse"], g["ref_xyz_coarse"], what="coarse points", **TIGHT)
    assert_close(asked["fine"], g["ref_xyz_fine"], what="fine points", **TIGHT)
    assert_close(rc, g["ref_rgb_coarse"], what="rgb_coarse", **TIGHT)
    assert_close(rf, g["ref_rgb_fine"], what="rgb_fine", **TIGHT)
    assert_close(depth, g["ref_depth"], what="depth", rtol=2e-6, atol=5e-7)


def test_raymarcher_golden_against_the_oracle_march():
    """tests/golden/raymarcher.npz (the reference's Raymarcher, renderers.py:292-358) re-derived on the
    CPU from oracle.lstm_march + the stub field: the fixture travels, and the oracle's loop is the
    reference's (it is also pinned bit for bit where the reference can be imported)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from field_stub import StubNet
    g = load_golden("raymarcher")
    steps = int(g["steps"])
    phi = StubNet()
    phi.load_state_dict({k[len("phi_"):].replace("__", "."): v for k, v in g.items() if k.startswith("phi_")})
    phi.encode(g["images"], g["src_pose"], float(g["focal"]))
    lstm, out_layer = torch.nn.LSTMCell(128, 16), torch.nn.Linear(16, 1)
    lstm.load_state_dict({k[len("state_lstm__"):]: v for k, v in g.items() if k.startswith("state_lstm__")})
    out_layer.load_state_dict({k[len("state_out_layer__"):]: v for k, v in g.items() if k.startswith("state_out_layer__")})
    ros, rds = O.world_rays(g["x_pix"], g["intrinsics"], g["cam2world"])
    world = O.lstm_march(ros, rds, g["init_distance"], phi, lstm, out_layer, steps)
    sb, r = g["x_pix"].shape[:2]
    out = phi(world.reshape(sb, -1, 3), viewdirs=rds.reshape(sb, -1, 3), coarse=True, return_features=False)
    rgb = out[..., :3].reshape(sb, r, 3)
    depth = O.camera_depth(world, g["cam2world"]).reshape(sb, r, -1)
    assert_close(rgb, g["ref_rgb"], rtol=2e-5, atol=2e-6, what="rgb")
    assert_close(depth, g["ref_depth"], rtol=2e-5, atol=2e-6, what="depth")
    (((rgb - 0.3) ** 2).mean() + 0.1 * depth.mean()).backward()
    for mod, prefix in ((lstm, "lstm__"), (out_layer, "out_layer__")):
        for name, p in mod.named_parameters():
            ref = g["ref_grad_" + prefix + name]
            scale = max(ref.abs().max().item(), 1e-9)
            assert_close(p.grad / scale, ref / scale, rtol=1e-4, atol=2e-5, what=f"grad {prefix}{name}")
