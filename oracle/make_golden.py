"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (build container only).

    python oracle/make_golden.py

Every array named ``ref_*`` in the fixtures is an output of the unmodified
functions/classes in /root/reference/renderers.py (imported through
``oracle/ref_shim.py``) on CPU fp32 (and fp64 where noted); everything else is
an input.  The reference draws its random numbers internally, so each case
seeds torch, runs the reference, then re-seeds and replays the same draw calls
in the reference's order (renderers.py:14, :41, :45, :63; :413 for the
adaptive renderer) to record the numbers it consumed.

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from fields import TinyField, TinyFeatureField, camera_setup  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **conv)
    print(f"{name}: {os.path.getsize(path)/1024:.1f} KiB")


def synth_rgbs(gen, shape_rk, sparse=True):
    """(…,K,4) rgbσ like the MLP's output: sigmoid colours, ReLU density with exact zeros."""
    rgb = torch.sigmoid(torch.randn(*shape_rk, 3, generator=gen))
    if sparse:
        sig = torch.relu(torch.randn(*shape_rk, 1, generator=gen)) * 30.0
    else:
        sig = torch.rand(*shape_rk, 1, generator=gen) * 5.0
    return torch.cat([rgb, sig], -1).contiguous()


def synth_z(gen, shape_rk, near=0.8, dup=0):
    z = near + torch.rand(*shape_rk, generator=gen)
    if dup:
        z[..., :dup] = near
    return torch.sort(z, -1).values.contiguous()


def case_coarse(ref):
    g = torch.Generator().manual_seed(11)
    sb, r = 2, 37
    out = {}
    # VolumeRenderer style: scalar near/far expanded with stride 0 (renderers.py:169)
    near = torch.tensor([0.8]).expand(sb, r)
    far = torch.tensor([1.8]).expand(sb, r)
    for k in (64, 20, 1):
        torch.manual_seed(100 + k)
        z = ref.sample_coarse(near, far, k, device="cpu")
        torch.manual_seed(100 + k)
        u = torch.rand(sb, r, k)
        out[f"u_k{k}"] = u
        out[f"ref_z_k{k}"] = z
    # AdaptiveVolumeRenderer style: per-ray interval d -/+ eps (renderers.py:492)
    d = 0.9 + 0.8 * torch.rand(sb, r, generator=g)
    torch.manual_seed(7)
    z = ref.sample_coarse(d - 0.15, d + 0.15, 20, device="cpu")
    torch.manual_seed(7)
    out["avr_u"] = torch.rand(sb, r, 20)
    out["avr_d"] = d
    out["ref_avr_z"] = z
    save("coarse", **out)


def case_composite(ref):
    g = torch.Generator().manual_seed(22)
    out = {}
    cases = {
        # name: (SB, R, K, dup, sparse, white_back)
        "k96": (1, 96, 96, 16, True, True),      # fine pass incl. the 16 duplicate z=near (renderers.py:255)
        "k64": (2, 48, 64, 0, True, True),       # coarse pass
        "k192": (1, 40, 192, 0, True, True),
        "k20": (1, 64, 20, 0, False, True),      # adaptive renderer
        "k20_noback": (1, 64, 20, 0, True, False),
        "k1": (1, 32, 1, 0, False, True),
        "k7": (1, 33, 7, 0, True, True),         # odd K, ragged-ish tail
        "dense_pos": (1, 64, 96, 0, False, True),  # all-positive sigma for strict gradient checks
    }
    for name, (sb, r, k, dup, sparse, wb) in cases.items():
        z = synth_z(g, (sb, r, k), dup=dup)
        rgbs = synth_rgbs(g, (sb, r, k), sparse=sparse)
        g_rgb = torch.randn(sb, r, 3, generator=g)
        g_depth = torch.randn(sb, r, 1, generator=g)
        g_w = torch.randn(sb, r, k, 1, generator=g)
        for dt, tag in ((torch.float32, ""), (torch.float64, "64")):
            zz = z.to(dt).clone().requires_grad_(True)
            xx = rgbs.to(dt).clone().requires_grad_(True)
            rgb, depth, w = ref.volume_integral(zz, xx[..., 3:4], xx[..., :3], white_back=wb)
            # loss as utils.py:364-377 touches it: rgb and depth only
            torch.autograd.backward([rgb, depth], [g_rgb.to(dt), g_depth.to(dt)], retain_graph=True)
            out[f"{name}_ref{tag}_rgb"] = rgb
            out[f"{name}_ref{tag}_depth"] = depth
            out[f"{name}_ref{tag}_w"] = w
            out[f"{name}_ref{tag}_d_rgbs"] = xx.grad.clone()
            out[f"{name}_ref{tag}_d_z"] = zz.grad.clone()
            xx.grad = None
            zz.grad = None
            torch.autograd.backward([rgb, depth, w], [g_rgb.to(dt), g_depth.to(dt), g_w.to(dt)])
            out[f"{name}_ref{tag}_d_rgbs_gw"] = xx.grad.clone()
            out[f"{name}_ref{tag}_d_z_gw"] = zz.grad.clone()
        out[f"{name}_z"] = z
        out[f"{name}_rgbs"] = rgbs
        out[f"{name}_g_rgb"] = g_rgb
        out[f"{name}_g_depth"] = g_depth
        out[f"{name}_g_w"] = g_w
        out[f"{name}_white_back"] = np.array(wb)
    # known-answer rows
    z = synth_z(g, (1, 8, 16))
    x = synth_rgbs(g, (1, 8, 16))
    x[0, 0, :, 3] = 0.0                # empty ray -> rgb = 1, depth = 0
    x[0, 1, :, 3] = 1e4                # very opaque: weights run through denormals
    x[0, 2, :, 3] = 0.0
    x[0, 2, 5, 3] = 1e6                # single opaque sample -> one-hot weights
    rgb, depth, w = ref.volume_integral(z, x[..., 3:4], x[..., :3])
    out.update(kat_z=z, kat_rgbs=x, kat_ref_rgb=rgb, kat_ref_depth=depth, kat_ref_w=w)
    save("composite", **out)


def case_fine(ref):
    g = torch.Generator().manual_seed(33)
    sb, r, kc, n = 1, 200, 64, 128
    near = torch.tensor([0.8]).expand(sb, r)
    far = torch.tensor([1.8]).expand(sb, r)
    w = (torch.rand(sb, r, kc, 1, generator=g) ** 6)
    w[0, 0] = 0.0                       # all-zero weights -> uniform pdf
    w[0, 1] = 1.0                       # uniform weights
    w[0, 2] = 0.0
    w[0, 2, 17] = 1.0                   # one spike
    torch.manual_seed(3)
    z = ref.sample_fine(near, far, n, w, device="cpu")
    torch.manual_seed(3)
    u = torch.rand(sb, r, n)
    u2 = torch.rand(sb, r, n)
    out = dict(w=w, u=u, u2=u2, ref_z=z)
    # adversarial draws cannot be injected into the reference (it draws inside), so
    # monkeypatch torch.rand for one call: still the reference's code doing the work.
    ua = u.clone()
    ua[:, :, 0] = 0.0
    ua[:, :, 1] = 2.0 ** -24
    ua[:, :, 2] = 1.0 - 2.0 ** -24
    ua[:, :, 3] = 0.5
    real_rand = torch.rand
    torch.rand = lambda *a, **k: ua.clone()
    try:
        torch.manual_seed(4)
        z_adv = ref.sample_fine(near, far, n, w, device="cpu")
        torch.manual_seed(4)
        u2_adv = torch.rand_like(ua)
    finally:
        torch.rand = real_rand
    out.update(u_adv=ua, u2_adv=u2_adv, ref_z_adv=z_adv)
    # per-ray near/far
    d = 0.9 + 0.8 * torch.rand(sb, r, generator=g)
    torch.manual_seed(5)
    z_pr = ref.sample_fine(d - 0.15, d + 0.15, 16, w, device="cpu")
    torch.manual_seed(5)
    out.update(pr_d=d, pr_u=torch.rand(sb, r, 16), pr_u2=torch.rand(sb, r, 16), ref_pr_z=z_pr)
    # the merge step as VolumeRenderer does it (renderers.py:254-258)
    torch.manual_seed(6)
    zc = ref.sample_coarse(near, far, kc, device="cpu")
    zd = torch.clamp(ref.sample_depth(torch.zeros(sb, r, 1), 16, 0.01), torch.tensor([0.8]), torch.tensor([1.8]))
    torch.manual_seed(6)
    uc = torch.rand(sb, r, kc)
    nrm = torch.randn(sb, r, 16)
    zs, _ = torch.sort(torch.cat([zc, z, zd], -1), -1)
    out.update(merge_uc=uc, merge_normals=nrm, ref_merge_zc=zc, ref_merge_zd=zd, ref_merge_sorted=zs)
    save("fine", **out)


def case_volume_renderer(ref):
    sb, r = 2, 64
    cam2world, intrinsics, x_pix = camera_setup(sb, r, seed=1)
    for name, (kc, nf, nd, wb) in {"default": (64, 32, 16, True), "small": (32, 16, 8, False)}.items():
        field = TinyField(seed=2)
        ren = ref.VolumeRenderer(0.8, 1.8, kc, nf, nd, 0.01, white_back=wb)
        ren.near, ren.far = ren.near.cpu(), ren.far.cpu()
        torch.manual_seed(40)
        with torch.no_grad():
            rc, rf, d0, d1 = ren(cam2world, intrinsics, x_pix, field)
        torch.manual_seed(40)
        ki = nf - nd
        draws = [torch.rand(sb, r, kc), torch.rand(sb, r, ki), torch.rand(sb, r, ki), torch.randn(sb, r, nd)]
        # gradients w.r.t. the field's parameters through the whole renderer (loss = utils.py:364-377 style)
        torch.manual_seed(40)
        field.zero_grad()
        rc2, rf2, d2, _ = ren(cam2world, intrinsics, x_pix, field)
        loss = ((rc2 - 0.3) ** 2).mean() + ((rf2 - 0.3) ** 2).mean() + 0.1 * d2.mean()
        loss.backward()
        grads = {f"{name}_ref_grad_{i}": p.grad.clone() for i, p in enumerate(field.parameters())}
        save(f"volume_renderer_{name}", cam2world=cam2world, intrinsics=intrinsics, x_pix=x_pix,
             u_coarse=draws[0], u_cdf=draws[1], u_bin=draws[2], normals=draws[3],
             ref_rgb_coarse=rc, ref_rgb_fine=rf, ref_depth=d0, ref_loss=loss.detach(),
             cfg=np.array([kc, nf, nd, int(wb)]), **grads)


def case_adaptive_renderer(ref):
    sb, r, ch = 1, 48, 32
    cam2world, intrinsics, x_pix = camera_setup(sb, r, seed=3)
    phi = TinyFeatureField(ch, seed=4)
    torch.manual_seed(9)
    ren = ref.AdaptiveVolumeRenderer(ch, raymarch_steps=3, epsilon=0.15, n_coarse=20, white_back=True)
    state = {k: v.clone() for k, v in ren.state_dict().items()}
    torch.manual_seed(50)
    rc, rgb, dc, depth = ren(cam2world, intrinsics, x_pix, phi)
    loss = ((rgb - 0.3) ** 2).mean() + 0.1 * depth.mean() + ((rc - 0.2) ** 2).mean()
    loss.backward()
    torch.manual_seed(50)
    init = torch.zeros((sb, r, 1)).normal_(mean=0.8, std=5e-2)      # renderers.py:413
    u = torch.rand(sb, r, 20)                                      # renderers.py:14 via :492
    save("adaptive_renderer", cam2world=cam2world, intrinsics=intrinsics, x_pix=x_pix,
         init_distance=init, u_coarse=u,
         ref_rgb_coarse=rc, ref_rgb=rgb, ref_depth_coarse=dc, ref_depth=depth,
         **{"state_" + k.replace(".", "__"): v for k, v in state.items()},
         **{"ref_grad_" + k.replace(".", "__"): p.grad.clone() for k, p in ren.named_parameters()},
         **{f"ref_phi_grad_{i}": p.grad.clone() for i, p in enumerate(phi.parameters())})


def case_adaptive_march(ref):
    """the reference's AdaptiveVolumeRenderer (march loop :411-435 and
    tail :489-509) around a radiance field that HAS a feature map — tests/field_stub.py's StubNet,
    whose stock forward is oracle/field_oracle.py (pinned bit for bit to NewPixelNeRFNet.forward) —
    so that the fused march kernel can be held to numbers the reference's renderer produced:
    outputs, and gradients of the LSTM head, of the encoder's parameters (through the feature map,
    i.e. d_latent) and of the MLPs (through the sample positions)."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    from field_stub import StubNet
    sb, r, steps = 2, 40, 4
    cam2world, intrinsics, x_pix = camera_setup(sb, r, seed=13)
    torch.manual_seed(14)
    phi = StubNet()
    g = torch.Generator().manual_seed(15)
    images = torch.rand(sb, 1, 3, 20, 16, generator=g) * 2 - 1
    src_pose = cam2world[:, :1].clone()                                   # the source view is the target view
    phi.encode(images, src_pose, 22.0)
    torch.manual_seed(16)
    ren = ref.AdaptiveVolumeRenderer(128, raymarch_steps=steps, epsilon=0.15, n_coarse=20, white_back=True)
    with torch.no_grad():                                                 # make the march move (default init barely does)
        ren.out_layer.weight.mul_(3.0)
    state = {k: v.clone() for k, v in ren.state_dict().items()}
    torch.manual_seed(51)
    rc, rgb, dc, depth = ren(cam2world, intrinsics, x_pix, phi)
    loss = ((rgb - 0.3) ** 2).mean() + 0.1 * depth.mean() + ((rc - 0.2) ** 2).mean() + 0.05 * dc.mean()
    loss.backward()
    torch.manual_seed(51)
    init = torch.zeros((sb, r, 1)).normal_(mean=0.8, std=5e-2)            # renderers.py:413
    u = torch.rand(sb, r, 20)                                             # renderers.py:14 via :492
    save("adaptive_march", cam2world=cam2world, intrinsics=intrinsics, x_pix=x_pix, images=images, src_pose=src_pose,
         focal=22.0, init_distance=init, u_coarse=u, steps=steps,
         ref_rgb_coarse=rc, ref_rgb=rgb, ref_depth_coarse=dc, ref_depth=depth, ref_loss=loss.detach(),
         **{"state_" + k.replace(".", "__"): v for k, v in state.items()},
         **{"phi_" + k.replace(".", "__"): v.clone() for k, v in phi.state_dict().items()},
         **{"ref_grad_" + k.replace(".", "__"): p.grad.clone() for k, p in ren.named_parameters()},
         **{"ref_phi_grad_" + k.replace(".", "__"): p.grad.clone() for k, p in phi.named_parameters()})


def case_raymarcher(ref):
    """The reference's Raymarcher (renderers.py:292-358: the LSTM march, then colour and depth at the point
    reached) around the same kind of field as case_adaptive_march: outputs and the gradients of the LSTM
    head, of the encoder (through the feature map) and of the coarse MLP."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    from field_stub import StubNet
    sb, r, steps = 2, 33, 5
    cam2world, intrinsics, x_pix = camera_setup(sb, r, seed=23)
    torch.manual_seed(24)
    phi = StubNet()
    g = torch.Generator().manual_seed(25)
    images = torch.rand(sb, 1, 3, 20, 16, generator=g) * 2 - 1
    src_pose = cam2world[:, :1].clone()
    phi.encode(images, src_pose, 22.0)
    torch.manual_seed(26)
    ren = ref.Raymarcher(128, steps)
    with torch.no_grad():
        ren.out_layer.weight.mul_(3.0)
    state = {k: v.clone() for k, v in ren.state_dict().items()}
    torch.manual_seed(52)
    rgb, none, depth, depth2 = ren(cam2world, intrinsics, x_pix, phi)
    assert none is None and depth is depth2
    loss = ((rgb - 0.3) ** 2).mean() + 0.1 * depth.mean()
    loss.backward()
    torch.manual_seed(52)
    init = torch.zeros((sb, r, 1)).normal_(mean=0.8, std=5e-2)            # renderers.py:320
    save("raymarcher", cam2world=cam2world, intrinsics=intrinsics, x_pix=x_pix, images=images, src_pose=src_pose,
         focal=22.0, init_distance=init, steps=steps, ref_rgb=rgb, ref_depth=depth, ref_loss=loss.detach(),
         **{"state_" + k.replace(".", "__"): v for k, v in state.items()},
         **{"phi_" + k.replace(".", "__"): v.clone() for k, v in phi.state_dict().items()},
         **{"ref_grad_" + k.replace(".", "__"): p.grad.clone() for k, p in ren.named_parameters()},
         **{"ref_phi_grad_" + k.replace(".", "__"): p.grad.clone() for k, p in phi.named_parameters() if p.grad is not None})


def case_geometry(ref):
    """utils.get_world_rays / depth_from_world and the point generation of renderers.py:171-175,
    run by the reference's own utils module (star-imported into renderers, renderers.py:1)."""
    g = torch.Generator().manual_seed(31)
    sb, r, k = 2, 75, 12
    cam2world, intrinsics, x_pix = camera_setup(sb, r, seed=5)
    # per-ray poses that differ (a general rigid pose per ray, as the API allows)
    jitter = torch.zeros(sb, r, 4, 4)
    jitter[..., :3, 3] = 0.05 * torch.randn(sb, r, 3, generator=g)
    cam2world = cam2world + jitter
    ros, rds = ref.get_world_rays(x_pix, intrinsics, cam2world)
    z = torch.sort(0.8 + torch.rand(sb, r, k, generator=g), -1).values
    pts = ros.unsqueeze(-2) + rds.unsqueeze(-2) * z.unsqueeze(-1)                # renderers.py:171
    viewdirs = rds.unsqueeze(-2).expand(sb, r, k, -1).reshape(sb, -1, 3)          # :174
    dist = 0.8 + torch.rand(sb, r, generator=g)
    world = (ros + rds * dist.unsqueeze(-1)).requires_grad_(True)
    depth = ref.depth_from_world(world, cam2world)                                # :274-275
    gd = torch.randn(sb, r, generator=g)
    depth.backward(gd)
    save("geometry", cam2world=cam2world, intrinsics=intrinsics, x_pix=x_pix, z=z, dist=dist, g_depth=gd,
         ref_ros=ros, ref_rds=rds, ref_pts=pts, ref_viewdirs=viewdirs, ref_depth=depth, ref_d_world=world.grad)


class _Recorder(torch.nn.Module):
    """Wraps the reference radiance field: records what the renderer asks and what it gets."""

    def __init__(self, net):
        super().__init__()
        self.net, self.calls = net, []

    def forward(self, xyz, viewdirs=None, coarse=True, **kw):
        out = self.net(xyz, coarse=coarse, viewdirs=viewdirs, **kw)
        out.retain_grad()
        self.calls.append((xyz.detach().clone(), viewdirs.detach().clone(), out))
        return out


def case_pixelnerf_replay(ref):
    """BASELINE.json config 1: the reference VolumeRenderer around the reference's own
    NewPixelNeRFNet (conf/default.conf model{} values, random init because there is no network
    for the ImageNet weights), single source view, 128x128 pixel grid subsampled to 512 rays,
    64 coarse + 32 fine (16 importance + 16 depth) samples.  The radiance field cannot travel to
    the GPU box (26 M parameters, reference code), so the fixture records every exchange at the
    renderer <-> field boundary: the points/viewdirs the renderer sent, the (SB, R*K, 4) buffers
    the field returned, and the gradient that came back into them.  tests/test_gpu_renderers.py
    replays the field side."""
    import models
    from ref_shim import Conf

    torch.manual_seed(0)
    conf = Conf(use_encoder=True, use_global_encoder=False, use_xyz=True, canon_xyz=False, use_code=True,
                code=dict(num_freqs=6, freq_factor=1.5, include_input=True), use_viewdirs=True,
                use_code_viewdirs=False, mlp_coarse=dict(type="resnet", n_blocks=3, d_hidden=512),
                mlp_fine=dict(type="resnet", n_blocks=3, d_hidden=512),
                encoder=dict(backbone="resnet34", pretrained=False, num_layers=4))
    net = models.make_new_model(conf)
    sl = 128
    cam2world, intrinsics, _ = camera_setup(1, 1, seed=0)
    x_all = ref.get_opencv_pixel_coordinates(sl, sl).reshape(1, sl * sl, 2)          # utils.py:339-356
    pick = torch.arange(0, sl * sl, 32) + (torch.arange(0, sl * sl, 32) // sl) % 32   # 512 rays over the frame
    x_pix = x_all[:, pick].contiguous()
    r = x_pix.shape[1]
    c2w = cam2world[:, :1].expand(1, r, 4, 4).contiguous()                            # train.py:150
    src = torch.rand(1, 1, 3, sl, sl) * 2 - 1
    focal = torch.tensor(131.25)
    with torch.no_grad():
        net.encode(src, cam2world[:, :1], focal)                                      # train.py:68
    ren = ref.VolumeRenderer.from_conf(Conf(near=0.8, far=1.8, n_coarse=64, n_fine=32, n_fine_depth=16,
                                            depth_std=0.01, white_back=True))
    rec = _Recorder(net)
    torch.manual_seed(77)
    rc, rf, depth, _ = ren(c2w, intrinsics, x_pix, rec)
    torch.manual_seed(77)   # replay the draws the reference consumed (renderers.py:14, :41, :45, :63)
    u_c = torch.rand_like(torch.empty(1, r, 64))
    u_cdf = torch.rand(1, r, 16)
    u_bin = torch.rand_like(u_cdf)
    normals = torch.randn_like(torch.empty(1, r, 16))
    gen = torch.Generator().manual_seed(5)
    g_rc, g_rf, g_d = torch.randn(1, r, 3, generator=gen), torch.randn(1, r, 3, generator=gen), torch.randn(1, r, generator=gen)
    torch.autograd.backward([rc, rf, depth], [g_rc, g_rf, g_d])
    (xyz_c, vd_c, out_c), (xyz_f, vd_f, out_f) = rec.calls
    save("pixelnerf_replay", cam2world=c2w, intrinsics=intrinsics, x_pix=x_pix,
         u_coarse=u_c, u_cdf=u_cdf, u_bin=u_bin, normals=normals, g_rgb_coarse=g_rc, g_rgb_fine=g_rf, g_depth=g_d,
         ref_xyz_coarse=xyz_c, ref_viewdirs_coarse=vd_c, field_out_coarse=out_c, ref_grad_out_coarse=out_c.grad,
         ref_xyz_fine=xyz_f, field_out_fine=out_f, ref_grad_out_fine=out_f.grad,
         ref_rgb_coarse=rc, ref_rgb_fine=rf, ref_depth=depth)


class _CaptureMLP(torch.nn.Module):
    """Stands in for NewPixelNeRFNet.mlp_coarse: keeps the (N, C + code) tensor the front end built."""

    def __init__(self):
        super().__init__()
        self.seen = None

    def forward(self, x, combine_inner_dims=(1,), **kw):
        self.seen = x
        return x.new_zeros(x.shape[0] // combine_inner_dims[0], 4) + 0 * x.sum()


def _source_poses(n, seed):
    """n camera-to-world poses at distance ~1.3 looking at the origin (dataset.py:85-86 flip)."""
    from fields import camera_setup
    return camera_setup(n, 1, seed=seed)[0][:, 0].contiguous()


def _run_field_front_end(net, xyz, viewdirs, g_out, dtype):
    """The reference's forward up to its MLP, and the gradients of <out, g_out>."""
    cap = _CaptureMLP()
    net.mlp_coarse = cap
    xyz = xyz.detach().to(dtype).clone().requires_grad_(True)
    viewdirs = viewdirs.detach().to(dtype).clone().requires_grad_(True)
    lat = net.encoder.latent.detach().to(dtype).requires_grad_(True)
    net.encoder.latent = lat
    net(xyz, coarse=True, viewdirs=viewdirs)
    out = cap.seen
    feats = net(xyz, coarse=True, viewdirs=viewdirs, return_features=True).detach()
    torch.autograd.backward([out], [g_out.to(dtype)])
    return out.detach(), feats, lat.grad, xyz.grad, viewdirs.grad


def case_field_inputs(ref):
    """NewPixelNeRFNet.forward between the renderer's points and the
    MLP (models.py:754-826), run by the reference itself with its MLP replaced by a recorder.

    field_inputs_c512: conf/default.conf's model around its own resnet34 encoder (random init, a
    24x24 source image -> 512 x 12 x 12 features), one object, two source views.
    field_inputs_small: the same code with the module state set by hand — 64 channels on a 8 x 10
    map, two objects x three views, per-object focal length and principal point, normalize_z off,
    and a third of the points projecting outside the map (border clamp).
    fp64 runs of the same reference code are stored as ref64_* (the yardstick for the gradients)."""
    import models
    from ref_shim import Conf

    torch.manual_seed(0)
    conf = Conf(use_encoder=True, use_global_encoder=False, use_xyz=True, canon_xyz=False, use_code=True,
                code=dict(num_freqs=6, freq_factor=1.5, include_input=True), use_viewdirs=True,
                use_code_viewdirs=False, mlp_coarse=dict(type="resnet", n_blocks=3, d_hidden=512),
                mlp_fine=dict(type="resnet", n_blocks=3, d_hidden=512),
                encoder=dict(backbone="resnet34", pretrained=False, num_layers=4))
    net = models.make_new_model(conf)
    g = torch.Generator().manual_seed(11)

    # ---- c512: the real encoder ----------------------------------------------------------
    sb, ns, b, sl = 1, 2, 80, 24
    src = torch.rand(sb, ns, 3, sl, sl, generator=g) * 2 - 1
    c2w = _source_poses(ns, seed=3).reshape(sb, ns, 4, 4)
    with torch.no_grad():
        net.encode(src, c2w, torch.tensor(131.25 / 128 * sl))                  # train.py:68
    xyz = torch.randn(sb, b, 3, generator=g) * 0.25
    vd = torch.nn.functional.normalize(torch.randn(sb, b, 3, generator=g), dim=-1)
    g_out = torch.randn(sb * ns * b, 512 + 42, generator=g)
    state = dict(poses=net.poses.clone(), focal=net.focal.clone(), c=net.c.clone(), image_shape=net.image_shape.clone(),
                 latent=net.encoder.latent.detach().clone(), latent_scaling=net.encoder.latent_scaling.clone(),
                 freqs=net.code._freqs.clone(), phases=net.code._phases.clone())
    out, feats, d_lat, d_xyz, d_vd = _run_field_front_end(net, xyz, vd, g_out, torch.float32)
    save("field_inputs_c512", xyz=xyz, viewdirs=vd, g_out=g_out, ns=ns, normalize_z=1, **state,
         ref_out=out, ref_features=feats, ref_d_latent=d_lat, ref_d_xyz=d_xyz, ref_d_viewdirs=d_vd)

    # ---- small: module state set by hand --------------------------------------------------
    sb, ns, b, ch, h, w = 2, 3, 50, 64, 8, 10
    nv = sb * ns
    c2w = _source_poses(nv, seed=7)
    rot = c2w[:, :3, :3].transpose(1, 2)
    net.poses = torch.cat((rot, -torch.bmm(rot, c2w[:, :3, 3:])), dim=-1)      # models.py:705-707
    net.num_views_per_obj = ns
    net.image_shape = torch.tensor([40.0, 32.0])                               # [W, H]
    net.focal = torch.tensor([[40.0, -40.0], [47.0, -44.0]])                   # per object, y negated (:721-722)
    net.c = torch.tensor([[20.0, 16.0], [18.5, 17.25]])
    net.encoder.latent = torch.randn(nv, ch, h, w, generator=g)
    ls = torch.tensor([float(w), float(h)])
    net.encoder.latent_scaling = ls / (ls - 1) * 2.0                           # :325-327
    net.latent_size = ch
    net.normalize_z = False
    xyz = torch.randn(sb, b, 3, generator=g) * 0.25
    xyz[:, ::3] *= 4.0                                                        # these leave the image
    vd = torch.nn.functional.normalize(torch.randn(sb, b, 3, generator=g), dim=-1)
    g_out = torch.randn(nv * b, ch + 42, generator=g)
    state = dict(poses=net.poses.clone(), focal=net.focal.clone(), c=net.c.clone(), image_shape=net.image_shape.clone(),
                 latent=net.encoder.latent.clone(), latent_scaling=net.encoder.latent_scaling.clone(),
                 freqs=net.code._freqs.clone(), phases=net.code._phases.clone())
    out, feats, d_lat, d_xyz, d_vd = _run_field_front_end(net, xyz, vd, g_out, torch.float32)
    # the same reference code in fp64
    net64 = net.double()
    for name in ("poses", "focal", "c", "image_shape"):
        setattr(net64, name, state[name].double())
    net64.encoder.latent = state["latent"].double()
    net64.encoder.latent_scaling = state["latent_scaling"].double()
    out64, _, d_lat64, d_xyz64, d_vd64 = _run_field_front_end(net64, xyz, vd, g_out, torch.float64)
    save("field_inputs_small", xyz=xyz, viewdirs=vd, g_out=g_out, ns=ns, normalize_z=0, **state,
         ref_out=out, ref_features=feats, ref_d_latent=d_lat, ref_d_xyz=d_xyz, ref_d_viewdirs=d_vd,
         ref64_out=out64, ref64_d_latent=d_lat64, ref64_d_xyz=d_xyz64, ref64_d_viewdirs=d_vd64)


def main():
    ref = ref_shim.load()
    torch.set_num_threads(1)   # one thread: reductions are order-stable across machines
    only = sys.argv[1:]        # optional: names of the cases to regenerate (e.g. field_inputs)
    if only:
        for name in only:
            globals()["case_" + name](ref)
        return
    case_coarse(ref)
    case_composite(ref)
    case_fine(ref)
    case_volume_renderer(ref)
    case_adaptive_renderer(ref)
    case_adaptive_march(ref)
    case_raymarcher(ref)
    case_geometry(ref)
    case_pixelnerf_replay(ref)
    case_field_inputs(ref)


if __name__ == "__main__":
    main()
