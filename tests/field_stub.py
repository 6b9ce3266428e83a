"""Shared by the CPU and GPU tests of the radiance-field front end (test infrastructure):
a module with the attribute names of the reference's NewPixelNeRFNet that the fused forward
touches (its stock forward is the oracle), and ray-ordered sample points."""
import math

import torch

import field_oracle as FO


def ray_ordered_case(sb=2, ns=2, rays=24, k=48, ch=128, h=12, w=9, seed=0):
    """Points ordered as the renderer sends them — K consecutive samples per ray — so that
    neighbouring rows share a texel cell and the kernels' register caches are actually reused."""
    from fields import camera_setup
    g = torch.Generator().manual_seed(seed)
    nv = sb * ns
    c2w = camera_setup(nv, 1, seed=seed + 1)[0][:, 0]
    rot = c2w[:, :3, :3].transpose(1, 2)
    poses = torch.cat((rot, -torch.bmm(rot, c2w[:, :3, 3:])), dim=-1).contiguous()
    o = torch.randn(sb, rays, 1, 3, generator=g) * 0.05 + torch.tensor([0.9, 0.3, 0.4])
    dirs = torch.nn.functional.normalize(-o + 0.2 * torch.randn(sb, rays, 1, 3, generator=g), dim=-1)
    z = torch.linspace(0.6, 1.4, k).view(1, 1, k, 1)
    xyz = (o + dirs * z).reshape(sb, rays * k, 3).contiguous()
    vd = dirs.expand(sb, rays, k, 3).reshape(sb, rays * k, 3).contiguous()
    ls = torch.tensor([float(w), float(h)])
    phases = torch.zeros(12)
    phases[1::2] = math.pi * 0.5
    return dict(xyz=xyz, viewdirs=vd, poses=poses, focal=torch.tensor([[1.1 * 4 * w, -1.1 * 4 * w]]),
                c=torch.tensor([[2.0 * w, 2.0 * h]]), image_shape=torch.tensor([4.0 * w, 4.0 * h]),
                latent=torch.randn(nv, ch, h, w, generator=g), latent_scaling=ls / (ls - 1) * 2.0,
                freqs=torch.repeat_interleave(1.5 * 2.0 ** torch.arange(0, 6), 2).view(1, -1, 1),
                phases=phases.view(1, -1, 1), ns=ns, g_out=torch.randn(nv * rays * k, ch + 42, generator=g))


class _Code(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.include_input = True
        phases = torch.zeros(12)
        phases[1::2] = math.pi * 0.5
        self.register_buffer("_freqs", torch.repeat_interleave(1.5 * 2.0 ** torch.arange(0, 6), 2).view(1, -1, 1))
        self.register_buffer("_phases", phases.view(1, -1, 1))


class _Encoder(torch.nn.Module):
    index_interp, index_padding = "bilinear", "border"

    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(3, 128, 3, padding=1)
        self.latent = None
        self.latent_scaling = None


class _Mlp(torch.nn.Module):
    def __init__(self, d_in):
        super().__init__()
        self.lin = torch.nn.Linear(d_in, 4)

    def forward(self, x, combine_inner_dims=(1,), combine_index=None, dim_size=None):
        y = self.lin(x)
        ns, b = combine_inner_dims
        return y.reshape(-1, ns, b, 4).mean(1).reshape(-1, 4)       # utils.combine_interleaved, "average"


class StubNet(torch.nn.Module):
    """The attributes of NewPixelNeRFNet that the fused forward touches (models.py:609-866)."""

    use_encoder = use_xyz = use_code = use_viewdirs = normalize_z = True
    use_global_encoder = use_code_viewdirs = stop_encoder_grad = False
    d_out = 4

    def __init__(self):
        super().__init__()
        self.encoder, self.code = _Encoder(), _Code()
        self.mlp_coarse, self.mlp_fine = _Mlp(128 + 42), _Mlp(128 + 42)
        self.num_views_per_obj = 1

    def encode(self, images, poses, focal):                         # models.py:682-737, shortened
        sb, ns = images.shape[:2]
        self.num_views_per_obj = ns
        self.encoder.latent = self.encoder.conv(images.reshape(-1, *images.shape[2:]))
        h, w = self.encoder.latent.shape[-2:]
        ls = torch.tensor([float(w), float(h)], device=images.device)
        self.encoder.latent_scaling = ls / (ls - 1) * 2.0
        poses = poses.reshape(-1, 4, 4)
        rot = poses[:, :3, :3].transpose(1, 2)
        self.poses = torch.cat((rot, -torch.bmm(rot, poses[:, :3, 3:])), dim=-1)
        self.image_shape = torch.tensor([float(images.shape[-1]), float(images.shape[-2])], device=images.device)
        self.focal = torch.tensor([[focal, -focal]], device=images.device)
        self.c = (self.image_shape * 0.5).unsqueeze(0)

    def forward(self, xyz, coarse=True, viewdirs=None, far=False, return_features=False):
        x = FO.field_inputs(xyz, viewdirs, self.poses, self.focal, self.c, self.image_shape, self.encoder.latent,
                            self.encoder.latent_scaling, self.code._freqs, self.code._phases, ns=self.num_views_per_obj,
                            features_only=return_features)
        if return_features:
            return x
        b = xyz.shape[1]
        out = (self.mlp_coarse if coarse else self.mlp_fine)(x, combine_inner_dims=(self.num_views_per_obj, b))
        out = out.reshape(-1, b, 4)
        return torch.cat([torch.sigmoid(out[..., :3]), torch.relu(out[..., 3:4])], -1).reshape(xyz.shape[0], b, -1)


class MapField(torch.nn.Module):
    """A radiance field that is nothing but a feature map (a parameter, so d_latent has somewhere to
    go) with the attributes the fused front end reads — for the LSTM march at 256 / 512 channels,
    where StubNet's conv encoder would only slow the test down."""

    use_encoder = use_xyz = use_code = use_viewdirs = normalize_z = True
    use_global_encoder = use_code_viewdirs = stop_encoder_grad = False
    d_out = 4
    num_views_per_obj = 1

    def __init__(self, sb, ch, h=12, w=10, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.encoder, self.code = _Encoder(), _Code()
        smooth = torch.nn.functional.interpolate(torch.randn(sb, ch, 4, 4, generator=g), size=(h, w), mode="bicubic", align_corners=True)
        self.map = torch.nn.Parameter(smooth * 0.4)
        ls = torch.tensor([float(w), float(h)])
        self.encoder.latent_scaling = ls / (ls - 1) * 2.0
        self.mlp_coarse = self.mlp_fine = None

    def place(self, cam2world, focal):
        """Source view = the given poses (SB, 4, 4); call after .to(device)."""
        dev = self.map.device
        rot = cam2world[:, :3, :3].transpose(1, 2)
        self.poses = torch.cat((rot, -torch.bmm(rot, cam2world[:, :3, 3:])), dim=-1).contiguous().to(dev)
        h, w = self.map.shape[-2:]
        self.image_shape = torch.tensor([4.0 * w, 4.0 * h], device=dev)
        self.focal = torch.tensor([[focal, -focal]], device=dev)
        self.c = (self.image_shape * 0.5).unsqueeze(0)
        self.encoder.latent_scaling = self.encoder.latent_scaling.to(dev)
        self.refresh()

    def refresh(self):
        self.encoder.latent = self.map * 1.0        # a non-leaf, like an encoder's output

    def forward(self, xyz, coarse=True, viewdirs=None, far=False, return_features=False):
        assert return_features
        return FO.field_inputs(xyz, viewdirs, self.poses, self.focal, self.c, self.image_shape, self.encoder.latent,
                               self.encoder.latent_scaling, self.code._freqs, self.code._phases, ns=1, features_only=True)
