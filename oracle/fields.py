"""Small deterministic stand-ins for the radiance field, shared by the golden
generator and the tests (TEST INFRASTRUCTURE ONLY).

The real PixelNeRF (models.py:609-910) stays the reference's own torch module
and is out of scope; the renderer only needs something that honours the
callback contract of SURVEY.md section 3.5:
``field(xyz (SB,N,3), viewdirs=(SB,N,3), coarse=bool) -> (SB,N,4)`` with
sigmoid colours and a ReLU density that has exact zeros (models.py:854-863),
plus ``return_features=True -> (SB,N,C)`` for the adaptive renderer's march
(models.py:739, 822-823).
"""
from __future__ import annotations

import math

import torch
from torch import nn


class TinyField(nn.Module):
    def __init__(self, hidden: int = 32, seed: int = 0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.l1 = nn.Linear(6, hidden)
        self.l2 = nn.Linear(hidden, hidden)
        self.l3 = nn.Linear(hidden, 4)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (1.5 if p.dim() > 1 else 0.3) / math.sqrt(max(p.shape[-1], 1)) * 2.0)

    def trunk(self, xyz, viewdirs):
        h = torch.cat([xyz * 3.0, viewdirs], -1)
        h = torch.sin(self.l1(h))
        return torch.tanh(self.l2(h))

    def forward(self, xyz, viewdirs=None, coarse=True, return_features=False):
        h = self.trunk(xyz, viewdirs)
        if return_features:
            return h
        o = self.l3(h)
        if not coarse:
            o = o * 1.25
        rgb = torch.sigmoid(o[..., :3])
        sigma = torch.relu(o[..., 3:4] * 20.0)
        return torch.cat([rgb, sigma], -1)


class TinyFeatureField(TinyField):
    """Feature width = ``channels`` so it can drive the adaptive renderer's LSTM."""

    def __init__(self, channels: int = 32, seed: int = 0):
        super().__init__(hidden=channels, seed=seed)


def camera_setup(sb: int, r: int, seed: int = 0):
    """Synthetic single-view rays as SURVEY.md section 8(d) C1 describes: SRN-cars
    intrinsics, a pose at distance 1.3 looking at the origin (OpenGL flip as
    dataset.py:85-86), expanded per ray as train.py:83/150 do."""
    g = torch.Generator().manual_seed(seed)
    x_pix = torch.rand(sb, r, 2, generator=g)
    f = 131.25 / 128.0
    intrinsics = torch.tensor([[f, 0.0, 0.5], [0.0, f, 0.5], [0.0, 0.0, 1.0]]).expand(sb, 3, 3).contiguous()
    poses = []
    for i in range(sb):
        th = 0.7 + 1.1 * i
        ph = 0.5 + 0.2 * i
        eye = 1.3 * torch.tensor([math.cos(th) * math.cos(ph), math.sin(th) * math.cos(ph), math.sin(ph)])
        fwd = -eye / eye.norm()
        up = torch.tensor([0.0, 0.0, 1.0])
        right = torch.linalg.cross(fwd, up)
        right = right / right.norm()
        true_up = torch.linalg.cross(right, fwd)
        c2w = torch.eye(4)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, true_up, -fwd, eye
        c2w = c2w @ torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0]))
        poses.append(c2w)
    cam2world = torch.stack(poses).unsqueeze(1).expand(sb, r, 4, 4).contiguous()
    return cam2world, intrinsics, x_pix
