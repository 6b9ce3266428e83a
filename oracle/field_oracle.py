"""CPU restatement of the radiance field's FRONT END — what ``NewPixelNeRFNet.forward`` computes
between receiving the renderer's sample points and calling its MLP.

TEST INFRASTRUCTURE ONLY: imported by ``tests/`` (and by ``oracle/make_golden.py``); never by the
package.  Pinned bit-for-bit to the imported reference by ``tests/test_oracle_vs_reference.py``
and to ``tests/golden/field_inputs_*.npz`` (outputs of the reference itself).

Follows /root/reference/models.py line by line, with the module state (``poses``, ``focal``,
``c``, ``image_shape``, the encoder's ``latent`` / ``latent_scaling``, the positional encoding's
``_freqs`` / ``_phases``) passed in as tensors:

    models.py:754-761   repeat over source views, xyz_rot = R p, xyz = xyz_rot + t
    models.py:763-781   what is encoded (normalize_z), PositionalEncoding.forward (:62-71)
    models.py:783-794   view directions into the view frame, concatenated after the code
    models.py:803-815   uv = -xy/z * focal + c; ConvEncoder.index (:256-279): uv * scale - 1,
                        F.grid_sample(bilinear, border, align_corners=True)
    models.py:817-829   latent -> (N, C); mlp_input = cat(latent, z_feature); return_features
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def repeat_views(t: torch.Tensor, ns: int) -> torch.Tensor:
    """utils.py:62-69 (repeat_interleave along dim 0)."""
    return t.unsqueeze(1).expand(-1, ns, *t.shape[1:]).reshape(-1, *t.shape[1:])


def positional_code(x: torch.Tensor, freqs: torch.Tensor, phases: torch.Tensor, include_input=True) -> torch.Tensor:
    """models.py:62-71; ``freqs`` / ``phases`` are the module's (1, 2F, 1) buffers."""
    n_rows = freqs.shape[1]
    embed = x.unsqueeze(1).repeat(1, n_rows, 1)
    embed = torch.sin(torch.addcmul(phases, embed, freqs))
    embed = embed.view(x.shape[0], -1)
    return torch.cat((x, embed), dim=-1) if include_input else embed


def field_inputs(xyz, viewdirs, poses, focal, c, image_shape, latent, latent_scaling, freqs, phases, ns=1,
                 include_input=True, normalize_z=True, use_viewdirs=True, features_only=False):
    """Returns the MLP's input (SB*NS*B, C + code) — or the (SB*NS*B, C) features alone.

    xyz, viewdirs (SB, B, 3); poses (SB*NS, 3, 4) world->view; focal, c (1 or SB, 2);
    image_shape (2,) = [W, H]; latent (SB*NS, C, H, W); latent_scaling (2,)."""
    sb, b, _ = xyz.shape
    p = repeat_views(xyz, ns)
    rot = torch.matmul(poses[:, None, :3, :3], p.unsqueeze(-1))[..., 0]
    cam = rot + poses[:, None, :3, 3]
    z_feature = (rot if normalize_z else cam).reshape(-1, 3)
    z_feature = positional_code(z_feature, freqs, phases, include_input)
    if use_viewdirs:
        d = repeat_views(viewdirs.reshape(sb, b, 3, 1), ns)
        d = torch.matmul(poses[:, None, :3, :3], d).reshape(-1, 3)
        z_feature = torch.cat((z_feature, d), dim=1)
    uv = -cam[:, :, :2] / cam[:, :, 2:]
    uv = uv * repeat_views(focal.unsqueeze(1), ns if focal.shape[0] > 1 else 1)
    uv = uv + repeat_views(c.unsqueeze(1), ns if c.shape[0] > 1 else 1)
    uv = uv * (latent_scaling / image_shape) - 1.0
    samples = F.grid_sample(latent, uv.unsqueeze(2), align_corners=True, mode="bilinear", padding_mode="border")
    feat = samples[:, :, :, 0].transpose(1, 2).reshape(-1, latent.shape[1])
    if features_only:
        return feat
    return torch.cat((feat, z_feature), dim=-1)
