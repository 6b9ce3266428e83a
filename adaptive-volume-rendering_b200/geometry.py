"""Ray setup and depth re-projection either side of the hot path.

These stay stock torch ops (SURVEY.md section 8f ranks them as the next rows to fuse); they
restate ``utils.get_world_rays`` (utils.py:315-336, via ``unproject`` :246-267) and
``utils.depth_from_world`` (utils.py:358-361) so the drop-in renderers are self-contained.
"""
from __future__ import annotations

import torch


def get_world_rays(xy_pix: torch.Tensor, intrinsics: torch.Tensor, cam2world: torch.Tensor):
    """xy_pix (SB,R,2), intrinsics (SB,3,3), cam2world (SB,R,4,4) -> origins, directions (SB,R,3)."""
    origins = cam2world[..., :3, -1]
    ones = torch.ones_like(xy_pix[..., :1])
    pix_h = torch.cat((xy_pix, ones), dim=-1)
    cam = torch.einsum("...ij,...kj->...ki", intrinsics.inverse(), pix_h)
    # unproject flips x, then scales by z = -1 (the camera looks down -z)
    cam = torch.cat((-cam[..., :1], cam[..., 1:]), dim=-1) * (-ones)
    cam = cam / torch.norm(cam, dim=-1).unsqueeze(-1)
    dirs_h = torch.cat((cam, torch.zeros_like(cam[..., :1])), dim=-1)
    world = torch.einsum("...ij,...j->...i", cam2world, dirs_h)
    return origins, world[..., :3]


def depth_from_world(world_coords: torch.Tensor, cam2world: torch.Tensor) -> torch.Tensor:
    """-z of the points in camera coordinates."""
    pts_h = torch.cat((world_coords, torch.ones_like(world_coords[..., :1])), dim=-1)
    cam = torch.einsum("...ij,...j->...i", torch.inverse(cam2world), pts_h)
    return -cam[..., 2]
