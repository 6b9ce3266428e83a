"""Ray setup and depth re-projection either side of the hot path, under the reference's names.

``utils.get_world_rays`` (utils.py:315-336, via ``unproject`` :246-267) and
``utils.depth_from_world`` (utils.py:358-361) as CUDA kernels (csrc/geometry.cu) behind
``ops.world_rays`` / ``ops.depth_from_world``: SURVEY.md section 8(f) row 2.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def get_world_rays(xy_pix: torch.Tensor, intrinsics: torch.Tensor, cam2world: torch.Tensor):
    """xy_pix (SB,R,2), intrinsics (SB,3,3), cam2world (SB,R,4,4) -> origins, directions (SB,R,3)."""
    return ops.world_rays(xy_pix, intrinsics, cam2world)


def depth_from_world(world_coords: torch.Tensor, cam2world: torch.Tensor, rds: Optional[torch.Tensor] = None,
                     dist: Optional[torch.Tensor] = None) -> torch.Tensor:
    """-z of the points in camera coordinates (the reference's two-argument call).  With
    ``rds``/``dist`` the points are ``world_coords + rds * dist`` formed inside the kernel
    (renderers.py:274-275, 508-509) instead of by a torch op."""
    return ops.depth_from_world(world_coords, rds, dist, cam2world)
