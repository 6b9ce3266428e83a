"""torch-facing operators over the C ABI: output allocation, stream/device plumbing and
``torch.autograd.Function`` wrappers.  No arithmetic happens here.

Shapes follow the reference (renderers.py): leading dims are ``(SB, R)``; the kernels
see them flattened to one ray axis.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr, require_cuda


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 + contiguous (no copy when it already is — the radiance field's output is)."""
    if t.dtype != torch.float32:
        raise _lib.AvrError(f"avr_b200 kernels are fp32 (as the reference); got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _bounds(near: torch.Tensor, far: torch.Tensor, n_rays: int):
    """(near, far, bound_stride).  The reference hands VolumeRenderer's scalar bounds
    over as stride-0 expands of a 1-element tensor (renderers.py:169); those stay one
    element (bound_stride 0), everything else becomes one value per ray."""
    def scalar(t):
        return t.numel() == 1 or all(s == 0 for s in t.stride())

    def first(t):
        return _f32c(t.as_strided((1,), (1,)))

    if scalar(near) and scalar(far):
        return first(near), first(far), 0
    near = _f32c(first(near).expand(n_rays) if scalar(near) else near.reshape(-1))
    far = _f32c(first(far).expand(n_rays) if scalar(far) else far.reshape(-1))
    if near.numel() != n_rays or far.numel() != n_rays:
        raise _lib.AvrError(f"near/far must hold 1 or {n_rays} values")
    return near, far, 1


# --------------------------------------------------------------------------- coarse
def coarse_sample_raw(near, far, bound_stride: int, u: torch.Tensor) -> torch.Tensor:
    require_cuda(near, far, u)
    u = _f32c(u)
    k = u.shape[-1]
    r = u.numel() // k
    z = torch.empty_like(u)
    with torch.cuda.device(u.device):
        check(_lib.load().avr_coarse_sample_fwd(ptr(near), ptr(far), bound_stride, ptr(u), r, k, ptr(z), _stream(u)),
              "avr_coarse_sample_fwd")
    return z


class CoarseSample(torch.autograd.Function):
    """z = sample_coarse(near, far) for per-ray bounds that carry grad (AdaptiveVolumeRenderer,
    renderers.py:490-494).  u is the explicit uniform draw."""

    @staticmethod
    def forward(ctx, near, far, u):
        near_c, far_c, stride = _bounds(near, far, u.numel() // u.shape[-1])
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(u)
        ctx.bshape = (near.shape, far.shape)
        ctx.stride = stride
        return coarse_sample_raw(near_c, far_c, stride, u)

    @staticmethod
    def backward(ctx, g_z):
        if g_z is None:
            return None, None, None
        (u,) = ctx.saved_tensors
        g_z = _f32c(g_z)
        u = _f32c(u)
        k = u.shape[-1]
        r = u.numel() // k
        d_near = torch.empty(r, dtype=torch.float32, device=u.device)
        d_far = torch.empty_like(d_near)
        with torch.cuda.device(u.device):
            check(_lib.load().avr_coarse_sample_bwd(ptr(g_z), ptr(u), r, k, ptr(d_near), ptr(d_far), _stream(u)),
                  "avr_coarse_sample_bwd")
        ns, fs = ctx.bshape
        # an expanded (stride-0) input still receives a per-ray gradient of its own shape;
        # autograd's expand backward does the reduction
        d_near = d_near.reshape(ns) if d_near.numel() == ns.numel() else d_near.sum().reshape(ns)
        d_far = d_far.reshape(fs) if d_far.numel() == fs.numel() else d_far.sum().reshape(fs)
        return d_near, d_far, None


# ------------------------------------------------------------------------ importance
def importance_sample(
    weights: torch.Tensor,              # (..., Kc) or (..., Kc, 1)
    near: torch.Tensor, far: torch.Tensor,
    u: torch.Tensor, u2: torch.Tensor,  # (..., n)
    z_coarse: Optional[torch.Tensor] = None,
    normals: Optional[torch.Tensor] = None, depth_std: float = 0.0,
    want_fine: bool = True, want_sorted: bool = False, want_cdf: bool = False, want_idx: bool = False,
):
    """sample_fine (+ sample_depth/clamp + cat/sort when ``want_sorted``) in one kernel.
    Returns a dict with the requested outputs.  Non-differentiable (the reference
    detaches the weights, renderers.py:36)."""
    if weights.dim() == u.dim() + 1:
        weights = weights.squeeze(-1)
    weights = _f32c(weights.detach())
    u, u2 = _f32c(u), _f32c(u2)
    require_cuda(weights, u, u2, near, far)
    kc, n = weights.shape[-1], u.shape[-1]
    r = weights.numel() // kc
    lead = weights.shape[:-1]
    near_c, far_c, stride = _bounds(near.detach(), far.detach(), r)
    nd = 0
    if want_sorted:
        if z_coarse is None:
            raise _lib.AvrError("want_sorted needs z_coarse")
        z_coarse = _f32c(z_coarse.detach())
        if normals is not None:
            normals = _f32c(normals)
            nd = normals.shape[-1]
    dev = weights.device
    out = {}
    z_fine = torch.empty(*lead, n, dtype=torch.float32, device=dev) if want_fine else None
    z_sorted = torch.empty(*lead, kc + n + nd, dtype=torch.float32, device=dev) if want_sorted else None
    cdf = torch.empty(*lead, kc + 1, dtype=torch.float32, device=dev) if want_cdf else None
    idx = torch.empty(*lead, n, dtype=torch.int32, device=dev) if want_idx else None
    with torch.cuda.device(dev):
        check(_lib.load().avr_importance_sample(
            ptr(weights), ptr(z_coarse), ptr(u), ptr(u2), ptr(normals) if nd else None, ptr(near_c), ptr(far_c), stride,
            r, kc, n, nd, float(depth_std), ptr(z_fine), ptr(z_sorted), ptr(cdf), ptr(idx), _stream(weights)),
            "avr_importance_sample")
    out.update(z_fine=z_fine, z_sorted=z_sorted, cdf=cdf, idx=idx)
    return out


class SortRays(torch.autograd.Function):
    """Ascending per-ray sort that routes gradients (torch.sort at renderers.py:494)."""

    @staticmethod
    def forward(ctx, z):
        require_cuda(z)
        zc = _f32c(z)
        k = zc.shape[-1]
        r = zc.numel() // k
        out = torch.empty_like(zc)
        perm = torch.empty(zc.shape, dtype=torch.int32, device=zc.device)
        with torch.cuda.device(zc.device):
            check(_lib.load().avr_sort_rays(ptr(zc), r, k, ptr(out), ptr(perm), _stream(zc)), "avr_sort_rays")
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(perm)
        ctx.mark_non_differentiable(perm)
        return out, perm

    @staticmethod
    def backward(ctx, g_out, _g_perm):
        if g_out is None:
            return None
        (perm,) = ctx.saved_tensors
        g_out = _f32c(g_out)
        k = g_out.shape[-1]
        g_in = torch.empty_like(g_out)
        with torch.cuda.device(g_out.device):
            check(_lib.load().avr_sort_rays_bwd(ptr(g_out), ptr(perm), g_out.numel() // k, k, ptr(g_in), _stream(g_out)),
                  "avr_sort_rays_bwd")
        return g_in


# -------------------------------------------------------------------------- composite
def composite_fwd_raw(rgbs: torch.Tensor, z: torch.Tensor, white_back: bool, infinity: float, want_w: bool = True,
                      depth_affine: Optional[torch.Tensor] = None):
    require_cuda(rgbs, z, depth_affine)
    rgbs, z = _f32c(rgbs), _f32c(z)
    k = z.shape[-1]
    r = z.numel() // k
    if rgbs.numel() != z.numel() * 4:
        raise _lib.AvrError(f"rgbs {tuple(rgbs.shape)} does not match z {tuple(z.shape)}")
    lead = z.shape[:-1]
    dev = z.device
    w = torch.empty(*lead, k, dtype=torch.float32, device=dev) if want_w else None
    rgb = torch.empty(*lead, 3, dtype=torch.float32, device=dev)
    depth = torch.empty(*lead, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().avr_composite_fwd_camera(ptr(rgbs), ptr(z), ptr(_affine(depth_affine, r)), r, k,
                                                   int(bool(white_back)), float(infinity), ptr(w), ptr(rgb), ptr(depth),
                                                   _stream(z)), "avr_composite_fwd")
    return rgb, depth, w


def _affine(depth_affine, r):
    if depth_affine is None:
        return None
    a = _f32c(depth_affine)
    if a.numel() != 2 * r:
        raise _lib.AvrError(f"depth_affine holds {a.numel() // 2} rays, the batch {r}")
    return a


def composite_bwd_raw(rgbs, z, g_rgb, g_depth, g_w, white_back: bool, infinity: float, want_dz: bool, depth_affine=None):
    rgbs, z = _f32c(rgbs), _f32c(z)
    k = z.shape[-1]
    r = z.numel() // k
    g_rgb = None if g_rgb is None else _f32c(g_rgb)
    g_depth = None if g_depth is None else _f32c(g_depth)
    g_w = None if g_w is None else _f32c(g_w)
    d_rgbs = torch.empty_like(rgbs)
    d_z = torch.empty_like(z) if want_dz else None
    with torch.cuda.device(z.device):
        check(_lib.load().avr_composite_bwd_camera(ptr(rgbs), ptr(z), ptr(_affine(depth_affine, r)), ptr(g_rgb), ptr(g_depth),
                                                   ptr(g_w), r, k, int(bool(white_back)), float(infinity), ptr(d_rgbs),
                                                   ptr(d_z), _stream(z)), "avr_composite_bwd")
    return d_rgbs, d_z


class Composite(torch.autograd.Function):
    """(rgb, depth, w) = volume_integral(z, rgbs) — renderers.py:69-119.

    Saves only its inputs; backward recomputes the transmittance.  ``w`` is returned
    with shape (..., K) (callers add the trailing 1 the reference has)."""

    @staticmethod
    def forward(ctx, rgbs, z, white_back: bool, infinity: float, want_w: bool, depth_affine=None):
        rgb, depth, w = composite_fwd_raw(rgbs, z, white_back, infinity, want_w, depth_affine)
        # an output nobody differentiates through must reach backward as None, not as a tensor of
        # zeros: a materialised g_w would push every training step of VolumeRenderer's coarse pass
        # (w_c only feeds the detached sampler, renderers.py:36) off the span kernel (api.cu)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(rgbs, z, depth_affine)
        ctx.cfg = (bool(white_back), float(infinity))
        if w is None:
            w = rgb.new_empty(0)
            ctx.mark_non_differentiable(w)
        return rgb, depth, w

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_w):
        rgbs, z, depth_affine = ctx.saved_tensors
        white_back, infinity = ctx.cfg
        if g_w is not None and g_w.numel() == 0:
            g_w = None
        if g_rgb is None and g_depth is None and g_w is None:
            return None, None, None, None, None, None
        want_dz = ctx.needs_input_grad[1]
        d_rgbs, d_z = composite_bwd_raw(rgbs, z, g_rgb, g_depth, g_w, white_back, infinity, want_dz, depth_affine)
        return (d_rgbs.view_as(rgbs) if ctx.needs_input_grad[0] else None), d_z, None, None, None, None


def composite(rgbs: torch.Tensor, z: torch.Tensor, white_back: bool = True, infinity: float = 1.8,
              want_w: bool = True, depth_affine: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """Differentiable compositing of a (..., K, 4) = (r,g,b,sigma) buffer along z (..., K).
    Returns rgb (...,3), depth (...), w (...,K) or None.  With ``depth_affine`` (..., 2) — the
    per-ray coefficients ``world_rays`` / ``rays_coarse_sample_points`` return — ``depth`` is the
    camera depth of the composited point (utils.depth_from_world, renderers.py:274-275) instead
    of the distance along the ray."""
    rgb, depth, w = Composite.apply(rgbs, z, white_back, infinity, want_w, depth_affine)
    return rgb, depth, (w if want_w else None)


# ----------------------------------------------------------------- packed (ragged) rays
def composite_packed_fwd_raw(rgbs, z, offsets, white_back, infinity, want_w=True):
    require_cuda(rgbs, z, offsets)
    rgbs, z = _f32c(rgbs), _f32c(z)
    if offsets.dtype != torch.int64:
        raise _lib.AvrError("offsets must be int64")
    offsets = offsets.contiguous()
    r = offsets.numel() - 1
    s = z.numel()
    dev = z.device
    w = torch.empty(s, dtype=torch.float32, device=dev) if want_w else None
    rgb = torch.empty(r, 3, dtype=torch.float32, device=dev)
    depth = torch.empty(r, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().avr_composite_fwd_packed(ptr(rgbs), ptr(z), ptr(offsets), r, s, int(bool(white_back)),
                                                   float(infinity), ptr(w), ptr(rgb), ptr(depth), _stream(z)),
              "avr_composite_fwd_packed")
    return rgb, depth, w


class CompositePacked(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgbs, z, offsets, white_back, infinity, want_w):
        rgb, depth, w = composite_packed_fwd_raw(rgbs, z, offsets, white_back, infinity, want_w)
        ctx.set_materialize_grads(False)   # see Composite.forward
        ctx.save_for_backward(rgbs, z, offsets)
        ctx.cfg = (bool(white_back), float(infinity))
        if w is None:
            w = rgb.new_empty(0)
            ctx.mark_non_differentiable(w)
        return rgb, depth, w

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_w):
        rgbs, z, offsets = ctx.saved_tensors
        white_back, infinity = ctx.cfg
        rgbs_c, z_c = _f32c(rgbs), _f32c(z)
        if g_w is not None and g_w.numel() == 0:
            g_w = None
        if g_rgb is None and g_depth is None and g_w is None:
            return None, None, None, None, None, None
        g_rgb = None if g_rgb is None else _f32c(g_rgb)
        g_depth = None if g_depth is None else _f32c(g_depth)
        g_w = None if g_w is None else _f32c(g_w)
        want_dz = ctx.needs_input_grad[1]
        # every sample belongs to a ray with count > 0, so the kernel writes every element
        d_rgbs = torch.empty_like(rgbs_c)
        d_z = torch.empty_like(z_c) if want_dz else None
        r = offsets.numel() - 1
        with torch.cuda.device(z.device):
            check(_lib.load().avr_composite_bwd_packed(ptr(rgbs_c), ptr(z_c), ptr(offsets), ptr(g_rgb), ptr(g_depth),
                                                       ptr(g_w), r, z_c.numel(), int(white_back), float(infinity),
                                                       ptr(d_rgbs), ptr(d_z), _stream(z)), "avr_composite_bwd_packed")
        return (d_rgbs.view_as(rgbs) if ctx.needs_input_grad[0] else None), d_z, None, None, None, None


def composite_packed(rgbs, z, offsets, white_back=True, infinity=1.8, want_w=True):
    """Compositing over packed rays: rgbs (S,4), z (S,), offsets (R+1,) int64."""
    rgb, depth, w = CompositePacked.apply(rgbs, z, offsets, white_back, infinity, want_w)
    return rgb, depth, (w if want_w else None)


def coarse_sample_packed(near, far, u, offsets):
    """Stratified depths for packed rays; ray r is stratified over its own count."""
    require_cuda(u, offsets)
    u = _f32c(u)
    r = offsets.numel() - 1
    near_c, far_c, stride = _bounds(near, far, r)
    z = torch.empty_like(u)
    with torch.cuda.device(u.device):
        check(_lib.load().avr_coarse_sample_fwd_packed(ptr(near_c), ptr(far_c), stride, ptr(u), ptr(offsets.contiguous()),
                                                       r, u.numel(), ptr(z), _stream(u)), "avr_coarse_sample_fwd_packed")
    return z


def importance_sample_packed(weights, z_coarse, near, far, u, u2, offsets, fine_offsets, max_coarse, max_fine,
                             want_sorted=True, want_cdf=False, want_idx=False):
    """Packed sample_fine + merge: returns (z_fine (Sf,), z_sorted (S+Sf,) or None) — and, when asked
    for, the cdf (S+R,: ray r's Kc_r+1 entries start at offsets[r]+r) and the int32 bin indices (Sf,)."""
    require_cuda(weights, u, u2, offsets, fine_offsets)
    weights, u, u2 = _f32c(weights.detach()), _f32c(u), _f32c(u2)
    z_coarse = _f32c(z_coarse.detach())
    r = offsets.numel() - 1
    near_c, far_c, stride = _bounds(near.detach(), far.detach(), r)
    z_fine = torch.empty_like(u)
    z_sorted = torch.empty(weights.numel() + u.numel(), dtype=torch.float32, device=u.device) if want_sorted else None
    cdf = torch.zeros(weights.numel() + r, dtype=torch.float32, device=u.device) if want_cdf else None
    idx = torch.zeros(u.numel(), dtype=torch.int32, device=u.device) if want_idx else None
    with torch.cuda.device(u.device):
        check(_lib.load().avr_importance_sample_packed(
            ptr(weights), ptr(z_coarse), ptr(u), ptr(u2), ptr(near_c), ptr(far_c), stride,
            ptr(offsets.contiguous()), ptr(fine_offsets.contiguous()), r, int(max_coarse), int(max_fine),
            ptr(z_fine), ptr(z_sorted), ptr(cdf), ptr(idx), _stream(u)), "avr_importance_sample_packed")
    if want_cdf or want_idx:
        return z_fine, z_sorted, cdf, idx
    return z_fine, z_sorted


# ------------------------------------------- ray setup / sample points / depth
def _rays3(t: torch.Tensor) -> torch.Tensor:
    return _f32c(t.reshape(-1, 3))


class RayPoints(torch.autograd.Function):
    """pts = ros + rds * z and the K-fold view-direction copy the radiance field is called with
    (renderers.py:171-175, 260-265, 496-500).  Returns pts (..., K, 3), viewdirs (..., K, 3).
    Backward: d_z in CUDA (the adaptive renderer's depths carry grad); gradients w.r.t. the ray
    origins/directions are per-ray reductions nobody on the reference's path asks for — they are
    formed with torch ops if requested."""

    @staticmethod
    def forward(ctx, ros, rds, z):
        require_cuda(ros, rds, z)
        zc = _f32c(z)
        k = zc.shape[-1]
        r = zc.numel() // k
        o, d = _rays3(ros), _rays3(rds)
        if o.shape[0] != r or d.shape[0] != r:
            raise _lib.AvrError(f"ros/rds hold {o.shape[0]}/{d.shape[0]} rays, z holds {r}")
        pts = torch.empty(*zc.shape, 3, dtype=torch.float32, device=zc.device)
        vd = torch.empty_like(pts)
        with torch.cuda.device(zc.device):
            check(_lib.load().avr_ray_points_fwd(ptr(o), ptr(d), ptr(zc), r, k, ptr(pts), ptr(vd), _stream(zc)),
                  "avr_ray_points_fwd")
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(d, zc)
        ctx.shapes = (ros.shape, rds.shape)
        return pts, vd

    @staticmethod
    def backward(ctx, g_pts, g_vd):
        d, zc = ctx.saved_tensors
        k = zc.shape[-1]
        r = zc.numel() // k
        d_z = d_o = d_d = None
        if g_pts is not None and ctx.needs_input_grad[2]:
            g = _f32c(g_pts)
            d_z = torch.empty_like(zc)
            with torch.cuda.device(zc.device):
                check(_lib.load().avr_ray_points_bwd(ptr(d), ptr(g), r, k, ptr(d_z), _stream(zc)), "avr_ray_points_bwd")
        if ctx.needs_input_grad[0] and g_pts is not None:
            d_o = g_pts.reshape(r, k, 3).sum(1).reshape(ctx.shapes[0])
        if ctx.needs_input_grad[1] and (g_pts is not None or g_vd is not None):
            acc = torch.zeros(r, 3, dtype=torch.float32, device=zc.device)
            if g_pts is not None:
                acc = acc + (g_pts.reshape(r, k, 3) * zc.reshape(r, k, 1)).sum(1)
            if g_vd is not None:
                acc = acc + g_vd.reshape(r, k, 3).sum(1)
            d_d = acc.reshape(ctx.shapes[1])
        return d_o, d_d, d_z


def ray_points(ros: torch.Tensor, rds: torch.Tensor, z: torch.Tensor):
    """(pts, viewdirs), each (..., K, 3), for rays (..., 3) and depths (..., K)."""
    return RayPoints.apply(ros, rds, z)


class RayPointsPacked(torch.autograd.Function):
    """Sample points and view directions for packed rays: ros, rds (R,3), z (S,), offsets (R+1,)
    -> pts, viewdirs (S,3).  Backward gives d_z (the adaptive depths carry grad)."""

    @staticmethod
    def forward(ctx, ros, rds, z, offsets):
        require_cuda(ros, rds, z, offsets)
        o, d, zc = _rays3(ros), _rays3(rds), _f32c(z)
        off = offsets.contiguous()
        r, s = off.numel() - 1, zc.numel()
        if o.shape[0] != r or d.shape[0] != r:
            raise _lib.AvrError(f"ros/rds hold {o.shape[0]}/{d.shape[0]} rays, offsets describe {r}")
        pts = torch.empty(s, 3, dtype=torch.float32, device=zc.device)
        vd = torch.empty_like(pts)
        with torch.cuda.device(zc.device):
            check(_lib.load().avr_ray_points_fwd_packed(ptr(o), ptr(d), ptr(zc), ptr(off), r, s, ptr(pts), ptr(vd), _stream(zc)),
                  "avr_ray_points_fwd_packed")
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(d, off)
        ctx.mark_non_differentiable(vd)
        return pts, vd

    @staticmethod
    def backward(ctx, g_pts, _g_vd):
        d, off = ctx.saved_tensors
        if g_pts is None or not ctx.needs_input_grad[2]:
            return None, None, None, None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            raise _lib.AvrError("RayPointsPacked differentiates w.r.t. the depths only")
        g = _f32c(g_pts)
        s = g.shape[0]
        d_z = torch.empty(s, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(_lib.load().avr_ray_points_bwd_packed(ptr(d), ptr(g), ptr(off), off.numel() - 1, s, ptr(d_z), _stream(g)),
                  "avr_ray_points_bwd_packed")
        return None, None, d_z, None


def ray_points_packed(ros, rds, z, offsets):
    return RayPointsPacked.apply(ros, rds, z, offsets)


def coarse_sample_points(near, far, bound_stride: int, u: torch.Tensor, ros: torch.Tensor, rds: torch.Tensor):
    """sample_coarse + point generation in one pass over the uniforms (renderers.py:169-175).
    Returns z (..., K), pts (..., K, 3), viewdirs (..., K, 3); non-differentiable (VolumeRenderer's
    bounds are constants)."""
    require_cuda(near, far, u, ros, rds)
    u = _f32c(u)
    k = u.shape[-1]
    r = u.numel() // k
    o, d = _rays3(ros.detach()), _rays3(rds.detach())
    z = torch.empty_like(u)
    pts = torch.empty(*u.shape, 3, dtype=torch.float32, device=u.device)
    vd = torch.empty_like(pts)
    with torch.cuda.device(u.device):
        check(_lib.load().avr_coarse_sample_points_fwd(ptr(near), ptr(far), bound_stride, ptr(u), ptr(o), ptr(d), r, k,
                                                       ptr(z), ptr(pts), ptr(vd), _stream(u)),
              "avr_coarse_sample_points_fwd")
    return z, pts, vd


def _camera_inputs(xy_pix, intrinsics, cam2world):
    sb = xy_pix.shape[0]
    x = _f32c(xy_pix.detach())
    c2w = _f32c(cam2world.detach())
    kinv = _f32c(intrinsics.detach().reshape(-1, 3, 3))
    if kinv.shape[0] != sb:
        kinv = _f32c(kinv.expand(sb, 3, 3))
    return x, kinv, c2w


def rays_coarse_sample_points(xy_pix, intrinsics, cam2world, near, far, bound_stride: int, u: torch.Tensor):
    """renderers.py:166-175 in one launch: get_world_rays + sample_coarse + sample points + view
    directions.  Returns ros, rds (SB,R,3), depth_affine (SB,R,2), z (SB,R,K), pts, viewdirs
    (SB,R,K,3); non-differentiable (poses, intrinsics and VolumeRenderer's bounds are data)."""
    require_cuda(xy_pix, intrinsics, cam2world, near, far, u)
    sb, n = xy_pix.shape[0], xy_pix.shape[1]
    x, kinv, c2w = _camera_inputs(xy_pix, intrinsics, cam2world)
    u = _f32c(u)
    k = u.shape[-1]
    dev = x.device
    ros = torch.empty(sb, n, 3, dtype=torch.float32, device=dev)
    rds = torch.empty_like(ros)
    aff = torch.empty(sb, n, 2, dtype=torch.float32, device=dev)
    z = torch.empty_like(u)
    pts = torch.empty(*u.shape, 3, dtype=torch.float32, device=dev)
    vd = torch.empty_like(pts)
    with torch.cuda.device(dev):
        check(_lib.load().avr_rays_coarse_sample_points_fwd(ptr(x), ptr(kinv), ptr(c2w), n, ptr(near), ptr(far), bound_stride,
                                                            ptr(u), sb * n, k, ptr(ros), ptr(rds), ptr(aff), ptr(z), ptr(pts),
                                                            ptr(vd), _stream(x)), "avr_rays_coarse_sample_points_fwd")
    return ros, rds, aff, z, pts, vd


def world_rays(xy_pix: torch.Tensor, intrinsics: torch.Tensor, cam2world: torch.Tensor, want_affine: bool = False):
    """utils.get_world_rays (utils.py:309-336): xy_pix (SB,R,2), intrinsics (SB,3,3), cam2world
    (SB,R,4,4) -> origins, unit directions (SB,R,3), all of it (including the 3x3 inverse of
    utils.py:263) in one kernel.  Non-differentiable (the reference's poses and intrinsics are data)."""
    require_cuda(xy_pix, intrinsics, cam2world)
    sb, n = xy_pix.shape[0], xy_pix.shape[1]
    x, kinv, c2w = _camera_inputs(xy_pix, intrinsics, cam2world)
    ros = torch.empty(sb, n, 3, dtype=torch.float32, device=x.device)
    rds = torch.empty_like(ros)
    aff = torch.empty(sb, n, 2, dtype=torch.float32, device=x.device) if want_affine else None
    with torch.cuda.device(x.device):
        check(_lib.load().avr_world_rays(ptr(x), ptr(kinv), ptr(c2w), sb * n, n, ptr(ros), ptr(rds), ptr(aff), _stream(x)),
              "avr_world_rays")
    return (ros, rds, aff) if want_affine else (ros, rds)


class DepthFromWorld(torch.autograd.Function):
    """depth = -(cam2world^-1 [p,1])_z with p = ros + rds*dist (dist given) or p = ros (dist None):
    utils.depth_from_world (utils.py:358-361) at renderers.py:274-275, 486, 508-509.  depth is
    affine in p, so backward is the saved per-ray row d depth/d p times the upstream gradient."""

    @staticmethod
    def forward(ctx, ros, rds, dist, cam2world):
        require_cuda(ros, cam2world)
        o = _rays3(ros)
        r = o.shape[0]
        d = _rays3(rds) if dist is not None else None
        t = _f32c(dist.reshape(-1)) if dist is not None else None
        c2w = _f32c(cam2world.detach().reshape(-1, 4, 4))
        if c2w.shape[0] != r:
            raise _lib.AvrError(f"cam2world holds {c2w.shape[0]} poses for {r} points")
        need_grad = any(ctx.needs_input_grad[:3])
        depth = torch.empty(ros.shape[:-1], dtype=torch.float32, device=o.device)
        row = torch.empty(r, 3, dtype=torch.float32, device=o.device) if need_grad else None
        with torch.cuda.device(o.device):
            check(_lib.load().avr_depth_from_world(ptr(o), ptr(d), ptr(t), ptr(c2w), r, ptr(depth), ptr(row), _stream(o)),
                  "avr_depth_from_world")
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(row, d, t)
        ctx.shapes = (ros.shape, None if rds is None else rds.shape, None if dist is None else dist.shape)
        return depth

    @staticmethod
    def backward(ctx, g):
        if g is None:
            return None, None, None, None
        row, d, t = ctx.saved_tensors
        gp = g.reshape(-1, 1) * row                      # dL/dp, (R,3)
        d_o = gp.reshape(ctx.shapes[0]) if ctx.needs_input_grad[0] else None
        d_d = d_t = None
        if t is not None:
            if ctx.needs_input_grad[1]:
                d_d = (gp * t.reshape(-1, 1)).reshape(ctx.shapes[1])
            if ctx.needs_input_grad[2]:
                d_t = (gp * d).sum(-1).reshape(ctx.shapes[2])
        return d_o, d_d, d_t, None


def depth_from_world(ros, rds, dist, cam2world):
    """Depth (SB,R) of ros + rds*dist (or of the points `ros` when dist is None) seen from cam2world."""
    return DepthFromWorld.apply(ros, rds, dist, cam2world)
