"""Radiance-field front end: the step between the renderer's sample
points and the reference's MLP, as one CUDA kernel each way (csrc/field_inputs.cu).

``NewPixelNeRFNet.forward`` (models.py:739-866) spends ~25 ATen calls turning ``xyz`` / ``viewdirs``
into the MLP's input — repeat over source views, two batched 3x3 products, the positional
encoding, the projection, ``F.grid_sample`` of the encoder's feature map, a transpose and two
concatenations — each a pass over (or a copy of) a (points, 512+) tensor.  ``field_inputs`` writes
the final ``(SB*NS*B, C + code)`` tensor directly; its backward returns the gradients of the
feature map (to train the encoder), of the points and of the view directions (the adaptive
renderer differentiates through its sample positions).

    net = models.make_new_model(conf)          # the reference's module, unchanged
    avr_b200.fuse_field_inputs(net)            # same object; forward() now uses the kernels
    net.encode(images, poses, focal, c=c)      # reference code
    rgbs = net(xyz, coarse=True, viewdirs=d)   # reference signature (models.py:739)

The encoder and the MLPs stay the reference's torch modules (BASELINE.json north_star); only
the glue between them is replaced.  There is no fallback: an unsupported configuration raises
``AvrError`` from ``fuse_field_inputs`` (the module is then left untouched).
"""
from __future__ import annotations

import ctypes
import types
from ctypes import c_float, c_int, c_int64, c_void_p
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import AvrError, check, ptr, require_cuda

MAX_SIN = 32  # AVR_FIELD_MAX_SIN


class FieldInputsDesc(ctypes.Structure):
    """Mirror of ``avr_field_inputs`` (include/avr_b200.h), field for field."""

    _fields_ = [
        ("xyz", c_void_p), ("viewdirs", c_void_p), ("poses", c_void_p), ("focal", c_void_p), ("c", c_void_p),
        ("latent", c_void_p), ("out", c_void_p),
        ("B", c_int64), ("NV", c_int64), ("NS", c_int), ("focal_per_obj", c_int), ("c_per_obj", c_int),
        ("scale_x", c_float), ("scale_y", c_float),
        ("C", c_int), ("H", c_int), ("W", c_int), ("n_sin", c_int),
        ("freqs", c_float * MAX_SIN), ("phases", c_float * MAX_SIN),
        ("include_input", c_int), ("normalize_z", c_int), ("use_viewdirs", c_int), ("features_only", c_int),
        ("g_out", c_void_p), ("d_latent", c_void_p), ("d_xyz", c_void_p), ("d_viewdirs", c_void_p),
    ]


@dataclass(frozen=True)
class FieldConfig:
    """Launch constants of the front end: the module state the reference keeps as Python values
    or tiny tensors, read once per ``encode`` (no device synchronisation per call)."""

    ns: int                         # source views per object (models.py:697)
    scale: tuple                    # latent_scaling / image_shape (models.py:268-270), (x, y)
    freqs: tuple = ()               # PositionalEncoding._freqs flattened (models.py:53-56)
    phases: tuple = ()              # PositionalEncoding._phases (models.py:58-60)
    include_input: bool = True
    normalize_z: bool = True
    use_viewdirs: bool = True

    def code_width(self) -> int:
        return (3 if self.include_input else 0) + 3 * len(self.freqs) + (3 if self.use_viewdirs else 0)


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise AvrError(f"{what}: the front end is fp32 (as the reference); got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _fill(cfg: FieldConfig, xyz, viewdirs, latent, poses, focal, c, features_only) -> FieldInputsDesc:
    sb, b, _ = xyz.shape
    nv, h, w, ch = latent.shape
    if nv != sb * cfg.ns or poses.shape != (nv, 3, 4):
        raise AvrError(f"field_inputs: {nv} feature maps / poses {tuple(poses.shape)} for {sb} objects x {cfg.ns} views")
    if len(cfg.freqs) != len(cfg.phases) or len(cfg.freqs) > MAX_SIN:
        raise AvrError("field_inputs: at most 16 encoding frequencies")
    for name, t in (("focal", focal), ("c", c)):
        if t.dim() != 2 or t.shape[1] != 2 or t.shape[0] not in (1, sb):
            raise AvrError(f"field_inputs: {name} must be (1, 2) or (SB, 2), got {tuple(t.shape)}")
    d = FieldInputsDesc()
    d.xyz, d.viewdirs, d.poses, d.focal, d.c, d.latent = ptr(xyz), ptr(viewdirs), ptr(poses), ptr(focal), ptr(c), ptr(latent)
    d.B, d.NV, d.NS = b, nv, cfg.ns
    d.focal_per_obj, d.c_per_obj = int(focal.shape[0] > 1), int(c.shape[0] > 1)
    d.scale_x, d.scale_y = cfg.scale
    d.C, d.H, d.W = ch, h, w
    d.n_sin = len(cfg.freqs)
    for i, (f, p) in enumerate(zip(cfg.freqs, cfg.phases)):
        d.freqs[i], d.phases[i] = f, p
    d.include_input, d.normalize_z = int(cfg.include_input), int(cfg.normalize_z)
    d.use_viewdirs, d.features_only = int(cfg.use_viewdirs), int(features_only)
    return d


class _FieldInputs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, viewdirs, latent, poses, focal, c, cfg: FieldConfig, features_only: bool):
        require_cuda(xyz, viewdirs, latent, poses, focal, c)
        xyz, latent = _f32c(xyz, "xyz"), _f32c(latent, "latent")
        poses, focal, c = _f32c(poses, "poses"), _f32c(focal, "focal"), _f32c(c, "c")
        use_vd = cfg.use_viewdirs and not features_only
        if use_vd:
            if viewdirs is None:
                raise AvrError("field_inputs: this configuration needs viewdirs (models.py:785)")
            viewdirs = _f32c(viewdirs.reshape(xyz.shape), "viewdirs")
        else:
            viewdirs = None
        d = _fill(cfg, xyz, viewdirs, latent, poses, focal, c, features_only)
        row = latent.shape[-1] + (0 if features_only else cfg.code_width())
        out = torch.empty((d.NV * d.B, row), dtype=torch.float32, device=xyz.device)
        d.out = ptr(out)
        with torch.cuda.device(xyz.device):
            check(_lib.load().avr_field_inputs_fwd(ctypes.byref(d), torch.cuda.current_stream(xyz.device).cuda_stream),
                  "avr_field_inputs_fwd")
        ctx.cfg, ctx.features_only = cfg, features_only
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(xyz, viewdirs, latent, poses, focal, c)
        return out

    @staticmethod
    def backward(ctx, g_out):
        xyz, viewdirs, latent, poses, focal, c = ctx.saved_tensors
        need_xyz, need_vd, need_lat = ctx.needs_input_grad[:3]
        need_vd = need_vd and viewdirs is not None
        if g_out is None or not (need_xyz or need_vd or need_lat):
            return (None,) * 8
        g_out = _f32c(g_out, "g_out")
        d = _fill(ctx.cfg, xyz, viewdirs, latent, poses, focal, c, ctx.features_only)
        d.g_out = ptr(g_out)
        # the launcher zeroes what it is handed, so empty() is enough
        d_lat = torch.empty_like(latent) if need_lat else None
        d_xyz = torch.empty_like(xyz) if need_xyz else None
        d_vd = torch.empty_like(viewdirs) if need_vd else None
        d.d_latent, d.d_xyz, d.d_viewdirs = ptr(d_lat), ptr(d_xyz), ptr(d_vd)
        with torch.cuda.device(xyz.device):
            check(_lib.load().avr_field_inputs_bwd(ctypes.byref(d), torch.cuda.current_stream(xyz.device).cuda_stream),
                  "avr_field_inputs_bwd")
        return d_xyz, d_vd, d_lat, None, None, None, None, None


def field_inputs(xyz: torch.Tensor, viewdirs: Optional[torch.Tensor], latent_nhwc: torch.Tensor, poses: torch.Tensor,
                 focal: torch.Tensor, c: torch.Tensor, cfg: FieldConfig, features_only: bool = False) -> torch.Tensor:
    """The MLP's input ``(SB*NS*B, C + code)`` (models.py:826) — or the features alone
    (``return_features=True``, models.py:828-829).

    xyz, viewdirs ``(SB, B, 3)``; latent_nhwc ``(SB*NS, H, W, C)`` (the encoder's ``latent``
    permuted to channels-last); poses ``(SB*NS, 3, 4)`` world->view; focal, c ``(1 or SB, 2)``."""
    return _FieldInputs.apply(xyz, viewdirs, latent_nhwc, poses, focal, c, cfg, features_only)


# ---------------------------------------------------------------------------------------------
# drop-in for the reference module
# ---------------------------------------------------------------------------------------------
def _config_of(net) -> FieldConfig:
    """Reads the launch constants off a NewPixelNeRFNet after ``encode`` (one small D2H copy)."""
    enc = net.encoder
    scale = (enc.latent_scaling.detach().float().cpu() / net.image_shape.detach().float().cpu()).tolist()
    freqs: Sequence[float] = ()
    phases: Sequence[float] = ()
    include_input = False
    if net.use_code:
        code = net.code
        freqs = code._freqs.detach().float().cpu().reshape(-1).tolist()
        phases = code._phases.detach().float().cpu().reshape(-1).tolist()
        include_input = bool(code.include_input)
    return FieldConfig(ns=int(net.num_views_per_obj), scale=(float(scale[0]), float(scale[1])), freqs=tuple(freqs),
                       phases=tuple(phases), include_input=include_input, normalize_z=bool(net.normalize_z),
                       use_viewdirs=bool(net.use_viewdirs))


def _check_supported(net) -> None:
    """The configurations the kernels implement: conf/default.conf's family."""
    problems = []
    if not getattr(net, "use_encoder", False):
        problems.append("use_encoder = False")
    if not getattr(net, "use_xyz", False):
        problems.append("use_xyz = False")
    if getattr(net, "use_global_encoder", False):
        problems.append("use_global_encoder = True")
    if getattr(net, "use_code", False) and getattr(net, "use_viewdirs", False) and getattr(net, "use_code_viewdirs", False):
        problems.append("use_code_viewdirs = True")
    if not getattr(net, "use_code", False):
        problems.append("use_code = False")
    enc = getattr(net, "encoder", None)
    if enc is None or getattr(enc, "index_interp", "bilinear") != "bilinear" or getattr(enc, "index_padding", "border") != "border":
        problems.append("encoder.index is not bilinear / border")
    if problems:
        raise AvrError("fuse_field_inputs: configuration not implemented by the kernels (" + "; ".join(problems) + ")")


def field_state(net):
    """(launch constants, channels-last feature map) of a fused radiance field, refreshed when
    ``encode`` has produced a new feature map.  Shared by the fused forward and the LSTM march."""
    latent = net.encoder.latent
    st = net._avr_field_state
    if st.get("latent") is not latent:          # encode() ran: new feature map, new camera state
        st["latent"] = latent
        st["cfg"] = _config_of(net)
        src = latent.detach() if net.stop_encoder_grad else latent
        st["nhwc"] = src.permute(0, 2, 3, 1).contiguous()
    return st["cfg"], st["nhwc"]


def _fused_forward(self, xyz, coarse=True, viewdirs=None, far=False, return_features=False):
    """``NewPixelNeRFNet.forward`` (models.py:739-866) with lines 754-826 done by ``field_inputs``."""
    sb, b, _ = xyz.shape
    cfg, nhwc = field_state(self)
    mlp_input = field_inputs(xyz, viewdirs, nhwc, self.poses, self.focal, self.c, cfg, features_only=return_features)
    if return_features:
        return mlp_input                        # (SB*NS*B, latent), models.py:828-829
    mlp = self.mlp_coarse if (coarse or self.mlp_fine is None) else self.mlp_fine
    out = mlp(mlp_input, combine_inner_dims=(self.num_views_per_obj, b), combine_index=None, dim_size=None)
    out = out.reshape(-1, b, self.d_out)        # models.py:853-862
    out = torch.cat([torch.sigmoid(out[..., :3]), torch.relu(out[..., 3:4])], dim=-1)
    return out.reshape(sb, b, -1)


def fuse_field_inputs(net):
    """Make a reference ``NewPixelNeRFNet`` compute its MLP input with the CUDA kernels.

    The module is modified in place (its ``forward`` is rebound; parameters, buffers, ``encode``
    and the state_dict are untouched, so checkpoints keep loading) and returned."""
    _check_supported(net)
    net._avr_field_state = {}
    net.forward = types.MethodType(_fused_forward, net)
    return net
