"""AdaptiveVolumeRenderer's LSTM ray march (renderers.py:411-435) as
one CUDA launch forward and one backward (csrc/lstm_march.cu).

The reference marches every ray ``steps`` times through
``phi(world, return_features=True)`` -> ``LSTMCell(C -> 16)`` -> ``Linear(16 -> 1)`` -> advance;
in torch that is ~25 launches per step.  ``lstm_march`` does the whole loop in a persistent kernel
when ``phi`` is a radiance field whose feature fetch the front-end kernels implement
(``field.fuse_field_inputs``) with one source view per object — the only case the reference's
reshapes admit (renderers.py:423, :430).  The parameter gradients of the two small layers are four
plain GEMMs over rows the backward kernel writes; they go through ``torch.matmul`` (cuBLAS).
"""
from __future__ import annotations

import ctypes
from ctypes import c_int, c_int64, c_void_p

import torch

from . import _lib
from ._lib import AvrError, check, ptr, require_cuda
from .field import FieldConfig, _f32c, _fill

HIDDEN = 16          # renderers.py:370
CHANNELS = (128, 256, 512)


class LstmMarchDesc(ctypes.Structure):
    """Mirror of ``avr_lstm_march`` (include/avr_b200.h), field for field."""

    _fields_ = [
        ("ros", c_void_p), ("rds", c_void_p), ("init_dist", c_void_p),
        ("R", c_int64), ("rays_per_obj", c_int64), ("steps", c_int),
        ("w_ih", c_void_p), ("w_hh", c_void_p), ("b_ih", c_void_p), ("b_hh", c_void_p), ("w_out", c_void_p), ("b_out", c_void_p),
        ("world", c_void_p),
        ("feats", c_void_p), ("gates", c_void_p), ("cells", c_void_p), ("hidden", c_void_p),
        ("g_world", c_void_p), ("d_gates", c_void_p), ("d_dist", c_void_p),
    ]


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class _LstmMarch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, latent, w_ih, w_hh, b_ih, b_hh, w_out, b_out, ros, rds, init, poses, focal, c, cfg: FieldConfig, steps: int):
        require_cuda(latent, w_ih, w_hh, b_ih, b_hh, w_out, b_out, ros, rds, init, poses, focal, c)
        sb, n_rays = ros.shape[0], ros.shape[1]
        r = sb * n_rays
        dev = ros.device
        latent, poses, focal, c = _f32c(latent, "latent"), _f32c(poses, "poses"), _f32c(focal, "focal"), _f32c(c, "c")
        ros_c, rds_c = _f32c(ros.reshape(r, 3), "ros"), _f32c(rds.reshape(r, 3), "rds")
        init_c = _f32c(init.reshape(r), "initial distance")
        ws = [_f32c(t, n) for t, n in ((w_ih, "lstm.weight_ih"), (w_hh, "lstm.weight_hh"), (b_ih, "lstm.bias_ih"),
                                       (b_hh, "lstm.bias_hh"), (w_out.reshape(-1), "out_layer.weight"), (b_out, "out_layer.bias"))]
        ch = latent.shape[-1]
        if ws[0].shape != (4 * HIDDEN, ch) or ws[1].shape != (4 * HIDDEN, HIDDEN) or ws[4].numel() != HIDDEN:
            raise AvrError(f"lstm_march: LSTMCell({ch} -> {HIDDEN}) + Linear({HIDDEN} -> 1) expected, got weight_ih "
                           f"{tuple(ws[0].shape)}, weight_hh {tuple(ws[1].shape)}, out_layer {tuple(w_out.shape)}")
        save = any(ctx.needs_input_grad)
        world = torch.empty(steps + 1, r, 3, dtype=torch.float32, device=dev)
        saved = [None] * 4
        if save:
            saved = [torch.empty(steps, r, w, dtype=torch.float32, device=dev) for w in (ch, 4 * HIDDEN, HIDDEN, HIDDEN)]
        fd = _fill(cfg, ros.reshape(sb, n_rays, 3), None, latent, poses, focal, c, True)
        m = LstmMarchDesc()
        m.ros, m.rds, m.init_dist = ptr(ros_c), ptr(rds_c), ptr(init_c)
        m.R, m.rays_per_obj, m.steps = r, n_rays, steps
        m.w_ih, m.w_hh, m.b_ih, m.b_hh, m.w_out, m.b_out = (ptr(t) for t in ws)
        m.world = ptr(world)
        m.feats, m.gates, m.cells, m.hidden = (ptr(t) for t in saved)
        with torch.cuda.device(dev):
            check(_lib.load().avr_lstm_march_fwd(ctypes.byref(fd), ctypes.byref(m), _stream(dev)), "avr_lstm_march_fwd")
        ctx.set_materialize_grads(False)
        ctx.cfg, ctx.steps, ctx.shape = cfg, steps, (sb, n_rays)
        if save:
            ctx.save_for_backward(latent, poses, focal, c, ros_c, rds_c, world, *ws, *saved)
        return world[steps].view(sb, n_rays, 3)

    @staticmethod
    def backward(ctx, g_world):
        if g_world is None:
            return (None,) * 15
        latent, poses, focal, c, ros_c, rds_c, world, w_ih, w_hh, b_ih, b_hh, w_out, b_out, feats, gates, cells, hidden = ctx.saved_tensors
        sb, n_rays = ctx.shape
        r, steps, dev = sb * n_rays, ctx.steps, g_world.device
        g = _f32c(g_world.reshape(r, 3), "g_world")
        d_gates = torch.empty(steps, r, 4 * HIDDEN, dtype=torch.float32, device=dev)
        d_dist = torch.empty(steps, r, dtype=torch.float32, device=dev)
        d_latent = torch.zeros_like(latent) if ctx.needs_input_grad[0] else None
        fd = _fill(ctx.cfg, ros_c.view(sb, n_rays, 3), None, latent, poses, focal, c, True)
        fd.d_latent = ptr(d_latent)
        m = LstmMarchDesc()
        m.ros, m.rds = ptr(ros_c), ptr(rds_c)
        m.R, m.rays_per_obj, m.steps = r, n_rays, steps
        m.w_ih, m.w_hh, m.b_ih, m.b_hh, m.w_out, m.b_out = (ptr(t) for t in (w_ih, w_hh, b_ih, b_hh, w_out, b_out))
        m.world = ptr(world)
        m.feats, m.gates, m.cells, m.hidden = ptr(feats), ptr(gates), ptr(cells), ptr(hidden)
        m.g_world, m.d_gates, m.d_dist = ptr(g), ptr(d_gates), ptr(d_dist)
        with torch.cuda.device(dev):
            check(_lib.load().avr_lstm_march_bwd(ctypes.byref(fd), ctypes.byref(m), _stream(dev)), "avr_lstm_march_bwd")
        # parameter gradients: GEMMs over the (step, ray) rows the kernel wrote
        da = d_gates.view(steps * r, 4 * HIDDEN)
        h_prev = torch.cat([hidden.new_zeros(1, r, HIDDEN), hidden[:-1]], 0).view(steps * r, HIDDEN)
        dd = d_dist.view(1, steps * r)
        d_w_ih = da.t() @ feats.view(steps * r, -1) if ctx.needs_input_grad[1] else None
        d_w_hh = da.t() @ h_prev if ctx.needs_input_grad[2] else None
        d_b = da.sum(0)
        d_w_out = (dd @ hidden.view(steps * r, HIDDEN)) if ctx.needs_input_grad[5] else None
        d_b_out = dd.sum().reshape(1) if ctx.needs_input_grad[6] else None
        return (d_latent, d_w_ih, d_w_hh, d_b if ctx.needs_input_grad[3] else None, d_b if ctx.needs_input_grad[4] else None,
                d_w_out, d_b_out, None, None, None, None, None, None, None, None)


def march_supported(phi, lstm, out_layer) -> bool:
    """True when ``phi``'s feature fetch is the front-end kernels' (``fuse_field_inputs``) with one
    source view per object and the LSTM head has the reference's shape (renderers.py:370-378)."""
    st = getattr(phi, "_avr_field_state", None)
    if st is None or not isinstance(lstm, torch.nn.LSTMCell) or not isinstance(out_layer, torch.nn.Linear):
        return False
    if getattr(phi, "num_views_per_obj", 0) != 1 or getattr(phi.encoder, "latent", None) is None:
        return False
    ch = phi.encoder.latent.shape[1]
    return (ch in CHANNELS and lstm.input_size == ch and lstm.hidden_size == HIDDEN and lstm.bias
            and out_layer.in_features == HIDDEN and out_layer.out_features == 1 and out_layer.bias is not None
            and phi.encoder.latent.is_cuda and lstm.weight_ih.is_cuda)


def lstm_march(ros: torch.Tensor, rds: torch.Tensor, init_dist: torch.Tensor, phi, lstm: torch.nn.LSTMCell,
               out_layer: torch.nn.Linear, steps: int) -> torch.Tensor:
    """``world_coords[-1]`` of renderers.py:411-435: ros, rds ``(SB, R, 3)``, init_dist ``(SB, R, 1)``
    -> ``(SB, R, 3)``.  Differentiable w.r.t. the LSTM head and the encoder's feature map."""
    from .field import field_state

    cfg, nhwc = field_state(phi)
    return _LstmMarch.apply(nhwc, lstm.weight_ih, lstm.weight_hh, lstm.bias_ih, lstm.bias_hh, out_layer.weight,
                            out_layer.bias, ros, rds, init_dist, phi.poses, phi.focal, phi.c, cfg, int(steps))
