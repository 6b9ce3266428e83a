"""Drop-in replacements for the reference's ``renderers.py`` hot path.

Same names, signatures and return conventions as the reference so that
``train.py``/``test.py``/``models.RadFieldAndRenderer`` (models.py:913-929) can use them
unchanged:

* free functions ``sample_coarse`` / ``sample_fine`` / ``sample_depth`` /
  ``volume_integral``                                     (renderers.py:4, 27, 56, 69)
* ``VolumeRenderer``                                      (renderers.py:121-289)
* ``AdaptiveVolumeRenderer``                              (renderers.py:360-557)

Differences, all additive: every function that draws random numbers accepts the draws as
optional keyword arguments (``u=``, ``u2=``, ``normals=``, ``draws=``) so tests and
benchmarks can feed the CUDA kernels and the oracle identical numbers; when they are
omitted the draws are made with the same torch calls, in the same order, as the reference
(SURVEY.md section 0.5).  Everything runs on CUDA (sm_100a); CPU tensors raise.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
from torch import nn

from . import ops
from ._lib import AvrError
from .geometry import depth_from_world, get_world_rays


def _to_gpu(t: torch.Tensor) -> torch.Tensor:
    # utils.to_gpu (utils.py:92-103): .cuda() when possible, unchanged otherwise
    try:
        return t.cuda()
    except Exception:
        return t


# ------------------------------------------------------------------ free functions
def sample_coarse(near_depth, far_depth, num_samples: int, device=None, infinity=-1, *, u: Optional[torch.Tensor] = None):
    """Stratified depths (SB, R, num_samples).  renderers.py:4-24."""
    lead = near_depth.shape
    if u is None:
        u = torch.rand(*lead, num_samples, dtype=torch.float32, device=near_depth.device if device is None else device)
    if near_depth.requires_grad or far_depth.requires_grad:
        z = ops.CoarseSample.apply(near_depth, far_depth, u)
    else:
        n, f, stride = ops._bounds(near_depth, far_depth, u.numel() // num_samples)
        z = ops.coarse_sample_raw(n, f, stride, u)
    if infinity != -1:  # never taken by the reference's callers; kept for signature parity (:16-19)
        tail = torch.broadcast_to(torch.tensor([infinity], device=z.device, dtype=z.dtype), z[..., :1].shape)
        z = torch.cat([z[..., 1:], tail], -1)
    return z


def sample_fine(near_depth, far_depth, num_samples: int, weights, device=None, *,
                u: Optional[torch.Tensor] = None, u2: Optional[torch.Tensor] = None):
    """Importance samples (SB, R, num_samples), unsorted.  renderers.py:27-54."""
    sb, r = weights.shape[0], weights.shape[1]
    dev = weights.device if device is None else device
    if u is None:
        u = torch.rand(sb, r, num_samples, dtype=torch.float32, device=dev)
    if u2 is None:
        u2 = torch.rand_like(u)
    return ops.importance_sample(weights, near_depth, far_depth, u, u2)["z_fine"]


def sample_depth(depth, num_samples: int, depth_std, *, normals: Optional[torch.Tensor] = None):
    """``randn * depth_std`` — the depth is NOT added, exactly like the reference
    (renderers.py:56-66); the caller clamps to [near, far]."""
    sb, r, _ = depth.shape
    if normals is None:
        normals = torch.randn(sb, r, num_samples, dtype=depth.dtype, device=depth.device)
    return normals * depth_std


def _same_layout(a: torch.Tensor, b: torch.Tensor) -> bool:
    if a.shape != b.shape or a.data_ptr() != b.data_ptr():
        return False
    return all(sa == sb for n, sa, sb in zip(a.shape, a.stride(), b.stride()) if n != 1)


def _as_rgbs(sigmas: torch.Tensor, radiances: torch.Tensor) -> torch.Tensor:
    """The reference passes sigma (...,K,1) and radiance (...,K,3) as two slices of one
    (…,K,4) buffer (renderers.py:177-178).  If that is what we were given, use the buffer
    itself (no copy, one gradient); otherwise interleave."""
    base = sigmas._base
    if base is not None and base is radiances._base and base.is_contiguous() \
            and base.numel() == sigmas.numel() * 4 and base.shape[-1] == 4:
        cand = base.view(*sigmas.shape[:-1], 4)
        if _same_layout(sigmas, cand[..., 3:4]) and _same_layout(radiances, cand[..., :3]):
            return cand
    return torch.cat([radiances, sigmas], dim=-1)


def volume_integral(z_vals, sigmas, radiances, white_back=True, infinity=1.8):
    """rgb (SB,R,3), depth (SB,R,1), weights (SB,R,K,1).  renderers.py:69-119."""
    rgbs = _as_rgbs(sigmas, radiances)
    rgb, depth, w = ops.composite(rgbs, z_vals, bool(white_back), float(infinity), want_w=True)
    return rgb, depth.unsqueeze(-1), w.unsqueeze(-1)


def volume_integral_rgbs(z_vals, rgbs, white_back=True, infinity=1.8, want_weights=True):
    """Same, taking the radiance field's (…, K, 4) output directly (the zero-copy path the
    renderers use).  ``want_weights=False`` skips the weight store (the fine pass discards
    them, renderers.py:270)."""
    rgb, depth, w = ops.composite(rgbs, z_vals, bool(white_back), float(infinity), want_w=want_weights)
    return rgb, depth.unsqueeze(-1), (w.unsqueeze(-1) if w is not None else None)


# ------------------------------------------------------------------ VolumeRenderer
class VolumeRenderer(nn.Module):
    """Coarse + importance-resampled volume renderer; no parameters or buffers
    (``state_dict()`` is empty, as the reference's)."""

    def __init__(self, near, far, n_coarse, n_fine, n_fine_depth, depth_std, white_back=True):
        super().__init__()
        self.near = _to_gpu(torch.tensor([near]))
        self.far = _to_gpu(torch.tensor([far]))
        self.n_coarse = n_coarse
        self.n_fine = n_fine
        self.n_fine_depth = n_fine_depth
        self.depth_std = depth_std
        self.white_back = white_back

    def forward(self, cam2world, intrinsics, x_pix, radiance_field: nn.Module,
                draws: Optional[Sequence[torch.Tensor]] = None) -> Tuple[torch.Tensor, ...]:
        sb, num_rays, _ = x_pix.shape
        dev = x_pix.device
        if not x_pix.is_cuda:
            raise AvrError("VolumeRenderer (avr_b200) needs CUDA inputs; there is no CPU fallback")
        kc, ki, kd = self.n_coarse, self.n_fine - self.n_fine_depth, self.n_fine_depth
        if draws is None:
            # the reference's draw order: renderers.py:14, :41, :45, :63
            draws = (
                torch.rand(sb, num_rays, kc, dtype=torch.float32, device=dev),
                torch.rand(sb, num_rays, ki, dtype=torch.float32, device=dev),
                torch.rand(sb, num_rays, ki, dtype=torch.float32, device=dev),
                torch.randn(sb, num_rays, kd, dtype=torch.float32, device=dev),
            )
        u_c, u_cdf, u_bin, normals = draws
        near, far = self.near.to(dev), self.far.to(dev)
        white_back = bool(self.white_back)

        # ray setup, coarse depths, sample points and the view-direction copy in one kernel   # :166-175
        ros, rds, aff, z_c, pts, vd = ops.rays_coarse_sample_points(x_pix, intrinsics, cam2world, near, far, 0, u_c)
        out = radiance_field(pts.view(sb, -1, 3), viewdirs=vd.view(sb, -1, 3), coarse=True)   # :173
        rgb_c, _dist_c, w_c = ops.composite(out.reshape(sb, num_rays, kc, 4), z_c, white_back, 1.8, want_w=True)  # :180

        # importance + "depth" resampling, merged and sorted in one kernel             # :252-258
        z_s = ops.importance_sample(w_c, near, far, u_cdf, u_bin, z_coarse=z_c,
                                    normals=normals if kd > 0 else None, depth_std=self.depth_std,
                                    want_fine=False, want_sorted=True)["z_sorted"]
        k = kc + self.n_fine
        pts, vd = ops.ray_points(ros, rds, z_s)                                       # :260-265
        out = radiance_field(pts.view(sb, -1, 3), viewdirs=vd.view(sb, -1, 3), coarse=False)  # :263
        # the fine composite returns the camera depth of ros + rds * dist itself              # :270-275
        rgb_f, depth, _ = ops.composite(out.reshape(sb, num_rays, k, 4), z_s, white_back, 1.8, want_w=False,
                                        depth_affine=aff)
        return rgb_c, rgb_f, depth, depth

    @classmethod
    def from_conf(cls, conf, white_back=True):
        return cls(
            near=conf.get_float("near", 0.8),
            far=conf.get_float("far", 1.8),
            n_coarse=conf.get_int("n_coarse", 32),
            n_fine=conf.get_int("n_fine", 16),
            n_fine_depth=conf.get_int("n_fine_depth", 8),
            depth_std=conf.get_float("depth_std", 0.01),
            white_back=conf.get_float("white_back", white_back),
        )


# ---------------------------------------------------------- AdaptiveVolumeRenderer
def _init_recurrent_weights(module: nn.Module):
    # utils.init_recurrent_weights (utils.py:108-117) only touches GRU/LSTM/RNN modules, so
    # for the LSTMCell used here it leaves torch's default init in place (and consumes no
    # RNG) — kept that way so a shared seed builds identical parameters.
    for m in module.modules():
        if type(m) in (nn.GRU, nn.LSTM, nn.RNN):
            for name, p in m.named_parameters():
                if "weight_ih" in name:
                    nn.init.kaiming_normal_(p.data)
                elif "weight_hh" in name:
                    nn.init.orthogonal_(p.data)
                elif "bias" in name:
                    p.data.fill_(0)


def _forget_gate_init(cell: nn.LSTMCell):
    for name, p in cell.named_parameters():
        if "bias" in name:
            n = p.size(0)
            p.data[n // 4: n // 2].fill_(1.0)


def _lstm_march(module, ros, rds, init, phi):
    """The LSTM ray march shared by ``Raymarcher`` (renderers.py:313-351) and ``AdaptiveVolumeRenderer``
    (:415-435): one persistent kernel when ``phi``'s feature fetch is the front-end kernels'
    (``fuse_field_inputs``, one source view per object); otherwise — an arbitrary ``phi`` callable —
    the reference's loop around ``phi``."""
    from . import march as _march

    sb, num_rays, _ = ros.shape
    if module.fused_march and _march.march_supported(phi, module.lstm, module.out_layer):
        return _march.lstm_march(ros, rds, init, phi, module.lstm, module.out_layer, module.steps)
    world = ros + rds * init
    state = None
    for _ in range(module.steps):
        v = phi(world.reshape(sb, -1, 3), viewdirs=rds.reshape(sb, -1, 3), return_features=True)
        state = module.lstm(v.reshape(-1, module.n_feature_channels), state)
        if state[0].requires_grad:
            state[0].register_hook(lambda x: x.clamp(min=-10, max=10))
        signed_distance = module.out_layer(state[0]).view(sb, num_rays, 1)
        world = world + rds * signed_distance
    return world


class Raymarcher(nn.Module):
    """The reference's LSTM sphere tracer (renderers.py:292-358): march, then colour and depth at the
    point reached — no sampling, no compositing.  Same constructor, ``from_conf(conf, raymarch_steps)``
    and 4-tuple ``(rgb, None, depth, depth)``; parameters ``lstm.*`` / ``out_layer.*`` as in the
    reference, so checkpoints load."""

    def __init__(self, num_feature_channels, raymarch_steps):
        super().__init__()
        self.n_feature_channels = num_feature_channels
        self.steps = raymarch_steps
        hidden_size = 16
        self.lstm = nn.LSTMCell(input_size=self.n_feature_channels, hidden_size=hidden_size)
        _init_recurrent_weights(self.lstm)
        _forget_gate_init(self.lstm)
        self.out_layer = nn.Linear(hidden_size, 1)
        self.counter = 0
        self.fused_march = True

    def forward(self, cam2world, intrinsics, xy_pix, phi, draws: Optional[Sequence[torch.Tensor]] = None):
        sb, num_rays, _ = xy_pix.shape
        if not xy_pix.is_cuda:
            raise AvrError("Raymarcher (avr_b200) needs CUDA inputs; there is no CPU fallback")
        ros, rds = get_world_rays(xy_pix, intrinsics=intrinsics, cam2world=cam2world)      # :318
        if draws is None:       # :320 draws on the CPU generator
            init = torch.zeros((sb, num_rays, 1)).normal_(mean=0.8, std=5e-2).to(xy_pix.device)
        else:
            (init,) = draws
        world = _lstm_march(self, ros, rds, init, phi)                                      # :321-343
        self.counter += 1
        out = phi(world.reshape(sb, -1, 3), viewdirs=rds.reshape(sb, -1, 3), coarse=True, return_features=False)  # :346
        rgb = out[..., :3].reshape(sb, num_rays, 3)
        final_depth = depth_from_world(world, cam2world).reshape(sb, num_rays, -1)          # :349
        return rgb, None, final_depth, final_depth

    @classmethod
    def from_conf(cls, conf, raymarch_steps):
        return cls(num_feature_channels=conf.get_int("num_feature_channels", 512), raymarch_steps=raymarch_steps)


class AdaptiveVolumeRenderer(nn.Module):
    """LSTM ray-march to a surface estimate d, then a thin stratified slab [d-eps, d+eps]
    composited with the fused kernels.  The march (renderers.py:411-435)
    is one persistent kernel when the radiance field allows it (``march``), the reference's loop
    otherwise; the tail (:489-509) is the hot path: coarse sampling with
    per-ray bounds that carry grad, a gradient-routing sort, and compositing with d_z."""

    def __init__(self, num_feature_channels, raymarch_steps, epsilon, n_coarse, white_back):
        super().__init__()
        self.epsilon = epsilon
        self.n_coarse = n_coarse
        self.white_back = white_back
        self.n_feature_channels = num_feature_channels
        self.steps = raymarch_steps
        hidden_size = 16
        self.lstm = nn.LSTMCell(input_size=self.n_feature_channels, hidden_size=hidden_size)
        _init_recurrent_weights(self.lstm)
        _forget_gate_init(self.lstm)
        self.out_layer = nn.Linear(hidden_size, 1)
        self.counter = 0
        self.fused_march = True      # False: always the reference's torch loop (A/B measurements)

    def forward(self, cam2world, intrinsics, xy_pix, phi, debug=False,
                draws: Optional[Sequence[torch.Tensor]] = None):
        sb, num_rays, _ = xy_pix.shape
        dev = xy_pix.device
        if not xy_pix.is_cuda:
            raise AvrError("AdaptiveVolumeRenderer (avr_b200) needs CUDA inputs; there is no CPU fallback")
        ros, rds, aff = ops.world_rays(xy_pix, intrinsics, cam2world, want_affine=True)  # :411
        if draws is None:
            # :413 draws the initial distance on the CPU generator, :14 (via :492) on the device
            init = torch.zeros((sb, num_rays, 1)).normal_(mean=0.8, std=5e-2).to(dev)
            u = None
        else:
            init, u = draws
        world = self.march(ros, rds, init, phi)                                         # :415-435

        out_c = phi(world.reshape(sb, -1, 3), viewdirs=rds.reshape(sb, -1, 3), coarse=True, return_features=False)
        rgb_coarse = out_c[..., :3].reshape(sb, num_rays, 3)                            # :485
        depth_coarse = depth_from_world(world, cam2world).reshape(sb, num_rays, -1)     # :486

        final_distance = (world[..., 0] - ros[..., 0]) / rds[..., 0]                    # :490
        if u is None:
            u = torch.rand(sb, num_rays, self.n_coarse, dtype=torch.float32, device=dev)
        z = sample_coarse(final_distance - self.epsilon, final_distance + self.epsilon,
                          self.n_coarse, device=dev, u=u)                               # :492
        z_sorted, _ = ops.SortRays.apply(z)                                             # :494
        pts, vd = ops.ray_points(ros, rds, z_sorted)                                    # :496-498 (d_z flows back)
        out = phi(pts.view(sb, -1, 3), coarse=False, viewdirs=vd.view(sb, -1, 3), return_features=False)  # :499
        rgb, depth, _ = ops.composite(out.reshape(sb, num_rays, self.n_coarse, 4), z_sorted,
                                      bool(self.white_back), 1.8, want_w=False, depth_affine=aff)  # :505-509
        dist = depth
        if debug:
            print("now AVR")
            print(f" pixel location is {xy_pix[0][64]}")
            print(f" distances are {z_sorted[0][64].squeeze()}")
            print(f" distance is {dist[0][64]}")
            print(f" color is {rgb[0][64][0]}")
            print(f" depth is {depth[0][64]}")
        return rgb_coarse, rgb, depth_coarse, depth

    def march(self, ros, rds, init, phi):
        """The LSTM ray march (renderers.py:415-435), see ``_lstm_march``."""
        return _lstm_march(self, ros, rds, init, phi)

    @classmethod
    def from_conf(cls, conf, white_back=False):
        return cls(
            num_feature_channels=conf.get_int("num_feature_channels", 512),
            raymarch_steps=conf.get_int("raymarch_steps", 10),
            epsilon=conf.get_float("epsilon", 0.05),
            n_coarse=conf.get_int("n_coarse", 20),
            white_back=conf.get_float("white_back", white_back),
        )
