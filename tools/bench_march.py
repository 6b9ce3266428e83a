"""the adaptive renderer's LSTM ray march (renderers.py:411-435) —
one persistent kernel (csrc/lstm_march.cu) against the reference's loop in eager torch on the same GPU
(whose feature fetch is already the fused front-end kernel, so the comparison isolates the loop).

    python tools/bench_march.py [--iters 20]

conf/default.conf's adaptive_renderer: raymarch_steps 10, 512 feature channels, 64 x 64 feature map;
2048 rays (train.py's batch) and 16384 rays (a 128 x 128 frame), forward (no_grad) and forward+backward.
"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch import nn  # noqa: E402


class _Code(nn.Module):
    def __init__(self):
        super().__init__()
        self.include_input = True
        phases = torch.zeros(12)
        phases[1::2] = math.pi * 0.5
        self.register_buffer("_freqs", torch.repeat_interleave(1.5 * 2.0 ** torch.arange(0, 6), 2).view(1, -1, 1))
        self.register_buffer("_phases", phases.view(1, -1, 1))


class _Encoder(nn.Module):
    index_interp, index_padding = "bilinear", "border"

    def __init__(self, sb, ch, h, w):
        super().__init__()
        # smooth features (an encoder's output is; white noise would make the march chaotic and the
        # fused-vs-loop agreement check meaningless)
        self.latent_param = nn.Parameter(torch.nn.functional.interpolate(torch.randn(sb, ch, 6, 6), size=(h, w), mode="bicubic",
                                                                         align_corners=True) * 0.3)
        self.latent = None
        self.register_buffer("latent_scaling", torch.tensor([w / (w - 1) * 2.0, h / (h - 1) * 2.0]))


class FeatureField(nn.Module):
    """The attributes of NewPixelNeRFNet the fused front end touches (models.py:609-866); the feature
    map is a parameter so that d_latent has somewhere to go."""

    use_encoder = use_xyz = use_code = use_viewdirs = normalize_z = True
    use_global_encoder = use_code_viewdirs = stop_encoder_grad = False
    d_out = 4
    num_views_per_obj = 1

    def __init__(self, sb, ch=512, h=64, w=64):
        super().__init__()
        self.encoder, self.code = _Encoder(sb, ch, h, w), _Code()
        self.mlp_coarse = self.mlp_fine = None
        eye = torch.eye(4)
        eye[2, 3] = 1.3
        c2w = eye @ torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0]))
        rot = c2w[:3, :3].t()
        self.register_buffer("poses", torch.cat((rot, -rot @ c2w[:3, 3:]), -1).expand(sb, 3, 4).contiguous())
        self.register_buffer("image_shape", torch.tensor([128.0, 128.0]))
        self.register_buffer("focal", torch.tensor([[131.25, -131.25]]))
        self.register_buffer("c", torch.tensor([[64.0, 64.0]]))

    def refresh(self):
        self.encoder.latent = self.encoder.latent_param * 1.0     # a non-leaf, like the encoder's output


def run(dev, iters=20):
    import avr_b200
    out = {}
    for sb, r, label in ((4, 512, "train.py batch, SB=4 x 512 rays"), (1, 16384, "128x128 frame")):
        torch.manual_seed(0)
        phi = FeatureField(sb).to(dev)
        phi.refresh()
        avr_b200.fuse_field_inputs(phi)
        ren = avr_b200.AdaptiveVolumeRenderer(512, raymarch_steps=10, epsilon=0.15, n_coarse=20, white_back=True).to(dev)
        g = torch.Generator(device=dev).manual_seed(1)
        ros = torch.tensor([0.0, 0.0, 1.3], device=dev).expand(sb, r, 3).contiguous()
        d = torch.randn(sb, r, 3, device=dev, generator=g) * 0.15 + torch.tensor([0.0, 0.0, -1.0], device=dev)
        rds = torch.nn.functional.normalize(d, dim=-1)
        init = 0.8 + 0.05 * torch.randn(sb, r, 1, device=dev, generator=g)
        params = list(ren.parameters()) + list(phi.parameters())

        def step(fused, train):
            ren.fused_march = fused
            if train:
                for p in params:
                    p.grad = None
                phi.refresh()
                ren.march(ros, rds, init, phi).square().mean().backward()
            else:
                with torch.no_grad():
                    ren.march(ros, rds, init, phi)

        def timeit(fused, train):
            for _ in range(3):
                step(fused, train)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(iters):
                step(fused, train)
            e1.record()
            torch.cuda.synchronize(dev)
            return {"wall_ms": round((time.perf_counter() - t0) / iters * 1e3, 4), "device_ms": round(e0.elapsed_time(e1) / iters, 4)}

        rec = {"rays": sb * r, "steps": 10, "channels": 512}
        for train in (False, True):
            key = "fwd_bwd" if train else "fwd"
            rec[key] = {"fused_kernel": timeit(True, train), "torch_loop": timeit(False, train)}
            rec[key]["speedup_wall"] = round(rec[key]["torch_loop"]["wall_ms"] / rec[key]["fused_kernel"]["wall_ms"], 2)
        # agreement of the two paths on this workload
        with torch.no_grad():
            ren.fused_march = True
            a = ren.march(ros, rds, init, phi)
            ren.fused_march = False
            b = ren.march(ros, rds, init, phi)
        rec["max_abs_diff_fused_vs_loop"] = float((a - b).abs().max())
        out[label] = rec
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    print(json.dumps(run(torch.device("cuda:0"), a.iters)))
