"""Radiance-field front end on one B200.

    python tools/bench_field.py [--rays 2048] [--samples 96] [--views 1] [--batches 8] [--iters 10]

Points are ordered as the renderer sends them (K consecutive samples per ray) for `--batches`
ray batches of conf/default.conf's shape (512-channel 64 x 64 feature map, 42-wide code).  Prints one
JSON line per kernel: ms, rows/s and the fraction of the HBM roofline, the algorithmic bytes being
the MLP input itself — (C + 42) * 4 B per (view, point) written by the forward pass and read (as
g_out) by the backward pass — plus 24 B of point/view-direction per point; the feature map (8 MB) is
L2-resident.  Next to it: the same work as the reference's torch ops on the same device
(oracle/field_oracle.py is NOT used here; the ops are spelled out below).
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import avr_b200  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def torch_front_end(xyz, vd, poses, focal, c, scale, latent, freqs, phases, ns):
    """models.py:754-826 as stock torch ops (what the reference runs on this device)."""
    sb, b, _ = xyz.shape
    p = xyz.unsqueeze(1).expand(-1, ns, -1, -1).reshape(-1, b, 3)
    rot = torch.matmul(poses[:, None, :3, :3], p.unsqueeze(-1))[..., 0]
    cam = rot + poses[:, None, :3, 3]
    zf = rot.reshape(-1, 3)
    emb = torch.sin(torch.addcmul(phases, zf.unsqueeze(1).repeat(1, freqs.shape[1], 1), freqs)).view(zf.shape[0], -1)
    zf = torch.cat((zf, emb), -1)
    d = vd.reshape(sb, b, 3, 1).unsqueeze(1).expand(-1, ns, -1, -1, -1).reshape(-1, b, 3, 1)
    zf = torch.cat((zf, torch.matmul(poses[:, None, :3, :3], d).reshape(-1, 3)), 1)
    uv = -cam[:, :, :2] / cam[:, :, 2:]
    uv = (uv * focal.unsqueeze(1) + c.unsqueeze(1)) * scale - 1.0
    s = F.grid_sample(latent, uv.unsqueeze(2), align_corners=True, mode="bilinear", padding_mode="border")
    return torch.cat((s[:, :, :, 0].transpose(1, 2).reshape(-1, latent.shape[1]), zf), -1)


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=2048)
    ap.add_argument("--samples", type=int, default=96)
    ap.add_argument("--views", type=int, default=1)
    ap.add_argument("--batches", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--raw-only", action="store_true", help="only the raw C-ABI kernel timings")
    return ap.parse_args(argv)


def run(a, dev=None):
    """Yields one record (dict) per measurement; `a` as from parse_args()."""
    dev = dev or torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    ch, h, w, ns = 512, 64, 64, a.views
    b = a.rays * a.samples
    # one pose per view at distance 1.3 looking at the origin; rays through the object
    poses = torch.zeros(ns, 3, 4, device=dev)
    for v in range(ns):
        th = 0.5 + 0.9 * v
        eye = 1.3 * torch.tensor([math.cos(th), math.sin(th), 0.4])
        fwd = -eye / eye.norm()
        right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0]))
        right = right / right.norm()
        up = torch.linalg.cross(right, fwd)
        rot = torch.stack([right, -up, fwd])                  # world -> view (OpenCV axes flipped as dataset.py:85-86)
        poses[v, :, :3] = rot.to(dev)
        poses[v, :, 3] = (-rot @ eye).to(dev)
    focal = torch.tensor([[131.25, -131.25]], device=dev)
    c = torch.tensor([[64.0, 64.0]], device=dev)
    ls = torch.tensor([float(w), float(h)], device=dev)
    scale = ls / (ls - 1) * 2.0 / torch.tensor([128.0, 128.0], device=dev)
    freqs = torch.repeat_interleave(1.5 * 2.0 ** torch.arange(0, 6), 2).view(1, -1, 1).to(dev)
    phases = torch.zeros(12)
    phases[1::2] = math.pi * 0.5
    phases = phases.view(1, -1, 1).to(dev)
    cfg = avr_b200.FieldConfig(ns=ns, scale=tuple(scale.tolist()), freqs=tuple(freqs.reshape(-1).tolist()),
                               phases=tuple(phases.reshape(-1).tolist()))
    latent = torch.randn(ns, ch, h, w, device=dev, generator=g)
    nhwc = latent.permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    sets = []
    for _ in range(a.batches):
        o = torch.randn(1, a.rays, 1, 3, device=dev, generator=g) * 0.05 + torch.tensor([0.9, 0.3, 0.9], device=dev)
        d = F.normalize(-o + 0.25 * torch.randn(1, a.rays, 1, 3, device=dev, generator=g), dim=-1)
        z = 0.8 + torch.sort(torch.rand(1, a.rays, a.samples, 1, device=dev, generator=g), 2).values
        xyz = (o + d * z).reshape(1, b, 3).contiguous().requires_grad_(True)
        vd = d.expand(1, a.rays, a.samples, 3).reshape(1, b, 3).contiguous().requires_grad_(True)
        sets.append((xyz, vd, torch.randn(ns * b, ch + 42, device=dev, generator=g)))
    rows = ns * b
    row_bytes = (ch + 42) * 4 + 24 / ns
    lib = avr_b200.load_library()
    from avr_b200 import field as F_
    import ctypes
    sp = torch.cuda.current_stream(dev).cuda_stream

    # ---- raw C-ABI launches with prebuilt descriptors: kernel time without the Python wrapper ----
    out = torch.empty(rows, ch + 42, device=dev)
    d_lat, d_xyz, d_vd = torch.empty_like(nhwc), torch.empty(1, b, 3, device=dev), torch.empty(1, b, 3, device=dev)
    descs = {"fwd": [], "bwd_all": [], "bwd_latent": [], "bwd_points": []}
    for xyz, vd, g_out in sets:
        for key in descs:
            d = F_._fill(cfg, xyz.detach(), vd.detach(), nhwc.detach(), poses, focal, c, False)
            d.out, d.g_out = out.data_ptr(), g_out.data_ptr()
            if key in ("bwd_all", "bwd_latent"):
                d.d_latent = d_lat.data_ptr()
            if key in ("bwd_all", "bwd_points"):
                d.d_xyz, d.d_viewdirs = d_xyz.data_ptr(), d_vd.data_ptr()
            descs[key].append(d)

    def raw(key):
        fn = lib.avr_field_inputs_fwd if key == "fwd" else lib.avr_field_inputs_bwd
        for d in descs[key][:2]:
            assert fn(ctypes.byref(d), sp) == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            for d in descs[key]:           # rotating inputs/outputs larger than L2
                fn(ctypes.byref(d), sp)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (a.iters * len(sets))

    pk = peak()
    tag = {k: os.environ.get(k) for k in ("AVR_FIELD_NOCACHE", "AVR_FIELD_BWD_SPLIT") if os.environ.get(k)}
    for key, name in (("fwd", "field_inputs_fwd"), ("bwd_all", "field_inputs_bwd (latent + points)"),
                      ("bwd_latent", "field_inputs_bwd (latent only)"), ("bwd_points", "field_inputs_bwd (points only)")):
        ms = raw(key)
        yield {"kernel": name, "rows": rows, "ms": round(ms, 4), "rows_per_s": rows / (ms * 1e-3),
               "GBps": round(rows * row_bytes / ms / 1e6, 1), "hbm_frac": round(rows * row_bytes / ms / 1e6 / pk, 4),
               "knobs": tag}
    if a.raw_only:
        return

    # ---- through the autograd API, next to the reference's torch ops on the same device -----------
    def timed(fn):
        for s_ in sets[:2]:
            fn(*s_)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            for s_ in sets:
                fn(*s_)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (a.iters * len(sets))

    outs = {}

    def fwd(xyz, vd, g_out):
        outs["o"] = avr_b200.field_inputs(xyz, vd, nhwc, poses, focal, c, cfg)

    def fwd_bwd(xyz, vd, g_out):
        o = avr_b200.field_inputs(xyz, vd, nhwc, poses, focal, c, cfg)
        torch.autograd.grad([o], [nhwc, xyz, vd], [g_out])

    lat_t = latent.clone().requires_grad_(True)

    def torch_fwd(xyz, vd, g_out):
        outs["t"] = torch_front_end(xyz, vd, poses, focal, c, scale, lat_t, freqs, phases, ns)

    def torch_fwd_bwd(xyz, vd, g_out):
        o = torch_front_end(xyz, vd, poses, focal, c, scale, lat_t, freqs, phases, ns)
        torch.autograd.grad([o], [lat_t, xyz, vd], [g_out])

    t_f, t_fb = timed(fwd), timed(fwd_bwd)
    tt_f, tt_fb = timed(torch_fwd), timed(torch_fwd_bwd)
    yield {"autograd_api": {"fwd_ms": round(t_f, 4), "fwd_bwd_ms": round(t_fb, 4)},
           "torch_eager_same_device": {"fwd_ms": round(tt_f, 4), "fwd_bwd_ms": round(tt_fb, 4)},
           "speedup_fwd": round(tt_f / t_f, 2), "speedup_fwd_bwd": round(tt_fb / t_fb, 2),
           "max_abs_diff_vs_torch_cuda": float((outs["o"] - outs["t"]).detach().abs().max())}


def main():
    for rec in run(parse_args()):
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
