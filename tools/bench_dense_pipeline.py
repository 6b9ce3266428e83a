"""The north-star sentence as one number: "fused sample + composite forward/backward at >= 70 % of B200 HBM
bandwidth".  The kernels of one VolumeRenderer render (conf/default.conf: 64 coarse + 16 importance + 16
"depth" samples, renderers.py:166-275) over 2^20 rays with synthetic radiance-field outputs between them:

    rays + coarse depths + points + view directions  ->  composite (weights)  ->  importance + depth samples
    + merge  ->  points + view directions  ->  composite (camera depth)  ->  both composites backward

    python tools/bench_dense_pipeline.py [--rays N] [--iters 5]

One JSON record per stage and one for the whole: ms, algorithmic bytes (SURVEY.md section 8d per stage),
fraction of the measured HBM roofline.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from avr_b200 import ops  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def synth(r, k, dev, g):
    x = torch.empty(1, r, k, 4, device=dev)
    x[..., :3] = torch.sigmoid(torch.randn(1, r, k, 3, device=dev, generator=g))
    x[..., 3] = torch.relu(torch.randn(1, r, k, device=dev, generator=g)) * 30
    return x


def run(rays, iters, dev, warm=2):
    kc, ki, kd = 64, 16, 16
    k = kc + ki + kd
    r = rays
    g = torch.Generator(device=dev).manual_seed(0)
    x_pix = torch.rand(1, r, 2, device=dev, generator=g)
    f = 131.25 / 128.0
    intr = torch.tensor([[[f, 0.0, 0.5], [0.0, f, 0.5], [0.0, 0.0, 1.0]]], device=dev)
    c2w = torch.eye(4, device=dev)
    c2w[2, 3] = 1.3
    c2w = (c2w @ torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0], device=dev))).expand(1, r, 4, 4).contiguous()
    near, far = torch.tensor([0.8], device=dev), torch.tensor([1.8], device=dev)
    u_c = torch.rand(1, r, kc, device=dev, generator=g)
    u_cdf, u_bin = torch.rand(1, r, ki, device=dev, generator=g), torch.rand(1, r, ki, device=dev, generator=g)
    normals = torch.randn(1, r, kd, device=dev, generator=g)
    x_c, x_f = synth(r, kc, dev, g), synth(r, k, dev, g)
    g_rgb, g_d = torch.randn(1, r, 3, device=dev, generator=g), torch.randn(1, r, device=dev, generator=g)
    st = {}

    def s_front():
        st["ros"], st["rds"], st["aff"], st["z_c"], st["pts"], st["vd"] = ops.rays_coarse_sample_points(x_pix, intr, c2w, near, far, 0, u_c)

    def s_comp_c():
        st["w_c"] = ops.composite_fwd_raw(x_c, st["z_c"], True, 1.8, True)[2]

    def s_imp():
        st["z_s"] = ops.importance_sample(st["w_c"], near, far, u_cdf, u_bin, z_coarse=st["z_c"], normals=normals, depth_std=0.01,
                                          want_fine=False, want_sorted=True)["z_sorted"]

    def s_pts():
        st["pts_f"] = ops.ray_points(st["ros"], st["rds"], st["z_s"])

    def s_comp_f():
        st["out"] = ops.composite_fwd_raw(x_f, st["z_s"], True, 1.8, False, st["aff"])

    def s_bwd_f():
        st["dx_f"] = ops.composite_bwd_raw(x_f, st["z_s"], g_rgb, g_d, None, True, 1.8, False, st["aff"])[0]

    def s_bwd_c():
        st["dx_c"] = ops.composite_bwd_raw(x_c, st["z_c"], g_rgb, None, None, True, 1.8, False)[0]

    stages = [
        ("rays + coarse sample + points + view directions (K=64)", s_front, r * (8 + 64 + 4 * kc + 32 + 28 * kc)),
        ("composite fwd, coarse (weights)", s_comp_c, r * (24 * kc + 16)),
        ("importance + depth samples + merge (64 -> 16 + 16)", s_imp, r * (12 * kc + 12 * ki + 8 * kd + 8)),
        ("points + view directions (K=96)", s_pts, r * (28 * k + 24)),
        ("composite fwd, fine (camera depth)", s_comp_f, r * (20 * k + 24)),
        ("composite bwd, fine", s_bwd_f, r * (36 * k + 24)),
        ("composite bwd, coarse", s_bwd_c, r * (36 * kc + 12)),
    ]
    # every stage drops its previous result BEFORE it runs, so the caching allocator hands the same block back
    # (with two generations alive, a 1.2 - 2.4 GB result can cost a cudaMalloc inside the timed region: 53 ms
    # instead of 0.48 ms for the fine points when this ran inside bench.py)
    keys = (("ros", "rds", "aff", "z_c", "pts", "vd"), ("w_c",), ("z_s",), ("pts_f",), ("out",), ("dx_f",), ("dx_c",))
    stages = [(name, (lambda fn=fn, ks=ks: ([st.pop(k_, None) for k_ in ks], fn())), nbytes)
              for (name, fn, nbytes), ks in zip(stages, keys)]
    for _ in range(1 + warm):
        for _, fn, _ in stages:
            fn()
    torch.cuda.synchronize(dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)] for _ in range(iters)]
    for it in range(iters):
        ev[it][0].record()
        for i, (_, fn, _) in enumerate(stages):
            fn()
            ev[it][i + 1].record()
    torch.cuda.synchronize(dev)
    pk = peak()
    total = 0
    for i, (name, _, nbytes) in enumerate(stages):
        ms = sum(ev[it][i].elapsed_time(ev[it][i + 1]) for it in range(iters)) / iters
        total += nbytes
        yield {"stage": name, "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1), "hbm_frac": round(nbytes / ms / 1e6 / pk, 4)}
    ms = sum(ev[it][0].elapsed_time(ev[it][-1]) for it in range(iters)) / iters
    yield {"pipeline": "one VolumeRenderer render, sampling + compositing forward and backward (64 + 16 + 16 samples), synthetic field outputs",
           "rays": r, "ms": round(ms, 3), "rays_per_s": r / (ms * 1e-3), "composited_samples_per_s": r * (kc + k) / (ms * 1e-3),
           "GBps": round(total / ms / 1e6, 1), "hbm_frac": round(total / ms / 1e6 / pk, 4)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1 << 20)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    for rec in run(a.rays, a.iters, torch.device("cuda:0")):
        print(json.dumps(rec), flush=True)
