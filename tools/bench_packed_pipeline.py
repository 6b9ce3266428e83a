"""BASELINE.json config 4: the adaptive ragged pipeline on packed rays.

    python tools/bench_packed_pipeline.py [--rays N] [--iters 5]

2^22 rays with per-ray sample counts 8..256 (packed offsets), per-ray bounds d -/+ 0.15
(AdaptiveVolumeRenderer-like), and the full path
    coarse sample -> composite (weights) -> importance sample K_r/2 + merge -> composite fwd + bwd
with synthetic radiance-field outputs between the stages.
Prints one JSON line per stage and one for the whole pipeline: ms, rays/s, samples/s and the
fraction of the HBM roofline (algorithmic bytes of section 8d / time / MEASURED_PEAKS.json).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import avr_b200  # noqa: E402
from avr_b200 import ops  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def synth_rgbs(n, dev, g):
    x = torch.empty(n, 4, device=dev)
    x[:, :3] = torch.sigmoid(torch.randn(n, 3, device=dev, generator=g))
    x[:, 3] = torch.relu(torch.randn(n, device=dev, generator=g)) * 30
    return x


def run(rays, iters, dev, warm=2):
    """Yields one record per stage and a final one for the whole pipeline."""
    lib = avr_b200.load_library()
    g = torch.Generator(device=dev).manual_seed(0)
    r = rays
    counts = torch.randint(8, 257, (r,), device=dev, generator=g)
    offsets = torch.zeros(r + 1, dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(counts, 0)
    s = int(offsets[-1])
    fo = torch.zeros(r + 1, dtype=torch.int64, device=dev)
    fo[1:] = torch.cumsum(counts // 2, 0)
    sf = int(fo[-1])
    mo = offsets + fo                      # merged layout
    d = 0.9 + 0.8 * torch.rand(r, device=dev, generator=g)
    near, far = d - 0.15, d + 0.15
    u = torch.rand(s, device=dev, generator=g)
    uf, uf2 = torch.rand(sf, device=dev, generator=g), torch.rand(sf, device=dev, generator=g)
    x_c = synth_rgbs(s, dev, g)
    x_f = synth_rgbs(s + sf, dev, g)
    g_rgb, g_d = torch.randn(r, 3, device=dev, generator=g), torch.randn(r, device=dev, generator=g)
    dx = torch.empty_like(x_f)
    sp = torch.cuda.current_stream(dev).cuda_stream
    state = {}

    def st_coarse():
        state["z"] = ops.coarse_sample_packed(near, far, u, offsets)

    def st_comp_c():
        state["w"] = ops.composite_packed_fwd_raw(x_c, state["z"], offsets, True, 1.8, True)[2]

    def st_imp():
        state["zs"] = ops.importance_sample_packed(state["w"], state["z"], near, far, uf, uf2, offsets, fo, 256, 128)[1]

    def st_comp_f():
        state["out"] = ops.composite_packed_fwd_raw(x_f, state["zs"], mo, True, 1.8, False)

    def st_bwd():
        rc = lib.avr_composite_bwd_packed(x_f.data_ptr(), state["zs"].data_ptr(), mo.data_ptr(), g_rgb.data_ptr(),
                                          g_d.data_ptr(), None, r, s + sf, 1, 1.8, dx.data_ptr(), None, sp)
        assert rc == 0, lib.avr_last_cuda_error()

    stages = [
        ("coarse_sample_packed", st_coarse, 8 * s + 24 * r),
        ("composite_fwd_packed (coarse, weights)", st_comp_c, 24 * s + 24 * r),
        ("importance_sample_packed + merge", st_imp, 12 * s + 12 * sf + 24 * r),
        ("composite_fwd_packed (fine)", st_comp_f, 20 * (s + sf) + 24 * r),
        ("composite_bwd_packed (fine)", st_bwd, 36 * (s + sf) + 24 * r),
    ]
    for _ in range(1 + warm):    # warm-up, also builds the state the later stages read
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
    total_bytes = 0
    for i, (name, _, nbytes) in enumerate(stages):
        ms = sum(ev[it][i].elapsed_time(ev[it][i + 1]) for it in range(iters)) / iters
        total_bytes += nbytes
        yield {"stage": name, "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1),
               "hbm_frac": round(nbytes / ms / 1e6 / pk, 4), "bytes": nbytes}
    ms = sum(ev[it][0].elapsed_time(ev[it][-1]) for it in range(iters)) / iters
    yield {"pipeline": "coarse -> composite -> importance+merge -> composite fwd+bwd (packed, counts 8..256)",
           "rays": r, "coarse_samples": s, "fine_samples": sf, "ms": round(ms, 3),
           "rays_per_s": r / (ms * 1e-3), "composited_samples_per_s": (2 * s + sf) / (ms * 1e-3),
           "GBps": round(total_bytes / ms / 1e6, 1), "hbm_frac": round(total_bytes / ms / 1e6 / pk, 4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1 << 22)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    for rec in run(a.rays, a.iters, torch.device("cuda:0")):
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
