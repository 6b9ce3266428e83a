"""Microbenchmarks of the sampling kernels and the packed pipeline (BASELINE.json configs 3, 4).

    python tools/bench_samplers.py [--rays N] [--iters 20] [--what coarse importance packed]

Prints one JSON line per kernel: ms, rays/s and the fraction of the HBM roofline
(algorithmic bytes of SURVEY.md section 8d / time / MEASURED_PEAKS.json copy bandwidth).
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


WARMUP = 3


def timeit(fn, iters):
    for _ in range(WARMUP):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, rays, bytes_per_ray, **kw):
    gbs = bytes_per_ray * rays / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "rays_per_s": rays / (ms * 1e-3), "GBps": round(gbs, 1),
                      "hbm_frac": round(gbs / peak(), 4), "bytes_per_ray": bytes_per_ray, **kw}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1 << 20)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--what", nargs="+", default=["coarse", "importance", "packed", "geometry", "adaptive"])
    ap.add_argument("--min-count", type=int, default=8, help="packed: shortest ray (BASELINE.json config 4: 8)")
    ap.add_argument("--warmup", type=int, default=3, help="untimed launches per kernel (1 for ncu captures)")
    a = ap.parse_args()
    global WARMUP
    WARMUP = a.warmup
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    r = a.rays
    near = torch.tensor([0.8], device=dev)
    far = torch.tensor([1.8], device=dev)
    if "coarse" in a.what:
        for k in (64, 96):
            u = torch.rand(1, r, k, device=dev, generator=g)
            ms = timeit(lambda: ops.coarse_sample_raw(near, far, 0, u), a.iters)
            report(f"coarse_sample K={k}", ms, r, 8 * k + 8)
    if "importance" in a.what:
        for kc, n, nd in ((64, 128, 0), (64, 16, 16)):
            w = torch.rand(1, r, kc, device=dev, generator=g) ** 6
            u = torch.rand(1, r, n, device=dev, generator=g)
            u2 = torch.rand(1, r, n, device=dev, generator=g)
            nrm = torch.randn(1, r, nd, device=dev, generator=g) if nd else None
            zc = ops.coarse_sample_raw(near, far, 0, torch.rand(1, r, kc, device=dev, generator=g))
            ms = timeit(lambda: ops.importance_sample(w, near, far, u, u2, z_coarse=zc, normals=nrm, depth_std=0.01,
                                                      want_fine=False, want_sorted=True), a.iters)
            report(f"importance_sample+merge {kc}->{n}+{nd}", ms, r, 12 * kc + 12 * n + 8 * nd + 8)
            ms = timeit(lambda: ops.importance_sample(w, near, far, u, u2, want_fine=True), a.iters)
            report(f"importance_sample only {kc}->{n}", ms, r, 4 * kc + 12 * n + 8)
    if "geometry" in a.what:
        import avr_b200
        lib = avr_b200.load_library()
        sp = torch.cuda.current_stream().cuda_stream
        ros = torch.randn(1, r, 3, device=dev, generator=g)
        rds = torch.nn.functional.normalize(torch.randn(1, r, 3, device=dev, generator=g), dim=-1)
        for k in (64, 96):
            u = torch.rand(1, r, k, device=dev, generator=g)
            z = ops.coarse_sample_raw(near, far, 0, u)
            pts = torch.empty(1, r, k, 3, device=dev)
            vd = torch.empty_like(pts)

            def points():
                rc = lib.avr_ray_points_fwd(ros.data_ptr(), rds.data_ptr(), z.data_ptr(), r, k, pts.data_ptr(), vd.data_ptr(), sp)
                assert rc == 0

            def fused():
                rc = lib.avr_coarse_sample_points_fwd(near.data_ptr(), far.data_ptr(), 0, u.data_ptr(), ros.data_ptr(),
                                                      rds.data_ptr(), r, k, z.data_ptr(), pts.data_ptr(), vd.data_ptr(), sp)
                assert rc == 0

            report(f"ray_points (pts + viewdirs) K={k}", timeit(points, a.iters), r, 28 * k + 24)
            report(f"coarse_sample + points fused K={k}", timeit(fused, a.iters), r, 32 * k + 32)
            ms = timeit(lambda: (ros.unsqueeze(-2) + rds.unsqueeze(-2) * z.unsqueeze(-1),
                                 rds.unsqueeze(-2).expand(1, r, k, 3).reshape(1, -1, 3)), a.iters)
            report(f"torch eager equivalent (renderers.py:171-174) K={k}", ms, r, 28 * k + 24)
            gp = torch.randn(1, r, k, 3, device=dev, generator=g)
            dz = torch.empty(1, r, k, device=dev)

            def bwd():
                rc = lib.avr_ray_points_bwd(rds.data_ptr(), gp.data_ptr(), r, k, dz.data_ptr(), sp)
                assert rc == 0

            report(f"ray_points_bwd (d_z) K={k}", timeit(bwd, a.iters), r, 16 * k + 12)
            del pts, vd, gp, dz
        x_pix = torch.rand(1, r, 2, device=dev, generator=g)
        c2w = torch.eye(4, device=dev).repeat(1, r, 1, 1) + 0.01 * torch.randn(1, r, 4, 4, device=dev, generator=g)
        kin = torch.tensor([[[1.025, 0.0, 0.5], [0.0, 1.025, 0.5], [0.0, 0.0, 1.0]]], device=dev)
        report("world_rays (get_world_rays)", timeit(lambda: ops.world_rays(x_pix, kin, c2w), a.iters), r, 8 + 64 + 24)
        dist_t = torch.rand(1, r, device=dev, generator=g)
        report("depth_from_world", timeit(lambda: ops.depth_from_world(ros, rds, dist_t, c2w), a.iters), r, 24 + 4 + 64 + 4)
        from avr_b200 import geometry as _g  # noqa: F401
        pts1 = ros + rds * dist_t.unsqueeze(-1)
        h = torch.cat((pts1, torch.ones_like(pts1[..., :1])), -1)
        ms = timeit(lambda: -torch.einsum("...ij,...j->...i", torch.inverse(c2w), h)[..., 2], a.iters)
        report("torch eager depth_from_world (utils.py:358-361)", ms, r, 24 + 4 + 64 + 4)
    if "adaptive" in a.what:
        # AdaptiveVolumeRenderer's tail (renderers.py:489-509): K = 20 samples in [d-eps, d+eps], depths carry grad
        import avr_b200
        lib = avr_b200.load_library()
        sp = torch.cuda.current_stream().cuda_stream
        k = 20
        d = 0.9 + 0.8 * torch.rand(1, r, device=dev, generator=g)
        nr, fr = (d - 0.15).reshape(-1).contiguous(), (d + 0.15).reshape(-1).contiguous()
        u = torch.rand(1, r, k, device=dev, generator=g)
        report(f"coarse_sample per-ray bounds K={k}", timeit(lambda: ops.coarse_sample_raw(nr, fr, 1, u), a.iters), r, 8 * k + 8)
        z = ops.coarse_sample_raw(nr, fr, 1, u)
        gz = torch.randn(1, r, k, device=dev, generator=g)
        dn, df = torch.empty(r, device=dev), torch.empty(r, device=dev)

        def cbwd():
            assert lib.avr_coarse_sample_bwd(gz.data_ptr(), u.data_ptr(), r, k, dn.data_ptr(), df.data_ptr(), sp) == 0

        report(f"coarse_sample_bwd (d_near, d_far) K={k}", timeit(cbwd, a.iters), r, 8 * k + 8)
        zs, perm = torch.empty_like(z), torch.empty(1, r, k, dtype=torch.int32, device=dev)

        def srt():
            assert lib.avr_sort_rays(z.data_ptr(), r, k, zs.data_ptr(), perm.data_ptr(), sp) == 0

        report(f"sort_rays (+perm) K={k}", timeit(srt, a.iters), r, 12 * k)
        x = torch.cat([torch.sigmoid(torch.randn(1, r, k, 3, device=dev, generator=g)),
                       torch.relu(torch.randn(1, r, k, 1, device=dev, generator=g)) * 30], -1).contiguous()
        report(f"composite_fwd K={k} (no weights)", timeit(lambda: ops.composite_fwd_raw(x, z, True, 1.8, False), a.iters), r, 20 * k + 16)
        g1, g2 = torch.randn(1, r, 3, device=dev, generator=g), torch.randn(1, r, device=dev, generator=g)
        report(f"composite_bwd K={k} with d_z", timeit(lambda: ops.composite_bwd_raw(x, z, g1, g2, None, True, 1.8, True), a.iters),
               r, 40 * k + 16)
        report(f"composite_bwd K={k} without d_z", timeit(lambda: ops.composite_bwd_raw(x, z, g1, g2, None, True, 1.8, False), a.iters),
               r, 36 * k + 16)
        del x, z, zs, perm, gz
    if "packed" in a.what:
        rp = min(r, 1 << 20)
        counts = torch.randint(a.min_count, 257, (rp,), device=dev, generator=g)
        offsets = torch.zeros(rp + 1, dtype=torch.int64, device=dev)
        offsets[1:] = torch.cumsum(counts, 0)
        s = int(offsets[-1])
        d = 0.9 + 0.8 * torch.rand(rp, device=dev, generator=g)
        nr, fr = d - 0.15, d + 0.15
        u = torch.rand(s, device=dev, generator=g)
        ms = timeit(lambda: ops.coarse_sample_packed(nr, fr, u, offsets), a.iters)
        report("coarse_sample_packed 8..256", ms, rp, 8 * s / rp + 16, samples=s)
        z = ops.coarse_sample_packed(nr, fr, u, offsets)
        x = torch.cat([torch.sigmoid(torch.randn(s, 3, device=dev, generator=g)),
                       torch.relu(torch.randn(s, 1, device=dev, generator=g)) * 30], -1)
        ms = timeit(lambda: ops.composite_packed_fwd_raw(x, z, offsets, True, 1.8, True), a.iters)
        report("composite_fwd_packed 8..256", ms, rp, 24 * s / rp + 24, samples=s)
        import avr_b200
        lib = avr_b200.load_library()
        rgb, depth, w = ops.composite_packed_fwd_raw(x, z, offsets, True, 1.8, True)
        g1, g2 = torch.randn_like(rgb), torch.randn_like(depth)
        dx = torch.empty_like(x)
        sp = torch.cuda.current_stream().cuda_stream

        def bwd():
            rc = lib.avr_composite_bwd_packed(x.data_ptr(), z.data_ptr(), offsets.data_ptr(), g1.data_ptr(), g2.data_ptr(),
                                              None, rp, s, 1, 1.8, dx.data_ptr(), None, sp)
            assert rc == 0

        ms = timeit(bwd, a.iters)
        report("composite_bwd_packed 8..256", ms, rp, 36 * s / rp + 24, samples=s)
        fo = torch.zeros(rp + 1, dtype=torch.int64, device=dev)
        fo[1:] = torch.cumsum(counts // 2, 0)
        sf = int(fo[-1])
        uf, uf2 = torch.rand(sf, device=dev, generator=g), torch.rand(sf, device=dev, generator=g)
        ms = timeit(lambda: ops.importance_sample_packed(w, z, nr, fr, uf, uf2, offsets, fo, 256, 128), a.iters)
        report("importance_sample_packed Kc 8..256 -> Kc/2", ms, rp, (12 * s + 12 * sf) / rp + 24, samples=s, fine=sf)


if __name__ == "__main__":
    main()
