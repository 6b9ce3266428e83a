"""Tuning sweep for the span kernels (experiments only): times composite fwd / bwd for
combinations of AVR_SPAN_WARPS x AVR_SPAN_STAGES (x AVR_SPAN_L) on synthetic rays.

    python tools/sweep_span.py [--rays N] [--k 96] [--iters 20]
"""
import argparse
import ctypes
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import avr_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1 << 20)
    ap.add_argument("--k", type=int, nargs="+", default=[96])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warps", type=int, nargs="+", default=[2, 3, 4, 6, 8])
    ap.add_argument("--stages", type=int, nargs="+", default=[3, 4])
    ap.add_argument("--ls", type=int, nargs="+", default=[0])
    ap.add_argument("--peak", type=float, default=6456.2)
    ap.add_argument("--no-w", action="store_true", help="forward without the weights output (the fine pass of a render)")
    a = ap.parse_args()
    lib = avr_b200.load_library()
    dev = torch.device("cuda:0")
    for k in a.k:
        rays = a.rays
        g = torch.Generator(device=dev).manual_seed(0)
        z = torch.sort(0.8 + torch.rand(rays, k, device=dev, generator=g), -1).values
        x = torch.cat([torch.sigmoid(torch.randn(rays, k, 3, device=dev, generator=g)),
                       torch.relu(torch.randn(rays, k, 1, device=dev, generator=g)) * 30], -1).contiguous()
        g_rgb = torch.randn(rays, 3, device=dev, generator=g)
        g_d = torch.randn(rays, device=dev, generator=g)
        w, rgb, depth, dx = torch.empty(rays, k, device=dev), torch.empty(rays, 3, device=dev), torch.empty(rays, device=dev), torch.empty_like(x)
        sp = torch.cuda.current_stream().cuda_stream
        ref = None
        for L, warps, stages in itertools.product(a.ls, a.warps, a.stages):
            lib.avr_set_option(b"AVR_SPAN_WARPS", warps, 0)
            lib.avr_set_option(b"AVR_SPAN_STAGES", stages, 0)
            lib.avr_set_option(b"AVR_SPAN_L", L or 0, 0 if L else 1)
            Lc, rpt, mr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
            if not lib.avr_composite_plan_info(rays, k, x.data_ptr(), z.data_ptr(), ctypes.byref(Lc), ctypes.byref(rpt), ctypes.byref(mr)):
                print(json.dumps({"k": k, "L": L, "skip": "no span plan"}))
                continue

            def fwd():
                rc = lib.avr_composite_fwd(x.data_ptr(), z.data_ptr(), rays, k, 1, 1.8, None if a.no_w else w.data_ptr(), rgb.data_ptr(), depth.data_ptr(), sp)
                assert rc == 0, (rc, lib.avr_last_cuda_error())

            def bwd():
                rc = lib.avr_composite_bwd(x.data_ptr(), z.data_ptr(), g_rgb.data_ptr(), g_d.data_ptr(), None, rays, k, 1, 1.8, dx.data_ptr(), None, sp)
                assert rc == 0, (rc, lib.avr_last_cuda_error())

            res = {"k": k, "L": Lc.value, "rays_per_tile": rpt.value, "warps": warps, "stages": stages}
            for name, fn, bpr in (("fwd", fwd, (20 if a.no_w else 24) * k + 16), ("bwd", bwd, 36 * k + 16)):
                try:
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(a.iters):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / a.iters
                    res[name + "_ms"] = round(ms, 4)
                    res[name + "_frac"] = round(bpr * rays / (ms * 1e-3) / 1e9 / a.peak, 4)
                except AssertionError as err:
                    res[name + "_err"] = str(err)
            chk = (rgb[:3000].clone(), w[:3000].clone(), dx[:3000].clone())
            if ref is None:
                ref = chk
            res["same_bits_as_first"] = all(torch.equal(p, q) for p, q in zip(chk, ref)) if Lc.value == ref_L(ref, Lc.value) else None
            print(json.dumps(res), flush=True)


_first_L = []


def ref_L(ref, L):
    if not _first_L:
        _first_L.append(L)
    return _first_L[0]


if __name__ == "__main__":
    main()
