"""BASELINE.json config 1 as a throughput number, through the drop-in API.

    python tools/bench_dropin.py [--iters 20]

Times ``avr_b200.VolumeRenderer.forward`` + ``backward`` (64 coarse + 32 fine samples, 16 of them
"depth" samples: conf/default.conf's normal_renderer) the way the reference drives it
(models.py:919-929 from train.py:108 / test.py:54) at the two sizes the reference uses:

* a full 128x128 frame, SB=1 x 16384 rays  (train.py:150, test.py:54)
* a training batch, SB=4 x 512 rays        (train.py:78-84, --ray_batch_size 512, train.py:202)

around a SMALL synthetic radiance field (so the renderer, not the MLP, is what is measured),
next to the reference's own op sequence (oracle/avr_oracle.py::render_volume = renderers.py:133-277
restated) running eagerly on the SAME GPU.  This is the launch-bound regime: what counts
is the number of launches and the host time per launch, so wall-clock per step (with a device
synchronise at the end of the batch of steps) is reported next to the device time.

Arms:  drop-in eager | drop-in replayed from a CUDA graph (whole fwd+bwd step captured once) |
       eager-GPU reference | the field alone (its two calls, forward + backward, same point counts).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402
from torch import nn  # noqa: E402


class SyntheticField(nn.Module):
    """Callback contract of SURVEY 3.5: (xyz (SB,N,3), viewdirs (SB,N,3), coarse) -> (SB,N,4) with
    sigmoid colours and a ReLU density (models.py:854-863).  Two small Linear layers: enough for
    gradients to reach parameters, small enough that the renderer dominates."""

    def __init__(self, hidden=32):
        super().__init__()
        self.l1 = nn.Linear(6, hidden)
        self.l2 = nn.Linear(hidden, 4)

    def forward(self, xyz, viewdirs=None, coarse=True, return_features=False):
        h = torch.sin(self.l1(torch.cat([xyz * 3.0, viewdirs], -1)))
        o = self.l2(h)
        return torch.cat([torch.sigmoid(o[..., :3]), torch.relu(o[..., 3:4] * 20.0)], -1)


class WideField(SyntheticField):
    """Same contract, an MLP of PixelNeRF's width (five 512-wide layers: ~2 MFLOP per point against
    ResnetFC's ~6, models.py:473-606) — the regime the renderer actually runs in."""

    def __init__(self, hidden=512, layers=4):
        super().__init__(hidden)
        self.body = nn.Sequential(*[m for _ in range(layers) for m in (nn.Linear(hidden, hidden), nn.ReLU())])

    def forward(self, xyz, viewdirs=None, coarse=True, return_features=False):
        h = self.body(torch.sin(self.l1(torch.cat([xyz * 3.0, viewdirs], -1))))
        o = self.l2(h)
        return torch.cat([torch.sigmoid(o[..., :3]), torch.relu(o[..., 3:4] * 20.0)], -1)


def camera(sb, r, dev):
    import math
    side = int(round(math.sqrt(r)))
    if side * side == r:        # utils.get_opencv_pixel_coordinates (utils.py:339-356), normalised
        ii, jj = torch.meshgrid(torch.arange(side), torch.arange(side), indexing="ij")
        x_pix = torch.stack([jj.float() / side, ii.float() / side], -1).reshape(1, r, 2).expand(sb, r, 2).contiguous()
    else:
        x_pix = torch.rand(sb, r, 2, generator=torch.Generator().manual_seed(0))
    f = 131.25 / 128.0
    intrinsics = torch.tensor([[f, 0.0, 0.5], [0.0, f, 0.5], [0.0, 0.0, 1.0]]).expand(sb, 3, 3).contiguous()
    c2w = torch.eye(4)
    c2w[:3, 3] = torch.tensor([0.0, 0.0, 1.3])
    c2w = c2w @ torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0]))
    cam2world = c2w.expand(sb, r, 4, 4).contiguous()
    return cam2world.to(dev), intrinsics.to(dev), x_pix.to(dev)


def _time(step, iters, dev, warm=5):
    for _ in range(warm):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    wall = (time.perf_counter() - t0) / iters * 1e3
    return {"wall_ms": round(wall, 4), "device_ms": round(e0.elapsed_time(e1) / iters, 4)}


def boundary_overhead(dev, calls=2000):
    """Host cost of ONE call through the boundary on a batch too small to matter (96 rays x 64):
    the ctypes call alone (outputs preallocated), the Python operator around it (argument checks,
    three torch.empty, stream lookup), and a stock ATen op on the same tensor as the yardstick.
    BASELINE.json's north_star names a torch C++ extension for this layer; the repo binds the C ABI
    with ctypes instead — these numbers are what that choice costs."""
    import avr_b200
    from avr_b200 import ops
    lib = avr_b200.load_library()
    r, k = 96, 64
    x = torch.rand(1, r, k, 4, device=dev)
    z = torch.sort(0.8 + torch.rand(1, r, k, device=dev), -1).values
    w, rgb, depth = torch.empty(1, r, k, device=dev), torch.empty(1, r, 3, device=dev), torch.empty(1, r, device=dev)
    sp = torch.cuda.current_stream(dev).cuda_stream

    def raw():
        lib.avr_composite_fwd(x.data_ptr(), z.data_ptr(), r, k, 1, 1.8, w.data_ptr(), rgb.data_ptr(), depth.data_ptr(), sp)

    def per_call(fn):
        for _ in range(50):
            fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(calls):
            fn()
        dt = time.perf_counter() - t0          # host time to ENQUEUE (the queue never fills at this size)
        torch.cuda.synchronize(dev)
        return round(dt / calls * 1e6, 2)

    return {"ctypes_call_us": per_call(raw),
            "python_operator_us": per_call(lambda: ops.composite_fwd_raw(x, z, True, 1.8, True)),
            "autograd_function_us": per_call(lambda: ops.composite(x, z, True, 1.8, True)),
            "aten_cumprod_us": per_call(lambda: torch.cumprod(z, -1)),
            "aten_sort_us": per_call(lambda: torch.sort(z, -1))}


def run(dev, iters=20, sizes=((1, 16384, "128x128 frame, SB=1"), (4, 512, "train.py batch, SB=4 x 512 rays"))):
    import avr_b200
    import avr_oracle as O
    from avr_b200 import _lib

    kc, kf, kd = 64, 32, 16
    out = {}
    for sb, r, label in sizes:
        c2w, intr, x_pix = camera(sb, r, dev)
        torch.manual_seed(0)
        field = SyntheticField().to(dev)
        renderer = avr_b200.VolumeRenderer(0.8, 1.8, kc, kf, kd, 0.01, True)
        params = list(field.parameters())

        def loss_of(o):
            return o[0].mean() + o[1].mean() + o[2].mean()      # rgb_coarse, rgb_fine, depth (utils.py:364-377)

        def dropin_step():
            for p in params:
                p.grad = None
            loss_of(renderer(c2w, intr, x_pix, field)).backward()

        def reference_step():
            for p in params:
                p.grad = None
            draws = O.draw_volume_randoms(sb, r, kc, kf, kd, device=dev)
            loss_of(O.render_volume(c2w, intr, x_pix, field, 0.8, 1.8, kc, kf, kd, 0.01, True, draws)).backward()

        def field_step():
            for p in params:
                p.grad = None
            tot = 0.0
            for k in (kc, kc + kf):
                pts = torch.empty(sb, r * k, 3, device=dev).uniform_(-1, 1)
                tot = tot + field(pts, viewdirs=pts, coarse=True).mean()
            tot.backward()

        rec = {"rays": sb * r, "samples": f"{kc}+{kf} ({kd} depth)"}
        _lib.dispatch_reset()
        rec["dropin_eager"] = _time(dropin_step, iters, dev)
        rec["dropin_dispatch"] = {k: v for k, v in _lib.dispatch_counters().items() if v}
        rec["reference_eager_gpu"] = _time(reference_step, iters, dev)
        rec["field_only"] = _time(field_step, iters, dev)
        # whole step from a CUDA graph: every kernel of the path is capturable (no sync, no allocation
        # outside torch's capture-aware allocator); the draws are made inside the graph
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    dropin_step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            for p in params:
                p.grad = None
            with torch.cuda.graph(graph):
                loss_of(renderer(c2w, intr, x_pix, field)).backward()
            rec["dropin_cuda_graph"] = _time(graph.replay, iters, dev)
            del graph
        except Exception as exc:  # the eager numbers must survive a capture problem
            rec["dropin_cuda_graph"] = {"error": f"{type(exc).__name__}: {exc}"}
        e, ref, fo = rec["dropin_eager"], rec["reference_eager_gpu"], rec["field_only"]
        rec["speedup_vs_reference_wall"] = round(ref["wall_ms"] / e["wall_ms"], 2)
        # what the renderer itself costs (kernels + ctypes + torch.empty + Python autograd), per step
        rec["renderer_ms"] = round(max(e["wall_ms"] - fo["wall_ms"], 0.0), 4)
        rec["reference_renderer_ms"] = round(max(ref["wall_ms"] - fo["wall_ms"], 0.0), 4)
        rec["renderer_share_of_step"] = round(rec["renderer_ms"] / e["wall_ms"], 3)
        rec["rays_per_s_wall"] = sb * r / (e["wall_ms"] * 1e-3)
        if sb * r <= 4096:
            # the same step around an MLP of PixelNeRF's width: the share of the boundary in a real step
            wide = WideField().to(dev)
            wparams = list(wide.parameters())

            def wide_step(render):
                for p in wparams:
                    p.grad = None
                if render:
                    loss_of(renderer(c2w, intr, x_pix, wide)).backward()
                else:
                    tot = 0.0
                    for k in (kc, kc + kf):
                        pts = torch.empty(sb, r * k, 3, device=dev).uniform_(-1, 1)
                        tot = tot + wide(pts, viewdirs=pts, coarse=True).mean()
                    tot.backward()

            a = _time(lambda: wide_step(True), max(iters // 2, 3), dev, warm=2)
            b = _time(lambda: wide_step(False), max(iters // 2, 3), dev, warm=2)
            rec["with_512_wide_mlp"] = {"dropin_eager": a, "field_only": b,
                                        "renderer_share_of_step": round(max(a["wall_ms"] - b["wall_ms"], 0.0) / a["wall_ms"], 4)}
            del wide
        out[label] = rec
    out["boundary_call_overhead"] = boundary_overhead(dev)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    print(json.dumps(run(torch.device("cuda:0"), a.iters)))
