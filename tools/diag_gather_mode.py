"""Where the fused all-gather's extra forward time comes from: the forward span kernel on ONE GPU, plain
vs gather mode with only a LOCAL target buffer (no NVLink traffic at all).

    python tools/diag_gather_mode.py
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import avr_b200  # noqa: E402

lib = avr_b200.load_library()
dev = torch.device("cuda:0")
rays, k = 1 << 20, 96
g = torch.Generator(device=dev).manual_seed(0)
z = torch.sort(0.8 + torch.rand(rays, k, device=dev, generator=g), -1).values
x = torch.rand(rays, k, 4, device=dev, generator=g)
w, rgb, depth = torch.empty(rays, k, device=dev), torch.empty(rays, 3, device=dev), torch.empty(rays, device=dev)
gathered = torch.empty(rays, 4, device=dev)
target = (ctypes.c_void_p * 1)(gathered.data_ptr())
sp = torch.cuda.current_stream(dev).cuda_stream


def plain():
    assert lib.avr_composite_fwd(x.data_ptr(), z.data_ptr(), rays, k, 1, 1.8, w.data_ptr(), rgb.data_ptr(), depth.data_ptr(), sp) == 0


def local_gather():
    assert lib.avr_composite_fwd_gather(x.data_ptr(), z.data_ptr(), rays, k, 1, 1.8, w.data_ptr(), rgb.data_ptr(), depth.data_ptr(),
                                        target, 1, 0, sp) == 0


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, fn in (("plain", plain), ("gather mode, local target only", local_gather), ("plain", plain), ("gather mode, local target only", local_gather)):
    print(json.dumps({"variant": name, "ms": round(timeit(fn), 4)}), flush=True)
local_gather()
torch.cuda.synchronize()
assert torch.equal(gathered[:, :3], rgb) and torch.equal(gathered[:, 3], depth)
