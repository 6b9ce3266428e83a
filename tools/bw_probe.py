"""Write-only / read-only / copy bandwidth of the box (torch kernels, CUDA events): what a write-heavy kernel can expect
against the copy figure in MEASURED_PEAKS.json.   python tools/bw_probe.py"""
import json
import torch

def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

n = 1 << 29  # 2 GiB of fp32
a = torch.empty(n, device="cuda")
b = torch.empty(n, device="cuda")
a.normal_()
res = {}
ms = timeit(lambda: b.copy_(a)); res["copy_GBps"] = 8 * n / ms / 1e6
ms = timeit(lambda: b.fill_(1.0)); res["fill_GBps"] = 4 * n / ms / 1e6
ms = timeit(lambda: b.zero_()); res["memset_GBps"] = 4 * n / ms / 1e6
ms = timeit(lambda: torch.sum(a)); res["sum_GBps"] = 4 * n / ms / 1e6
ms = timeit(lambda: torch.mul(a[: n // 8], 2.0, out=b[: n // 8].view(-1))); res["scale_small_GBps"] = 8 * (n // 8) / ms / 1e6
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
