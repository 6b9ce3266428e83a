"""Sweep of the host-buffer pipeline (avr_composite_fwd_bwd_host, the call bench.py times for `e2e`):
slots in flight x chunk size.  PCIe-bound: 2.03 GB up + 2.03 GB down per 2^20 x 96 step.

    python tools/sweep_host_chunks.py
"""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import avr_b200  # noqa: E402

lib = avr_b200.load_library()
rays, k = 1 << 20, 96
g = torch.Generator().manual_seed(0)
hx = torch.rand(rays, k, 4, generator=g).pin_memory()
hz = torch.sort(0.8 + torch.rand(rays, k, generator=g), -1).values.pin_memory()
hg = torch.randn(rays, 3, generator=g).pin_memory()
hd = torch.randn(rays, generator=g).pin_memory()
o_rgb, o_depth = torch.empty(rays, 3).pin_memory(), torch.empty(rays).pin_memory()
o_w, o_dx = torch.empty(rays, k).pin_memory(), torch.empty(rays, k, 4).pin_memory()
for slots in (2, 3, 4, 6, 8):
    for mib in (12, 24, 48, 96):
        lib.avr_set_option(b"AVR_HOST_SLOTS", slots, 0)
        lib.avr_set_option(b"AVR_HOST_CHUNK_MIB", mib, 0)
        ws = ctypes.c_void_p()
        assert lib.avr_host_workspace_create(k, 0, ctypes.byref(ws)) == 0

        def step():
            rc = lib.avr_composite_fwd_bwd_host(ws, hx.data_ptr(), hz.data_ptr(), hg.data_ptr(), hd.data_ptr(), rays, k, 1, 1.8,
                                                o_rgb.data_ptr(), o_depth.data_ptr(), o_w.data_ptr(), o_dx.data_ptr())
            assert rc == 0

        step()
        step()
        t0 = time.perf_counter()
        for _ in range(4):
            step()
        dt = (time.perf_counter() - t0) / 4
        print(json.dumps({"slots": slots, "chunk_mib": mib, "ms": round(dt * 1e3, 2), "Mrays_per_s": round(rays / dt / 1e6, 2),
                          "GBps_each_way": round(rays * (20 * k + 16) / dt / 1e9, 1)}), flush=True)
        lib.avr_host_workspace_destroy(ws)
