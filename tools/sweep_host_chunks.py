import ctypes, sys, time, torch
sys.path.insert(0, '/root/repo')
import avr_b200
lib = avr_b200.load_library()
rays, k = 1 << 20, 96
g = torch.Generator().manual_seed(0)
hx = torch.rand(rays, k, 4, generator=g).pin_memory(); hz = torch.sort(0.8 + torch.rand(rays, k, generator=g), -1).values.pin_memory()
hg = torch.randn(rays, 3, generator=g).pin_memory(); hd = torch.randn(rays, generator=g).pin_memory()
o_rgb = torch.empty(rays, 3).pin_memory(); o_depth = torch.empty(rays).pin_memory(); o_dx = torch.empty(rays, k, 4).pin_memory()
for chunk in (4096, 8192, 16384, 32768, 65536, 131072, 262144):
    ws = ctypes.c_void_p()
    assert lib.avr_host_workspace_create(k, chunk, ctypes.byref(ws)) == 0
    def step():
        rc = lib.avr_composite_fwd_bwd_host(ws, hx.data_ptr(), hz.data_ptr(), hg.data_ptr(), hd.data_ptr(), rays, k, 1, 1.8,
                                            o_rgb.data_ptr(), o_depth.data_ptr(), None, o_dx.data_ptr())
        assert rc == 0
    step(); step()
    t0 = time.perf_counter()
    for _ in range(5): step()
    dt = (time.perf_counter() - t0) / 5
    print(chunk, round(dt * 1e3, 2), "ms", round(rays / dt / 1e6, 2), "Mrays/s", flush=True)
    lib.avr_host_workspace_destroy(ws)
