"""The host-buffer entry point bench.py times for the end-to-end figure."""
import ctypes

import pytest
import torch

import avr_oracle as O
from conftest import assert_close

pytestmark = pytest.mark.gpu


def test_fwd_bwd_host_matches_device_path(dev):
    import avr_b200
    from avr_b200 import ops
    lib = avr_b200.load_library()
    r, k = 20000, 96
    g = torch.Generator().manual_seed(0)
    z = torch.sort(0.8 + torch.rand(r, k, generator=g), -1).values.pin_memory()
    x = torch.cat([torch.sigmoid(torch.randn(r, k, 3, generator=g)), torch.relu(torch.randn(r, k, 1, generator=g)) * 30], -1).pin_memory()
    g_rgb = torch.randn(r, 3, generator=g).pin_memory()
    g_d = torch.randn(r, generator=g).pin_memory()
    rgb = torch.empty(r, 3).pin_memory()
    depth = torch.empty(r).pin_memory()
    dx = torch.empty(r, k, 4).pin_memory()
    w = torch.empty(r, k).pin_memory()
    for chunk in (0, 4096, 7000):
        rgb.zero_(), depth.zero_(), dx.zero_()
        ws = ctypes.c_void_p()
        assert lib.avr_host_workspace_create(k, chunk, ctypes.byref(ws)) == 0, lib.avr_last_cuda_error()
        for _ in range(2):      # the workspace is reusable
            rc = lib.avr_composite_fwd_bwd_host(ws, x.data_ptr(), z.data_ptr(), g_rgb.data_ptr(), g_d.data_ptr(), r, k, 1, 1.8,
                                                rgb.data_ptr(), depth.data_ptr(), w.data_ptr() if chunk else None, dx.data_ptr())
            assert rc == 0, lib.avr_last_cuda_error()
        assert lib.avr_composite_fwd_bwd_host(ws, x.data_ptr(), z.data_ptr(), None, None, r, k + 1, 1, 1.8,
                                              rgb.data_ptr(), depth.data_ptr(), None, dx.data_ptr()) == -1
        assert lib.avr_host_workspace_destroy(ws) == 0
        xd = x.to(dev).requires_grad_(True)
        a, b, wd = ops.composite(xd, z.to(dev), True, 1.8, want_w=True)
        torch.autograd.backward([a, b], [g_rgb.to(dev), g_d.to(dev)])
        if chunk == 0:      # one chunk: the very same launches as the device path
            assert torch.equal(rgb, a.cpu()) and torch.equal(depth, b.cpu()) and torch.equal(dx, xd.grad.cpu())
        # chunking moves rays to other lanes of a warp tile, which regroups the scan: equal to
        # rounding, not to the bit
        assert_close(rgb, a, rtol=2e-6, atol=2e-7, what="rgb")
        assert_close(depth, b, rtol=2e-6, atol=2e-7, what="depth")
        if chunk:
            assert_close(w, wd, rtol=2e-6, atol=2e-7, what="weights")
        assert_close(dx[..., :3], xd.grad[..., :3], rtol=2e-6, atol=2e-7, what="d_rgb")
        assert_close(dx[..., -1, 3] / 1e10, xd.grad[..., -1, 3] / 1e10, rtol=1e-5, atol=1e-6, what="d_sigma[-1]/1e10")
        assert_close(dx[..., :-1, 3], xd.grad[..., :-1, 3], rtol=1e-5, atol=1e-6, what="d_sigma[:-1]")
    want = O.composite_rgbs(z[:512].unsqueeze(0), x[:512].unsqueeze(0), True)
    assert_close(rgb[:512], want[0][0], what="rgb vs oracle")
