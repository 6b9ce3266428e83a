"""CPU-side checks of the boundary: the library loads, exports exactly what the header
declares, rejects bad arguments, and the host side refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    import avr_b200
    return avr_b200.load_library()


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "avr_b200.h")).read()
    return sorted(set(re.findall(r"^AVR_API [\w\* ]+?\b(avr_\w+)\(", text, flags=re.M)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from avr_b200 import _lib
    declared = _header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/avr_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == declared
    # parameter counts of the ctypes table follow the header
    text = open(os.path.join(ROOT, "include", "avr_b200.h")).read()
    for name in declared:
        m = re.search(r"\b%s\(([^;]*?)\);" % name, text, flags=re.S)
        args = m.group(1).strip()
        n = 0 if args == "void" else len(args.split(","))
        assert n == len(_lib.PROTOTYPES[name][1]), name


def test_abi_version_and_strings(lib):
    assert lib.avr_abi_version() == 2
    assert lib.avr_status_string(0) == b"ok" and lib.avr_status_string(-1) == b"bad argument"


def test_argument_checks_need_no_gpu(lib):
    # bad shapes / null pointers are refused before anything touches a device
    assert lib.avr_composite_fwd(None, None, 10, 96, 1, 1.8, None, None, None, None) == -1
    assert lib.avr_composite_fwd(None, None, -1, 96, 1, 1.8, None, None, None, None) == -1
    assert lib.avr_composite_fwd(None, None, 0, 96, 1, 1.8, None, None, None, None) == 0      # empty input
    assert lib.avr_composite_bwd(None, None, None, None, None, 0, 96, 1, 1.8, None, None, None) == 0
    assert lib.avr_coarse_sample_fwd(None, None, 2, None, 4, 8, None, None) == -1
    assert lib.avr_importance_sample(None, None, None, None, None, None, None, 0, 0, 64, 16, 0, 0.0, None, None, None, None, None) == 0
    assert lib.avr_sort_rays(None, 4, 0, None, None, None) == -1
    buf = (ctypes.c_float * 64)()
    misaligned = ctypes.addressof(buf) + 4
    assert lib.avr_composite_fwd(misaligned, ctypes.addressof(buf), 1, 4, 1, 1.8, None, ctypes.addressof(buf), ctypes.addressof(buf), None) == -1


def test_options_and_dispatch_counters_are_host_state(lib):
    from avr_b200 import _lib
    a = 1 << 20
    assert lib.avr_set_option(b"AVR_NO_SUCH_SWITCH", 1, 0) == -1
    L = ctypes.c_int()
    assert lib.avr_composite_plan_info(1 << 20, 96, a, a, ctypes.byref(L), None, None) == 1 and L.value == 9
    _lib.set_option("AVR_SPAN_L", 13)            # what AVR_SPAN_L=13 in the environment would do, without getenv per launch
    assert lib.avr_composite_plan_info(1 << 20, 96, a, a, ctypes.byref(L), None, None) == 1 and L.value == 13
    _lib.set_option("AVR_SPAN_L", None)
    assert lib.avr_composite_plan_info(1 << 20, 96, a, a, ctypes.byref(L), None, None) == 1 and L.value == 9
    _lib.dispatch_reset()
    c = _lib.dispatch_counters()
    assert set(c) == set(_lib.DISPATCH_NAMES) and not any(c.values())
    # bad arguments of the signalled gather and of the wait are refused before any launch
    assert lib.avr_gather_wait(None, 2, 1, None, None) == -1
    assert lib.avr_gather_wait(a, 0, 1, None, None) == -1
    assert lib.avr_composite_fwd_gather_signal(a, a, 96, 96, 1, 1.8, None, a, a, None, 2, 0, 0, None, 2, 0, 1, None, None) == -1


def test_span_planner_is_host_logic(lib):
    a = 1 << 20   # any 16-byte aligned address; the planner only looks at alignment
    for k in (96, 64, 192, 20, 128, 8, 33):
        assert lib.avr_composite_plan(1 << 20, k, a, a) == 1, k
    assert lib.avr_composite_plan(1 << 20, 417, a, a) == 0       # longer than any warp tile
    assert lib.avr_composite_plan(1 << 20, 96, a + 4, a) == 0
    assert lib.avr_composite_plan(2, 96, a, a) == 0               # fewer rays than one tile
    lib.avr_set_force_generic(1)
    assert lib.avr_composite_plan(1 << 20, 96, a, a) == 0
    lib.avr_set_force_generic(0)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    import avr_b200
    assert lib.avr_device_check() == -3
    z = torch.rand(1, 4, 8).sort(-1).values
    x = torch.rand(1, 4, 8, 4)
    with pytest.raises(avr_b200.AvrError):
        avr_b200.volume_integral(z, x[..., 3:4], x[..., :3])
    with pytest.raises(avr_b200.AvrError):
        avr_b200.sample_coarse(torch.tensor([0.8]).expand(1, 4), torch.tensor([1.8]).expand(1, 4), 8)
    ren = avr_b200.VolumeRenderer(0.8, 1.8, 8, 4, 2, 0.01)
    with pytest.raises(avr_b200.AvrError):
        ren(torch.eye(4).expand(1, 4, 4, 4), torch.eye(3).expand(1, 3, 3), torch.rand(1, 4, 2), lambda *a, **k: None)


def test_host_side_view_detection():
    from avr_b200.renderers import _as_rgbs
    out = torch.rand(2, 5 * 8, 4)
    sig = out[..., 3].view(2, 5, 8, 1)
    rad = out[..., :3].view(2, 5, 8, 3)
    packed = _as_rgbs(sig, rad)
    assert packed.data_ptr() == out.data_ptr() and packed.shape == (2, 5, 8, 4)      # zero-copy
    sep = _as_rgbs(sig.clone(), rad.clone())
    assert sep.data_ptr() != out.data_ptr() and torch.equal(sep, out.view(2, 5, 8, 4))


def test_bounds_classification():
    from avr_b200.ops import _bounds
    n = torch.tensor([0.8]).expand(2, 6)
    f = torch.tensor([1.8]).expand(2, 6)
    a, b, s = _bounds(n, f, 12)
    assert s == 0 and a.numel() == 1 and float(a) == pytest.approx(0.8)
    d = torch.rand(2, 6)
    a, b, s = _bounds(d - 0.1, d + 0.1, 12)
    assert s == 1 and a.numel() == 12 and a.is_contiguous()
    a, b, s = _bounds(n, d, 12)
    assert s == 1 and a.numel() == 12 and torch.all(a == 0.8)
