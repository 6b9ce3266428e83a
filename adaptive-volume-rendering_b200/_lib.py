"""ctypes binding of libavr_b200.so — the only way the host side reaches the kernels.

There is deliberately no fallback: if the shared library is missing, cannot be
loaded, or the current device is not an sm_100 part, every op raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libavr_b200.so")

AVR_OK = 0
ABI_VERSION = 2

_P = c_void_p  # every device pointer crosses as a plain address

# name -> (restype, argtypes); mirrors include/avr_b200.h one to one
PROTOTYPES = {
    "avr_abi_version": (c_int, []),
    "avr_status_string": (c_char_p, [c_int]),
    "avr_last_cuda_error": (c_char_p, []),
    "avr_device_check": (c_int, []),
    "avr_composite_plan": (c_int, [c_int64, c_int, _P, _P]),
    "avr_composite_plan_info": (c_int, [c_int64, c_int, _P, _P, ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                                        ctypes.POINTER(c_int64)]),
    "avr_set_force_generic": (None, [c_int]),
    "avr_set_option": (c_int, [c_char_p, c_int, c_int]),
    "avr_dispatch_counters": (c_int, [ctypes.POINTER(c_int64), c_int]),
    "avr_dispatch_reset": (None, []),
    "avr_coarse_sample_fwd": (c_int, [_P, _P, c_int, _P, c_int64, c_int, _P, _P]),
    "avr_coarse_sample_bwd": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P]),
    "avr_importance_sample": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, c_int,
                                      c_float, _P, _P, _P, _P, _P]),
    "avr_sort_rays": (c_int, [_P, c_int64, c_int, _P, _P, _P]),
    "avr_sort_rays_bwd": (c_int, [_P, _P, c_int64, c_int, _P, _P]),
    "avr_composite_fwd": (c_int, [_P, _P, c_int64, c_int, c_int, c_float, _P, _P, _P, _P]),
    "avr_composite_fwd_gather": (c_int, [_P, _P, c_int64, c_int, c_int, c_float, _P, _P, _P,
                                         ctypes.POINTER(c_void_p), c_int, c_int64, _P]),
    "avr_composite_fwd_gather_multicast": (c_int, [_P, _P, c_int64, c_int, c_int, c_float, _P, _P, _P, _P, c_int64, _P]),
    "avr_composite_fwd_gather_signal": (c_int, [_P, _P, c_int64, c_int, c_int, c_float, _P, _P, _P,
                                                ctypes.POINTER(c_void_p), c_int, c_int, c_int64,
                                                ctypes.POINTER(c_void_p), c_int, c_int, ctypes.c_uint32, _P, _P]),
    "avr_gather_wait": (c_int, [_P, c_int, ctypes.c_uint32, _P, _P]),
    "avr_gather_push_rows": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int64, c_int64, _P]),
    "avr_composite_bwd": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_float, _P, _P, _P]),
    "avr_composite_fwd_packed": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, c_float, _P, _P, _P, _P]),
    "avr_composite_bwd_packed": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_float, _P, _P, _P]),
    "avr_coarse_sample_fwd_packed": (c_int, [_P, _P, c_int, _P, _P, c_int64, c_int64, _P, _P]),
    "avr_importance_sample_packed": (c_int, [_P, _P, _P, _P, _P, _P, c_int, _P, _P, c_int64, c_int, c_int,
                                             _P, _P, _P, _P, _P]),
    "avr_ray_points_fwd": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P, _P]),
    "avr_ray_points_bwd": (c_int, [_P, _P, c_int64, c_int, _P, _P]),
    "avr_ray_points_fwd_packed": (c_int, [_P, _P, _P, _P, c_int64, c_int64, _P, _P, _P]),
    "avr_ray_points_bwd_packed": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P]),
    "avr_coarse_sample_points_fwd": (c_int, [_P, _P, c_int, _P, _P, _P, c_int64, c_int, _P, _P, _P, _P]),
    "avr_world_rays": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, _P, _P]),
    "avr_rays_coarse_sample_points_fwd": (c_int, [_P, _P, _P, c_int64, _P, _P, c_int, _P, c_int64, c_int, _P, _P, _P, _P, _P,
                                                  _P, _P]),
    "avr_composite_fwd_camera": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_float, _P, _P, _P, _P]),
    "avr_composite_bwd_camera": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_float, _P, _P, _P]),
    "avr_depth_from_world": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P]),
    "avr_field_inputs_fwd": (c_int, [_P, _P]),
    "avr_field_inputs_bwd": (c_int, [_P, _P]),
    "avr_lstm_march_fwd": (c_int, [_P, _P, _P]),
    "avr_lstm_march_bwd": (c_int, [_P, _P, _P]),
    "avr_host_workspace_create": (c_int, [c_int, c_int64, ctypes.POINTER(c_void_p)]),
    "avr_host_workspace_destroy": (c_int, [_P]),
    "avr_composite_fwd_bwd_host": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_float, _P, _P, _P, _P]),
}

_lib = None


class AvrError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the library (once).  Raises AvrError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AvrError(
            f"{LIB_PATH} is missing: build it with `python adaptive-volume-rendering_b200/build.py` "
            "(or __graft_entry__.build()). There is no CPU or eager fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.avr_abi_version() != ABI_VERSION:
        raise AvrError(f"ABI mismatch: library {lib.avr_abi_version()}, binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != AVR_OK:
        lib = load()
        msg = lib.avr_status_string(status).decode()
        detail = lib.avr_last_cuda_error().decode()
        raise AvrError(f"{what}: {msg} ({status})" + (f" [{detail}]" if detail and status in (-2, -3, -5) else ""))


DISPATCH_NAMES = ("fwd_span", "fwd_wray", "fwd_generic", "fwd_span_packed", "bwd_span", "bwd_wray", "bwd_generic",
                  "bwd_span_packed", "importance_bins", "importance_grp", "importance_reg", "importance_smem")


def dispatch_counters() -> dict:
    """Calls served per kernel family since the last ``dispatch_reset`` (process-wide)."""
    buf = (c_int64 * len(DISPATCH_NAMES))()
    load().avr_dispatch_counters(buf, len(DISPATCH_NAMES))
    return dict(zip(DISPATCH_NAMES, (int(v) for v in buf)))


def dispatch_reset() -> None:
    load().avr_dispatch_reset()


def set_option(name: str, value: int | None) -> None:
    """Override one of the library's A/B switches (None: back to the default)."""
    check(load().avr_set_option(name.encode(), 0 if value is None else int(value), 1 if value is None else 0),
          f"avr_set_option({name})")


def ptr(t) -> int | None:
    """Device (or host) address of a tensor, None for None."""
    return None if t is None else t.data_ptr()


def current_stream_ptr(device) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise AvrError(
                "avr_b200 ops run on CUDA tensors only (sm_100a kernels; there is no CPU fallback) — "
                f"got a tensor on {t.device}"
            )
