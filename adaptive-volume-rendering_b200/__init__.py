"""avr_b200 — B200 (sm_100a) ray-sampling + volume-compositing path of
yankeesong/adaptive-volume-rendering, behind the reference's own renderer API.

    from avr_b200 import VolumeRenderer, AdaptiveVolumeRenderer, volume_integral, ...

The kernels live in ``csrc/`` and are reached only through the C ABI in
``include/avr_b200.h`` (``lib/libavr_b200.so``, loaded with ctypes by ``_lib``).
"""
from ._lib import AvrError, LIB_PATH, load as load_library  # noqa: F401
from .renderers import (  # noqa: F401
    AdaptiveVolumeRenderer,
    Raymarcher,
    VolumeRenderer,
    sample_coarse,
    sample_depth,
    sample_fine,
    volume_integral,
    volume_integral_rgbs,
)
from . import ops, geometry, field, march  # noqa: F401
from .march import lstm_march  # noqa: F401
from .field import FieldConfig, field_inputs, fuse_field_inputs  # noqa: F401
from .dropin import accelerate, convert_renderer  # noqa: F401

__all__ = [
    "AdaptiveVolumeRenderer", "Raymarcher", "VolumeRenderer", "sample_coarse", "sample_depth", "sample_fine",
    "volume_integral", "volume_integral_rgbs", "ops", "geometry", "field", "FieldConfig", "field_inputs",
    "fuse_field_inputs", "lstm_march", "march", "accelerate", "convert_renderer", "AvrError", "load_library", "LIB_PATH",
]
