"""One-call switch of a reference model to the kernels.

``train.py:255-276`` builds ``RadFieldAndRenderer(rf, renderer)`` (models.py:912-929) from the
reference's own classes.  ``accelerate(model)`` swaps ``model.renderer`` for the avr_b200 renderer
with the same configuration — for the adaptive renderer the ``lstm`` / ``out_layer`` modules are the
SAME objects, so parameters, optimizer state and checkpoints carry over — and lets ``model.rf`` build
its MLP input with the front-end kernels where its configuration is one they implement
(``field.fuse_field_inputs``).  Nothing else about the model changes.
"""
from __future__ import annotations

from ._lib import AvrError
from .field import fuse_field_inputs
from .renderers import AdaptiveVolumeRenderer, Raymarcher, VolumeRenderer


def convert_renderer(renderer):
    """The avr_b200 renderer equivalent to a reference ``VolumeRenderer`` (renderers.py:121-289) or
    ``AdaptiveVolumeRenderer`` (renderers.py:360-509) instance.  avr_b200 renderers pass through."""
    if isinstance(renderer, (VolumeRenderer, AdaptiveVolumeRenderer, Raymarcher)):
        return renderer
    if all(hasattr(renderer, a) for a in ("lstm", "out_layer", "steps")) and not hasattr(renderer, "epsilon"):
        new = Raymarcher(renderer.n_feature_channels, renderer.steps)       # renderers.py:292-358
        new.lstm, new.out_layer = renderer.lstm, renderer.out_layer          # shared, not copied
        new.train(renderer.training)
        return new
    if all(hasattr(renderer, a) for a in ("lstm", "out_layer", "steps", "epsilon", "n_coarse")):
        new = AdaptiveVolumeRenderer(renderer.n_feature_channels, renderer.steps, renderer.epsilon, renderer.n_coarse,
                                     renderer.white_back)
        new.lstm, new.out_layer = renderer.lstm, renderer.out_layer      # shared, not copied
        new.train(renderer.training)
        return new
    if all(hasattr(renderer, a) for a in ("near", "far", "n_coarse", "n_fine", "n_fine_depth", "depth_std")):
        new = VolumeRenderer(float(renderer.near), float(renderer.far), renderer.n_coarse, renderer.n_fine,
                             renderer.n_fine_depth, renderer.depth_std, getattr(renderer, "white_back", True))
        new.train(renderer.training)
        return new
    raise AvrError(f"convert_renderer: {type(renderer).__name__} is neither of the reference's volume renderers")


def accelerate(model, fuse_field: bool = True):
    """``model``: a reference ``RadFieldAndRenderer`` (attributes ``rf``, ``renderer``).  Modified in
    place and returned.  A radiance field whose configuration the front-end kernels do not implement
    is left as it is (its torch path keeps working with the new renderer)."""
    model.renderer = convert_renderer(model.renderer)
    if fuse_field:
        try:
            fuse_field_inputs(model.rf)
        except AvrError:
            pass
    return model
