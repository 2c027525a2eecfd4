"""JAX registration of the four custom-call targets -- what a jaxsplat maintainer would import
instead of the reference's ``_jaxsplat`` pybind module.

Import-guarded: JAX is not installed in the build image, so this module cannot be executed there;
it is kept deliberately thin (the torch twin in ``_project.py`` / ``_rasterize.py`` is what the
test-suite drives).  It mirrors jaxsplat/_project/impl.py:12-13 and jaxsplat/_rasterize/impl.py.
"""
from __future__ import annotations

from . import _ffi

try:  # pragma: no cover - JAX absent in this image
    import jax  # noqa: F401
    from jax.lib import xla_client
    HAVE_JAX = True
except Exception:  # ImportError or an incompatible jaxlib
    HAVE_JAX = False


def registrations():
    """drop-in for ``_jaxsplat.registrations()`` (reference lib/ffi.cu:4-13)"""
    return _ffi.registrations()


def make_descriptor(num_points, img_shape, f, c, glob_scale, clip_thresh, block_width) -> bytes:
    """drop-in for ``_jaxsplat.make_descriptor(...)`` (reference lib/ffi.cu:15-52)"""
    return _ffi.make_descriptor(num_points, img_shape, f, c, glob_scale, clip_thresh, block_width)


def register_targets() -> None:
    """registers project_fwd/bwd and rasterize_fwd/bwd for platform "gpu", exactly like
    jaxsplat/_project/impl.py:12-13 does with the reference's capsules"""
    if not HAVE_JAX:
        raise ImportError("jax is not installed: the JAX binding cannot be used in this environment "
                          "(use the torch twin jaxsplat_b200.render/project/rasterize)")
    for name, capsule in registrations().items():  # pragma: no cover
        xla_client.register_custom_call_target(name, capsule, platform="gpu")
