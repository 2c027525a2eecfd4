"""JAX host layer over the C ABI -- what a jaxsplat maintainer would ship on a current JAX.

The reference's Python layer (jaxsplat/_project, jaxsplat/_rasterize) is written against JAX
internals that have since been removed (``jax.core.Primitive`` + ``xla.apply_primitive``,
``jaxlib.hlo_helpers.custom_call``, ``xla_client.register_custom_call_target``;
jaxsplat/_project/impl.py:1-19, lowering.py:3).  There are therefore two ways onto this library:

* unchanged reference code on the JAX version it was written for: put
  ``jaxsplat_b200/dropin/`` on ``sys.path`` -- its compiled ``_jaxsplat`` module replaces the
  reference's pybind module one for one;
* this module, for a current JAX: the same public functions ``render / project / rasterize`` with
  the reference's signatures, output shapes / dtypes (jaxsplat/_project/abstract.py:7-22,
  jaxsplat/_rasterize/abstract.py:7-24) and ``custom_vjp`` structure
  (jaxsplat/_project/__init__.py:96-208, jaxsplat/_rasterize/__init__.py:100-190), lowered with
  ``jax.ffi.ffi_call`` onto the four legacy custom-call targets of ``libjaxsplat_b200.so``.

Import-guarded and NOT executed in this repository's test-suite: JAX cannot be installed in the
build image (no network, no wheel).  The torch twin (``jaxsplat_b200.render`` ...) is what the
tests and the benchmark drive; this file is deliberately a thin, table-driven transcription of it.
"""
from __future__ import annotations

from functools import partial

from . import _ffi

try:  # pragma: no cover - JAX absent in this image
    import jax
    import jax.numpy as jnp
    HAVE_JAX = hasattr(jax, "ffi")
except Exception:  # ImportError or an incompatible jaxlib
    HAVE_JAX = False

# XLA custom-call API versions of the legacy (untyped) interface
_API_VERSION_ORIGINAL = 1          # fn(stream, buffers, opaque, opaque_len)
_API_VERSION_STATUS_RETURNING = 2  # ... + XlaCustomCallStatus*  (the jsb200_*_status targets)

_registered = False


def registrations(status_returning: bool = False):
    """drop-in for ``_jaxsplat.registrations()`` (reference lib/ffi.cu:4-13)"""
    if not status_returning:
        return _ffi.registrations()
    import ctypes

    from . import _lib
    L = _lib.lib()
    new = ctypes.pythonapi.PyCapsule_New
    new.restype = ctypes.py_object
    new.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
    return {name: new(ctypes.cast(getattr(L, f"jsb200_{name}_status"), ctypes.c_void_p).value,
                      b"xla._CUSTOM_CALL_TARGET", None) for name in _ffi.TARGETS}


def make_descriptor(num_points, img_shape, f, c, glob_scale, clip_thresh, block_width) -> bytes:
    """drop-in for ``_jaxsplat.make_descriptor(...)`` (reference lib/ffi.cu:15-52)"""
    return _ffi.make_descriptor(num_points, img_shape, f, c, glob_scale, clip_thresh, block_width)


def register_targets() -> None:
    """registers project_fwd/bwd and rasterize_fwd/bwd for CUDA, like
    jaxsplat/_project/impl.py:12-13 does with the reference's capsules; the status-returning
    variants, so that a failing call raises XlaRuntimeError instead of aborting the process"""
    global _registered
    if not HAVE_JAX:
        raise ImportError("jax (with jax.ffi) is not installed: the JAX binding cannot be used in this "
                          "environment (use the torch twin jaxsplat_b200.render/project/rasterize)")
    if _registered:  # pragma: no cover
        return
    for name, capsule in registrations(status_returning=True).items():  # pragma: no cover
        jax.ffi.register_ffi_target(name, capsule, platform="CUDA", api_version=0)
    _registered = True  # pragma: no cover


# ---- the four primitives ------------------------------------------------------------------------
def _call(name, operands, result_specs, opaque):  # pragma: no cover - needs JAX + a GPU
    register_targets()
    out = jax.ffi.ffi_call(name, tuple(jax.ShapeDtypeStruct(s, d) for s, d in result_specs),
                           custom_call_api_version=_API_VERSION_STATUS_RETURNING,
                           legacy_backend_config=opaque)(*operands)
    return tuple(out)


def _project_fwd(desc, mean3ds, scales, quats, viewmat):  # pragma: no cover
    n, f32 = desc["num_points"], jnp.float32
    specs = [((n, 6), f32), ((n, 2), f32), ((n, 1), f32), ((n, 1), jnp.int32), ((n, 3), f32),
             ((n, 1), f32), ((n, 1), jnp.uint32), ((n, 1), jnp.uint32)]
    # (cov3ds, xys, depths, radii, conics, compensation, num_tiles_hit, cum_tiles_hit)
    return _call("project_fwd", (mean3ds, scales, quats, viewmat), specs, make_descriptor(**desc))


def _project_bwd(desc, mean3ds, scales, quats, viewmat, cov3ds, xys, radii, conics, compensation,
                 v_compensation, v_xy, v_depth, v_conic):  # pragma: no cover
    n, f32 = desc["num_points"], jnp.float32
    specs = [((n, 3), f32), ((n, 3), f32), ((n, 4), f32), ((n, 3), f32), ((n, 6), f32)]
    # (v_mean3d, v_scale, v_quat, v_cov2d, v_cov3d)
    return _call("project_bwd", (mean3ds, scales, quats, viewmat, cov3ds, xys, radii, conics,
                                 compensation, v_compensation, v_xy, v_depth, v_conic), specs,
                 make_descriptor(**desc))


def _raster_desc(num_points, img_shape, block_width) -> bytes:
    # the reference's rasterize lowerings pass zeros for f, c, glob_scale, clip_thresh
    # (jaxsplat/_rasterize/lowering.py:28-36)
    return make_descriptor(num_points, img_shape, (0.0, 0.0), (0.0, 0.0), 0.0, 0.0, block_width)


def _rasterize_fwd(num_points, img_shape, block_width, colors, opacities, background, xys, depths,
                   radii, conics, cum_tiles_hit):  # pragma: no cover
    H, W = img_shape
    specs = [((H, W, 1), jnp.float32), ((H, W, 1), jnp.int32), ((H, W, 3), jnp.float32)]
    # (final_Ts, final_idx, img)
    return _call("rasterize_fwd", (colors, opacities, background, xys, depths, radii, conics,
                                   cum_tiles_hit), specs,
                 _raster_desc(num_points, img_shape, block_width))


def _rasterize_bwd(num_points, img_shape, block_width, colors, opacities, background, xys, depths,
                   radii, conics, cum_tiles_hit, final_Ts, final_idx, v_img, v_img_alpha):  # pragma: no cover
    n, f32 = num_points, jnp.float32
    specs = [((n, 3), f32), ((n, 1), f32), ((n, 2), f32), ((n, 2), f32), ((n, 3), f32)]
    # (v_colors, v_opacity, v_xy, v_xy_abs, v_conic)
    return _call("rasterize_bwd", (colors, opacities, background, xys, depths, radii, conics,
                                   cum_tiles_hit, final_Ts, final_idx, v_img, v_img_alpha), specs,
                 _raster_desc(num_points, img_shape, block_width))


# ---- public API (reference jaxsplat/__init__.py:6-73) ---------------------------------------------
if HAVE_JAX:  # pragma: no cover

    @partial(jax.custom_vjp, nondiff_argnums=(0,))
    def _project(static, mean3ds, scales, quats, viewmat):
        return _project_vjp_fwd(static, mean3ds, scales, quats, viewmat)[0]

    def _project_vjp_fwd(static, mean3ds, scales, quats, viewmat):
        desc = dict(static)
        (cov3ds, xys, depths, radii, conics, compensation, num_tiles_hit,
         cum_tiles_hit) = _project_fwd(desc, mean3ds, scales, quats, viewmat)
        primals = (xys, depths, radii, conics, num_tiles_hit, cum_tiles_hit, compensation)
        residuals = (mean3ds, scales, quats, viewmat, cov3ds, xys, radii, conics, compensation)
        return primals, residuals

    def _project_vjp_bwd(static, residuals, cotangents):
        v_xy, v_depth, _, v_conic, _, _, v_compensation = cotangents
        v_mean3d, v_scale, v_quat, _, _ = _project_bwd(dict(static), *residuals, v_compensation,
                                                       v_xy, v_depth, v_conic)
        return v_mean3d, v_scale, v_quat, None   # no gradient for viewmat, like the reference

    _project.defvjp(_project_vjp_fwd, _project_vjp_bwd)

    @partial(jax.custom_vjp, nondiff_argnums=(0,))
    def _rasterize(static, colors, opacities, background, xys, depths, radii, conics,
                   cum_tiles_hit):
        return _rasterize_vjp_fwd(static, colors, opacities, background, xys, depths, radii,
                                  conics, cum_tiles_hit)[0]

    def _rasterize_vjp_fwd(static, *operands):
        final_Ts, final_idx, img = _rasterize_fwd(*static, *operands)
        return (img, 1 - final_Ts), (*operands, final_Ts, final_idx)

    def _rasterize_vjp_bwd(static, residuals, cotangents):
        v_img, v_img_alpha = cotangents
        v_colors, v_opacity, v_xy, _, v_conic = _rasterize_bwd(*static, *residuals, v_img,
                                                               v_img_alpha)
        # gradients for (colors, opacities, xys, conics); none for background, depths, radii,
        # cum_tiles_hit -- the reference's structure (jaxsplat/_rasterize/__init__.py:187)
        return v_colors, v_opacity, None, v_xy, None, None, v_conic, None

    _rasterize.defvjp(_rasterize_vjp_fwd, _rasterize_vjp_bwd)


def project(mean3ds, scales, quats, viewmat, *, img_shape, f, c, glob_scale, clip_thresh,
            block_width):  # pragma: no cover
    """reference jaxsplat/_project/__init__.py:56-92: returns
    (xys, depths, radii, conics, num_tiles_hit, cum_tiles_hit)"""
    static = (("num_points", int(mean3ds.shape[0])), ("img_shape", tuple(img_shape)),
              ("f", tuple(f)), ("c", tuple(c)), ("glob_scale", float(glob_scale)),
              ("clip_thresh", float(clip_thresh)), ("block_width", int(block_width)))
    return _project(static, mean3ds, scales, quats, viewmat)[:6]


def rasterize(colors, opacities, background, xys, depths, radii, conics, cum_tiles_hit, *,
              img_shape, block_width):  # pragma: no cover
    """reference jaxsplat/_rasterize/__init__.py:48-82: returns the (H, W, 3) image"""
    static = (int(colors.shape[0]), tuple(img_shape), int(block_width))
    return _rasterize(static, colors, opacities, background, xys, depths, radii, conics,
                      cum_tiles_hit)[0]


def render(means3d, scales, quats, colors, opacities, *, viewmat, background, img_shape, f, c,
           glob_scale, clip_thresh, block_size=16):  # pragma: no cover
    """reference jaxsplat/__init__.py:6-70"""
    xys, depths, radii, conics, _, cum_tiles_hit = project(
        means3d, scales, quats, viewmat, img_shape=img_shape, f=f, c=c, glob_scale=glob_scale,
        clip_thresh=clip_thresh, block_width=block_size)
    return rasterize(colors, opacities, background, xys, depths, radii, conics, cum_tiles_hit,
                     img_shape=img_shape, block_width=block_size)
