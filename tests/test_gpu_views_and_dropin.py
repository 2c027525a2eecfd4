"""GPU: the batched-camera entry point (jsb200_render_views_fwd/bwd behind
``jaxsplat_b200.render_views``) against the single-view path, and the four primitives driven
through the compiled ``_jaxsplat`` module / the status-returning targets."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import grad_close, load_golden
import gpu_util as G
import jaxsplat_b200 as jsb
from jaxsplat_b200 import _ffi, _lib, synthetic
from test_dropin_module import load_dropin

pytestmark = pytest.mark.gpu


def _scene_and_cameras(num_views, n=6000, shape=(144, 208), block=16):
    s = synthetic.benchmark_scene(n, shape, scale_max=0.25, seed=5)
    s.block_size = block
    cams = torch.stack(synthetic.ring_cameras(num_views, seed=3)).contiguous()
    return s.to(G.DEV), cams.to(G.DEV)


def _per_view(sc, cams, targets):
    """reference semantics: a Python loop over render(), gradients summed over the views"""
    params = [p.clone().requires_grad_(True) for p in sc.params()]
    imgs = []
    kw = sc.render_kwargs()
    for b in range(cams.shape[0]):
        kw["viewmat"] = cams[b]
        imgs.append(jsb.render(*params, **kw))
    imgs = torch.stack(imgs)
    loss = ((imgs - targets) ** 2).mean()
    loss.backward()
    return imgs.detach(), [p.grad for p in params]


def _batched(sc, cams, targets):
    params = [p.clone().requires_grad_(True) for p in sc.params()]
    kw = sc.render_kwargs()
    del kw["viewmat"]
    imgs = jsb.render_views(*params, viewmats=cams, **kw)
    loss = ((imgs - targets) ** 2).mean()
    loss.backward()
    return imgs.detach(), [p.grad for p in params]


@pytest.mark.parametrize("num_views,slots,block", [(1, 8, 16), (5, 8, 16), (5, 2, 16), (3, 8, 8)])
def test_render_views_equals_the_loop_over_render(num_views, slots, block):
    sc, cams = _scene_and_cameras(num_views, block=block)
    H, W = sc.img_shape
    targets = torch.rand((num_views, H, W, 3), generator=torch.Generator().manual_seed(4)).to(G.DEV)
    _lib.set_option("view_slots", slots)
    try:
        ref_imgs, ref_grads = _per_view(sc, cams, targets)
        before = _lib.stats()
        imgs, grads = _batched(sc, cams, targets)
        after = _lib.stats()
    finally:
        _lib.set_option("view_slots", 8)
    assert torch.equal(imgs, ref_imgs)  # bit-identical, view by view
    assert float((imgs[0] - imgs[-1]).abs().max()) > 0 or num_views == 1  # the cameras do differ
    for a, b in zip(grads, ref_grads):
        ok, err, scale = grad_close(G.npy(a), G.npy(b), rel=1e-4)  # summation order only
        assert ok, (err, scale)
    # the backward reused one forward binning per workspace slot and re-binned the other views
    hits = after["binning_cache_hits"] - before["binning_cache_hits"]
    runs = after["binning_runs"] - before["binning_runs"]
    assert hits == min(num_views, slots)
    assert runs == num_views + (num_views - hits)


def test_render_views_without_tokens_and_empty_batch():
    sc, cams = _scene_and_cameras(2)
    n = sc.means3d.shape[0]
    H, W = sc.img_shape
    kw = dict(num_points=n, img_shape=sc.img_shape, f=sc.f, c=sc.c, glob_scale=sc.glob_scale,
              clip_thresh=sc.clip_thresh, block_width=16)
    from jaxsplat_b200 import _ops
    args = (sc.means3d, sc.scales, sc.quats, cams, sc.colors, sc.opacities, sc.background)
    res, tokens = _ops.render_views_fwd(*args, **kw)
    v = torch.rand((2, H, W, 3), generator=torch.Generator().manual_seed(8)).to(G.DEV)
    with_tokens = _ops.render_views_bwd(*args, *res[:8], v, None, tokens=tokens, **kw)
    stale = _ops.render_views_bwd(*args, *res[:8], v, None, tokens=[0, 0], **kw)  # re-bins both
    for a, b in zip(with_tokens, stale):
        ok, err, scale = grad_close(G.npy(a), G.npy(b), rel=1e-4)
        assert ok, (err, scale)
    # zero views: gradients are zeros, images empty
    res0, tok0 = _ops.render_views_fwd(sc.means3d, sc.scales, sc.quats, cams[:0], sc.colors,
                                       sc.opacities, sc.background, **kw)
    assert res0[8].shape == (0, H, W, 3) and tok0 == []
    g0 = _ops.render_views_bwd(sc.means3d, sc.scales, sc.quats, cams[:0], sc.colors, sc.opacities,
                               sc.background, *res0[:8], v[:0], None, tokens=[], **kw)
    assert all(float(g.abs().max()) == 0.0 for g in g0)


def test_primitives_through_the_compiled_module_match_the_golden_fixture():
    """`_jaxsplat.registrations()` + `_jaxsplat.make_descriptor` driven exactly like the
    reference's capsules: outputs equal the fixture the reference's kernels produced"""
    m = load_dropin()
    table = _ffi.CustomCallTable(m.registrations(), "_jaxsplat (dropin)", check=_lib.check_error,
                                 make_descriptor=m.make_descriptor)
    g = load_golden("g_small16")
    inp = G.golden_inputs(g)
    P = G.project_fwd(inp, g["kw"], table=table)
    for k in ("radii", "num_tiles_hit", "cum_tiles_hit"):
        assert np.array_equal(G.npy(P[k]), g[k]), k
    for k in ("xys", "depths", "conics", "cov3ds", "compensation"):
        assert G.bits_equal(G.npy(P[k]), g[k]) == 0, k
    R = G.rasterize_fwd(inp, P, g["rkw"], table=table)
    assert np.array_equal(G.npy(R["final_idx"]), g["final_idx"])
    assert float(np.abs(G.npy(R["img"]) - g["img"]).max()) <= 1e-5
    B = G.rasterize_bwd(inp, P, R, G.t(g["v_img"]), G.t(g["v_img_alpha"]), g["rkw"], table=table)
    for k in ("v_colors", "v_opacity", "v_xy", "v_conic"):
        ok, err, scale = grad_close(G.npy(B[k]), g[k], rel=1e-4)
        assert ok, (k, err, scale)


def test_status_returning_targets_compute_the_same():
    g = load_golden("g_small16")
    inp = G.golden_inputs(g)
    P = G.project_fwd(inp, g["kw"])
    n = inp["means3d"].shape[0]
    ins = [G.t(inp[k]) for k in ("means3d", "scales", "quats", "viewmat")]
    outs = [torch.empty_like(P[k]) for k in ("cov3ds", "xys", "depths", "radii", "conics",
                                             "compensation", "num_tiles_hit", "cum_tiles_hit")]
    bufs = (ctypes.c_void_p * 12)(*[x.data_ptr() for x in ins + outs])
    d = _ffi.make_descriptor(n, **{k: g["kw"][k] for k in ("img_shape", "f", "c", "glob_scale",
                                                           "clip_thresh", "block_width")})
    status = ctypes.create_string_buffer(64)
    _lib.lib().jsb200_project_fwd_status(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
                                         bufs, d, len(d), ctypes.cast(status, ctypes.c_void_p))
    _lib.check_error("project_fwd_status")
    torch.cuda.synchronize()
    for o, k in zip(outs, ("cov3ds", "xys", "depths", "radii", "conics", "compensation",
                           "num_tiles_hit", "cum_tiles_hit")):
        assert torch.equal(o, P[k]), k


def test_gradient_sink_receives_the_gradients_in_place():
    """set_gradient_sink: the fused backward writes the five parameter gradients straight into
    caller-owned views of one flat buffer (what the data-parallel driver all-reduces), autograd
    adopts those views as .grad, and the values are those of the ordinary path"""
    sc, _ = _scene_and_cameras(1)
    H, W = sc.img_shape
    n = sc.means3d.shape[0]
    target = torch.rand((H, W, 3), generator=torch.Generator().manual_seed(6)).to(G.DEV)

    def grads(sink):
        params = [p.clone().requires_grad_(True) for p in sc.params()]
        jsb.set_gradient_sink(sink)
        try:
            img = jsb.render(*params, **sc.render_kwargs())
            ((img - target) ** 2).mean().backward()
        finally:
            jsb.set_gradient_sink(None)
        return params

    ref = grads(None)
    flat = torch.full((14 * n + 5 * n,), float("nan"), device=G.DEV)
    views, off = [], 0
    for p in sc.params():
        views.append(flat[off:off + p.numel()].view(p.shape))
        off += p.numel()
    got = grads(views + [flat[14 * n:]])
    for p, v, r in zip(got, views, ref):
        assert p.grad.data_ptr() == v.data_ptr()          # adopted, not copied
        ok, err, scale = grad_close(G.npy(v), G.npy(r.grad), rel=1e-4)
        assert ok, (err, scale)
    assert bool(torch.isfinite(flat[:14 * n]).all())
    # without the scratch tail the library allocates its own (same values)
    flat2 = torch.zeros(14 * n, device=G.DEV)
    views2, off = [], 0
    for p in sc.params():
        views2.append(flat2[off:off + p.numel()].view(p.shape))
        off += p.numel()
    grads(views2)
    ok, err, scale = grad_close(G.npy(flat2), G.npy(flat[:14 * n]), rel=1e-4)
    assert ok, (err, scale)
