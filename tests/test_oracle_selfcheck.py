"""CPU: self-consistency of the oracle itself (tier 3): the analytic gradients that
oracle/oracle.c restates from the reference (backward.cu:153-570) against central finite
differences of its own forward, on a small scene.  This pins the oracle's backward independently
of the golden fixtures (which pin it against the reference's CUDA kernels)."""
import numpy as np
import pytest

from jaxsplat_b200 import synthetic
from oracle import cpu


def _loss(s_np, target):
    kw = dict(img_shape=s_np["img_shape"], f=s_np["f"], c=s_np["c"], glob_scale=s_np["glob_scale"],
              clip_thresh=s_np["clip_thresh"], block_width=s_np["block_size"])
    rkw = dict(img_shape=s_np["img_shape"], block_width=s_np["block_size"])
    P = cpu.project_fwd(s_np["means3d"], s_np["scales"], s_np["quats"], s_np["viewmat"], **kw)
    R = cpu.rasterize_fwd(s_np["colors"], s_np["opacities"], s_np["background"], P["xys"],
                          P["depths"], P["radii"], P["conics"], P["cum_tiles_hit"], **rkw)
    return float(np.mean((R["img"].astype(np.float64) - target) ** 2))


@pytest.mark.parametrize("which,key,tol", [("means3d", "v_mean3d", 0.10), ("colors", "v_colors", 1e-3),
                                            ("opacities", "v_opacity", 0.15), ("scales", "v_scale", 0.20)])
def test_oracle_gradient_matches_finite_differences(which, key, tol):
    """<grad, d> vs (L(p + eps d) - L(p - eps d)) / 2 eps along random directions d, as a fraction
    of |grad|_2 (the spread of a directional derivative along a unit-normal direction), median over
    five directions.  The rendered loss is only piecewise smooth -- the 1/255 alpha cut-off, the
    T <= 1e-4 termination and the integer radii move with the parameters, and the reference's
    gradient ignores those jumps by design -- so single directions can be far off (hence the
    median) and growing a splat (scales, opacities) is held looser than moving it; colours are
    exactly linear."""
    # the scene of the GPU twin of this test (tests/test_gpu_api_and_scale.py::test_directional_derivative)
    scene = synthetic.benchmark_scene(400, (64, 64), scale_max=0.5, seed=9)
    s_np = cpu.scene_to_numpy(scene)
    target = np.random.RandomState(3).rand(64, 64, 3).astype(np.float32)
    out = cpu.render_fwd_bwd(s_np, target)
    g = out[key].astype(np.float64).reshape(-1)
    gnorm = float(np.linalg.norm(g))
    assert gnorm > 0
    eps, errs = 1e-3, []
    for seed in range(5):
        d = np.random.RandomState(100 + seed).randn(*s_np[which].shape).astype(np.float32)
        analytic = float((g * d.astype(np.float64).reshape(-1)).sum())
        vals = []
        for sgn in (+1.0, -1.0):
            moved = dict(s_np)
            moved[which] = (s_np[which] + np.float32(sgn * eps) * d).astype(np.float32)
            vals.append(_loss(moved, target))
        errs.append(abs((vals[0] - vals[1]) / (2 * eps) - analytic) / gnorm)
    assert float(np.median(errs)) <= tol, (which, errs)


def test_oracle_loss_is_deterministic_and_thread_count_independent():
    scene = synthetic.benchmark_scene(500, (64, 64), scale_max=0.3, seed=2)
    s_np = cpu.scene_to_numpy(scene)
    target = np.random.RandomState(0).rand(64, 64, 3).astype(np.float32)
    a = cpu.render_fwd_bwd(s_np, target)
    b = cpu.render_fwd_bwd(s_np, target)
    assert a["loss"] == b["loss"]
    assert np.array_equal(a["img"], b["img"]) and np.array_equal(a["final_idx"], b["final_idx"])
    for k in ("v_colors", "v_opacity", "v_mean3d"):
        assert np.allclose(a[k], b[k], rtol=1e-5, atol=1e-9), k
