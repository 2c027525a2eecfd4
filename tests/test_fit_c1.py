"""BASELINE.json configs[0] (examples/single_image.py fit): 10k Gaussians, 256x256, block_size 16,
Adam lr 0.01.  The CPU half runs the C oracle; the GPU half runs the product path and has to land
on the same loss curve (final MSE within 1 % relative)."""
import numpy as np
import pytest

import fit_util
from jaxsplat_b200 import synthetic


def _c1():
    scene = synthetic.make_config("C1")
    target = np.random.RandomState(0).rand(256, 256, 3).astype(np.float32)
    return scene, target


def test_oracle_fit_descends():
    scene, target = _c1()
    losses = fit_util.fit_cpu(scene, target, steps=12)
    assert all(np.isfinite(losses))
    assert losses[-1] < 0.8 * losses[0]
    assert all(b < a for a, b in zip(losses, losses[1:]))


@pytest.mark.gpu
def test_config1_fit_100_adam_steps_matches_oracle_run():
    scene, target = _c1()
    cpu_losses = fit_util.fit_cpu(scene, target, steps=100)
    gpu_losses = fit_util.fit_gpu(scene, target, steps=100)
    # same starting point: the first losses agree to float32 summation noise
    assert abs(gpu_losses[0] - cpu_losses[0]) <= 1e-5 * cpu_losses[0]
    # the curves stay together (MUFU.EX2 vs exp2f decision flips perturb single pixels only)
    for k in (1, 5, 10, 25, 50, 99):
        assert abs(gpu_losses[k] - cpu_losses[k]) <= 1e-2 * cpu_losses[k], (k, gpu_losses[k], cpu_losses[k])
    assert gpu_losses[-1] < 0.7 * gpu_losses[0]


@pytest.mark.gpu
def test_native_adam_matches_torch_adam():
    import torch
    from jaxsplat_b200.optim import Adam
    gen = torch.Generator().manual_seed(0)
    shapes = [(1000, 3), (1000, 4), (777, 1), (5,)]
    a = [torch.randn(s, generator=gen).cuda().requires_grad_(True) for s in shapes]
    b = [x.detach().clone().requires_grad_(True) for x in a]
    oa, ob = Adam(a, lr=0.01), torch.optim.Adam(b, lr=0.01)
    for step in range(7):
        grads = [torch.randn(s, generator=gen).cuda() * (10.0 ** (step - 3)) for s in shapes]
        for x, y, g in zip(a, b, grads):
            x.grad, y.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
        for x, y in zip(a, b):
            assert torch.allclose(x, y, rtol=2e-6, atol=1e-7), (step, (x - y).abs().max())


@pytest.mark.gpu
def test_flat_adam_reads_the_reduced_gradient_buffer():
    """Adam.for_flat: one kernel over the 14*N-float FlatParams buffer, gradients read
    from the flat all-reduced buffer in place, equals the per-array Adam on the same gradients"""
    import torch
    from jaxsplat_b200 import multiview
    from jaxsplat_b200.optim import Adam
    gen = torch.Generator().manual_seed(1)
    n = 1003  # 3n, 6n, 10n, 13n: array boundaries that are not 16-byte aligned
    shapes = [(n, 3), (n, 3), (n, 4), (n, 3), (n, 1)]
    init = [torch.randn(s, generator=gen).cuda() for s in shapes]
    flat = multiview.FlatParams(init)
    ref = [x.clone().requires_grad_(True) for x in init]
    of, orf = Adam.for_flat(flat, lr=0.01), Adam(ref, lr=0.01)
    assert flat.reducer.mode == "single"
    for step in range(5):
        flat.zero_grad()
        for q, y in zip(flat.params, ref):
            g = torch.randn(q.shape, generator=gen).cuda() * (10.0 ** (step - 2))
            q.grad.add_(g)   # what backward passes accumulate into the flat buffer's views
            y.grad = g.clone()
        flat.reducer.all_reduce()  # single rank: no-op
        of.step()
        orf.step()
        for q, y in zip(flat.params, ref):
            assert torch.allclose(q.detach(), y.detach(), rtol=2e-6, atol=1e-7), step
    # the parameter views still alias the flat buffer the kernel updated
    assert flat.params[0].data_ptr() == flat.flat.data_ptr()


@pytest.mark.gpu
def test_config1_fit_with_the_native_adam_step():
    scene, target = _c1()
    ref = fit_util.fit_gpu(scene, target, steps=30)
    got = fit_util.fit_gpu(scene, target, steps=30, native_adam=True)
    for k in (0, 1, 10, 29):
        assert abs(got[k] - ref[k]) <= 2e-3 * ref[k], (k, got[k], ref[k])


@pytest.mark.gpu
def test_native_mse_loss_and_gradient_match_torch():
    """losses.mse_loss_and_grad (one kernel) against the eager expressions: the cotangent bit for
    bit, the loss to float tolerance; sizes that are and are not multiples of 4, a misaligned view"""
    import torch
    from jaxsplat_b200.losses import mse_loss_and_grad
    gen = torch.Generator().manual_seed(5)
    for shape, off in (((1080, 1920, 3), 0), ((37, 53, 3), 0), ((5,), 0), ((64, 64, 3), 1)):
        n = 1
        for d in shape:
            n *= d
        a = torch.rand(n + off, generator=gen).cuda()[off:].view(shape)
        b = torch.rand(n + off, generator=gen).cuda()[off:].view(shape)
        loss, v = mse_loss_and_grad(a, b)
        diff = a - b
        assert torch.equal(v, diff * (2.0 / diff.numel()))
        ref = float((diff.double() ** 2).mean())
        assert abs(float(loss) - ref) <= 1e-5 * ref
        # an 8-bit target decoded inside the kernel == the decoded float target
        t8 = (torch.rand(n + off, generator=gen) * 256).clamp(0, 255).to(torch.uint8).cuda()[off:].view(shape)
        loss8, v8 = mse_loss_and_grad(a, t8)
        lossf, vf = mse_loss_and_grad(a, t8.float().mul_(1.0 / 255.0))
        assert torch.equal(v8, vf)
        assert abs(float(loss8) - float(lossf)) <= 1e-6 * float(lossf)
        # the partial sums are added in CTA order by the CTA that draws the last ticket: the same
        # bits on every launch, back to back without a memset, and on a second stream (own scratch)
        again = [mse_loss_and_grad(a, b)[0] for _ in range(5)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            other = [mse_loss_and_grad(a, b)[0] for _ in range(3)]
        side.synchronize()
        for l in again + other:
            assert torch.equal(l.view(torch.int32), loss.view(torch.int32))
