"""Runs ON the GPU box (seconds): jsb200_clear_async on a side stream, then render() fwd+bwd with the
accumulators cleared that way against the backward's own memset."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jaxsplat_b200 as jsb  # noqa: E402
from jaxsplat_b200 import _lib, synthetic  # noqa: E402

dev = "cuda:0"
x = torch.full((1 << 20,), 3.0, device=dev)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
_lib.clear_async(side.cuda_stream, x.data_ptr(), x.numel() * 4)
side.synchronize()
assert float(x.abs().max()) == 0.0
sc = synthetic.benchmark_scene(150000, (480, 640), scale_max=0.1, seed=1).to(dev)


def run():
    params = [p.clone().requires_grad_(True) for p in sc.params()]
    img = jsb.render(*params, **sc.render_kwargs())
    (img ** 2).mean().backward()
    return [p.grad.clone() for p in params]


jsb.set_preclear(False)
a = run()
jsb.set_preclear(True)
b, b2 = run(), run()
torch.cuda.synchronize()
errs = [float((p - q).abs().max() / (q.abs().max() + 1e-30)) for p, q in zip(b + b2, a + a)]
assert all(torch.isfinite(g).all() for g in b + b2) and max(errs) < 1e-4, errs
print("preclear ok: max rel err %.2e, library errors %d (0 = no fall-back to a torch fill)"
      % (max(errs), _lib.lib().jsb200_error_count()))
