"""host-side cost of a fwd+bwd step (cProfile over a small scene, where the GPU work is shorter than
the host's): python tools/host_profile.py [config=C1] [steps=300]"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jaxsplat_b200 as jsb  # noqa: E402
from jaxsplat_b200 import losses, synthetic  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
sc = synthetic.make_config(cfg).to("cuda:0")
H, W = sc.img_shape
target = torch.rand((H, W, 3), generator=torch.Generator().manual_seed(1)).to("cuda:0")
params = [p.clone().requires_grad_(True) for p in sc.params()]
kw = sc.render_kwargs()


def step():
    for p in params:
        p.grad = None
    img = jsb.render(*params, **kw)
    loss, v = losses.mse_loss_and_grad(img.detach(), target)
    img.backward(gradient=v)


if os.environ.get("JSB_AUTOGRAD_SINGLE_THREAD") == "1":
    torch.autograd.set_multithreading_enabled(False)  # the backward runs in the calling thread
for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    step()
torch.cuda.synchronize()
print("ms per step (wall):", 1e3 * (time.perf_counter() - t0) / steps)
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
