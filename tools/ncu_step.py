"""a few fwd+bwd steps of a BASELINE config through the public API, for ncu captures"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jaxsplat_b200 as jsb  # noqa: E402
from jaxsplat_b200 import synthetic  # noqa: E402

from jaxsplat_b200 import _lib  # noqa: E402

for kv in filter(None, os.environ.get("JSB_OPTS", "").split(",")):  # e.g. JSB_OPTS=supercols=3,fwdsuper=0
    k, v = kv.split("=")
    _lib.set_option(k, int(v))
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = synthetic.make_config(cfg).to("cuda:0")
H, W = sc.img_shape
target = torch.rand((H, W, 3), generator=torch.Generator().manual_seed(1)).to("cuda:0")
params = [p.clone().requires_grad_(True) for p in sc.params()]
for _ in range(steps):
    for p in params:
        p.grad = None
    img = jsb.render(*params, **sc.render_kwargs())
    ((img - target) ** 2).mean().backward()
torch.cuda.synchronize()
