#!/bin/bash
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/r02z_pytest.log 2>&1; tail -4 $OUT/r02z_pytest.log
# block_size 8 at C2: tuned kernels vs the generic ones
python - <<'PY' > gpurun_out/r02z_block8.json
import json, torch, sys, os
sys.path.insert(0, os.getcwd())
import jaxsplat_b200 as jsb
from jaxsplat_b200 import synthetic, _lib
out = {}
for cfg in ("C2", "C3"):
    s = synthetic.make_config(cfg); s.block_size = 8
    sc = s.to("cuda:0"); H, W = s.img_shape
    tgt = torch.rand((H, W, 3), generator=torch.Generator().manual_seed(1)).to("cuda:0")
    for mode in ("tuned", "generic"):
        _lib.set_option("fwdwarp", 1 if mode == "tuned" else 0)
        params = [p.clone().requires_grad_(True) for p in sc.params()]
        def step():
            for p in params: p.grad = None
            img = jsb.render(*params, **sc.render_kwargs())
            ((img - tgt) ** 2).mean().backward()
        for _ in range(5): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): step()
        e1.record(); torch.cuda.synchronize()
        out[f"{cfg}_block8_{mode}_forward"] = e0.elapsed_time(e1) / 20
    _lib.set_option("fwdwarp", 1)
print(json.dumps(out))
PY
cat gpurun_out/r02z_block8.json
