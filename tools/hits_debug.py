"""debug: hit-driven backward (bwdhits 1), all-entries replay (2) and the batch kernel (0) on a scene"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpu_util as G
from jaxsplat_b200 import _lib, synthetic
import test_gpu_oracle_and_edges as T

for name, fn in (("adversarial", T._adversarial_scene), ("C1", lambda: synthetic.make_config("C1")),
                 ("small", lambda: synthetic.benchmark_scene(60, (40, 40), scale_max=0.3, seed=3)),
                 ("C2_30k", lambda: synthetic.make_config("C2", num_points=30000))):
    s = fn()
    inp = G.scene_inputs(s)
    kw, rkw = G.scene_kw(s)
    H, W = s.img_shape
    target = np.random.RandomState(2).rand(H, W, 3).astype(np.float32)
    res = {}
    for mode in (0, 1, 2):
        _lib.set_option("bwdhits", mode)
        res[mode] = T._full_gpu(inp, kw, rkw, target, None)
    _lib.set_option("bwdhits", 1)
    for k in ("v_colors", "v_opacity", "v_xy", "v_conic"):
        sc = np.abs(res[0][k]).max()
        e1 = np.abs(res[1][k] - res[0][k]).max()
        e2 = np.abs(res[2][k] - res[0][k]).max()
        bad = np.argwhere(np.abs(res[1][k] - res[0][k]).reshape(res[0][k].shape[0], -1).max(1) > 1e-4 * sc)
        print(name, k, "scale %.3e hits err %.3e all-entries err %.3e  bad gaussians %d" % (sc, e1, e2, len(bad)), bad[:5].ravel())
