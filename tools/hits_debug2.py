"""debug: which (pixel, entry) pairs the backward needs are missing from the forward's hit words"""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpu_util as G
from jaxsplat_b200 import _lib, synthetic

s = synthetic.benchmark_scene(60, (40, 40), scale_max=0.3, seed=3)
inp = G.scene_inputs(s)
kw, rkw = G.scene_kw(s)
H, W = s.img_shape
P = G.project_fwd(inp, kw)
R = G.rasterize_fwd(inp, P, rkw)
I = _lib.stats()["last_num_intersects"]
tx, ty = (W + 15) // 16, (H + 15) // 16
got, ids, bins, _ = G.debug_binning(I, tx * ty)
words = torch.zeros(4 * (I // 32 + tx * ty + 1), dtype=torch.int32, device="cuda")
nw = _lib.lib().jsb200_debug_hit_words(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
                                       ctypes.c_void_p(words.data_ptr()), words.numel())
torch.cuda.synchronize()
print("I", I, "words", nw)
words = words.cpu().numpy().view(np.uint32)
ids = ids.cpu().numpy(); bins = bins.cpu().numpy()
xys = P["xys"].cpu().numpy(); conics = P["conics"].cpu().numpy(); opac = inp["opacities"].reshape(-1)
fidx = R["final_idx"].cpu().numpy().reshape(H, W)
missing = 0; needed = 0; detail = []
for t in range(tx * ty):
    x, y = bins[t]
    if y <= x: continue
    cap = (y >> 5) - (x >> 5) + 1
    base = 4 * ((x >> 5) + t)
    tyi, txi = divmod(t, tx)
    for q in range(4):
        for e in range(x, y):
            g = ids[e]
            need = False
            for py in range(tyi * 16 + (q >> 1) * 8, tyi * 16 + (q >> 1) * 8 + 8):
                for px in range(txi * 16 + (q & 1) * 8, txi * 16 + (q & 1) * 8 + 8):
                    if py >= H or px >= W or e > fidx[py, px]: continue
                    dx = xys[g, 0] - (px + 0.5); dy = xys[g, 1] - (py + 0.5)
                    sg = 0.5 * (conics[g, 0] * dx * dx + conics[g, 2] * dy * dy) + conics[g, 1] * dx * dy
                    if sg < 0: continue
                    if min(0.99, opac[g] * np.exp(-sg)) < 1.0 / 255 * 1.001: continue
                    need = True
            if need:
                needed += 1
                rel = e - x
                w = words[base + q * cap + (rel >> 5)]
                if not (w >> (rel & 31)) & 1:
                    missing += 1
                    detail.append((t, q, rel, hex(w)))
print("needed", needed, "missing", missing)
print(detail[:40])
for t in range(tx * ty):
    x, y = bins[t]
    cap = (y >> 5) - (x >> 5) + 1
    base = 4 * ((x >> 5) + t)
    print("tile", t, "range", x, y, "cap", cap, "base", base, [hex(v) for v in words[base:base + 4 * cap]])
