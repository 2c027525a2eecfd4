"""visits per quadrant warp (popcount of the forward's hit words), C3 by default"""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpu_util as G
from jaxsplat_b200 import _lib, _ops, synthetic

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
s = synthetic.make_config(cfg).to("cuda")
n = s.means3d.shape[0]
H, W = s.img_shape
outs, tok = _ops.render_fwd(s.means3d, s.scales, s.quats, s.viewmat, s.colors, s.opacities, s.background,
                            num_points=n, img_shape=s.img_shape, f=s.f, c=s.c, glob_scale=s.glob_scale,
                            clip_thresh=s.clip_thresh, block_width=16)
I = _lib.stats()["last_num_intersects"]
tx, ty = (W + 15) // 16, (H + 15) // 16
got, ids, bins, _ = G.debug_binning(I, tx * ty)
words = torch.zeros(4 * (I // 32 + tx * ty + 1), dtype=torch.int32, device="cuda")
nw = _lib.lib().jsb200_debug_hit_words(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
                                       ctypes.c_void_p(words.data_ptr()), words.numel())
torch.cuda.synchronize()
words = words.cpu().numpy().view(np.uint32)
bins = bins.cpu().numpy()
fidx = outs[7].view(H, W).cpu().numpy()
pop = np.array([bin(int(w)).count("1") for w in range(256)], dtype=np.int64)
def popcount(a):
    a = a.astype(np.uint32)
    return pop[a & 255] + pop[(a >> 8) & 255] + pop[(a >> 16) & 255] + pop[(a >> 24) & 255]
visits = []
walked = []
for t in range(tx * ty):
    x, y = bins[t]
    if y <= x:
        continue
    cap = (y >> 5) - (x >> 5) + 1
    base = 4 * ((x >> 5) + t)
    tyi, txi = divmod(t, tx)
    for q in range(4):
        y0, x0 = tyi * 16 + (q >> 1) * 8, txi * 16 + (q & 1) * 8
        blk = fidx[y0:y0 + 8, x0:x0 + 8]
        if blk.size == 0:
            continue
        top = min(int(blk.max()), y - 1)
        if top < x:
            visits.append(0); walked.append(0); continue
        g = (top - x) >> 5
        w = words[base + q * cap: base + q * cap + g + 1].copy()
        w[-1] &= np.uint32((2 << ((top - x) & 31)) - 1 if ((top - x) & 31) < 31 else 0xFFFFFFFF)
        visits.append(int(popcount(w).sum())); walked.append(top - x + 1)
v = np.array(visits); wk = np.array(walked)
print(cfg, "quadrants", len(v), "total visits", v.sum(), "mean", v.mean(), "median", np.median(v),
      "p90", np.percentile(v, 90), "p99", np.percentile(v, 99), "max", v.max())
print("walked entries per quadrant: mean", wk.mean(), "p99", np.percentile(wk, 99), "max", wk.max())
print("histogram of visits:", np.histogram(v, bins=[0, 10, 20, 40, 80, 160, 320, 640, 1280, 100000])[0])
