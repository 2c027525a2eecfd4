"""BASELINE.json configs[0]: the reference's examples/single_image.py fit (10k Gaussians, 256x256,
block_size 16, 100 Adam steps, lr 0.01) restated twice over the same seeded initialisation:

  * fit_cpu : numpy Adam around the C oracle's fwd+bwd (the "CPU oracle" run of the config);
  * fit_gpu : torch.optim.Adam around jaxsplat_b200.render + autograd (the product path).

Parameterisation as in examples/single_image.py:142-150 (render_fn): quats normalised, colors and
opacities through a sigmoid, means and scales raw.  optax.adam(lr) == Adam(b1=.9, b2=.999,
eps=1e-8) without weight decay, which is torch.optim.Adam's default.  The ground-truth image is
seeded uniform noise (no image files travel with the repo).
"""
import numpy as np


def raw_params(scene):
    """raw (pre-activation) parameters whose activations are exactly the Scene's arrays"""
    def logit(x):
        x = np.clip(x.astype(np.float64), 1e-6, 1 - 1e-6)
        return np.log(x / (1 - x)).astype(np.float32)
    s = scene
    return dict(means3d=s.means3d.numpy().copy(), scales=s.scales.numpy().copy(),
                quats=s.quats.numpy().copy(), colors=logit(s.colors.numpy()),
                opacities=logit(s.opacities.numpy()))


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(np.float32)


def fit_cpu(scene, target, steps=100, lr=0.01):
    from oracle import cpu
    p = raw_params(scene)
    m = {k: np.zeros_like(v, dtype=np.float64) for k, v in p.items()}
    v = {k: np.zeros_like(v_, dtype=np.float64) for k, v_ in p.items()}
    base = cpu.scene_to_numpy(scene)
    losses = []
    b1, b2, eps = 0.9, 0.999, 1e-8
    for t in range(1, steps + 1):
        norm = np.linalg.norm(p["quats"], axis=1, keepdims=True)
        q = (p["quats"] / norm).astype(np.float32)
        col, opa = _sigmoid(p["colors"]), _sigmoid(p["opacities"])
        s = dict(base, means3d=p["means3d"], scales=p["scales"], quats=q, colors=col, opacities=opa)
        r = cpu.render_fwd_bwd(s, target)
        losses.append(r["loss"])
        vq = r["v_quat"].reshape(q.shape)
        g = dict(means3d=r["v_mean3d"].reshape(p["means3d"].shape),
                 scales=r["v_scale"].reshape(p["scales"].shape),
                 quats=(vq - q * np.sum(q * vq, axis=1, keepdims=True)) / norm,
                 colors=r["v_colors"].reshape(col.shape) * col * (1 - col),
                 opacities=r["v_opacity"].reshape(opa.shape) * opa * (1 - opa))
        for k in p:
            gk = g[k].astype(np.float64)
            m[k] = b1 * m[k] + (1 - b1) * gk
            v[k] = b2 * v[k] + (1 - b2) * gk * gk
            step = lr * (m[k] / (1 - b1 ** t)) / (np.sqrt(v[k] / (1 - b2 ** t)) + eps)
            p[k] = (p[k].astype(np.float64) - step).astype(np.float32)
    return losses


def fit_gpu(scene, target, steps=100, lr=0.01, device="cuda:0", native_adam=False):
    """native_adam: jaxsplat_b200.optim.Adam (one fused kernel per array) instead of torch's"""
    import torch

    import jaxsplat_b200 as jsb
    p = {k: torch.from_numpy(v).to(device).requires_grad_(True) for k, v in raw_params(scene).items()}
    if native_adam:
        from jaxsplat_b200.optim import Adam
        opt = Adam(list(p.values()), lr=lr)
    else:
        opt = torch.optim.Adam(list(p.values()), lr=lr)
    kw = scene.to(device).render_kwargs()
    tgt = torch.from_numpy(target).to(device)
    losses = []
    for _ in range(steps):
        if native_adam:
            opt.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)
        q = p["quats"] / p["quats"].norm(dim=1, keepdim=True)
        img = jsb.render(p["means3d"], p["scales"], q, torch.sigmoid(p["colors"]),
                         torch.sigmoid(p["opacities"]), **kw)
        loss = ((img - tgt) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return losses
