"""CPU, world_size 2, gloo: the view-parallel plumbing (sharding, flat gradient buffer, one
all-reduce) gives the same gradient as a single process rendering every view.  The render
function here is a differentiable torch stand-in -- the product has no CPU renderer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jaxsplat_b200 import multiview


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_render(means, scales, quats, colors, opac, view):
    # any smooth function of all five parameter arrays and the view
    a = (means @ view[:3, :3].T + view[:3, 3]).tanh().sum(0)
    b = (scales * colors).sum(0) + quats[:, :3].sin().sum(0) * opac.mean()
    return (a + b).view(1, 1, 3).expand(4, 5, 3)


def _make(seed=0, n=50):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.rand((n, 3), generator=g), torch.rand((n, 3), generator=g),
          torch.rand((n, 4), generator=g), torch.rand((n, 3), generator=g),
          torch.rand((n, 1), generator=g)]
    views = [torch.eye(4) + 0.1 * torch.rand((4, 4), generator=g) for _ in range(5)]
    targets = [torch.rand((4, 5, 3), generator=g) for _ in range(5)]
    return ps, views, targets


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ps, views, targets = _make()
    flat = multiview.FlatParams(ps)
    mine = multiview.shard_views(len(views), rank, world)
    loss = multiview.multiview_step(flat, [views[i] for i in mine], [targets[i] for i in mine],
                                    _fake_render, num_views_total=len(views))
    if rank == 0:
        torch.save({"grad": flat.grad.clone(), "loss": loss}, out)
    dist.destroy_process_group()


def test_shard_views_partition():
    for world in (1, 2, 3, 8):
        seen = sorted(i for r in range(world) for i in multiview.shard_views(64, r, world))
        assert seen == list(range(64))
        sizes = [len(multiview.shard_views(64, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gradient_equals_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    ps, views, targets = _make()
    flat = multiview.FlatParams(ps)
    loss = multiview.multiview_step(flat, views, targets, _fake_render, num_views_total=len(views),
                                    reduce=False)
    assert torch.allclose(got["grad"], flat.grad, rtol=1e-5, atol=1e-7)
    assert torch.allclose(got["loss"], loss, rtol=1e-6)
    # gradient views alias the flat buffer
    for g, p in zip(flat.grads(), flat.params):
        assert g.data_ptr() == p.grad.data_ptr()


def test_tile_row_bands_partition():
    for tiles_y in (1, 7, 68, 135):
        for world in (1, 2, 3, 8):
            if world > tiles_y:
                continue
            bands = multiview.tile_row_bands(tiles_y, world)
            assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == tiles_y
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a1 > a0
            assert bands[-1][1] > bands[-1][0]
    w = [1] * 20
    w[3] = 50
    bands = multiview.tile_row_bands(20, 4, weights=w)
    assert bands[0][0] == 0 and bands[-1][1] == 20 and all(b > a for a, b in bands)


def test_reduced_buffer_scratch_tail_follows_the_reduced_floats():
    """single process: the per-rank scratch of a SwitchReducedBuffer starts exactly where the
    reduced floats end (the render backward relies on that to clear its accumulators at once)"""
    buf = multiview.SwitchReducedBuffer(14 * 10, "cpu", scratch_numel=5 * 10)
    assert buf.mode == "single" and buf.tensor.numel() == 140 and buf.scratch.numel() == 50
    assert buf.scratch.data_ptr() == buf.tensor.data_ptr() + 140 * 4
    buf.tensor.fill_(1.0)
    buf.all_reduce()  # no-op on one rank
    assert float(buf.tensor.sum()) == 140.0 and float(buf.scratch.abs().sum()) == 0.0
    assert multiview.SwitchReducedBuffer(7, "cpu").scratch.numel() == 0


def _fake_band_render(means, scales, quats, colors, opac, tile_rows=None, H=48, W=8):
    """a differentiable stand-in whose rows depend on all five arrays; like the product's band
    render it only produces the rows of its band (the others stay zero and must not be read)"""
    rows = torch.arange(H, dtype=torch.float32).view(H, 1, 1)
    base = (means.sum(0) * scales.mean(0)).view(1, 1, 3) + quats[:, :3].cos().mean(0).view(1, 1, 3)
    img = (base * (1.0 + 0.01 * rows) + (colors * opac).sum(0).view(1, 1, 3) * torch.sin(rows)).expand(H, W, 3)
    if tile_rows is None:
        return img
    mask = torch.zeros((H, 1, 1))
    mask[tile_rows[0] * 16:min(H, tile_rows[1] * 16)] = 1.0
    return img * mask


def _band_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ps, _, _ = _make(seed=3)
    target = torch.rand((48, 8, 3), generator=torch.Generator().manual_seed(9))
    flat = multiview.FlatParams(ps)
    band = multiview.tile_row_bands(3, world)[rank]
    loss = multiview.band_step(flat, target, band, _fake_band_render)
    if rank == 0:
        torch.save({"grad": flat.grad.clone(), "loss": loss}, out)
    dist.destroy_process_group()


def test_two_rank_band_step_equals_the_full_frame(tmp_path):
    """screen-tile-row sharding of one frame over two gloo ranks: summed gradient and loss equal
    the single-process full-frame values"""
    out = str(tmp_path / "band.pt")
    mp.spawn(_band_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    ps, _, _ = _make(seed=3)
    target = torch.rand((48, 8, 3), generator=torch.Generator().manual_seed(9))
    qs = [p.clone().requires_grad_(True) for p in ps]
    full = ((_fake_band_render(*qs) - target) ** 2).mean()
    full.backward()
    ref = torch.cat([q.grad.reshape(-1) for q in qs])
    assert torch.allclose(got["grad"], ref, rtol=1e-5, atol=1e-7)
    full_loss = float(full.detach())
    assert abs(float(got["loss"]) - full_loss) <= 1e-6 * max(1.0, abs(full_loss))


def test_tile_row_weights_equal_the_oracle_row_histogram():
    """multiview.tile_row_weights (planning pass of the tile-row sharding) against the oracle:
    per tile row, the number of intersections in the oracle's tile bins (reference
    forward.cu:135-208 restated in oracle/oracle.c); exact integers, block sizes 16 and 8, an
    image whose last tile row is half covered, Gaussians partly and wholly off screen"""
    import numpy as np

    from jaxsplat_b200 import synthetic
    from oracle import cpu
    for bw, shape, smax in ((16, (120, 200), 0.3), (8, (72, 90), 0.15), (16, (40, 64), 2.0)):
        s = synthetic.benchmark_scene(3000, shape, scale_max=smax, seed=4)
        P = cpu.project_fwd(s.means3d.numpy(), s.scales.numpy(), s.quats.numpy(), s.viewmat.numpy(),
                            img_shape=shape, f=s.f, c=s.c, glob_scale=1.0, clip_thresh=0.01,
                            block_width=bw)
        B = cpu.sort_and_bin(P["xys"], P["depths"], P["radii"], P["cum_tiles_hit"],
                             img_shape=shape, block_width=bw)
        tx, ty = cpu.tiles_of(shape, bw)
        per_tile = (B["tile_bins"][:, 1] - B["tile_bins"][:, 0]).reshape(ty, tx)
        want = per_tile.sum(1)
        got = multiview.tile_row_weights(torch.from_numpy(P["xys"]), torch.from_numpy(P["radii"]),
                                         shape, bw)
        assert got.shape == (ty,)
        assert np.array_equal(got.numpy().astype(np.int64), want.astype(np.int64)), (bw, shape)
        assert int(got.sum()) == int(P["cum_tiles_hit"][-1, 0]) == B["num_intersects"]
        # and the bands built from them are balanced to within twice the heaviest row (greedy cut)
        bands = multiview.tile_row_bands(ty, 2, weights=got.tolist())
        loads = [float(got[a:b].sum()) for a, b in bands]
        assert abs(loads[0] - loads[1]) <= 2.0 * float(got.max()) + 1e-9


def _run_tool(args, nproc=1):
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable]
    if nproc > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                "--master-addr", "127.0.0.1", "--master-port", str(_free_port())]
    cmd += [os.path.join(root, "tools", "multigpu_bench.py")] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout  # ONE JSON line, on rank 0
    return json.loads(lines[0])


def test_multigpu_bench_tool_dry_run_shards_consistently():
    """tools/multigpu_bench.py (the C4 view-batch / C5 tile-row-band lines) in its CPU dry mode
    (gloo, stand-in renderer): a 5-view batch in chunks of 2 and a banded frame give the same
    loss and gradient norm on one rank and on two"""
    for mode in ("views", "bands"):
        args = [mode, "--dry", "--steps", "1", "--warmup", "1", "--views", "5", "--chunk", "2"]
        one = _run_tool(args)
        two = _run_tool(args, nproc=2)
        assert one["n_gpus"] == 1 and two["n_gpus"] == 2 and two["reduction"] == "gloo"
        assert abs(one["loss"] - two["loss"]) <= 1e-5 * abs(one["loss"])
        assert abs(one["grad_l2"] - two["grad_l2"]) <= 1e-5 * abs(one["grad_l2"])
    assert two["bands"][0][0] == 0 and two["bands"][-1][1] == 3 and len(two["bands"]) == 2


def test_flat_adam_argument_checks():
    from jaxsplat_b200.optim import Adam
    with pytest.raises(TypeError):
        Adam([torch.zeros(4)])
    with pytest.raises((TypeError, ValueError)):
        Adam([], grads=[torch.zeros(4)])
