"""View-parallel driver: one process per GPU, Gaussians replicated, cameras sharded, ONE
all-reduce of the flat 14*N-float parameter gradient per step (BASELINE config 4).

The reference has no multi-device code at all; this is the data-parallel wrapper either side of
the hot path.  Backend-agnostic on purpose: NCCL on GPUs, gloo in the CPU tests (where the
``render_fn`` is a stand-in, since the product has no CPU renderer).
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    """round-robin: rank r renders views r, r+world, ...  (balanced to within one view)"""
    return list(range(rank, num_views, world))


def tile_row_bands(tiles_y: int, world: int, weights=None) -> List[tuple]:
    """Screen-tile-row sharding of ONE large frame (BASELINE config 5): contiguous
    bands of tile rows, one per rank, Gaussians replicated.  `weights` (per tile row, e.g. the
    number of intersections in that row) balances the bands by work instead of by rows."""
    if weights is None:
        weights = [1.0] * tiles_y
    total = float(sum(weights))
    bands, start, acc = [], 0, 0.0
    for r in range(world):
        target = total * (r + 1) / world
        end = start
        while end < tiles_y and (acc + weights[end] <= target + 1e-9 or end == start):
            acc += weights[end]
            end += 1
        # leave at least one row for every remaining rank
        end = min(end, tiles_y - (world - 1 - r))
        if r == world - 1:
            end = tiles_y
        end = max(end, start + 1) if start < tiles_y else start
        bands.append((start, end))
        start = end
    return bands


class FlatParams:
    """the five parameter arrays as views into one flat buffer, with their gradients as views
    into a second flat buffer so the reduction is a single collective call"""

    def __init__(self, params: Sequence[torch.Tensor]):
        self.shapes = [tuple(p.shape) for p in params]
        self.sizes = [p.numel() for p in params]
        self.flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
        self.grad = torch.zeros_like(self.flat)
        self.params, off = [], 0
        for shape, sz in zip(self.shapes, self.sizes):
            q = self.flat[off:off + sz].view(shape).requires_grad_(True)
            q.grad = self.grad[off:off + sz].view(shape)
            self.params.append(q)
            off += sz

    def zero_grad(self):
        self.grad.zero_()

    def grads(self):
        out, off = [], 0
        for shape, sz in zip(self.shapes, self.sizes):
            out.append(self.grad[off:off + sz].view(shape))
            off += sz
        return out


def multiview_step(flat: FlatParams, views: Sequence, targets: Sequence[torch.Tensor],
                   render_fn: Callable, *, num_views_total: int, reduce: bool = True):
    """accumulates d(mean over ALL views of MSE)/d(params) for this rank's views into flat.grad,
    then sums across ranks.  Returns the rank-local sum of per-view losses."""
    flat.zero_grad()
    total = torch.zeros((), dtype=torch.float32, device=flat.flat.device)
    for view, target in zip(views, targets):
        img = render_fn(*flat.params, view)
        loss = ((img - target) ** 2).mean() / num_views_total
        loss.backward()
        total = total + loss.detach()
    if reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat.grad)
        dist.all_reduce(total)
    return total
