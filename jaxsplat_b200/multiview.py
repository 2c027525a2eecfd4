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


def tile_row_weights(xys: torch.Tensor, radii: torch.Tensor, img_shape, block_size: int = 16):
    """Intersections per screen-tile row from the public outputs of ``project`` -- the `weights`
    of ``tile_row_bands`` ("balance bands by intersection count, not rows, using the
    per-row histogram").  Same tile box as the reference's get_tile_bbox (helpers.h:9-32): float32
    ``xy / bw -+ radius / bw (+ 1)``, truncated toward zero, clamped to the tile grid; Gaussians
    with ``radii <= 0`` hit nothing (forward.cu:145).  A planning pass outside the hot path (a few
    elementwise torch ops on whatever device the inputs live on); its sum equals
    ``cum_tiles_hit[-1]``.  Returns a float64 tensor of ``ceil(H / block_size)`` entries."""
    H, W = int(img_shape[0]), int(img_shape[1])
    tiles_x, tiles_y = (W + block_size - 1) // block_size, (H + block_size - 1) // block_size
    bw = float(block_size)
    xy = xys.reshape(-1, 2).to(torch.float32)
    r = radii.reshape(-1).to(torch.float32)
    centre, tr = xy / bw, (r / bw).unsqueeze(1)
    # out-of-range floats saturate in the reference's cvt.rzi.s32.f32; the clamp to the tile grid
    # that follows makes a float clamp before the cast equivalent
    big = float(max(tiles_x, tiles_y) + 1)
    lo = torch.nan_to_num(centre - tr, nan=0.0).clamp(-1.0, big).trunc().to(torch.int64)
    hi = torch.nan_to_num(centre + tr + 1.0, nan=0.0).clamp(-1.0, big).trunc().to(torch.int64)
    x0, x1 = lo[:, 0].clamp(0, tiles_x), hi[:, 0].clamp(0, tiles_x)
    y0, y1 = lo[:, 1].clamp(0, tiles_y), hi[:, 1].clamp(0, tiles_y)
    width = (x1 - x0).clamp_min(0).to(torch.float64)
    width = torch.where((radii.reshape(-1) > 0) & (y1 > y0), width, torch.zeros_like(width))
    # difference array over the rows: +width at y0, -width at y1, prefix sum
    delta = torch.zeros(tiles_y + 1, dtype=torch.float64, device=xy.device)
    delta.index_add_(0, y0, width)
    delta.index_add_(0, y1, -width)
    return delta.cumsum(0)[:tiles_y]


class SwitchReducedBuffer:
    """A flat float32 buffer that every rank of a node allocates symmetrically and that is summed
    across ranks IN the NVSwitch by one native kernel (csrc/collective.cu:
    multimem.ld_reduce + multimem.st, barriers inside the kernel) instead of an NCCL call.

    torch's symmetric-memory allocator is only the plumbing (allocation, handle exchange,
    multicast mapping).  Where the node has no multicast-capable switch (or a single rank) the
    buffer is an ordinary tensor and ``all_reduce`` is ``torch.distributed.all_reduce``:
    ``self.mode`` says which ("nvls-multimem" / "nccl" / "gloo" / "single")."""

    FLAG_BYTES = 4096

    def __init__(self, numel: int, device, group=None, scratch_numel: int = 0):
        """``scratch_numel`` floats of per-rank scratch (``self.scratch``) directly follow the
        reduced floats in memory; they are never meaningfully reduced (see
        ``jaxsplat_b200.set_gradient_sink``: the render backward keeps its two intermediate
        cotangents there, so that it clears all four accumulators with one memset)."""
        self.numel = int(numel)
        self.mode = "single"
        self._epoch = 0
        self._hdl = None
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if distributed:
            self.mode = dist.get_backend(group)
        self._group = group
        padded = (self.numel + 3) // 4 * 4
        total = (self.numel + int(scratch_numel) + 3) // 4 * 4   # flags start here (16-byte aligned)
        if distributed and torch.device(device).type == "cuda" and self.mode == "nccl":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                g = group if group is not None else dist.group.WORLD
                raw = symm_mem.empty(total + self.FLAG_BYTES // 4, dtype=torch.float32, device=device)
                hdl = symm_mem.rendezvous(raw, g.group_name)
                # (two ranks: NCCL's direct exchange is faster than a trip through the switch,
                # 0.129 vs 0.164 ms for 56 MB, profiles/r02m_allreduce_2gpu_switch_vs_nccl.json)
                import os
                min_world = 2 if os.environ.get("JSB_SWITCH_ALLREDUCE") == "1" else 3
                if hdl.multicast_ptr and min_world <= hdl.world_size <= 16:
                    raw.zero_()
                    hdl.barrier(channel=0)
                    self._raw, self._hdl = raw, hdl
                    self._padded = padded
                    self._flags_off = total * 4
                    self.tensor = raw[:self.numel]
                    self.scratch = raw[self.numel:self.numel + int(scratch_numel)]
                    self.mode = "nvls-multimem"
            except Exception:  # no symmetric memory / multicast on this system: NCCL it is
                self._hdl = None
        if self._hdl is None:
            raw = torch.zeros(self.numel + int(scratch_numel), dtype=torch.float32, device=device)
            self.tensor = raw[:self.numel]
            self.scratch = raw[self.numel:]

    def all_reduce(self):
        """in-place sum over the ranks, enqueued on the current stream"""
        if self.mode == "single":
            return
        if self._hdl is None:
            dist.all_reduce(self.tensor, group=self._group)
            return
        import ctypes

        from . import _lib
        hdl, raw = self._hdl, self._raw
        delta = raw.data_ptr() - int(hdl.buffer_ptrs[hdl.rank])
        peers = (ctypes.c_void_p * hdl.world_size)(*[int(p) + delta for p in hdl.buffer_ptrs])
        self._epoch += 1
        rc = _lib.lib().jsb200_multimem_allreduce_f32(
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
            ctypes.c_void_p(int(hdl.multicast_ptr) + delta), peers, self._padded,
            self._flags_off, hdl.rank, hdl.world_size, self._epoch)
        if rc != 0:
            _lib.check_error("multimem_allreduce")


class FlatParams:
    """the five parameter arrays as views into one flat buffer, with their gradients as views
    into a second flat buffer so the reduction is a single collective call"""

    def __init__(self, params: Sequence[torch.Tensor], group=None):
        self.shapes = [tuple(p.shape) for p in params]
        self.sizes = [p.numel() for p in params]
        self.flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
        self.reducer = SwitchReducedBuffer(self.flat.numel(), self.flat.device, group)
        self.grad = self.reducer.tensor
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
        flat.reducer.all_reduce()
        dist.all_reduce(total)
    return total


def band_step(flat: FlatParams, target: torch.Tensor, band: tuple, render_fn: Callable, *,
              block_size: int = 16, reduce: bool = True):
    """One fwd+bwd of ONE large frame sharded by screen-tile rows (BASELINE config 5):
    this rank renders only the tile rows ``band = (begin, end)`` (``render_fn(*params,
    tile_rows=band)`` -> (H, W, 3), e.g. ``jaxsplat_b200.render`` with its keyword arguments bound),
    takes the full-frame MSE restricted to its rows (same normalisation, so the ranks' losses add up
    to the full-frame loss), and the per-Gaussian gradients are summed across ranks with the one
    flat all-reduce.  Returns the rank-local share of the loss (summed over ranks when reducing).
    The image stays sharded: rows outside the band are not written by the band render."""
    flat.zero_grad()
    H, W = target.shape[0], target.shape[1]
    y0, y1 = band[0] * block_size, min(H, band[1] * block_size)
    img = render_fn(*flat.params, tile_rows=band)
    loss = ((img[y0:y1] - target[y0:y1]) ** 2).sum() / float(H * W * target.shape[2])
    loss.backward()
    total = loss.detach().clone()
    if reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        flat.reducer.all_reduce()
        dist.all_reduce(total)
    return total
