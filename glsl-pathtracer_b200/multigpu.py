"""Multi-GPU host logic: one process per GPU, sample-range sharding, one reduce of the accumulation buffers per readback.

The path shards naturally (every pixel-sample is independent, reference tile.glsl:41-75), so there is no data-path collective:
rank r renders the sample passes {first + r, first + r + world, ...} of the full frame with the exact per-pixel RNG seeds the
single-GPU run would have used for those passes (`ptb_render_samples(first + r, n, world)`), and the per-rank running sums are
combined with ONE `reduce(SUM)` (NCCL over NVLink on GPUs, gloo in the CPU tests) when the image is read back.
"""
from __future__ import annotations
import numpy as np


def shard_passes(first: int, total: int, rank: int, world: int):
    """Sample passes [first, first+total) split round-robin: returns (first_r, count_r, stride) for `ptb_render_samples`."""
    if world < 1 or not (0 <= rank < world) or total < 0 or first < 1:
        raise ValueError("bad shard arguments")
    count = (total - rank + world - 1) // world if total > rank else 0
    return first + rank, count, world


def passes_of(first: int, total: int, rank: int, world: int):
    f, c, s = shard_passes(first, total, rank, world)
    return [f + i * s for i in range(c)]


def reduce_accum(accum_tensor, dst: int = 0, group=None, scratch=None):
    """Sum of the per-rank running sums, delivered on rank `dst` in a SCRATCH tensor (returned; `scratch` is reused if given).
    The running sums themselves are never modified: each rank's buffer stays the partial sum of its own passes, so the image can be
    read back any number of times while rendering goes on (an in-place reduce would fold the other ranks' earlier passes into rank
    dst's buffer again at every readback).  `accum_tensor` views the accumulation buffer: a CUDA tensor over `ptb_accum_device_ptr`
    with the nccl backend, a CPU tensor with gloo.  On ranks other than `dst` the returned tensor's content is unspecified."""
    import torch
    import torch.distributed as dist
    if scratch is None or scratch.shape != accum_tensor.shape or scratch.device != accum_tensor.device:
        scratch = torch.empty_like(accum_tensor)
    scratch.copy_(accum_tensor)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(scratch, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return scratch


class DeviceAccumView:
    """Zero-copy torch view of a context's accumulation buffer (float32 [h, w, 4]) through __cuda_array_interface__."""

    def __init__(self, ctx):
        ptr, nbytes = ctx.accum_device_ptr()
        w, h = ctx.size
        assert nbytes == w * h * 16
        self.__cuda_array_interface__ = {"shape": (h, w, 4), "typestr": "<f4", "data": (ptr, False), "version": 3}

    def tensor(self, device):
        import torch
        return torch.as_tensor(self, device=device)


class ShardedRenderer:
    """`Renderer`-like driver for N ranks: every rank holds a full scene replica and renders its share of each batch of passes."""

    def __init__(self, ctx, rank: int, world: int, reduce_fn=reduce_accum):
        self.ctx, self.rank, self.world, self.reduce_fn = ctx, rank, world, reduce_fn
        self.next_pass = 1
        self._scratch = None

    def reduced(self, accum_tensor, dst: int = 0):
        """The summed image so far on rank `dst` (scratch tensor); safe to call between render() batches."""
        self._scratch = self.reduce_fn(accum_tensor, dst=dst, scratch=self._scratch)
        return self._scratch

    def render(self, total_passes: int):
        f, c, s = shard_passes(self.next_pass, total_passes, self.rank, self.world)
        if c:
            self.ctx.render_samples(f, c, s)
        self.next_pass += total_passes

    def samples_total(self) -> int:
        return self.next_pass - 1
