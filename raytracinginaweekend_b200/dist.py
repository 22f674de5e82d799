"""One process per GPU: the reference's own work split (src/lib/rendering.rs:161-252) across ranks.

The reference gives every worker thread the WHOLE image at spp/T samples and then averages the T planes
(`split_work_tasks`, `merge_planes`).  Here the T sample lanes are dealt round-robin to the ranks
(lane id % world_size == rank); each rank's hand-written CUDA path renders its lanes and leaves
   sum over its lanes of (lane sum / lane samples)
on its GPU; the film is then combined with ONE collective — an NCCL all-reduce over NVLink whose reduction is a
pre-multiplied sum, i.e. the 1/T of merge_planes is folded into the collective (`combine_film`).  (Single-process callers
use rt_render_multi instead: a fixed-order peer-memory reduce fused with finalize, bit-identical for any GPU count.)  The Philox streams are keyed by (pixel, global sample index), so the sample set does not depend
on the number of ranks.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from .api import RenderMode, Size2i, World


def n_tasks(thread_count: int, samples_per_pixel: int) -> int:
    """Number of lanes that get at least one sample (split_work_tasks stops at the first empty task)."""
    return min(thread_count, samples_per_pixel)


def lane_partition(thread_count: int, samples_per_pixel: int, rank: int, world_size: int):
    """(lane_first, lane_stride, lanes owned) for `rank`."""
    t = n_tasks(thread_count, samples_per_pixel)
    owned = len(range(rank, t, world_size))
    return rank, world_size, owned


def lane_samples(thread_count: int, samples_per_pixel: int, lane_id: int) -> int:
    """rendering.rs:222-237: whole + (id < remainder)."""
    return samples_per_pixel // thread_count + (1 if lane_id < samples_per_pixel % thread_count else 0)


def combine_film(plane, mult: float, world_size: int, group=None):
    """The film combine of the multi-process path: ONE NCCL all-reduce whose reduction is a pre-multiplied sum, so the `x 1/T` of
    merge_planes (rendering.rs:239-252) is folded into the collective (ncclRedOpCreatePreMulSum) — no extra pass over the film.
    With one rank the library's own scale kernel runs (rt_film_scale_device).  CUDA tensors only: the product path has no CPU code."""
    import torch.distributed as dist
    from . import _cabi
    import torch
    if world_size > 1:
        dist.all_reduce(plane, op=dist._make_nccl_premul_sum(float(mult)), group=group)
    else:
        _cabi.check(_cabi.lib().rt_film_scale_device(plane.data_ptr(), plane.numel(), float(mult), torch.cuda.current_stream(plane.device).cuda_stream))
    return plane


def render_distributed(world: Optional[World], image_size: Size2i, thread_count: int, samples_per_pixel: int, max_depth: int,
                       render_mode: RenderMode = RenderMode.Default, seed: int = 0x5EED, *, group=None, device=None,
                       plane_fn: Optional[Callable] = None, return_stats: bool = False):
    """Render with every rank of the default (or given) process group; returns the full film on every rank as a
    torch tensor [height, width, 3] on `device`.

    plane_fn(lane_first, lane_stride) -> tensor: overrides the per-rank plane producer (tests on CPU/gloo substitute
    the oracle here; the product default is the CUDA path and fails loudly without it).
    """
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    lane_first, lane_stride, _owned = lane_partition(thread_count, samples_per_pixel, rank, world_size)
    stats = {}
    if plane_fn is None:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        scene = world.device_scene(device.index or 0)
        plane = torch.empty((image_size.height, image_size.width, 3), dtype=torch.float32, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        stats = scene.render_planes_device(plane.data_ptr(), lane_first, lane_stride, image_size.width, image_size.height,
                                           samples_per_pixel, max_depth, thread_count, seed, int(render_mode), stream)
    else:
        plane = plane_fn(lane_first, lane_stride)
    mult = np.float32(1.0) / np.float32(n_tasks(thread_count, samples_per_pixel))  # merge_planes multiplier, rendering.rs:240
    if plane_fn is None:
        combine_film(plane, float(mult), world_size, group)  # the film combine: one collective per render, scale folded in
    else:  # tests (CPU tensors, gloo): the same two steps spelled out
        if world_size > 1:
            dist.all_reduce(plane, op=dist.ReduceOp.SUM, group=group)
        plane = plane * float(mult)
    return (plane, stats) if return_stats else plane
