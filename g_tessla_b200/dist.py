"""Frame sharding across ranks (one process per GPU) -- the multi-GPU form of the reference's OpenMP frame loop.

Frames are independent (reference src/gta_tri.c:106-296), so rank r of W owns the contiguous range
[r*ceil(F/W), min(F, (r+1)*ceil(F/W))) -- the same partition gta_b200_tessellate_multi uses inside one process --
and the only communication is the final gather of the per-frame results (<= 24 B per frame).
"""
from __future__ import annotations

import numpy as np


def frame_range(nframes: int, rank: int, world: int) -> tuple[int, int]:
    per = (nframes + world - 1) // world
    lo = min(nframes, rank * per)
    return lo, min(nframes, lo + per)


def gather_frames(local, nframes: int, group=None):
    """The path's only exchange: every rank holds the per-frame values of its own frame_range (a 1-D torch tensor, on
    the device with NCCL, on the host with gloo) and receives the full [nframes] tensor."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = (nframes + world - 1) // world
    loc = torch.zeros(per, dtype=local.dtype, device=local.device)
    loc[:len(local)] = local
    parts = [torch.zeros_like(loc) for _ in range(world)]
    dist.all_gather(parts, loc, group=group)
    return torch.cat(parts)[:nframes]


def tessellate_sharded(xyz, box, espace, flags, compute=None, group=None, device=None):
    """Every rank passes the same full arrays (or at least its own slice filled in); returns the full per-frame
    result dict on every rank. `compute(xyz, box, espace, flags) -> dict(area, areabox[, area2d], status)` defaults
    to the CUDA library; tests on CPU-only machines inject a stand-in to exercise the sharding/gather logic."""
    import torch
    import torch.distributed as dist
    if compute is None:
        from . import api
        compute = lambda x, b, e, f: api.tessellate(x, b, e, f)   # noqa: E731
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    F = len(xyz)
    lo, hi = frame_range(F, rank, world)
    r = compute(xyz[lo:hi], None if box is None else box[lo:hi], espace, flags) if hi > lo else None
    keys = ["area", "areabox", "status"] + (["area2d"] if (flags & 2) else [])
    out = {}
    for k in keys:
        loc = torch.zeros(hi - lo, dtype=torch.float64, device=device)
        if r is not None:
            loc[:] = torch.as_tensor(np.asarray(r[k], dtype=np.float64), device=device)
        full = gather_frames(loc, F, group=group).cpu().numpy()
        out[k] = full.astype(np.int32) if k == "status" else full
    return out
