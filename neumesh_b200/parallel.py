"""Multi-GPU rendering: rays shard, tables replicate, one all-gather assembles the image.

The reference's strategy is ``nn.DataParallel`` over rays (``models/trainer.py:39-42``: scatter rays, replicate the
module, gather on device 0).  Here: one process per GPU; every rank holds the full mesh / vertex tables / MLPs
(read-only at inference), renders an interleaved slice of the ray range and contributes it to ONE
``all_gather_into_tensor`` of the packed ``[rays, C]`` output tile (C = 5, or 8 with normals) over NCCL/NVLink.
There is no data-path collective besides that gather (rays are independent).
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.distributed as dist

PACK_KEYS = (("rgb", 3), ("depth_volume", 1), ("mask_volume", 1), ("normals_volume", 3))


def shard_range(n_rays: int, rank: int, world: int):
    """Contiguous, balanced partition of [0, n_rays): sizes differ by at most one (kept for callers that need
    contiguous blocks; rendering uses the interleaved ``shard_slice``)."""
    base, rem = divmod(n_rays, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_slice(rank: int, world: int) -> slice:
    """Interleaved partition: rank r renders rays r, r + world, r + 2 world, ...  Per-ray cost is very uneven (rays
    that hit the object have ~75 live samples, rays that miss have none), so contiguous image bands would leave the
    ranks holding the object's band as stragglers; an interleaved split gives every rank the same mix.  Spatial
    coherence inside a rank is restored by the library's Morton ordering of its rays."""
    return slice(rank, None, world)


def shard_count(n_rays: int, rank: int, world: int) -> int:
    return (n_rays - rank + world - 1) // world if n_rays > rank else 0


def gather_image(part: dict, n_rays: int, rank: int, world: int) -> "OrderedDict[str, torch.Tensor]":
    """part: this rank's outputs for its ``shard_slice`` -> full-image outputs (caller order) on every rank, through
    ONE ``all_gather_into_tensor`` of the packed ``[rays, C]`` tile."""
    keys = [(k, c) for k, c in PACK_KEYS if k in part]
    width = sum(c for _, c in keys)
    per = -(-n_rays // world)  # ceil: equal-size slots so a single all_gather_into_tensor suffices
    mine = shard_count(n_rays, rank, world)
    ref = part[keys[0][0]]
    tile = torch.zeros(per, width, dtype=torch.float32, device=ref.device)
    col = 0
    for k, c in keys:
        tile[:mine, col:col + c] = part[k].reshape(mine, c)
        col += c
    if world == 1:
        full = tile[None]
    else:
        full = torch.empty(world, per, width, dtype=torch.float32, device=ref.device)
        dist.all_gather_into_tensor(full.view(world * per, width), tile)
    # element (rank r, slot i) is ray i * world + r: transpose, flatten, drop the padding at the end
    flat = full.transpose(0, 1).reshape(world * per, width)[:n_rays]
    out = OrderedDict()
    col = 0
    for k, c in keys:
        v = flat[:, col:col + c]
        out[k] = v.reshape(n_rays).contiguous() if c == 1 else v.contiguous()
        col += c
    return out


def render_sharded(rays_o, rays_d, model, **render_kwargs):
    """Render this rank's slice with the fused path and all-gather the image.  Single process: plain render."""
    from .renderer import render_fused
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = rays_o.reshape(-1, 3).shape[0]
    sl = shard_slice(rank, world)
    part = render_fused(rays_o.reshape(-1, 3)[sl].contiguous(), rays_d.reshape(-1, 3)[sl].contiguous(), model,
                        **render_kwargs)
    return gather_image(part, n, rank, world)


def render_sharded_local(o_part, d_part, model, n_rays, rank, world, **render_kwargs):
    """As ``render_sharded`` when the caller already holds only this rank's slice of the rays."""
    from .renderer import render_fused
    part = render_fused(o_part, d_part, model, **render_kwargs)
    return gather_image(part, n_rays, rank, world)
