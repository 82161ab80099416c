"""Multi-GPU rendering: rays shard, tables replicate, one all-gather assembles the image.

The reference's strategy is ``nn.DataParallel`` over rays (``models/trainer.py:39-42``: scatter rays, replicate the
module, gather on device 0).  Here: one process per GPU; every rank holds the full mesh / vertex tables / MLPs
(read-only at inference), renders a block-cyclic slice of the ray range and contributes it to ONE
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
    contiguous blocks; rendering uses the block-cyclic ``shard_indices``)."""
    base, rem = divmod(n_rays, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


SHARD_BLOCK = 128  # rays per block of the block-cyclic distribution


def shard_indices(n_rays: int, rank: int, world: int, block: int = SHARD_BLOCK, device=None) -> torch.Tensor:
    """Block-cyclic partition: blocks of ``block`` consecutive rays are dealt round-robin to the ranks.

    Per-ray cost is very uneven (rays that hit the object have ~75 live samples, rays that miss have none), so
    contiguous image bands would leave the ranks that hold the object's band as stragglers, while a ray-by-ray
    interleave would destroy the locality the octree walk relies on.  Blocks of 128 consecutive rays keep neighbouring
    pixels together and still give every rank the same mix; inside a rank the library re-orders its rays along a
    Morton curve anyway."""
    idx = torch.arange(n_rays, device=device)
    return idx[(idx // block) % world == rank]


def shard_count(n_rays: int, rank: int, world: int, block: int = SHARD_BLOCK) -> int:
    full, rem = divmod(n_rays, block * world)
    return full * block + min(max(rem - rank * block, 0), block)


def gather_image(part: dict, n_rays: int, rank: int, world: int) -> "OrderedDict[str, torch.Tensor]":
    """part: this rank's outputs for its ``shard_indices`` -> full-image outputs (caller order) on every rank, through
    ONE ``all_gather_into_tensor`` of the packed ``[rays, C]`` tile."""
    keys = [(k, c) for k, c in PACK_KEYS if k in part]
    width = sum(c for _, c in keys)
    counts = [shard_count(n_rays, r, world) for r in range(world)]
    per = max(max(counts), 1)  # equal-size slots so a single all_gather_into_tensor suffices
    mine = counts[rank]
    ref = part[keys[0][0]]
    tile = torch.zeros(per, width, dtype=torch.float32, device=ref.device)
    col = 0
    for k, c in keys:
        tile[:mine, col:col + c] = part[k].reshape(mine, c)
        col += c
    if world == 1:
        flat = tile[:n_rays]
    else:
        full = torch.empty(world, per, width, dtype=torch.float32, device=ref.device)
        dist.all_gather_into_tensor(full.view(world * per, width), tile)
        flat = torch.empty(n_rays, width, dtype=torch.float32, device=ref.device)
        for r in range(world):
            flat[shard_indices(n_rays, r, world, device=ref.device)] = full[r, :counts[r]]
    out = OrderedDict()
    col = 0
    for k, c in keys:
        v = flat[:, col:col + c]
        out[k] = v.reshape(n_rays).contiguous() if c == 1 else v.contiguous()
        col += c
    return out


def gather_image_contiguous(part: dict, world: int) -> "OrderedDict[str, torch.Tensor]":
    """As ``gather_image`` when every rank rendered an equally long CONTIGUOUS slice of the ray list (whole frames of a
    multi-frame step): rank r's rows land at [r * n, (r + 1) * n) - one ``all_gather_into_tensor``, no re-ordering."""
    keys = [(k, c) for k, c in PACK_KEYS if k in part]
    width = sum(c for _, c in keys)
    ref = part[keys[0][0]]
    n = ref.shape[0]
    tile = torch.cat([part[k].reshape(n, c) for k, c in keys], dim=1).contiguous()
    if world == 1 or not dist.is_initialized():
        flat = tile
    else:
        flat = torch.empty(world * n, width, dtype=torch.float32, device=ref.device)
        dist.all_gather_into_tensor(flat, tile)
    out = OrderedDict()
    col = 0
    for k, c in keys:
        v = flat[:, col:col + c]
        out[k] = v.reshape(-1).contiguous() if c == 1 else v.contiguous()
        col += c
    return out


def render_sharded(rays_o, rays_d, model, **render_kwargs):
    """Render this rank's slice with the fused path and all-gather the image.  Single process: plain render."""
    from .renderer import render_fused
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = rays_o.reshape(-1, 3).shape[0]
    sl = shard_indices(n, rank, world, device=rays_o.device)
    part = render_fused(rays_o.reshape(-1, 3)[sl].contiguous(), rays_d.reshape(-1, 3)[sl].contiguous(), model,
                        **render_kwargs)
    return gather_image(part, n, rank, world)


def render_sharded_local(o_part, d_part, model, n_rays, rank, world, **render_kwargs):
    """As ``render_sharded`` when the caller already holds only this rank's slice of the rays."""
    from .renderer import render_fused
    part = render_fused(o_part, d_part, model, **render_kwargs)
    return gather_image(part, n_rays, rank, world)
