"""Surface rendering by ray casting - drop-in for ``models/ray_casting.py`` (``root_finding_surface_points`` :45-200,
``sphere_tracing_surface_points`` :203-227, ``surface_render`` :228-320; unused by the reference's own entry points but
part of the rendering path BASELINE.json names).

The field queries run on the fused density kernels (``forward_density_only`` of a ``neumesh_b200.NeuMesh`` / the fused
``NeuS`` teacher); the per-ray search for the first outside-to-inside sign change over the ``N_steps`` proposals is one
CUDA kernel (``nmb_first_crossing``); the secant refinement (``run_secant_method``) keeps the reference's arithmetic.

Protocol notes.  The reference's ``surface_render`` expects a UNISURF-style model (``model.implicit_surface`` and
``model.forward -> (colour, _, nablas)``), which none of its shipped models provides; here any model with the field
protocol of this package works: the surface function is ``model.implicit_surface.forward`` when present, otherwise
``model.forward_density_only``; colour and normals come from ``model.forward`` / ``model.forward_with_nablas``."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib


def run_secant_method(f_low, f_high, d_low, d_high, rays_o_masked, rays_d_masked, implicit_surface_query_fn, n_secant_steps,
                      logit_tau):
    """models/ray_casting.py:12-38."""
    d_pred = -f_low * (d_high - d_low) / (f_high - f_low) + d_low
    for _ in range(n_secant_steps):
        p_mid = rays_o_masked + d_pred.unsqueeze(-1) * rays_d_masked
        with torch.no_grad():
            f_mid = implicit_surface_query_fn(p_mid).squeeze(-1) - logit_tau
        low = f_mid < 0
        d_low = torch.where(low, d_pred, d_low)
        f_low = torch.where(low, f_mid, f_low)
        d_high = torch.where(low, d_high, d_pred)
        f_high = torch.where(low, f_high, f_mid)
        d_pred = -f_low * (d_high - d_low) / (f_high - f_low) + d_low
    return d_pred


def _surface_fn(model):
    if hasattr(model, "implicit_surface"):
        return model.implicit_surface.forward
    return lambda x: model.forward_density_only(x).squeeze(-1)


def root_finding_surface_points(surface_query_fn, rays_o, rays_d, near=0.0, far=6.0, batched=True, batched_info={},
                                N_steps=256, logit_tau=0.0, method="secant", N_secant_steps=8, fill_inf=True):
    """-> (d_pred [(B), N], pt_pred [(B), N, 3], mask, mask_sign_change); rays_d already normalised."""
    with torch.no_grad():
        dev = rays_o.device
        if not batched:
            rays_o, rays_d = rays_o.unsqueeze(0), rays_d.unsqueeze(0)
        B, N = rays_o.shape[0], rays_o.shape[-2]
        near_t = near if torch.is_tensor(near) else near * torch.ones(rays_o.shape[:-1], device=dev)
        far_t = far if torch.is_tensor(far) else far * torch.ones(rays_o.shape[:-1], device=dev)
        near_t, far_t = near_t.reshape(B, N).float(), far_t.reshape(B, N).float()
        t = torch.linspace(0.0, 1.0, N_steps, device=dev)[None, None, :]
        d_prop = near_t[..., None] * (1 - t) + far_t[..., None] * t
        pts = rays_o.unsqueeze(-2) + d_prop.unsqueeze(-1) * rays_d.unsqueeze(-2)
        val = surface_query_fn(pts).reshape(B * N, N_steps).float().contiguous()
        n = B * N
        if val.is_cuda:
            outs = [torch.empty(n, device=dev) for _ in range(4)]
            masks = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(3)]
            nf, ff = near_t.reshape(-1).contiguous(), far_t.reshape(-1).contiguous()
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().nmb_first_crossing(_lib.ptr(val), n, N_steps, float(logit_tau), _lib.ptr(nf),
                                                         _lib.ptr(ff), *[_lib.ptr(o) for o in outs],
                                                         *[_lib.ptr(m) for m in masks], _lib.stream_ptr(dev)))
            d_low, f_low, d_high, f_high = outs
            mask, mask_sign_change, first_free = [m.bool() for m in masks]
        else:
            raise RuntimeError("neumesh_b200.ray_casting needs CUDA tensors: the kernels have no CPU implementation")
        ro, rd = rays_o.reshape(n, 3), rays_d.reshape(n, 3)
        if method == "secant" and bool(mask.any()):
            d_pred = run_secant_method(f_low[mask], f_high[mask], d_low[mask], d_high[mask], ro[mask], rd[mask],
                                       surface_query_fn, N_secant_steps, logit_tau)
        else:
            d_pred = torch.ones(int(mask.sum()), device=dev)
        pt_pred = torch.ones(n, 3, device=dev)
        pt_pred[mask] = ro[mask] + d_pred.unsqueeze(-1) * rd[mask]
        d_out = torch.ones(n, device=dev)
        d_out[mask] = d_pred
        d_out[~mask] = float("inf") if fill_inf else far_t.reshape(-1)[~mask]
        d_out[~first_free] = 0     # the first proposal is occupied: depth 0
        shape = (B, N) if batched else (N,)
        return (d_out.reshape(shape), pt_pred.reshape(*shape, 3), mask.reshape(shape), mask_sign_change.reshape(shape))


def sphere_tracing_surface_points(implicit_surface, rays_o, rays_d, near=0.0, far=6.0, batched=True, batched_info={},
                                  N_iters=20):
    """models/ray_casting.py:203-227; ``implicit_surface``: module with ``forward(x) -> sdf [...]`` or a callable."""
    fn = implicit_surface.forward if hasattr(implicit_surface, "forward") else implicit_surface
    d = torch.ones(rays_o.shape[:-1], device=rays_o.device) * near
    mask = torch.ones_like(d, dtype=torch.bool)
    with torch.no_grad():
        for _ in range(N_iters):
            val = fn(rays_o + rays_d * d[..., None])
            d = torch.where(mask, d + val.reshape(d.shape), d)
            mask = mask & ~(d > far) & ~(d < 0)
    return d, rays_o + rays_d * d[..., None], mask


def surface_render(rays_o, rays_d, model, calc_normal=True, rayschunk=8192, netchunk=1048576, batched=True,
                   use_view_dirs=True, show_progress=False, ray_casting_algo="root_finding", ray_casting_cfgs={},
                   **not_used_kwargs):
    """models/ray_casting.py:228-320 -> (colors, depths, extras{implicit_nablas, mask_surface, normals_surface})."""
    with torch.no_grad():
        flat = [rays_d.shape[0], -1, 3] if batched else [-1, 3]
        dimb = 1 if batched else 0
        rays_o = torch.reshape(rays_o, flat).float()
        rays_d = F.normalize(torch.reshape(rays_d, flat).float(), dim=-1)
        surf = _surface_fn(model)
        pieces = []
        for i in range(0, rays_o.shape[dimb], rayschunk):
            sl = (slice(None), slice(i, i + rayschunk)) if batched else (slice(i, i + rayschunk),)
            o, d = rays_o[sl], rays_d[sl]
            if ray_casting_algo == "root_finding":
                depth, pt, mask, _ = root_finding_surface_points(surf, o, d, batched=batched, **ray_casting_cfgs)
            elif ray_casting_algo == "sphere_tracing":
                depth, pt, mask = sphere_tracing_surface_points(surf, o, d, batched=batched, **ray_casting_cfgs)
            else:
                raise NotImplementedError
            out = model.forward(pt, d if use_view_dirs else None)
            if len(out) == 3:            # UNISURF-style protocol of the reference's function: (colour, _, nablas)
                color, _, nablas = out
            else:                        # field protocol of this package: (sdf, colour) + forward_with_nablas
                color = out[1]
                nablas = model.forward_with_nablas(pt)[1]
            color = torch.where(mask[..., None], color, torch.zeros_like(color))
            pieces.append((color, depth, nablas, mask))
        colors, depths, nablas, masks = (torch.cat([p[k] for p in pieces], dimb) for k in range(4))
        extras = OrderedDict([("implicit_nablas", nablas), ("mask_surface", masks)])
        if calc_normal:
            normals = F.normalize(nablas, dim=-1)
            extras["normals_surface"] = torch.where(masks[..., None], normals, torch.zeros_like(normals))
        return colors, depths, extras
