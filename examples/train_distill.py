#!/usr/bin/env python
"""A few distillation training steps with the drop-ins (the role of the reference's ``train.py`` + ``models/trainer.py``):
NeuMesh student rendered under autograd with ``perturb=True`` (fused CUDA sampling cascade + ``FusedFieldFn`` forward /
backward kernels), frozen NeuS teacher evaluated under ``no_grad`` on the library GEMMs, the Trainer's losses
(``models/trainer.py:173-262``: image L1, eikonal, mask BCE, density / colour distillation, indicator regulariser), Adam.

    python examples/train_distill.py --steps 20 --rays 512
"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neumesh_b200 as nb  # noqa: E402
from neumesh_b200 import synth  # noqa: E402

TEACHER_KW = dict(variance_init=0.05, speed_factor=10.0, W_geo_feat=256, obj_bounding_radius=1.0,
                  surface_cfg=dict(embed_multires=6, radius_init=0.5, geometric_init=True, D=8, W=256, skips=[4]),
                  radiance_cfg=dict(embed_multires=-1, embed_multires_view=4, use_view_dirs=True, D=4, W=256, skips=[]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--rays", type=int, default=512)
    ap.add_argument("--level", type=int, default=6)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = synth.ModelConfig()
    mesh = synth.icosphere_mesh(args.level, seed=0)
    student = nb.NeuMesh(nb.MeshGrid(mesh, dev), **cfg.model_kwargs())
    student.load_state_dict(synth.make_state_dict(mesh, cfg, seed=1))
    student = student.to(dev).train()
    teacher = nb.NeuS(**TEACHER_KW).to(dev).eval()       # sphere-initialised sdf of radius 0.5 (geometric init)
    opt = torch.optim.Adam(student.parameters(), lr=5e-4)
    normals0 = student.mesh_grid.get_vertex_normal_torch().detach().clone()
    g = torch.Generator().manual_seed(0)
    kw = dict(calc_normal=True, white_bkgd=False, bounded_near_far=True, detailed_output=True, samples_output=True,
              perturb=True, rayschunk=4096)
    n0, t0 = nb._lib.launch_count(), time.time()
    for it in range(args.steps):
        o, d = synth.frame_rays(400, 400, view=it % 90)
        sel = torch.randint(0, o.shape[0], (args.rays,), generator=g)
        o, d = o[sel].to(dev), d[sel].to(dev)
        target_rgb = torch.rand(args.rays, 3, generator=g).to(dev)          # stands in for the ground-truth pixels
        target_mask = (torch.rand(args.rays, generator=g) > 0.5).float().to(dev)
        rgb, depth, ex = nb.volume_render(o, d, student, **kw)
        with torch.no_grad():
            gt_sdf, gt_rad = teacher(ex["xyz"], ex["dirs"])
        nab = ex["implicit_nablas"].norm(dim=-1)
        losses = {
            "img": F.l1_loss(rgb, target_rgb),
            "eikonal": 0.1 * F.mse_loss(nab, torch.ones_like(nab)),
            "mask": 0.1 * F.binary_cross_entropy(ex["mask_volume"].clamp(1e-3, 1 - 1e-3), target_mask),
            "density": 0.5 * F.l1_loss(ex["density"], gt_sdf.unsqueeze(-1)),
            "color": 0.5 * F.mse_loss(ex["colors"], gt_rad),
            "indicator_reg": 0.01 * F.mse_loss(student.indicator_vector, normals0),
        }
        loss = sum(losses.values())
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if it % 5 == 0 or it + 1 == args.steps:
            print(f"step {it:3d}  loss {loss.item():.4f}  " + "  ".join(f"{k} {v.item():.4f}" for k, v in losses.items()))
    torch.cuda.synchronize()
    print(f"{args.steps} steps of {args.rays} rays in {time.time() - t0:.2f} s; "
          f"{nb._lib.launch_count() - n0} launches of this library's kernels")


if __name__ == "__main__":
    main()
