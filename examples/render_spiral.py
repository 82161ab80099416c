#!/usr/bin/env python
"""Render a few frames of the synthetic spiral with the fused CUDA path and write PNGs (the role of the reference's
``render.py``: spiral track -> get_rays -> renderer -> images).

    python examples/render_spiral.py --views 4 --size 400 --out /tmp/spiral
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neumesh_b200 as nb  # noqa: E402
from neumesh_b200 import synth  # noqa: E402
from neumesh_b200.renderer import get_rays, pack_bgr8  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=4)
    ap.add_argument("--size", type=int, default=400)
    ap.add_argument("--level", type=int, default=6, help="icosphere subdivision level (V = 10 * 4^level + 2)")
    ap.add_argument("--out", default="spiral_out")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = synth.ModelConfig()
    mesh = synth.icosphere_mesh(args.level, seed=0)
    model = nb.NeuMesh(nb.MeshGrid(mesh, dev), **cfg.model_kwargs())
    model.load_state_dict(synth.make_state_dict(mesh, cfg, seed=1))
    model = model.to(dev).eval()
    renderer = nb.SingleRenderer(model)
    os.makedirs(args.out, exist_ok=True)
    H = W = args.size
    f = 1111.1 * W / 800.0
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32)
    import cv2
    for i, pose in enumerate(synth.spiral_poses(n_views=args.views)):
        rays_o, rays_d = get_rays(pose, K, H, W, device=dev)
        with torch.no_grad():
            rgb, depth, extras = renderer(rays_o[None], rays_d[None], batched=True, calc_normal=True, white_bkgd=True,
                                          detailed_output=False, perturb=False)
        img_bgr = pack_bgr8(rgb[0]).reshape(H, W, 3).cpu().numpy()   # clamp, x255, uint8, BGR on the device
        nrm = ((extras["normals_volume"][0].reshape(H, W, 3).cpu().numpy() * 0.5 + 0.5).clip(0, 1) * 255).astype(np.uint8)
        cv2.imwrite(os.path.join(args.out, f"rgb_{i:03d}.png"), img_bgr)
        cv2.imwrite(os.path.join(args.out, f"normal_{i:03d}.png"), nrm[..., ::-1])
        print(f"view {i}: mean acc {extras['mask_volume'].mean().item():.3f}")


if __name__ == "__main__":
    main()
