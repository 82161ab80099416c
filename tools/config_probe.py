"""Throughput of the two non-headline BASELINE.json configs on one GPU (diagnostic; bench.py stays on the headline
workload): config 3 = 800x800 frame, 256-d vertex codes; config 5 = 2.6 M vertices, 256 samples per ray."""
import sys
import time

import torch

sys.path.insert(0, ".")
from neumesh_b200 import synth  # noqa: E402
import neumesh_b200 as nb  # noqa: E402
from neumesh_b200.renderer import render_fused  # noqa: E402


def run(tag, level, cfg, kw, n_side=800, chunk=1 << 19):
    dev = torch.device("cuda:0")
    t0 = time.time()
    mesh = synth.icosphere_mesh(level, seed=0)
    sd = synth.make_state_dict(mesh, cfg, seed=1)
    mg = nb.MeshGrid(mesh, dev)
    model = nb.NeuMesh(mg, mlp_engine="tcgen05", **cfg.model_kwargs())
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    o, d = synth.frame_rays(n_side, n_side, view=0)
    o, d = o.to(dev), d.to(dev)
    setup = time.time() - t0
    with torch.no_grad():
        for _ in range(2):
            out = render_fused(o, d, model, chunk=chunk, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            out = render_fused(o, d, model, chunk=chunk, **kw)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print(f"{tag}: V {mesh.vertices.shape[0]} codes {cfg.geometry_dim}/{cfg.color_dim} rays {o.shape[0]} "
          f"samples/ray {kw.get('N_samples', 64) + kw.get('N_importance', 64)}: {ms:.1f} ms/frame "
          f"{o.shape[0] / ms * 1e3:.0f} rays/s (setup {setup:.1f} s, mean acc {out['mask_volume'].mean():.3f})", flush=True)


if __name__ == "__main__":
    base = dict(calc_normal=True, white_bkgd=True, bounded_near_far=True, detailed_output=False)
    run("config 3", 7, synth.ModelConfig(geometry_dim=256, color_dim=256), base)
    run("config 5", 9, synth.ModelConfig(), dict(base, N_samples=128, N_importance=128, N_upsample_iters=4))
