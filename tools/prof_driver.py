"""Small, deterministic workload for ncu captures: one fused render of N rays of spiral frame 0 (V = 163 842)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import neumesh_b200 as nb
from neumesh_b200 import synth
from neumesh_b200.renderer import render_fused

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
engine = sys.argv[2] if len(sys.argv) > 2 else "tcgen05"
dev = torch.device("cuda:0")
cfg = synth.ModelConfig()
mesh = synth.icosphere_mesh(7, seed=0)
sd = synth.make_state_dict(mesh, cfg, seed=1)
model = nb.NeuMesh(nb.MeshGrid(mesh, dev), mlp_engine=engine, **cfg.model_kwargs())
model.load_state_dict(sd)
model = model.to(dev).eval()
o, d = synth.frame_rays(800, 800, view=0)
sel = torch.linspace(0, o.shape[0] - 1, n).long()
o, d = o[sel].to(dev), d[sel].to(dev)
with torch.no_grad():
    out = render_fused(o, d, model, calc_normal=True, white_bkgd=True, bounded_near_far=True)
torch.cuda.synchronize()
print("rendered", n, "rays; mean acc", out["mask_volume"].mean().item())
