"""GPU probe: per-kernel-class time of one fused render for row-major vs spatially (Morton) sorted rays."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import neumesh_b200 as nb
from neumesh_b200 import synth, _lib
from neumesh_b200.renderer import render_fused

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
dev = torch.device("cuda:0")
cfg = synth.ModelConfig()
mesh = synth.icosphere_mesh(7, seed=0)
sd = synth.make_state_dict(mesh, cfg, seed=1)
model = nb.NeuMesh(nb.MeshGrid(mesh, dev), **cfg.model_kwargs())
model.load_state_dict(sd)
model = model.to(dev).eval()
o, d = synth.frame_rays(800, 800, view=0)
o, d = o[200 * 800: 200 * 800 + n], d[200 * 800: 200 * 800 + n]     # contiguous rows through the object


def part1by2(v):
    v = v.astype(np.uint64) & 0x3FF
    v = (v | (v << 16)) & 0x30000FF
    v = (v | (v << 8)) & 0x300F00F
    v = (v | (v << 4)) & 0x30C30C3
    v = (v | (v << 2)) & 0x9249249
    return v


dn = torch.nn.functional.normalize(d, dim=-1)
mid = (o + dn * (-(o * dn).sum(-1, keepdim=True))).numpy()
q = np.clip(((mid + 1.0) * 0.5 * 1023).astype(np.int64), 0, 1023)
key = (part1by2(q[:, 0]) << 2) | (part1by2(q[:, 1]) << 1) | part1by2(q[:, 2])
perm = torch.from_numpy(np.argsort(key, kind="stable"))
kw = dict(calc_normal=True, white_bkgd=True, bounded_near_far=True)
for name, oo, dd in (("row-major", o, d), ("morton-sorted", o[perm], d[perm])):
    oo, dd = oo.to(dev), dd.to(dev)
    with torch.no_grad():
        render_fused(oo, dd, model, **kw)
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        _lib.profile_collect()
        render_fused(oo, dd, model, **kw)
        torch.cuda.synchronize()
        prof = _lib.profile_collect()
        _lib.profile_enable(False)
    print(name, {k: round(v["ms"], 2) for k, v in prof.items() if v["launches"]}, flush=True)
