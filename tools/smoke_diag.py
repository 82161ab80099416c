import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import neumesh_b200 as nb
from neumesh_b200 import synth
from oracle.field import FieldOracle
dev = torch.device("cuda:0")
cfg = synth.ModelConfig()
mesh = synth.icosphere_mesh(4, seed=0)
sd = synth.make_state_dict(mesh, cfg, seed=1)
oracle = FieldOracle(mesh.vertices, sd, cfg)
o, d = synth.frame_rays(24, 24, view=1)
dn = torch.nn.functional.normalize(d, dim=-1)
for engine in ("fp32", "tcgen05", "tcgen05_f16"):
    model = nb.NeuMesh(nb.MeshGrid(mesh, dev), mlp_engine=engine, **cfg.model_kwargs())
    model.load_state_dict(sd); model = model.to(dev).eval()
    with torch.no_grad():
        rgb, depth, ex = nb.volume_render(o.to(dev), d.to(dev), model, detailed_output=True, calc_normal=True, white_bkgd=True, bounded_near_far=True)
        z = ex["d_all"].cpu()
        pts = o[:, None, :] + z[..., None] * dn[:, None, :]
        ref = oracle.forward_density_only(pts).squeeze(-1)
        got = ex["implicit_surface"].cpu()
        direct = model.forward_density_only(pts.to(dev)).squeeze(-1).cpu()
        ds_o, idx_o, _ = oracle.compute_distance(pts.reshape(-1, 3))
        ds_c, idx_c, _ = model.compute_distance(pts.reshape(-1, 3).to(dev))
    err = (got - ref).abs(); e2 = (direct - ref).abs()
    k = err.argmax(); r, s = divmod(int(k), z.shape[1])
    print(f"[{engine}] carried sdf vs oracle: max {err.max():.3e} at ray {r} sample {s}: z {z[r,s]:.6f} sdf {got[r,s]:.6f} ref {ref[r,s]:.6f} direct {direct[r,s]:.6f} ds {ds_o.reshape(z.shape)[r,s].item():.5f}; "
          f"direct eval vs oracle max {e2.max():.3e}; frac err>5e-6 {(err>5e-6).float().mean():.2e}; idx mismatch rows {(idx_o != idx_c.cpu()).any(-1).float().mean():.2e}; ds err {(ds_o - ds_c.cpu()).abs().max():.2e}")
    big = (err > 5e-6).nonzero()[:8]
    for rr, ss in big.tolist():
        print("    ray", rr, "sample", ss, "z", float(z[rr, ss]), "got", float(got[rr, ss]), "ref", float(ref[rr, ss]), "direct", float(direct[rr, ss]))
