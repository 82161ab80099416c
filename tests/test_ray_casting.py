"""``neumesh_b200.ray_casting`` against the UNMODIFIED ``models/ray_casting.py``
(``tests/golden/ray_casting_small.npz``: root finding over the reference NeuMesh field)."""
import os

import numpy as np
import pytest
import torch

import helpers
from neumesh_b200 import synth

pytestmark = pytest.mark.gpu


def test_root_finding_and_surface_render_vs_reference_golden(golden_dir):
    import neumesh_b200 as nb
    from neumesh_b200 import ray_casting as rc
    g = dict(np.load(os.path.join(golden_dir, "ray_casting_small.npz"), allow_pickle=False))
    dev = torch.device("cuda:0")
    cfg = synth.ModelConfig()
    mesh = synth.icosphere_mesh(4, seed=int(g["seed"]))
    sd = synth.make_state_dict(mesh, cfg, seed=int(g["seed"]) + 1)
    assert helpers.state_digest(sd) == str(g["state_digest"])
    model = helpers.cuda_model(mesh, cfg, sd)
    o, d = torch.from_numpy(g["rays_o"]).to(dev), torch.from_numpy(g["rays_d"]).to(dev)
    fn = lambda x: model.forward_density_only(x).squeeze(-1)   # noqa: E731
    n0 = nb._lib.launch_count()
    dp, pt, mask, msc = rc.root_finding_surface_points(fn, o.clone(), d.clone(), near=1.5, far=3.5, batched=False,
                                                       N_steps=128, N_secant_steps=8)
    assert nb._lib.launch_count() > n0
    m_ref = torch.from_numpy(g["mask"])
    agree = (mask.cpu() == m_ref).float().mean().item()
    both = mask.cpu() & m_ref
    err = (dp.cpu() - torch.from_numpy(g["d_pred"]))[both].abs().max().item()
    print(f"root finding: masks agree on {agree:.4f} of {m_ref.numel()} rays ({int(m_ref.sum())} hits); "
          f"max |depth - reference| on common hits {err:.2e}")
    assert agree >= 0.995 and err < 2e-5
    assert torch.equal(torch.isinf(dp.cpu()), torch.isinf(torch.from_numpy(g["d_pred"]))) or agree < 1.0
    # surface_render on top (field protocol of this package)
    colors, depths, ex = rc.surface_render(o, d, model, batched=False, ray_casting_cfgs=dict(near=1.5, far=3.5, N_steps=128))
    assert colors.shape == (400, 3) and ex["normals_surface"].shape == (400, 3) and ex["mask_surface"].dtype == torch.bool
    hit = ex["mask_surface"]
    assert (colors[~hit] == 0).all() and ((ex["normals_surface"][hit].norm(dim=-1) - 1).abs() < 1e-4).all()
    # sphere tracing converges to the same surface where both find it
    ds, pts, ms = rc.sphere_tracing_surface_points(fn, o, d, near=1.5, far=3.5, batched=False, N_iters=40)
    common = ms & mask
    assert (ds[common] - dp[common]).abs().median() < 1e-3
