"""Shared builders for the tests: identical synthetic inputs for the oracle, the reference and the CUDA path."""
from __future__ import annotations

import hashlib

import numpy as np
import torch

from neumesh_b200 import synth


def state_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def golden_case(path):
    """-> (npz dict, mesh, cfg, state_dict, render kwargs) of a committed golden file."""
    g = dict(np.load(path, allow_pickle=False))
    kw = {k[3:]: bool(v) for k, v in g.items() if k.startswith("kw_")}
    if "nonabla" in path:
        cfg = synth.ModelConfig(enable_nablas_input=False, ln_s=0.4, learn_indicator_weight=True)
    else:
        cfg = synth.ModelConfig()
    mesh = synth.icosphere_mesh(int(g["level"]), seed=int(g["seed"]))
    sd = synth.make_state_dict(mesh, cfg, seed=int(g["seed"]) + 1)
    assert state_digest(sd) == str(g["state_digest"]), "synthetic state_dict is not reproducible on this platform"
    return g, mesh, cfg, sd, kw


def oracle_field(mesh, cfg, sd, dtype=torch.float32):
    from oracle.field import FieldOracle
    return FieldOracle(mesh.vertices, sd, cfg, dtype=dtype)


def cuda_model(mesh, cfg, sd, engine="tcgen05", device="cuda:0"):
    import neumesh_b200 as nb
    mg = nb.MeshGrid(mesh, torch.device(device))
    kw = cfg.model_kwargs()
    model = nb.NeuMesh(mg, mlp_engine=engine, **kw)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval()


def sample_points(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    k = n // 2
    radii = torch.cat([0.5 + 0.05 * torch.randn(k, generator=g), 0.15 + 1.2 * torch.rand(n - k, generator=g)])
    view = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    return dirs * radii[:, None], view


from oracle.mesh_grid import OracleMeshGrid  # noqa: E402,F401  (CPU stand-in for MeshGrid; test infrastructure)


def train_loss(rgb, depth, extras):
    """A scalar that touches everything the Trainer's losses touch (models/trainer.py:197-262): colour, mask/depth and
    the eikonal term on ``implicit_nablas`` (which needs the double backward through the geometry MLP)."""
    nab = extras["implicit_nablas"]
    eik = ((nab.norm(dim=-1) - 1.0) ** 2).mean()
    return rgb.mean() + 0.5 * extras["mask_volume"].mean() + 0.1 * depth.mean() + 0.1 * eik


TRAIN_KW = dict(calc_normal=True, white_bkgd=False, bounded_near_far=True, detailed_output=True, perturb=False)
GRAD_KEYS = ["geometry_features", "color_features", "indicator_vector", "ln_s", "pts_linears.0.weight_v",
             "pts_linears.2.0.weight_g", "density_linear.weight_v", "views_linears.0.weight", "color_linear.0.bias"]


def texture_edit_case(seed=40):
    """Inputs of the texture-editing case: a main model, two reference models on other
    meshes, two overlapping painted regions on the main mesh, transferred colour codes and main->reference rotations.
    Deterministic in ``seed``; shared by ``tests/golden/make_golden.py`` and the tests."""
    g = torch.Generator().manual_seed(seed)
    cfg = synth.ModelConfig()
    main_mesh = synth.icosphere_mesh(4, seed=seed)
    main_sd = synth.make_state_dict(main_mesh, cfg, seed=seed + 1)
    refs = []
    for j, level in enumerate((3, 2)):
        m = synth.icosphere_mesh(level, seed=seed + 10 + j)
        refs.append((m, synth.make_state_dict(m, cfg, seed=seed + 20 + j)))
    v = torch.from_numpy(main_mesh.vertices).float()
    masks = torch.stack([v[:, 0] > 0.1, v[:, 2] > 0.25])                       # [2, V] bool, overlapping regions
    codes = torch.randn(v.shape[0], cfg.color_dim, generator=g)                # transferred colour codes
    rots = []
    for _ in range(2):
        q, _r = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
        if torch.det(q) < 0:
            q[:, 0] = -q[:, 0]
        T = torch.eye(4)
        T[:3, :3] = q.float()
        T[:3, 3] = torch.randn(3, generator=g) * 0.1
        rots.append(T)
    return dict(cfg=cfg, main_mesh=main_mesh, main_sd=main_sd, refs=refs, masks=masks, codes=codes, T=rots)


def texture_edit_oracle(case, dtype=torch.float32):
    from oracle.texture import TextureEditOracle
    main = oracle_field(case["main_mesh"], case["cfg"], case["main_sd"], dtype)
    refs = [oracle_field(m, case["cfg"], sd, dtype) for m, sd in case["refs"]]
    rot = torch.stack([T[:3, :3] for T in case["T"]])
    return TextureEditOracle(main, refs, case["masks"], case["codes"], rot)


NEUS_KW = dict(variance_init=0.05, speed_factor=10.0, W_geo_feat=256, obj_bounding_radius=1.0,
               surface_cfg=dict(embed_multires=6, radius_init=0.5, geometric_init=True, D=8, W=256, skips=[4]),
               radiance_cfg=dict(embed_multires=-1, embed_multires_view=4, use_view_dirs=True, D=4, W=256, skips=[]))


def neus_state_dict(model, seed=50):
    """Deterministic parameters for a NeuS teacher (reference or drop-in: same keys, same shapes), by key name.
    A sphere-like sdf: the model's own geometric initialisation is kept for the structure, then perturbed."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(model.state_dict()):
        v = model.state_dict()[k]
        if k.endswith("weight_v"):
            sd[k] = torch.randn(v.shape, generator=g) / (v.shape[-1] ** 0.5)
        elif k.endswith("weight_g"):
            sd[k] = 0.8 + 0.4 * torch.rand(v.shape, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.05 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = v.clone()
    return sd
