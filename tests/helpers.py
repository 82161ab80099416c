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
