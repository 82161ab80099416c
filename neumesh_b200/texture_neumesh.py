"""``TextureEditableNeuMesh`` - drop-in for ``editing/texture_neumesh/texture_neumesh.py``.

Texture swap / fill: the main model supplies geometry and its own colour everywhere; inside each painted region the
colour comes from a reference model's colour network evaluated on the main mesh's neighbours with the painted
vertices' (transferred) colour codes, blended by how much of a point's interpolation weight sits on painted vertices.

With ``neumesh_b200.NeuMesh`` models and grad mode off every field evaluation below runs in the CUDA library:
``forward(..., nablas_only=True, return_ds=True)`` -> ``nmb_field_forward_ex`` and both ``forward_color`` calls ->
``nmb_field_color`` (the reference model's colour network reading the edited code table through ``color_table``).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class TextureEditableNeuMesh(nn.Module):
    """Same constructor and model protocol as the reference class (``texture_neumesh.py:8-122``)."""

    def __init__(self, main_model, ref_models, main_editing_masks, main_editing_colorfeats, T_r_m_list=None):
        super().__init__()
        self.main_model = main_model
        self.ref_models = nn.ModuleList(ref_models)
        self.register_buffer("main_editing_masks", main_editing_masks)          # [n_ref, V_main] bool
        self.register_buffer("main_editing_colorfeats", main_editing_colorfeats)  # [V_main, color_dim]
        if T_r_m_list is not None:   # main -> reference frame transforms; only the rotation acts on directions
            self.register_buffer("rot_s_m", torch.stack([T[:3, :3] for T in T_r_m_list], dim=0))
            self.register_buffer("t_s_m", torch.stack([T[:3, 3] for T in T_r_m_list], dim=0))
        else:
            self.rot_s_m = None
            self.t_s_m = None
        self.enable_nablas_input = main_model.enable_nablas_input

    # ---- geometry: the main model's (texture_neumesh.py:40-50) ----
    def compute_distance(self, xyz):
        return self.main_model.compute_distance(xyz)

    def forward_s(self):
        return self.main_model.forward_s()

    def forward_density_only(self, xyz):
        return self.main_model.forward_density_only(xyz)

    def forward_with_nablas(self, xyz):
        return self.main_model.forward_with_nablas(xyz)

    # ---- colour: main colour, over-painted region by region (texture_neumesh.py:52-122) ----
    def forward(self, xyz, view_dirs, need_nablas=True, nablas_only=False):
        main = self.main_model
        sdf, nabla, ds, idx, w = main.forward(xyz, view_dirs, need_nablas=need_nablas, nablas_only=True, return_ds=True)
        out = main.forward_color(ds, view_dirs, main.color_features, indices=idx, weights=w, nabla=nabla).clone()
        for i, ref in enumerate(self.ref_models):
            painted = self.main_editing_masks[i][idx]                 # [..., 8] neighbour is a painted vertex
            w_paint = (w * painted).sum(dim=-1)
            w_rest = (w * (painted == False)).sum(dim=-1)             # noqa: E712  (kept as a separate sum, as there)
            region = w_paint > 0
            total = w_paint + w_rest
            a_paint = (w_paint / total)[region]
            a_rest = (w_rest / total)[region]
            w_ref = w * painted
            w_ref = w_ref / (w_ref.sum(dim=-1, keepdim=True) + 1e-8)   # painted neighbours only, renormalised
            if self.rot_s_m is not None:
                R = self.rot_s_m[i]
                dirs_ref = torch.matmul(R, view_dirs.unsqueeze(-1)).squeeze(-1)
                nabla_ref = torch.matmul(R, nabla.unsqueeze(-1)).squeeze(-1)
            else:
                dirs_ref, nabla_ref = view_dirs, nabla
            if bool(region.any()):
                c_ref = ref.forward_color(ds[region], dirs_ref[region], self.main_editing_colorfeats,
                                          indices=idx[region], weights=w_ref[region], nabla=nabla_ref[region])
                out[region] = out[region] * a_rest.unsqueeze(-1) + c_ref * a_paint.unsqueeze(-1)
        return sdf, out
