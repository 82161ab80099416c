"""CPU restatement of ``TextureEditableNeuMesh.forward`` (``editing/texture_neumesh/texture_neumesh.py:52-122``) -
TEST INFRASTRUCTURE (see ``oracle/__init__.py``); pinned against the verbatim reference class by
``tests/golden/texture_edit_small.npz`` (``tests/golden/make_golden.py``).

Geometry (sdf, nabla, bounded near/far distance) is the main model's.  Colour: the main model's colour, and for every
reference model i, on the points whose neighbours include painted vertices (mask i):

    a_paint = sum_k w_k [k painted] / sum_k w_k            a_rest = sum_k w_k [k not painted] / sum_k w_k
    w_ref_k = w_k [k painted] / (sum_k w_k [k painted] + 1e-8)
    c_ref   = colour network of reference model i (ds, R_i dirs, edited code table, main-mesh indices, w_ref, R_i nabla)
    colour  = colour * a_rest + c_ref * a_paint
"""
from __future__ import annotations

import torch

from .field import FieldOracle, positional_encoding


class TextureEditOracle:
    def __init__(self, main: FieldOracle, refs: "list[FieldOracle]", masks: torch.Tensor, edited_codes: torch.Tensor,
                 rotations: "torch.Tensor | None" = None):
        self.main, self.refs = main, refs
        self.masks = masks.bool()                       # [n_ref, V_main]
        self.codes = edited_codes.float()               # [V_main, color_dim]
        self.rot = rotations                            # [n_ref, 3, 3] or None
        self.enable_nablas_input = main.enable_nablas_input
        self.speed_factor = main.speed_factor

    # geometry protocol: the main model's (texture_neumesh.py:40-50)
    def compute_distance(self, xyz):
        return self.main.compute_distance(xyz)

    def forward_s(self):
        return self.main.forward_s()

    def forward_density_only(self, xyz):
        return self.main.forward_density_only(xyz)

    def forward_with_nablas(self, xyz):
        return self.main.forward_with_nablas(xyz)

    def forward(self, xyz, view_dirs):
        main = self.main
        sdf, nabla, d_emb, ds, idx, w = main._sdf_nabla(xyz)                      # :66-72
        colour = main._color_from(d_emb, view_dirs, idx, w, nabla).clone()        # :73-80
        for i, ref in enumerate(self.refs):
            painted = self.masks[i][idx]                                          # :86-92
            w_paint = (w * painted).sum(-1)
            w_rest = (w * (~painted)).sum(-1)
            region = w_paint > 0
            total = w_paint + w_rest
            a_paint, a_rest = (w_paint / total)[region], (w_rest / total)[region]
            w_ref = w * painted                                                   # :98-99
            w_ref = w_ref / (w_ref.sum(-1, keepdim=True) + 1e-8)
            if self.rot is not None:                                              # :102-108
                R = self.rot[i].to(view_dirs.dtype)
                dirs_r = torch.matmul(R, view_dirs.unsqueeze(-1)).squeeze(-1)
                nabla_r = torch.matmul(R, nabla.unsqueeze(-1)).squeeze(-1)
            else:
                dirs_r, nabla_r = view_dirs, nabla
            if bool(region.any()):                                                # :109-121
                d_emb_r = positional_encoding(ds[region], ref.cfg.multires_d)
                c_ref = ref._color_from(d_emb_r, dirs_r[region], idx[region], w_ref[region], nabla_r[region],
                                        table=self.codes)
                colour[region] = colour[region] * a_rest.unsqueeze(-1) + c_ref * a_paint.unsqueeze(-1)
        return sdf, colour
