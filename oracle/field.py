"""CPU restatement of the NeuMesh field model - TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Restates, function by function, ``models/mesh_grid.py:88-144`` (``compute_distance_frnn``),
``models/frameworks/neumesh/neumesh.py`` (``interpolation`` :11-13, ``forward*`` :113-174, ``_forward_density``
:204-237, ``_forward_color`` :239-260, ``compute_distance`` :262-273) and ``models/base.py:52-70`` (``Embedder``) as
plain fp32 torch-CPU tensor code over a flat parameter dict (the reference's ``state_dict``).

Pinned against the verbatim-imported reference by ``tests/test_oracle.py::test_oracle_vs_unmodified_reference`` (run in the build
container) and against ``tests/golden/*.npz`` everywhere else.  The KNN itself (third-party ``frnn``) is
**parity unpinned** - see ``oracle/knn.py``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import knn as _knn


def positional_encoding(x: torch.Tensor, n_freqs: int) -> torch.Tensor:
    """``Embedder.forward`` (models/base.py:52-70) with ``get_embedder`` settings (:73-87): output is
    [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)], each block spanning all input dims."""
    if n_freqs < 0:
        return x
    parts = [x]
    for k in range(n_freqs):
        f = float(2.0 ** k)
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, dim=-1)


def blend_rows(table: torch.Tensor, idx: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """``interpolation`` (neumesh.py:11-13): sum_k table[idx_k] * w_k."""
    return (table[idx] * w[..., None]).sum(dim=-2)


def mesh_distance(xyz: torch.Tensor, vertices: torch.Tensor, indicator: torch.Tensor, w1, K: int = 8):
    """``MeshGrid.compute_distance_frnn`` (models/mesh_grid.py:88-144).

    xyz [M,3] -> ds [M,1] (differentiable in xyz / indicator / w1), idx [M,K] int64, w [M,K] (detached).
    In float64 ("truth" mode of the tests) the neighbours are still the fp32 selection; their distances are then
    re-evaluated in float64.
    """
    d2, idx = _knn.knn_exact(xyz, vertices, K)
    if xyz.dtype == torch.float64:
        d2 = ((xyz.detach()[:, None, :] - vertices[idx]) ** 2).sum(-1)
    dist = d2.sqrt()  # mesh_grid.py:123
    w = 1.0 / (dist + 1e-7)  # :124
    w = w / w.sum(dim=-1, keepdim=True)  # :125
    v = xyz[:, None, :] - vertices[idx]  # :134  dir_vec
    rho = torch.norm(v, dim=-1, keepdim=True)  # :135
    mid = (indicator[idx] * w1 + v * rho) / (w1 + rho)  # :136
    ds = (w[..., None] * (v * mid).sum(dim=-1, keepdim=True)).sum(dim=-2)  # :137-142
    return ds, idx, w


class FieldOracle:
    """The NeuMesh model protocol the renderer uses, over a raw ``state_dict``."""

    def __init__(self, vertices, state_dict, cfg, dtype=torch.float32):
        """``dtype=torch.float64`` evaluates the same fp32 parameters in double precision: the "truth" against which
        the tests measure how accurate each fp32 implementation (this oracle, the CUDA kernels) is."""
        self.cfg = cfg
        self.dtype = dtype
        self.vertices = torch.as_tensor(vertices, dtype=torch.float32).to(dtype).contiguous()
        self.p = {k: v.detach().clone().float().to(dtype) for k, v in state_dict.items()}
        self.speed_factor = cfg.speed_factor
        self.enable_nablas_input = cfg.enable_nablas_input

    # -- parameter views --------------------------------------------------------------------------------------
    def _wn(self, prefix):
        """old-style ``torch.nn.utils.weight_norm`` (dim=0): W = g * v / ||v||_row (neumesh.py:77,81,101)."""
        v, g = self.p[prefix + ".weight_v"], self.p[prefix + ".weight_g"]
        # torch._weight_norm is the primitive torch.nn.utils.weight_norm itself evaluates (same rounding)
        return torch._weight_norm(v, g, 0), self.p[prefix + ".bias"]

    def geo_layers(self):
        names = ["pts_linears.0"] + [f"pts_linears.{i}.0" for i in range(2, self.cfg.D_density + 1)]
        return [self._wn(n) for n in names], self._wn("density_linear")

    def color_layers(self):
        names = ["views_linears.0"] + [f"views_linears.{i}.0" for i in range(2, self.cfg.D_color + 1)]
        return ([(self.p[n + ".weight"], self.p[n + ".bias"]) for n in names],
                (self.p["color_linear.0.weight"], self.p["color_linear.0.bias"]))

    def indicator_weight(self):
        if self.cfg.learn_indicator_weight:
            return torch.sigmoid(self.p["indicator_weight_raw"])  # neumesh.py:173-174
        return 0.1  # neumesh.py:266-268

    def forward_s(self):
        return torch.exp(self.p["ln_s"] * self.speed_factor)  # neumesh.py:170-171

    # -- distance ---------------------------------------------------------------------------------------------
    def compute_distance(self, xyz):
        flat = xyz.reshape(-1, 3)
        ds, idx, w = mesh_distance(flat, self.vertices, self.p["indicator_vector"], self.indicator_weight())
        lead = xyz.shape[:-1]
        return ds.reshape(*lead, -1), idx.reshape(*lead, -1), w.reshape(*lead, -1)

    # -- geometry branch --------------------------------------------------------------------------------------
    def _sdf_from(self, ds, idx, w):
        """neumesh.py:213-218.  Returns (sdf [...,1], d_emb)."""
        c = self.cfg
        d_emb = positional_encoding(ds, c.multires_d)
        fg = blend_rows(self.p["geometry_features"], idx, w)
        h = torch.cat([d_emb, positional_encoding(fg, c.multires_fg)], dim=-1)
        hidden, (w_out, b_out) = self.geo_layers()
        for wl, bl in hidden:
            h = F.softplus(F.linear(h, wl, bl), beta=100)  # nn.Softplus(beta=100), threshold 20
        return F.linear(h, w_out, b_out), d_emb

    def forward_density_only(self, xyz):
        ds, idx, w = self.compute_distance(xyz)
        return self._sdf_from(ds, idx, w)[0]

    def _sdf_nabla(self, xyz):
        """neumesh.py:147-154 + :223-237: nabla = d sdf / d xyz by reverse-mode autograd (idx, w detached)."""
        x = xyz.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            ds, idx, w = self.compute_distance(x)
            sdf, d_emb = self._sdf_from(ds, idx, w)
            (nabla,) = torch.autograd.grad(sdf, x, torch.ones_like(sdf))
        return sdf.detach(), nabla.detach(), d_emb.detach(), ds.detach(), idx, w

    def forward_with_nablas(self, xyz):
        sdf, nabla, *_ = self._sdf_nabla(xyz)
        return sdf, nabla

    # -- colour branch ----------------------------------------------------------------------------------------
    def _color_from(self, d_emb, view_dirs, idx, w, nabla, table=None):
        """neumesh.py:239-260: input = [nabla?, PE_8(ds), PE_4(view), PE_2(ft)].  ``table`` = the ``color_features``
        argument of ``forward_color`` (neumesh.py:156-168); default: the model's own codes."""
        c = self.cfg
        parts = []
        if self.enable_nablas_input:
            parts.append(nabla)
        parts.append(d_emb)
        parts.append(positional_encoding(view_dirs, c.multires_view))
        ft = blend_rows(self.p["color_features"] if table is None else table.to(self.dtype), idx, w)
        parts.append(positional_encoding(ft, c.multires_ft))
        h = torch.cat(parts, dim=-1)
        hidden, (w_out, b_out) = self.color_layers()
        for wl, bl in hidden:
            h = torch.relu(F.linear(h, wl, bl))
        return torch.sigmoid(F.linear(h, w_out, b_out))

    def forward(self, xyz, view_dirs):
        """neumesh.py:113-138 with the defaults the renderer uses (need_nablas=True)."""
        sdf, nabla, d_emb, ds, idx, w = self._sdf_nabla(xyz)
        rgb = self._color_from(d_emb, view_dirs, idx, w, nabla)
        return sdf, rgb
