"""TEST-ONLY torch implementation of the primitives interface of ``neumesh_b200/train_ops.py`` (``CudaPrims``).

Two uses: (1) on CPU, injected into ``field_forward`` / ``field_backward`` / ``FusedFieldFn`` so that the SEQUENCING of the
training op is checked against autograd without a GPU; (2) on the GPU, as the per-kernel reference each ``nmb_tr_*``
kernel is compared with on random inputs.  Nothing in the product imports this module."""
from __future__ import annotations

import torch


def pe(x, L):
    out = [x]
    for k in range(L):
        out += [torch.sin(x * 2.0 ** k), torch.cos(x * 2.0 ** k)]
    return torch.cat(out, -1)


def pe_d1(x, L):
    out = [torch.ones_like(x)]
    for k in range(L):
        f = 2.0 ** k
        out += [f * torch.cos(x * f), -f * torch.sin(x * f)]
    return torch.cat(out, -1)


def pe_d2(x, L):
    out = [torch.zeros_like(x)]
    for k in range(L):
        f = 2.0 ** k
        out += [-f * f * torch.sin(x * f), -f * f * torch.cos(x * f)]
    return torch.cat(out, -1)


def softplus_terms(z):
    sat = z * 100 > 20
    e = torch.exp(100 * torch.where(sat, torch.zeros_like(z), z))
    sp = torch.where(sat, z, torch.log1p(e) / 100)
    s1 = torch.where(sat, torch.ones_like(z), e / (1 + e))
    s2 = torch.where(sat, torch.zeros_like(z), 100 * s1 * (1 - s1))
    return sp, s1, s2


class TorchPrims:
    def __init__(self, device="cpu", dtype=torch.float32):
        self.dev, self.dtype = torch.device(device), dtype

    def empty(self, *shape, dtype=None):
        return torch.full(shape, float("nan"), dtype=dtype or self.dtype, device=self.dev)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=self.dtype, device=self.dev)

    @staticmethod
    def _view(T, ld, kc, rows, cols):
        """rows x cols view of element (r, c) at T[r*ld + c] if kc else T[c*ld + r]"""
        flat = T.reshape(-1)
        return flat.as_strided((rows, cols), (ld, 1) if kc else (1, ld))

    def gemm(self, A, lda, a_kc, B, ldb, b_kc, Cm, ldc, M, N, K, bias=None, epilogue=0, mask=None, ldmask=0,
             accumulate=False):
        Av = self._view(A, lda, a_kc, M, K)                       # A(m,k)
        Bv = self._view(B, ldb, b_kc, N, K).t() if b_kc else self._view(B, ldb, True, K, N)   # B(k,n)
        R = Av @ Bv
        if bias is not None:
            R = R + bias[None, :N]
        if epilogue == 1:
            R = R.clamp_min(0)
        elif epilogue == 2:
            R = torch.where(self._view(mask, ldmask, True, M, N) > 0, R, torch.zeros_like(R))
        Cv = self._view(Cm, ldc, True, M, N)
        if accumulate:
            Cv += R
        else:
            Cv.copy_(R)

    def _mesh(self, spec, t):
        x, idx, w, w1 = t["xyz"], t["idx"], t["w"], t["w1"]
        p, n = t["vertices"][idx], t["indicator_vector"][idx]
        v = x[:, None, :] - p
        rho = v.norm(dim=-1, keepdim=True)
        D = w1 + rho
        a = (v * n).sum(-1, keepdim=True)
        dot = (w1 * a + rho ** 3) / D
        safe = torch.where(rho > 0, rho, torch.ones_like(rho))
        inv_rD = torch.where(rho > 0, 1.0 / (safe * D), torch.zeros_like(rho))
        return v, n, rho, D, a, dot, inv_rD

    def prep(self, spec, t):
        v, n, rho, D, a, dot, inv_rD = self._mesh(spec, t)
        w, w1, idx = t["w"], t["w1"], t["idx"]
        ds = (w[..., None] * dot).sum(-2)
        gk = (w1 * n + 3 * rho * v) / D - dot * inv_rD * v
        t["ds"].copy_(ds[:, 0])
        t["G"].copy_((w[..., None] * gk).sum(-2))
        fg = (t["geometry_features"][idx] * w[..., None]).sum(-2)
        ft = (t["color_features"][idx] * w[..., None]).sum(-2)
        t["Xg"].zero_()
        t["Xg"][:, :spec.Kg] = torch.cat([pe(ds, spec.Ld), pe(fg, spec.Lfg)], -1)
        t["T0"].zero_()
        t["T0"][:, :spec.chd] = pe_d1(ds, spec.Ld)
        t["Xc"].zero_()
        t["Xc"][:, spec.offd:spec.Kc] = torch.cat([pe(ds, spec.Ld), pe(t["dirs"], spec.Lv), pe(ft, spec.Lft)], -1)

    def softplus_fwd(self, z, a, h, t):
        sp, s1, _ = softplus_terms(z)
        h.copy_(sp)
        t.copy_(s1 * a)

    def softplus_bwd(self, z, a, bh, bt, bz, ba):
        _, s1, s2 = softplus_terms(z)
        ba_v = bt * s1
        bz.copy_(bh * s1 + bt * a * s2)
        ba.copy_(ba_v)

    def geo_out_fwd(self, h, t, w_out, b_out, G, sdf, g, nabla, Xc):
        sdf.copy_(h @ w_out.reshape(-1) + b_out.reshape(-1)[0])
        g.copy_(t @ w_out.reshape(-1))
        nabla.copy_(g[:, None] * G)
        if Xc is not None:
            Xc[:, :3] = nabla

    def color_out_fwd(self, c, w_out, b_out, rgb):
        rgb.copy_(torch.sigmoid(c @ w_out.t() + b_out))

    def color_out_bwd(self, b_rgb, rgb, c, w_out, bz, dw_out, db_out):
        bo = b_rgb * rgb * (1 - rgb)
        bz.copy_((bo @ w_out) * (c > 0).to(c.dtype))
        dw_out += bo.t() @ c
        db_out += bo.sum(0)

    def colsum(self, X, out):
        out += X.sum(0)

    def geo_out_bwd(self, b_sdf, b_nabla, bXc, G, g, h, t, w_out, bh, bt, b_G, dw_out, db_out):
        M = G.shape[0]
        bn = torch.zeros(M, 3, dtype=G.dtype, device=G.device)
        if b_nabla is not None:
            bn = bn + b_nabla
        if bXc is not None:
            bn = bn + bXc[:, :3]
        bs = torch.zeros(M, dtype=G.dtype, device=G.device) if b_sdf is None else b_sdf
        bg = (bn * G).sum(-1)
        b_G.copy_(bn * g[:, None])
        wv = w_out.reshape(-1)
        bh.copy_(bs[:, None] * wv[None])
        bt.copy_(bg[:, None] * wv[None])
        dw_out += (bs @ h + bg @ t).reshape(dw_out.shape)
        db_out += bs.sum()

    def input_bwd(self, spec, t, bXg, bT0, bXc, b_G, d_fg, d_fc, d_ind, d_w1):
        v, n, rho, D, a, dot, inv_rD = self._mesh(spec, t)
        w, w1, idx = t["w"], t["w1"], t["idx"]
        ds = t["ds"][:, None]
        M = ds.shape[0]
        d1 = pe_d1(ds, spec.Ld)
        b_ds = (bXg[:, :spec.chd] * d1).sum(-1, keepdim=True) + (bXc[:, spec.offd:spec.offd + spec.chd] * d1).sum(-1, keepdim=True) \
            + (bT0[:, :spec.chd] * pe_d2(ds, spec.Ld)).sum(-1, keepdim=True)
        fg = t["Xg"][:, spec.chd:spec.chd + spec.Fg]
        b_fg = (bXg[:, spec.chd:spec.Kg] * pe_d1(fg, spec.Lfg)).reshape(M, 1 + 2 * spec.Lfg, spec.Fg).sum(1)
        offt = spec.offd + spec.chd + spec.chv
        ft = t["Xc"][:, offt:offt + spec.Fc]
        b_ft = (bXc[:, offt:spec.Kc] * pe_d1(ft, spec.Lft)).reshape(M, 1 + 2 * spec.Lft, spec.Fc).sum(1)
        d_fg.index_add_(0, idx.reshape(-1), (w[..., None] * b_fg[:, None, :]).reshape(-1, spec.Fg))
        d_fc.index_add_(0, idx.reshape(-1), (w[..., None] * b_ft[:, None, :]).reshape(-1, spec.Fc))
        wk = w[..., None]
        bGv = (b_G[:, None, :] * v).sum(-1, keepdim=True)
        bGn = (b_G[:, None, :] * n).sum(-1, keepdim=True)
        s = (b_ds[:, None, :] - bGv * inv_rD) * (w1 / D)
        b_n = wk * (s * v + (w1 / D) * b_G[:, None, :])
        d_ind.index_add_(0, idx.reshape(-1), b_n.reshape(-1, 3))
        if d_w1 is not None:
            ddot = rho * (a - rho ** 2) / D ** 2
            dgk = bGn / D - (w1 * bGn + 3 * rho * bGv) / D ** 2 - (ddot * bGv * inv_rD - dot * bGv * inv_rD / D)
            d_w1 += (wk * (b_ds[:, None, :] * ddot + dgk)).sum()
