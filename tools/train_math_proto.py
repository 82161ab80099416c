"""Blueprint of the fused training-path field op (csrc/train.cu): forward + hand-derived backward in float64 torch,
checked against autograd of a plain restatement of the reference's field (neumesh.py:204-260, mesh_grid.py:121-144).

The CUDA kernels implement exactly the formulas of ``manual_backward`` below; this script is the derivation record and
the CPU check of the derivation (run: ``python tools/train_math_proto.py``)."""
import math

import torch

torch.manual_seed(0)
DT = torch.float64


def pe(x, L):
    out = [x]
    for k in range(L):
        out += [torch.sin(x * 2.0 ** k), torch.cos(x * 2.0 ** k)]
    return torch.cat(out, -1)


def softplus100(z):
    return torch.where(z * 100 > 20, z, torch.log1p(torch.exp(100 * z.clamp_max(0.25))) / 100)


def reference_forward(P, x, dirs, idx, w):
    """Plain autograd restatement.  P: dict of parameters; idx, w detached neighbour lists."""
    x = x.clone().requires_grad_(True)
    p = P["verts"][idx]                      # [M,8,3]
    n = P["ind"][idx]
    w1 = P["w1"]
    v = x[:, None, :] - p
    rho = v.norm(dim=-1, keepdim=True)
    m = (n * w1 + v * rho) / (w1 + rho)
    ds = (w * (v * m).sum(-1)).sum(-1, keepdim=True)       # [M,1]
    fg = (P["Fg"][idx] * w[..., None]).sum(-2)
    h = torch.cat([pe(ds, 8), pe(fg, 2)], -1)
    for W, b in P["geo"]:
        h = softplus100(h @ W.t() + b)
    sdf = h @ P["geo_out_w"].t() + P["geo_out_b"]
    nabla = torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=True)[0]
    ft = (P["Fc"][idx] * w[..., None]).sum(-2)
    c = torch.cat([nabla, pe(ds, 8), pe(dirs, 4), pe(ft, 2)], -1)
    for W, b in P["col"]:
        c = torch.relu(c @ W.t() + b)
    rgb = torch.sigmoid(c @ P["col_out_w"].t() + P["col_out_b"])
    return sdf, nabla, rgb


def pe_grad(x, L):
    """d PE(x) / dx, same layout as pe()."""
    out = [torch.ones_like(x)]
    for k in range(L):
        f = 2.0 ** k
        out += [f * torch.cos(x * f), -f * torch.sin(x * f)]
    return torch.cat(out, -1)


def pe_grad2(x, L):
    out = [torch.zeros_like(x)]
    for k in range(L):
        f = 2.0 ** k
        out += [-f * f * torch.sin(x * f), -f * f * torch.cos(x * f)]
    return torch.cat(out, -1)


def manual_forward(P, x, dirs, idx, w):
    S = {}
    p, n, w1 = P["verts"][idx], P["ind"][idx], P["w1"]
    v = x[:, None, :] - p
    rho = v.norm(dim=-1, keepdim=True)
    D = w1 + rho
    a = (v * n).sum(-1, keepdim=True)
    dot = (w1 * a + rho ** 3) / D
    ds = (w[..., None] * dot).sum(-2)                       # [M,1]
    safe = torch.where(rho > 0, rho, torch.ones_like(rho))
    gk = (w1 * n + 3 * rho * v) / D - torch.where(rho > 0, dot / (safe * D), torch.zeros_like(rho)) * v
    G = (w[..., None] * gk).sum(-2)                         # [M,3] = grad_x ds
    fg = (P["Fg"][idx] * w[..., None]).sum(-2)
    X = torch.cat([pe(ds, 8), pe(fg, 2)], -1)
    T = torch.cat([pe_grad(ds, 8), torch.zeros_like(pe(fg, 2))], -1)   # dX / d ds
    hs, ts, zs, as_ = [X], [T], [], []
    h, t = X, T
    for W, b in P["geo"]:
        z = h @ W.t() + b
        a_l = t @ W.t()
        s = torch.sigmoid(100 * z)
        s = torch.where(z * 100 > 20, torch.ones_like(s), s)
        h = softplus100(z)
        t = s * a_l
        zs.append(z); as_.append(a_l); hs.append(h); ts.append(t)
    sdf = h @ P["geo_out_w"].t() + P["geo_out_b"]
    g = t @ P["geo_out_w"].t()                               # [M,1] = d sdf / d ds
    nabla = g * G
    ft = (P["Fc"][idx] * w[..., None]).sum(-2)
    C = torch.cat([nabla, pe(ds, 8), pe(dirs, 4), pe(ft, 2)], -1)
    cs, czs = [C], []
    c = C
    for W, b in P["col"]:
        z = c @ W.t() + b
        c = torch.relu(z)
        czs.append(z); cs.append(c)
    o = c @ P["col_out_w"].t() + P["col_out_b"]
    rgb = torch.sigmoid(o)
    S.update(v=v, rho=rho, D=D, a=a, dot=dot, n=n, ds=ds, G=G, fg=fg, ft=ft, hs=hs, ts=ts, zs=zs, as_=as_, g=g, cs=cs,
             czs=czs, rgb=rgb)
    return sdf, nabla, rgb, S


def manual_backward(P, S, idx, w, b_sdf, b_nabla, b_rgb):
    """Upstream gradients (b_*) -> dict of parameter gradients."""
    out = {}
    M = idx.shape[0]
    # ---- colour MLP ----
    rgb = S["rgb"]
    bo = b_rgb * rgb * (1 - rgb)
    out["col_out_w"] = bo.t() @ S["cs"][-1]
    out["col_out_b"] = bo.sum(0)
    bc = bo @ P["col_out_w"]
    out["col"] = []
    for l in reversed(range(len(P["col"]))):
        W, _ = P["col"][l]
        bz = bc * (S["czs"][l] > 0).to(DT)
        out["col"].insert(0, (bz.t() @ S["cs"][l], bz.sum(0)))
        bc = bz @ W
    bC = bc                                                   # [M, 3 + 17 + 27 + 5 Fc]
    b_nab = b_nabla + bC[:, :3]
    b_ds = (bC[:, 3:20] * pe_grad(S["ds"], 8)).sum(-1, keepdim=True)
    Fc = S["ft"].shape[1]
    b_ft = (bC[:, 47:47 + 5 * Fc] * pe_grad(S["ft"], 2)).reshape(M, 5, Fc).sum(1)
    # ---- nabla = g * G ----
    b_g = (b_nab * S["G"]).sum(-1, keepdim=True)
    b_G = b_nab * S["g"]
    # ---- geometry MLP: value and tangent chains ----
    Wo = P["geo_out_w"]
    out["geo_out_w"] = b_sdf.t() @ S["hs"][-1] + b_g.t() @ S["ts"][-1]
    out["geo_out_b"] = b_sdf.sum(0)
    bh = b_sdf @ Wo
    bt = b_g @ Wo
    out["geo"] = []
    for l in reversed(range(len(P["geo"]))):
        W, _ = P["geo"][l]
        z, a_l = S["zs"][l], S["as_"][l]
        s = torch.sigmoid(100 * z)
        sat = z * 100 > 20
        s1 = torch.where(sat, torch.ones_like(s), s)                        # softplus'
        s2 = torch.where(sat, torch.zeros_like(s), 100 * s * (1 - s))       # softplus''
        ba = bt * s1
        bz = bh * s1 + bt * a_l * s2
        out["geo"].insert(0, (bz.t() @ S["hs"][l] + ba.t() @ S["ts"][l], bz.sum(0)))
        bh = bz @ W
        bt = ba @ W
    bX, bT = bh, bt
    b_ds = b_ds + (bX[:, :17] * pe_grad(S["ds"], 8)).sum(-1, keepdim=True) \
        + (bT[:, :17] * pe_grad2(S["ds"], 8)).sum(-1, keepdim=True)
    Fg = S["fg"].shape[1]
    b_fg = (bX[:, 17:17 + 5 * Fg] * pe_grad(S["fg"], 2)).reshape(M, 5, Fg).sum(1)
    # ---- scatter into the vertex tables ----
    V = P["Fg"].shape[0]
    out["Fg"] = torch.zeros(V, Fg, dtype=DT).index_add_(0, idx.reshape(-1), (w[..., None] * b_fg[:, None, :]).reshape(-1, Fg))
    out["Fc"] = torch.zeros(V, Fc, dtype=DT).index_add_(0, idx.reshape(-1), (w[..., None] * b_ft[:, None, :]).reshape(-1, Fc))
    # ---- mesh distance: ds = sum_k w_k dot_k, G = sum_k w_k gk ----
    v, rho, D, a, dot, n, w1 = S["v"], S["rho"], S["D"], S["a"], S["dot"], S["n"], P["w1"]
    wk = w[..., None]
    pos = (rho > 0).to(DT)
    safe = torch.where(rho > 0, rho, torch.ones_like(rho))
    bGv = (b_G[:, None, :] * v).sum(-1, keepdim=True)        # b_G . v_k
    bGn = (b_G[:, None, :] * n).sum(-1, keepdim=True)
    # d dot / d n = w1 v / D ;  gk = (w1 n + 3 rho v)/D - dot v / (rho D)
    ddot_dn = w1 * v / D
    b_n = wk * (b_ds[:, None, :] * ddot_dn + (w1 / D) * b_G[:, None, :] - pos * bGv / (safe * D) * ddot_dn)
    out["ind"] = torch.zeros(V, 3, dtype=DT).index_add_(0, idx.reshape(-1), b_n.reshape(-1, 3))
    ddot_dw1 = rho * (a - rho ** 2) / D ** 2
    # b_G . d gk / d w1
    dgk_dw1 = bGn / D - (w1 * bGn + 3 * rho * bGv) / D ** 2 - pos * (ddot_dw1 * bGv / (safe * D) - dot * bGv / (safe * D ** 2))
    out["w1"] = (wk * (b_ds[:, None, :] * ddot_dw1 + dgk_dw1)).sum()
    return out


def main():
    V, M, Fg, Fc, Wd = 200, 64, 32, 32, 256
    P = {"verts": torch.randn(V, 3, dtype=DT) * 0.5, "ind": torch.randn(V, 3, dtype=DT),
         "w1": torch.tensor(0.13, dtype=DT), "Fg": torch.randn(V, Fg, dtype=DT), "Fc": torch.randn(V, Fc, dtype=DT)}
    def lin(i, o, s=1.0):
        return (torch.randn(o, i, dtype=DT) * s / math.sqrt(i), torch.randn(o, dtype=DT) * 0.02)
    P["geo"] = [lin(17 + 5 * Fg, Wd, 0.3), lin(Wd, Wd, 0.3), lin(Wd, Wd, 0.3)]
    P["geo_out_w"], P["geo_out_b"] = torch.randn(1, Wd, dtype=DT) / 16, torch.randn(1, dtype=DT) * 0.1
    P["col"] = [lin(3 + 17 + 27 + 5 * Fc, Wd), lin(Wd, Wd), lin(Wd, Wd), lin(Wd, Wd)]
    P["col_out_w"], P["col_out_b"] = torch.randn(3, Wd, dtype=DT) / 16, torch.randn(3, dtype=DT) * 0.1
    leaves = [P["ind"], P["w1"], P["Fg"], P["Fc"], P["geo_out_w"], P["geo_out_b"], P["col_out_w"], P["col_out_b"]]
    leaves += [t for pr in P["geo"] + P["col"] for t in pr]
    for t in leaves:
        t.requires_grad_(True)
    x = torch.randn(M, 3, dtype=DT) * 0.5
    x[0] = P["verts"][5].detach()        # a query exactly on a vertex (rho = 0 branch)
    dirs = torch.nn.functional.normalize(torch.randn(M, 3, dtype=DT), dim=-1)
    d2 = ((x[:, None, :] - P["verts"][None]) ** 2).sum(-1)
    dk, idx = torch.topk(d2, 8, largest=False)
    w = 1.0 / (dk.sqrt() + 1e-7)
    w = (w / w.sum(-1, keepdim=True)).detach()
    sdf, nabla, rgb = reference_forward(P, x, dirs, idx, w)
    b_sdf, b_nabla, b_rgb = torch.randn_like(sdf), torch.randn_like(nabla), torch.randn_like(rgb)
    loss = (sdf * b_sdf).sum() + (nabla * b_nabla).sum() + (rgb * b_rgb).sum()
    grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    ref = {"ind": grads[0], "w1": grads[1], "Fg": grads[2], "Fc": grads[3], "geo_out_w": grads[4], "geo_out_b": grads[5],
           "col_out_w": grads[6], "col_out_b": grads[7]}
    rest = grads[8:]
    ref["geo"] = [(rest[2 * i], rest[2 * i + 1]) for i in range(3)]
    ref["col"] = [(rest[6 + 2 * i], rest[6 + 2 * i + 1]) for i in range(4)]
    with torch.no_grad():
        sdf_m, nabla_m, rgb_m, S = manual_forward(P, x, dirs, idx, w)
        got = manual_backward(P, S, idx, w, b_sdf, b_nabla, b_rgb)
    print("forward  : sdf %.2e  nabla %.2e  rgb %.2e" % ((sdf_m - sdf).abs().max(), (nabla_m - nabla).abs().max(),
                                                          (rgb_m - rgb).abs().max()))
    worst = 0.0
    for k in ("ind", "w1", "Fg", "Fc", "geo_out_w", "geo_out_b", "col_out_w", "col_out_b"):
        e = ((got[k] - ref[k]).abs().max() / ref[k].abs().max().clamp_min(1e-30)).item()
        worst = max(worst, e)
        print(f"grad {k:10s}: rel max err {e:.2e}")
    for name in ("geo", "col"):
        for i, ((gw, gb), (rw, rb)) in enumerate(zip(got[name], ref[name])):
            e = max(((gw - rw).abs().max() / rw.abs().max()).item(), ((gb - rb).abs().max() / rb.abs().max()).item())
            worst = max(worst, e)
            print(f"grad {name}[{i}]    : rel max err {e:.2e}")
    assert worst < 1e-9, worst
    print("manual backward == autograd (float64)")


if __name__ == "__main__":
    main()
