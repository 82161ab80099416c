"""Numerical study for the round-2 MLP engine (CPU, torch): error of the geometry MLP's sdf against a float64 evaluation
of the same fp32 parameters, for three operand schemes with fp32 accumulation -

* ``fp32``     : plain fp32 linear layers (what the oracle / the reference compute),
* ``3xTF32``   : a = hi + lo with tf32 (10-bit mantissa) parts, D = a_lo b_hi + a_hi b_lo + a_hi b_hi  (the current engine),
* ``fp16x3``   : fp16 (10-bit mantissa) parts with exponent management so that ONE accumulator suffices:
                 B' = 2^8 W;  D' = a_hi B'_hi + (2^11 a_lo) (2^-11 B'_hi) + a_hi B'_lo;  z = 2^-8 D'
                 - three f16-kind MMAs per K step run at twice the tf32 rate and read half the operand bytes;
* ``fp16x3p``  : the same without the 2^11 scaling of a_lo (B then needs only two images, exactly like 3xTF32).

Products of two 11-bit significands are exact in fp32, so emulating each MMA as an fp32 matmul of the rounded operands
differs from the tensor core only in accumulation order.  Usage: python tools/split_precision_study.py
"""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import helpers  # noqa: E402
from neumesh_b200 import synth  # noqa: E402
from oracle.field import positional_encoding, blend_rows  # noqa: E402


def tf32_rna(x):
    """round-to-nearest, ties away, to a 10-bit mantissa (cvt.rna.tf32.f32)."""
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


def lin_fp32(a, w, b):
    return torch.nn.functional.linear(a, w, b)


def lin_3xtf32(a, w, b):
    ah, bh = tf32_rna(a), tf32_rna(w)
    al, bl = tf32_rna(a - ah), tf32_rna(w - bh)
    return (al @ bh.T + ah @ bl.T) + ah @ bh.T + b


def lin_fp16x3(a, w, b):
    f16 = lambda t: t.half().float()   # noqa: E731  round-to-nearest-even, subnormals kept
    wp = w * 256.0
    bh = f16(wp)
    bl = f16(wp - bh)
    bh_down = f16(bh * 2.0 ** -11)
    ah = f16(a)
    al_s = f16((a - ah) * 2048.0)
    d = (al_s @ bh_down.T + ah @ bl.T) + ah @ bh.T
    return d * (1.0 / 256.0) + b


def lin_fp16x3_plain(a, w, b):
    """as fp16x3 but with the low part of A left unscaled (it may go subnormal: absolute error <= 2^-25 per element),
    so B needs only its hi / lo images: D' = a_lo B'_hi + a_hi B'_lo + a_hi B'_hi"""
    f16 = lambda t: t.half().float()   # noqa: E731
    wp = w * 256.0
    bh = f16(wp)
    bl = f16(wp - bh)
    ah = f16(a)
    al = f16(a - ah)
    d = (al @ bh.T + ah @ bl.T) + ah @ bh.T
    return d * (1.0 / 256.0) + b


def sdf_with(lin, f, x):
    c = f.cfg
    ds, idx, w = f.compute_distance(x)
    h = torch.cat([positional_encoding(ds, c.multires_d),
                   positional_encoding(blend_rows(f.p["geometry_features"], idx, w), c.multires_fg)], dim=-1)
    hidden, (w_out, b_out) = f.geo_layers()
    for wl, bl in hidden:
        h = torch.nn.functional.softplus(lin(h, wl, bl), beta=100)
    return torch.nn.functional.linear(h, w_out, b_out)


def main():
    torch.set_num_threads(8)
    for dims in ((32, 32), (256, 256)):
        cfg = synth.ModelConfig(geometry_dim=dims[0], color_dim=dims[1])
        mesh = synth.icosphere_mesh(5, seed=0)
        sd = synth.make_state_dict(mesh, cfg, seed=1)
        f32 = helpers.oracle_field(mesh, cfg, sd)
        f64 = helpers.oracle_field(mesh, cfg, sd, torch.float64)
        x, _ = helpers.sample_points(20000, seed=4)
        truth = f64.forward_density_only(x.double())
        print(f"vertex codes {dims[0]}-d, 20 000 points, |sdf| max {truth.abs().max():.2f}")
        for name, lin in (("fp32", lin_fp32), ("3xTF32", lin_3xtf32), ("fp16x3", lin_fp16x3),
                          ("fp16x3p", lin_fp16x3_plain)):
            e = (sdf_with(lin, f32, x).double() - truth).abs()
            print(f"  {name:7s} max-abs {e.max():.3e}  mean {e.mean():.3e}  p99 {e.flatten().quantile(0.99):.3e}")


if __name__ == "__main__":
    main()
