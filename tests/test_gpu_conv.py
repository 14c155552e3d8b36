"""GPU parity of the convolution kernels: exact-fp32 SIMT kernel vs torch CPU (fp64), tensor-core
kernel vs the same, for the eight layer geometries of model.py:27-34."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GEOM = {"d1": (0, 32, 2, 16, 1, 2), "d2": (0, 8, 1, 2, 2, 2), "d3": (0, 8, 2, 1, 2, 2), "d4": (0, 4, 2, 1, 2, 4),
        "u4": (1, 5, 2, 1, 4, 2), "u3": (1, 8, 2, 1, 4, 2), "u2": (1, 8, 1, 2, 4, 2), "u1": (1, 32, 2, 16, 4, 2)}
LENS = {"d1": 136, "d2": 69, "d3": 66, "d4": 31, "u4": 15, "u3": 31, "u2": 66, "u1": 69}


def _ref(kind, x_cl, w, s, p):
    x = x_cl.double().permute(0, 2, 1).cpu()
    y = (F.conv_transpose1d if kind else F.conv1d)(x, w.double().cpu(), None, s, p)
    return y.permute(0, 2, 1).contiguous()


def _case(layer, C, L_in, B, seed=0):
    kind, k, s, p, cim, com = GEOM[layer]
    C_in, C_out = C * cim, C * com
    g = torch.Generator().manual_seed(seed)
    rows = (L_in + 7) // 8 * 8
    x = torch.zeros(B, rows, C_in)
    x[:, :L_in] = torch.randn(B, L_in, C_in, generator=g)
    w = torch.randn((C_in, C_out, k) if kind else (C_out, C_in, k), generator=g) / (C_in * k) ** 0.5
    return kind, k, s, p, C_in, C_out, rows, x.cuda(), w.cuda()


@pytest.mark.parametrize("layer", list(GEOM))
@pytest.mark.parametrize("C,scale", [(8, 1), (16, 3)])
def test_simt_conv_matches_torch(layer, C, scale):
    from phasegen import ops
    L_in, B = LENS[layer] * scale + (scale - 1), 2
    kind, k, s, p, C_in, C_out, rows, x, w = _case(layer, C, L_in, B)
    d = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, ops.PG_PREC_FP32_SIMT)
    _, _, ws = ops.pack_weight(w, kind, want_tc=False, want_simt=True)
    y = torch.empty(B, d.L_out, C_out, device="cuda")
    ops.conv_simt(d, x, ws, y)
    ref = _ref(kind, x[:, :L_in], w, s, p)
    assert y.shape == ref.shape
    assert float((y.cpu().double() - ref).norm() / ref.norm()) < 2e-6


@pytest.mark.parametrize("layer", list(GEOM))
@pytest.mark.parametrize("prec,tol", [("bf16x3", 3e-5), ("bf16", 1.5e-2), ("f16x3", 3e-5), ("f16x2", 5e-4), ("f16", 8e-4)])
def test_tc_conv_matches_torch(layer, prec, tol):
    from phasegen import ops
    from phasegen._lib import PRECISIONS
    C, B = 64, 3
    L_in = LENS[layer]
    kind, k, s, p, C_in, C_out, rows, x, w = _case(layer, C, L_in, B, seed=1)
    pdt = torch.float16 if prec.startswith("f16") else torch.bfloat16
    d = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, PRECISIONS[prec], taps_per_group=1)
    hi, lo, _ = ops.pack_weight(w, kind, plane_dtype=pdt)
    xh = x.to(pdt); xl = (x - xh.float()).to(pdt)
    y = torch.full((B, d.L_out, C_out), float("nan"), device="cuda")
    P = ops.conv_stat_parts(d)
    st = torch.zeros(B, P, C_out, 4, device="cuda")
    ops.conv_tc(d, xh, xl if prec not in ("bf16", "f16") else None, hi, lo if prec.endswith("x3") else None, y, st)
    torch.cuda.synchronize()
    ref = _ref(kind, x[:, :L_in], w, s, p)
    assert not torch.isnan(y).any()
    assert float((y.cpu().double() - ref).norm() / ref.norm()) < tol
    # statistics records: counts add up, combined moments match the tensor's
    n = st[..., 0].sum(1)
    assert bool((n == d.L_out).all())
    mean = (st[..., 0] * st[..., 1]).sum(1) / n
    m2 = (st[..., 2] + st[..., 0] * (st[..., 1] - mean[:, None]) ** 2).sum(1)
    yd = y.double()
    assert float((mean.double() - yd.mean(1)).abs().max()) < 1e-5
    assert float(((m2 / n).double() - yd.var(1, unbiased=False)).abs().max() / yd.var(1, unbiased=False).max()) < 1e-4


def test_tc_conv_full_width_tiles_and_persistence():
    """BASELINE-shape time axis (two 176-wide position tiles, both output phases) and more tiles
    than SMs, so every CTA loops over several tiles and both TMEM accumulators are used."""
    from phasegen import ops
    for layer, C, L_in, B in (("u1", 64, 349, 40), ("d1", 64, 696, 40), ("d2", 128, 349, 24)):
        kind, k, s, p, C_in, C_out, rows, x, w = _case(layer, C, L_in, B, seed=2)
        d = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, ops.PG_PREC_BF16X3, taps_per_group=1)
        ds = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, ops.PG_PREC_FP32_SIMT)
        hi, lo, ws = ops.pack_weight(w, kind, True, True)
        xh = x.to(torch.bfloat16); xl = (x - xh.float()).to(torch.bfloat16)
        y = torch.full((B, d.L_out, C_out), float("nan"), device="cuda")
        ys = torch.empty_like(y)
        ops.conv_tc(d, xh, xl, hi, lo, y, None)
        ops.conv_simt(ds, x, ws, ys)
        torch.cuda.synchronize()
        assert not torch.isnan(y).any(), layer
        assert float((y - ys).norm() / ys.norm()) < 3e-5, layer


def test_tc_conv_clip_groups_and_clip_bundles(monkeypatch):
    """Tile-order groups (forced small through the PG_TC_CLIP_GROUP test hook) and several clips
    per tile (short time axes), with a batch that is not a multiple of the bundle size."""
    from phasegen import ops
    monkeypatch.setenv("PG_TC_CLIP_GROUP", "2")
    for layer, C, L_in, B in (("d4", 128, 171, 7), ("u4", 128, 85, 7), ("d2", 64, 69, 11), ("u1", 64, 69, 5), ("d3", 64, 30, 19)):
        kind, k, s, p, C_in, C_out, rows, x, w = _case(layer, C, L_in, B, seed=3)
        d = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, ops.PG_PREC_BF16X3, taps_per_group=16)
        ds = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, ops.PG_PREC_FP32_SIMT)
        hi, lo, ws = ops.pack_weight(w, kind, True, True)
        xh = x.to(torch.bfloat16); xl = (x - xh.float()).to(torch.bfloat16)
        y = torch.full((B, d.L_out, C_out), float("nan"), device="cuda")
        ys = torch.empty_like(y)
        P = ops.conv_stat_parts(d)
        st = torch.zeros(B, P, C_out, 4, device="cuda")
        ops.conv_tc(d, xh, xl, hi, lo, y, st)
        ops.conv_simt(ds, x, ws, ys)
        torch.cuda.synchronize()
        assert not torch.isnan(y).any(), layer
        assert float((y - ys).norm() / ys.norm()) < 3e-5, layer
        assert bool((st[..., 0].sum(1) == d.L_out).all()), layer


@pytest.mark.parametrize("pair", [1, 2])
@pytest.mark.parametrize("layer,C,L_in,B,tpg", [
    ("d1", 128, 136, 3, 16), ("d2", 128, 69, 3, 16), ("d3", 128, 66, 5, 16), ("d4", 64, 31, 9, 16),
    ("u4", 128, 15, 9, 16), ("u3", 128, 31, 5, 16), ("u2", 128, 66, 3, 1), ("u1", 128, 69, 3, 16),
    ("u1", 128, 349, 40, 16), ("d1", 128, 696, 21, 16), ("d2", 128, 349, 24, 16), ("u4", 128, 85, 7, 16)])
def test_tc_conv_cta_pairs(layer, C, L_in, B, tpg, pair):
    """cta_group::2 tiles (pair = 2: required) against single-CTA tiles (pair = 1) and the exact SIMT
    kernel: every layer geometry, multi-tile time axes, clip bundles, persistent loops, statistics."""
    from phasegen import ops
    kind, k, s, p, C_in, C_out, rows, x, w = _case(layer, C, L_in, B, seed=4)
    ds = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, ops.PG_PREC_FP32_SIMT)
    hi, lo, ws = ops.pack_weight(w, kind, True, True)
    xh = x.to(torch.bfloat16); xl = (x - xh.float()).to(torch.bfloat16)
    ys = torch.empty(B, ds.L_out, C_out, device="cuda")
    ops.conv_simt(ds, x, ws, ys)
    for prec, tol in ((ops.PG_PREC_BF16X3, 3e-5), (ops.PG_PREC_BF16, 1.5e-2)):
        three = prec == ops.PG_PREC_BF16X3
        d = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, prec, taps_per_group=tpg, cta_pair=pair)
        y = torch.full((B, d.L_out, C_out), float("nan"), device="cuda")
        st = torch.zeros(B, ops.conv_stat_parts(d), C_out, 4, device="cuda")
        ops.conv_tc(d, xh, xl if three else None, hi, lo if three else None, y, st)
        torch.cuda.synchronize()
        assert not torch.isnan(y).any(), (layer, prec)
        assert float((y - ys).norm() / ys.norm()) < tol, (layer, prec)
        assert bool((st[..., 0].sum(1) == d.L_out).all()), (layer, prec)
        n = st[..., 0].sum(1)
        mean = (st[..., 0] * st[..., 1]).sum(1) / n
        assert float((mean - y.mean(1)).abs().max()) < 1e-4, (layer, prec)


def test_tc_conv_pair_needs_256_channels():
    from phasegen import ops
    x = torch.zeros(1, 32, 64, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(8, 128, 64, device="cuda", dtype=torch.bfloat16)
    y = torch.zeros(1, 29, 128, device="cuda")
    d = ops.conv_desc(0, 1, 64, 128, 32, 8, 1, 2, 32, 64, ops.PG_PREC_BF16X3, cta_pair=2)
    with pytest.raises(RuntimeError, match="256"):
        ops.conv_tc(d, x, x, w, w, y, None)


def test_conv_argument_errors():
    from phasegen import ops
    x = torch.zeros(1, 32, 48, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(8, 128, 48, device="cuda", dtype=torch.bfloat16)
    y = torch.zeros(1, 29, 128, device="cuda")
    d = ops.conv_desc(0, 1, 48, 128, 32, 8, 1, 2, 32, 48, ops.PG_PREC_BF16X3)
    with pytest.raises(RuntimeError, match="C_in"):
        ops.conv_tc(d, x, x, w, w, y, None)
    d = ops.conv_desc(0, 1, 64, 128, 32, 8, 3, 2, 32, 64, ops.PG_PREC_BF16X3)
    with pytest.raises(RuntimeError, match="stride"):
        ops.conv_tc(d, x, x, w, w, y, None)
