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


@pytest.mark.parametrize("prec", ["bf16x3", "f16x3", "f16x2"])
@pytest.mark.parametrize("layer,C,L_in,B,pair", [
    ("d1", 128, 696, 5, 0), ("d4", 128, 171, 7, 0), ("d1", 64, 136, 3, 1),                      # no norm: PG_EPI_ACT
    ("d2", 128, 349, 5, 0), ("d3", 128, 346, 5, 0), ("u4", 128, 171, 5, 0), ("u3", 128, 171, 5, 0), ("u2", 128, 346, 5, 0),
    ("u2", 64, 346, 3, 1), ("u3", 64, 171, 6, 1), ("d3", 128, 66, 9, 0), ("u4", 128, 135, 9, 0), ("d2", 128, 69, 7, 0),
    ("u3", 128, 121, 5, 0)])                                    # two phases x 128 columns: two clips per whole-clip tile (f16x2)
def test_tc_conv_fused_epilogues(layer, C, L_in, B, pair, prec):
    """The fused epilogues (PG_EPI_ACT; PG_EPI_NORM_ACT with whole-clip tiles: two position tiles, two output phases,
    merged short clips, CTA pairs and single CTAs) against the two-pass form built from the exact SIMT convolution:
    per-clip statistics, affine, LeakyReLU / ReLU fan-out, written at a channel offset into wider operand buffers."""
    from phasegen import ops
    from phasegen._lib import PG_DT_BF16_SPLIT, PG_DT_F16_SPLIT, PG_EPI_ACT, PG_EPI_NORM_ACT, PRECISIONS
    kind, k, s, p, C_in, C_out, rows, x, w = _case(layer, C, L_in, B, seed=5)
    has_norm = layer not in ("d1", "d4")
    mode = PG_EPI_NORM_ACT if has_norm else PG_EPI_ACT
    f16 = prec.startswith("f16")
    pdt = torch.float16 if f16 else torch.bfloat16
    d = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, PRECISIONS[prec], taps_per_group=16, cta_pair=pair)
    if not ops.conv_epilogue_supported(d, mode):
        # only case: single-CTA tiles (128 output channels) load full-width strips; two of them x two planes x two slots
        # do not fit shared memory, so the executor keeps the two-pass form for that layer
        assert (layer, pair, L_in) == ("u2", 1, 346) and prec != "bf16"
        return
    ds = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, ops.PG_PREC_FP32_SIMT)
    hi, lo, ws = ops.pack_weight(w, kind, True, True, plane_dtype=pdt)
    xh = x.to(pdt); xl = (x - xh.float()).to(pdt)
    ys = torch.empty(B, ds.L_out, C_out, device="cuda")
    ops.conv_simt(ds, x, ws, ys)
    g = torch.Generator().manual_seed(6)
    gamma = (1.0 + 0.2 * torch.randn(C_out, generator=g)).cuda()
    beta = (0.3 * torch.randn(C_out, generator=g)).cuda()
    h = ys.double()
    if has_norm:
        mean = h.mean(1, keepdim=True); var = h.var(1, unbiased=False, keepdim=True)
        h = (h - mean) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double()
    L_out, rows_o = ds.L_out, (ds.L_out + 7) // 8 * 8
    wide, off = 2 * C_out, C_out                                  # second destination: right half of a concat buffer
    flag = torch.zeros(1, device="cuda", dtype=torch.int32)
    dt = PG_DT_F16_SPLIT if f16 else PG_DT_BF16_SPLIT
    a_hi = torch.zeros(B, rows_o, C_out, device="cuda", dtype=pdt); a_lo = torch.zeros_like(a_hi)
    c_hi = torch.zeros(B, rows_o, wide, device="cuda", dtype=pdt); c_lo = torch.zeros_like(c_hi)
    epi = ops.conv_epilogue(mode, ops.act_dst(a_hi, a_lo, rows_o * C_out, C_out, 0, dt, 0.2, flag),
                            ops.act_dst(c_hi, c_lo, rows_o * wide, wide, off, dt, 0.0, flag),
                            gamma if has_norm else None, beta if has_norm else None, 1e-5)
    ops.conv_tc(d, xh, xl, hi, lo if prec.endswith("x3") else None, None, None, epi)
    torch.cuda.synchronize()
    tol = 2e-3 if prec == "f16x2" else 2e-4
    got_a = (a_hi.double() + a_lo.double())[:, :L_out]
    got_c = (c_hi.double() + c_lo.double())[:, :L_out, off:]
    ref_a = torch.where(h > 0, h, 0.2 * h); ref_c = torch.clamp(h, min=0)
    assert float((got_a - ref_a).norm() / ref_a.norm()) < tol, (layer, "leaky destination")
    assert float((got_c - ref_c).norm() / ref_c.norm()) < tol, (layer, "relu destination at a channel offset")
    assert float(c_hi[:, :, :off].abs().max()) == 0 and float(a_hi[:, L_out:].abs().max()) == 0   # nothing written outside
    assert int(flag.item()) == 0


def test_fused_norm_epilogue_needs_a_whole_clip_per_tile():
    """u1 at the BASELINE time axis has 4 x 176 = 704 accumulator columns per clip: no fused norm (the ISTFT applies it)."""
    from phasegen import ops
    from phasegen._lib import PG_EPI_ACT, PG_EPI_NORM_ACT
    kind, k, s, p, cim, com = GEOM["u1"]
    d = ops.conv_desc(kind, 4, 128 * cim, 128 * com, 349, k, s, p, 352, 128 * cim, ops.PG_PREC_F16X2)
    assert not ops.conv_epilogue_supported(d, PG_EPI_NORM_ACT) and ops.conv_epilogue_supported(d, PG_EPI_ACT)
    x = torch.zeros(4, 352, 128 * cim, device="cuda", dtype=torch.float16)
    w = torch.zeros(k, 128 * com, 128 * cim, device="cuda", dtype=torch.float16)
    o = torch.zeros(4, 696, 128 * com, device="cuda", dtype=torch.float16)
    epi = ops.conv_epilogue(PG_EPI_NORM_ACT, ops.act_dst(o, o, 696 * 128 * com, 128 * com, 0, 5, 0.0))
    with pytest.raises(RuntimeError, match="whole clip|accumulator columns"):
        ops.conv_tc(d, x, x, w, None, None, None, epi)


def test_fp16_range_guard_fires_in_every_writer():
    """Values beyond 65504 written into fp16 operand planes set the sticky flag (bn_act, fused conv epilogue, input
    transpose, weight cast); the same values into bf16 planes do not."""
    from phasegen import ops
    from phasegen._lib import PG_DT_BF16_SPLIT, PG_DT_F16_SPLIT, PG_EPI_ACT
    B, L, Cn = 2, 16, 64
    y = torch.randn(B, L, Cn, device="cuda"); y[1, 3, 5] = 1.0e5
    for dt, pdt, want in ((PG_DT_F16_SPLIT, torch.float16, 1), (PG_DT_BF16_SPLIT, torch.bfloat16, 0)):
        flag = torch.zeros(1, device="cuda", dtype=torch.int32)
        hi = torch.zeros(B, L, Cn, device="cuda", dtype=pdt); lo = torch.zeros_like(hi)
        ops.bn_act(y, B, L, Cn, L, Cn, None, False, ops.act_dst(hi, lo, L * Cn, Cn, 0, dt, 1.0, flag))
        assert int(flag.item()) == want
        flag.zero_()
        ops.transpose(y.permute(0, 2, 1).contiguous(), dst_hi=hi, dst_lo=lo, dst_batch_stride=L * Cn, dst_ld=Cn, range_flag=flag)
        assert int(flag.item()) == want
        flag.zero_()
        ops.cast_split(y, hi, lo, flag)
        assert int(flag.item()) == want
    # fused epilogue: a convolution whose output exceeds the fp16 range
    kind, k, s, p, C_in, C_out, rows, x, w = _case("d4", 64, 31, 2, seed=7)
    w = w * 3.0e4
    hi, lo, _ = ops.pack_weight(w, kind, plane_dtype=torch.float16)
    xh = x.half(); xl = (x - xh.float()).half()
    d = ops.conv_desc(kind, 2, C_in, C_out, 31, k, s, p, rows, C_in, ops.PG_PREC_F16X3)
    flag = torch.zeros(1, device="cuda", dtype=torch.int32)
    o_hi = torch.zeros(2, 16, C_out, device="cuda", dtype=torch.float16); o_lo = torch.zeros_like(o_hi)
    ops.conv_tc(d, xh, xl, hi, lo, None, None,
                ops.conv_epilogue(PG_EPI_ACT, ops.act_dst(o_hi, o_lo, 16 * C_out, C_out, 0, PG_DT_F16_SPLIT, 0.0, flag)))
    assert int(flag.item()) == 1


def test_kernels_write_only_inside_their_outputs():
    """Own bounds check (compute-sanitizer is closed on this pool, profiles/r02_sanitizer_closed.log): every output of the
    tensor-core / STFT / ISTFT kernels sits inside a larger sentinel-filled allocation; after the launch the guard regions
    before and after the output, and the padding rows inside it, must be untouched."""
    from phasegen import ops
    from phasegen._lib import PG_SPEC_CARTESIAN, PG_STFT_REIM
    SENT = 12345.0

    def guarded(shape, dtype=torch.float32, pad=4096):
        n = int(np.prod(shape))
        buf = torch.full((n + 2 * pad,), SENT, device="cuda", dtype=dtype)
        return buf, buf[pad:pad + n].view(*shape)

    def intact(buf, n, pad=4096):
        return bool((buf[:pad] == SENT).all()) and bool((buf[pad + n:] == SENT).all())

    for layer, C, L_in, B, pair in (("u1", 128, 69, 3, 0), ("d2", 64, 349, 5, 1), ("u3", 128, 31, 5, 0)):
        kind, k, s, p, C_in, C_out, rows, x, w = _case(layer, C, L_in, B, seed=9)
        d = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, ops.PG_PREC_BF16X3, taps_per_group=16, cta_pair=pair)
        hi, lo, _ = ops.pack_weight(w, kind)
        xh = x.to(torch.bfloat16); xl = (x - xh.float()).to(torch.bfloat16)
        ybuf, y = guarded((B, d.L_out, C_out))
        P = ops.conv_stat_parts(d)
        sbuf, st = guarded((B, P, C_out, 4))
        ops.conv_tc(d, xh, xl, hi, lo, y, st)
        torch.cuda.synchronize()
        assert intact(ybuf, y.numel()) and intact(sbuf, st.numel()), layer
        assert not bool((y == SENT).any()), layer                     # ... and every element inside was written
        # weight gradient of the same layer
        grows = (d.L_out + 7) // 8 * 8
        g = torch.zeros(B, grows, C_out, device="cuda"); g[:, :d.L_out] = torch.randn(B, d.L_out, C_out, device="cuda")
        gh = g.bfloat16(); gl = (g - gh.float()).bfloat16()
        wbuf, dw = guarded((k, C_out, C_in))
        ops.wgrad_tc(d, xh, xl, gh, gl, grows, dw)
        torch.cuda.synchronize()
        assert intact(wbuf, dw.numel()) and not bool((dw == SENT).any()), layer
    # STFT / ISTFT (library-allocated outputs are exact-size torch tensors: check through the C ABI with guarded buffers)
    import ctypes as C_
    from phasegen import _lib
    n_fft, hop, T, B = 512, 128, 24, 3
    N = (T - 1) * hop
    wv = torch.randn(B, N, device="cuda") * 0.1
    abuf, a = guarded((B, T, n_fft // 2)); bbuf, b = guarded((B, T, n_fft // 2))
    _lib.call("pg_stft", ops._ptr(wv), B, N, n_fft, hop, ops._ptr(ops.twiddle(n_fft, wv.device)), PG_STFT_REIM, ops._ptr(a), ops._ptr(b),
              None, None, 0, 0, ops._stream())
    obuf, out = guarded((B, N))
    pk = torch.zeros(B, device="cuda"); bad = torch.zeros(B, device="cuda", dtype=torch.int32)
    _lib.call("pg_istft", ops._ptr(a), ops._ptr(b), PG_SPEC_CARTESIAN, B, T, n_fft, hop, ops._ptr(ops.twiddle(n_fft, wv.device)),
              ops._ptr(out), ops._ptr(pk), ops._ptr(bad), None, 0, ops._stream())
    torch.cuda.synchronize()
    assert intact(abuf, a.numel()) and intact(bbuf, b.numel()) and intact(obuf, out.numel())
    assert not bool((a == SENT).any()) and not bool((out == SENT).any())
