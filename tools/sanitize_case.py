"""Small invocations of every hand-rolled-protocol kernel (mbarrier / cluster / TMEM: conv_tc single + pair, raw + fused
epilogues, merged tiles; wgrad_tc; stft / istft / stitch) for `compute-sanitizer --tool memcheck|racecheck|synccheck`.
Each result is also checked against the exact SIMT kernel / round trip, so a sanitizer run is a correctness run too."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unet-phasegen_b200")]

import torch  # noqa: E402

from phasegen import ops  # noqa: E402
from phasegen._lib import PG_DT_F16_SPLIT, PG_EPI_ACT, PG_EPI_NORM_ACT, PG_SPEC_CARTESIAN, PG_STFT_REIM  # noqa: E402

GEOM = {"d1": (0, 32, 2, 16, 1, 2), "d3": (0, 8, 2, 1, 2, 2), "u3": (1, 8, 2, 1, 4, 2), "u2": (1, 8, 1, 2, 4, 2)}


def conv_case(layer, C, L_in, B, pair, prec, fused):
    kind, k, s, p, cim, com = GEOM[layer]
    C_in, C_out = C * cim, C * com
    rows = (L_in + 7) // 8 * 8
    g = torch.Generator().manual_seed(1)
    x = torch.zeros(B, rows, C_in); x[:, :L_in] = torch.randn(B, L_in, C_in, generator=g)
    w = torch.randn((C_in, C_out, k) if kind else (C_out, C_in, k), generator=g) / (C_in * k) ** 0.5
    x, w = x.cuda(), w.cuda()
    f16 = prec in (ops.PG_PREC_F16X3, ops.PG_PREC_F16X2)
    pdt = torch.float16 if f16 else torch.bfloat16
    hi, lo, ws = ops.pack_weight(w, kind, True, True, plane_dtype=pdt)
    xh = x.to(pdt); xl = (x - xh.float()).to(pdt)
    d = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, prec, taps_per_group=16, cta_pair=pair)
    ds = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, ops.PG_PREC_FP32_SIMT)
    ys = torch.empty(B, ds.L_out, C_out, device="cuda")
    ops.conv_simt(ds, x, ws, ys)
    if not fused:
        y = torch.empty(B, d.L_out, C_out, device="cuda")
        st = torch.zeros(B, ops.conv_stat_parts(d), C_out, 4, device="cuda")
        ops.conv_tc(d, xh, xl, hi, lo, y, st)
        err = float((y - ys).norm() / ys.norm())
    else:
        mode = PG_EPI_ACT if layer == "d1" else PG_EPI_NORM_ACT
        assert ops.conv_epilogue_supported(d, mode), (layer, "fused epilogue not supported")
        rows_o = (d.L_out + 7) // 8 * 8
        o_hi = torch.zeros(B, rows_o, C_out, device="cuda", dtype=pdt); o_lo = torch.zeros_like(o_hi)
        flag = torch.zeros(1, device="cuda", dtype=torch.int32)
        ops.conv_tc(d, xh, xl, hi, lo, None, None, ops.conv_epilogue(mode, ops.act_dst(o_hi, o_lo, rows_o * C_out, C_out, 0, PG_DT_F16_SPLIT, 0.0, flag)))
        h = ys
        if mode == PG_EPI_NORM_ACT:
            h = (ys - ys.mean(1, keepdim=True)) / torch.sqrt(ys.var(1, unbiased=False, keepdim=True) + 1e-5)
        ref = torch.clamp(h, min=0)
        got = (o_hi.float() + o_lo.float())[:, :d.L_out]
        err = float((got - ref).norm() / ref.norm())
    torch.cuda.synchronize()
    print(f"conv {layer} C={C} L_in={L_in} B={B} pair={pair} prec={prec} fused={fused}: rel err {err:.2e}", flush=True)
    assert err < 2e-3


def wgrad_case(C, L_in, B):
    kind, k, s, p, cim, com = GEOM["u2"]
    C_in, C_out = C * cim, C * com
    rows = (L_in + 7) // 8 * 8
    g = torch.Generator().manual_seed(2)
    x = torch.zeros(B, rows, C_in); x[:, :L_in] = torch.randn(B, L_in, C_in, generator=g)
    d = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, ops.PG_PREC_BF16X3)
    grows = (d.L_out + 7) // 8 * 8
    gr = torch.zeros(B, grows, C_out); gr[:, :d.L_out] = torch.randn(B, d.L_out, C_out, generator=g)
    x, gr = x.cuda(), gr.cuda()
    xh = x.bfloat16(); xl = (x - xh.float()).bfloat16(); gh = gr.bfloat16(); gl = (gr - gh.float()).bfloat16()
    dw = torch.zeros(k, C_out, C_in, device="cuda"); dws = torch.zeros_like(dw)
    ops.wgrad_tc(d, xh, xl, gh, gl, grows, dw)
    ds = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, ops.PG_PREC_FP32_SIMT)
    ops.wgrad_simt(ds, x, gr, grows, dws)
    torch.cuda.synchronize()
    err = float((dw - dws).norm() / dws.norm())
    print(f"wgrad u2 C={C} L_in={L_in} B={B}: rel err {err:.2e}", flush=True)
    assert err < 1e-4


def stft_case():
    n_fft, hop, T, B = 1024, 256, 24, 2
    w = torch.randn(B, (T - 1) * hop, device="cuda") * 0.1
    re, im = ops.stft(w, n_fft, hop, PG_STFT_REIM)
    back, _ = ops.istft(re, im, PG_SPEC_CARTESIAN, n_fft, hop, normalize=False, check_finite=True)
    ref = torch.stft(w, n_fft, hop, window=torch.hann_window(n_fft, device="cuda"), return_complex=True)[:, 1:].permute(0, 2, 1)
    e1 = float((torch.complex(re, im) - ref).norm() / ref.norm())
    from phasegen import longform
    wins = torch.randn(5, 1000, device="cuda")
    y = longform.stitch(wins, list(range(5)), 5, 1000 + 4 * 504, 8, frames=126)   # (frames-1)*hop = 1000 samples per window, step 504
    torch.cuda.synchronize()
    print(f"stft rel err {e1:.2e}; istft(stft(x)) finite {bool(torch.isfinite(back).all())}; stitch finite {bool(torch.isfinite(y).all())}", flush=True)
    assert e1 < 1e-4


if __name__ == "__main__":
    torch.cuda.set_device(0)
    conv_case("d3", 128, 66, 3, pair=0, prec=ops.PG_PREC_BF16X3, fused=False)       # CTA pair, merged clips, raw + statistics
    conv_case("d3", 64, 66, 3, pair=1, prec=ops.PG_PREC_BF16X3, fused=False)        # single CTAs
    conv_case("d1", 128, 136, 3, pair=0, prec=ops.PG_PREC_F16X2, fused=True)        # fused activation epilogue, two clips per tile
    conv_case("u2", 128, 346, 2, pair=0, prec=ops.PG_PREC_F16X2, fused=True)        # whole-clip tile (two position tiles), fused norm
    conv_case("u3", 128, 171, 2, pair=0, prec=ops.PG_PREC_F16X3, fused=True)        # whole-clip tile (two output phases), fused norm
    wgrad_case(64, 30, 3)
    stft_case()
    print("SANITIZE CASES OK")
