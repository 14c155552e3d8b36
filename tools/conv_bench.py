"""Per-layer timing of the tensor-core convolution at the bench shape (CUDA events, warm).

    python tools/conv_bench.py [--B 256] [--C 512] [--T 696] [--prec bf16x3] [--reps 5] [--nb 0] [--tpg 16]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unet-phasegen_b200")]

import torch  # noqa: E402

from phasegen import ops  # noqa: E402
from phasegen._lib import PRECISIONS  # noqa: E402

GEOM = {"d1": (0, 32, 2, 16, 1, 2), "d2": (0, 8, 1, 2, 2, 2), "d3": (0, 8, 2, 1, 2, 2), "d4": (0, 4, 2, 1, 2, 4),
        "u4": (1, 5, 2, 1, 4, 2), "u3": (1, 8, 2, 1, 4, 2), "u2": (1, 8, 1, 2, 4, 2), "u1": (1, 32, 2, 16, 4, 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--C", type=int, default=512)
    ap.add_argument("--T", type=int, default=696)
    ap.add_argument("--prec", default="bf16x3")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--nb", type=int, default=0)
    ap.add_argument("--tpg", type=int, default=16)
    ap.add_argument("--layers", default="d1,d2,d3,d4,u4,u3,u2,u1")
    ap.add_argument("--phase-only", type=int, default=1)
    ap.add_argument("--epi", type=int, default=0, help="1 = the fused epilogue the inference executor would pick for the layer "
                    "(ACT for d1/d4, per-clip NORM_ACT for the norm layers where a whole clip fits a tile, raw for u1)")
    a = ap.parse_args()
    T = a.T
    L1 = T // 2 + 1; L2 = L1 - 3; L3 = L2 // 2 - 2; L4 = (L3 - 1) // 2
    lin = {"d1": T, "d2": L1, "d3": L2, "d4": L3, "u4": L4, "u3": L3, "u2": L2, "u1": L1}
    tot_ms = tot_fl = 0.0
    mix = {"f16mix": {"d1": "f16x2", "u1": "f16x2", "u2": "f16x2"}, "f16mix1": {"d1": "f16x2", "u1": "f16", "u2": "f16x2"}}
    prec_arg = a.prec
    for name in a.layers.split(","):
        a.prec = mix[prec_arg].get(name, "f16x3") if prec_arg in mix else prec_arg
        kind, k, s, p, cim, com = GEOM[name]
        C_in, C_out = a.C * cim, a.C * com
        if name == "u1" and a.phase_only:
            C_out //= 2
        L_in = lin[name]
        rows = (L_in + 7) // 8 * 8
        pdt = torch.float16 if a.prec.startswith("f16") else torch.bfloat16
        xf = torch.zeros(a.B, rows, C_in, device="cuda")
        xf[:, :L_in] = torch.randn(a.B, L_in, C_in, device="cuda").clamp_min(0)     # post-ReLU-like activations
        x = xf.to(pdt)
        xl = (xf - x.float()).to(pdt)
        w = torch.randn((C_in, C_out, k) if kind else (C_out, C_in, k), device="cuda") / (C_in * k) ** 0.5
        hi, lo, _ = ops.pack_weight(w, kind, plane_dtype=pdt)
        d = ops.conv_desc(kind, a.B, C_in, C_out, L_in, k, s, p, rows, C_in, PRECISIONS[a.prec],
                          taps_per_group=a.tpg, max_clips_per_tile=a.nb)
        y = torch.empty(a.B, d.L_out, C_out, device="cuda")
        st = torch.empty(a.B, ops.conv_stat_parts(d), C_out, 4, device="cuda")
        terms = {"bf16x3": 3, "f16x3": 3, "f16x2": 2, "bf16": 1, "f16": 1}[a.prec]
        three = terms
        epi, tag = None, "raw"
        if a.epi and name != "u1":
            from phasegen._lib import PG_DT_BF16_SPLIT, PG_DT_F16_SPLIT, PG_EPI_ACT, PG_EPI_NORM_ACT
            mode = PG_EPI_ACT if name in ("d1", "d4") else PG_EPI_NORM_ACT
            if ops.conv_epilogue_supported(d, mode):
                rows_o = (d.L_out + 7) // 8 * 8
                dt = PG_DT_F16_SPLIT if pdt == torch.float16 else PG_DT_BF16_SPLIT
                o1 = torch.zeros(2, a.B, rows_o, C_out, device="cuda", dtype=pdt)
                o2 = torch.zeros(2, a.B, rows_o, 2 * C_out, device="cuda", dtype=pdt)
                two = name in ("d1", "d2", "d3")
                gam = torch.ones(C_out, device="cuda"); bet = torch.zeros(C_out, device="cuda")
                d0 = ops.act_dst(o1[0], o1[1], rows_o * C_out, C_out, 0, dt, 0.2) if name.startswith("d") else \
                    ops.act_dst(o2[0], o2[1], rows_o * 2 * C_out, 2 * C_out, C_out, dt, 0.0)
                d1_ = ops.act_dst(o2[0], o2[1], rows_o * 2 * C_out, 2 * C_out, 0, dt, 0.0) if two else None
                epi = ops.conv_epilogue(mode, d0, d1_, gam if mode == PG_EPI_NORM_ACT else None, bet if mode == PG_EPI_NORM_ACT else None)
                tag = "act" if mode == PG_EPI_ACT else "norm+act"
        run = lambda: ops.conv_tc(d, x, xl if terms >= 2 else None, hi, lo if terms == 3 else None,
                                  None if epi is not None else y, None if epi is not None else st, epi)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        L_macs = d.L_out if kind == 0 else L_in
        fl = 2.0 * C_in * C_out * k * L_macs * a.B
        tot_ms += ms; tot_fl += fl
        print(f"{name}: C_in={C_in:5d} C_out={C_out:5d} L_in={L_in:4d} L_out={d.L_out:4d} {ms:8.3f} ms  "
              f"{fl / ms / 1e9:7.1f} TFLOP/s algorithmic  x{three} = {fl * three / ms / 1e9:7.1f} tensor  [{a.prec} {tag}]", flush=True)
    print(f"total {tot_ms:.3f} ms  {tot_fl / tot_ms / 1e9:.1f} TFLOP/s algorithmic")


if __name__ == "__main__":
    main()
