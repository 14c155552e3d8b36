"""GPU probe: tensor-core convolution vs the exact-fp32 SIMT convolution, one configuration per
subprocess (a trapped kernel kills the CUDA context).  Writes a table to stdout.

    python tools/tc_probe.py            # all geometries x strip modes
    python tools/tc_probe.py one <layer> <tpg> <bo> <prec> <C> <T> <B>
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unet-phasegen_b200")]

GEOM = {  # name: (kind, k, s, p, cin_mult, cout_mult)   -- model.py:27-34
    "d1": (0, 32, 2, 16, 1, 2), "d2": (0, 8, 1, 2, 2, 2), "d3": (0, 8, 2, 1, 2, 2), "d4": (0, 4, 2, 1, 2, 4),
    "u4": (1, 5, 2, 1, 4, 2), "u3": (1, 8, 2, 1, 4, 2), "u2": (1, 8, 1, 2, 4, 2), "u1": (1, 32, 2, 16, 4, 2),
}


def one(layer, tpg, bo, prec, C, L_in, B):
    import torch
    from phasegen import ops
    from phasegen._lib import PRECISIONS
    kind, k, s, p, cim, com = GEOM[layer]
    C_in, C_out = C * cim, C * com
    torch.manual_seed(0)
    dev = "cuda"
    rows = (L_in + 7) // 8 * 8
    x = torch.zeros(B, rows, C_in, device=dev)
    x[:, :L_in] = torch.randn(B, L_in, C_in, device=dev)
    w = torch.randn((C_in, C_out, k) if kind else (C_out, C_in, k), device=dev) / (C_in * k) ** 0.5
    hi, lo, _ = ops.pack_weight(w, kind, True, False)
    _, _, ws = ops.pack_weight(w, kind, False, True)
    xh = x.to(torch.bfloat16); xl = (x - xh.float()).to(torch.bfloat16)
    d_tc = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, PRECISIONS[prec], taps_per_group=tpg, base_offset_mode=bo)
    d_si = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, 0)
    L_out = d_tc.L_out
    y_tc = torch.full((B, L_out, C_out), float("nan"), device=dev)
    y_si = torch.empty(B, L_out, C_out, device=dev)
    P = ops.conv_stat_parts(d_tc)
    st = torch.zeros(B, P, C_out, 4, device=dev)
    ops.conv_simt(d_si, x, ws, y_si)
    ops.conv_tc(d_tc, xh, xl if prec == "bf16x3" else None, hi, lo if prec == "bf16x3" else None, y_tc, st)
    torch.cuda.synchronize()
    err = ((y_tc - y_si).norm() / y_si.norm()).item()
    nan = int(torch.isnan(y_tc).sum().item())
    # stats check: combine partial records and compare with the direct per-(clip, channel) moments
    n = st[..., 0].sum(1); mean = (st[..., 0] * st[..., 1]).sum(1) / n
    m2 = (st[..., 2] + st[..., 0] * (st[..., 1] - mean[:, None]) ** 2).sum(1)
    mref = y_si.mean(1); vref = y_si.var(1, unbiased=False)
    e_mean = ((mean - mref).abs().max() / y_si.abs().max()).item()
    e_var = ((m2 / n - vref).abs().max() / vref.max()).item()
    print(f"RESULT {layer} tpg={tpg} bo={bo} {prec} C={C} L_in={L_in} L_out={L_out} B={B}: rel_l2={err:.3e} nan={nan} "
          f"n_ok={bool((n == L_out).all())} mean_err={e_mean:.2e} var_err={e_var:.2e}")


def wgrad(layer, prec, C, L_in, B):
    """tensor-core weight gradient (MN-major operands) vs the exact SIMT one."""
    import torch
    from phasegen import ops
    from phasegen._lib import PRECISIONS
    kind, k, s, p, cim, com = GEOM[layer]
    C_in, C_out = C * cim, C * com
    torch.manual_seed(0)
    dev = "cuda"
    rows = (L_in + 7) // 8 * 8
    x = torch.zeros(B, rows, C_in, device=dev)
    x[:, :L_in] = torch.randn(B, L_in, C_in, device=dev)
    d = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, PRECISIONS[prec])
    L_out = d.L_out
    grows = (L_out + 7) // 8 * 8
    g = torch.zeros(B, grows, C_out, device=dev)
    g[:, :L_out] = torch.randn(B, L_out, C_out, device=dev)
    xh = x.to(torch.bfloat16); xl = (x - xh.float()).to(torch.bfloat16)
    gh = g.to(torch.bfloat16); gl = (g - gh.float()).to(torch.bfloat16)
    three = prec == "bf16x3"
    dw_tc = torch.full((k, C_out, C_in), float("nan"), device=dev)
    dw_si = torch.empty(k, C_out, C_in, device=dev)
    ds = ops.conv_desc(kind, B, C_in, C_out, L_in, k, s, p, rows, C_in, 0)
    ops.wgrad_simt(ds, x, g, grows, dw_si)
    ops.wgrad_tc(d, xh, xl if three else None, gh, gl if three else None, grows, dw_tc)
    torch.cuda.synchronize()
    err = ((dw_tc - dw_si).norm() / dw_si.norm()).item()
    # independent check of the SIMT result through torch autograd on the GPU (fp64)
    import torch.nn.functional as F
    xc = x[:, :L_in].double().permute(0, 2, 1)
    w = torch.zeros((C_in, C_out, k) if kind else (C_out, C_in, k), device=dev, dtype=torch.float64, requires_grad=True)
    y = (F.conv_transpose1d if kind else F.conv1d)(xc, w, None, s, p)
    (y * g[:, :L_out].double().permute(0, 2, 1)).sum().backward()
    ref = w.grad.permute(2, 1, 0) if kind else w.grad.permute(2, 0, 1)
    e_si = ((dw_si.double() - ref).norm() / ref.norm()).item()
    print(f"RESULT wgrad {layer} {prec} C={C} L_in={L_in} L_out={L_out} B={B}: tc_vs_simt={err:.3e} "
          f"nan={int(torch.isnan(dw_tc).sum())} simt_vs_autograd={e_si:.3e}")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        a = sys.argv[2:]
        one(a[0], int(a[1]), int(a[2]), a[3], int(a[4]), int(a[5]), int(a[6]))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "wgrad1":
        a = sys.argv[2:]
        wgrad(a[0], a[1], int(a[2]), int(a[3]), int(a[4]))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "wgrad":
        lens = {"d1": 136, "d2": 69, "d3": 66, "d4": 31, "u4": 15, "u3": 31, "u2": 66, "u1": 69}
        cfgs = [(l, "bf16x3", 64, lens[l], 3) for l in GEOM] + [("d1", "bf16", 64, 136, 3), ("u1", "bf16x3", 128, 65, 5),
                                                                   ("d2", "bf16x3", 128, 349, 2), ("d4", "bf16", 128, 29, 4)]
        for c in cfgs:
            cmd = [sys.executable, os.path.abspath(__file__), "wgrad1"] + [str(v) for v in c]
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
                lines = [l for l in (r.stdout + r.stderr).splitlines() if l.startswith("RESULT") or "timeout" in l or "rror" in l]
                print("\n".join(lines[:6]) if lines else f"NO OUTPUT {c} rc={r.returncode} {(r.stderr or '')[-300:]}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"TIMEOUT {c}", flush=True)
        return
    # reference lengths for T=136 (C=64): d1 136->69, d2 69->66, d3 66->31, d4 31->15, u4 15->31, u3 31->66, u2 66->69, u1 69->136
    lens = {"d1": 136, "d2": 69, "d3": 66, "d4": 31, "u4": 15, "u3": 31, "u2": 66, "u1": 69}
    configs = []
    for layer in GEOM:
        configs.append((layer, 1, 0, "bf16x3", 64, lens[layer], 2))
    for layer in ("d1", "u1", "d2", "u4"):
        configs.append((layer, 16, 0, "bf16x3", 64, lens[layer], 2))
        configs.append((layer, 16, 1, "bf16x3", 64, lens[layer], 2))
    configs.append(("d1", 1, 0, "bf16", 64, 136, 2))
    # a long time axis (two position tiles, N = 176) and many tiles per CTA (persistence, phases)
    configs.append(("u1", 1, 0, "bf16x3", 64, 349, 3))
    configs.append(("d1", 1, 0, "bf16x3", 64, 696, 3))
    configs.append(("d2", 1, 0, "bf16x3", 128, 349, 40))
    for c in configs:
        cmd = [sys.executable, os.path.abspath(__file__), "one"] + [str(v) for v in c]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
            lines = [l for l in (r.stdout + r.stderr).splitlines() if l.startswith("RESULT") or "timeout" in l or "rror" in l]
            print("\n".join(lines[:6]) if lines else f"NO OUTPUT {c} rc={r.returncode}", flush=True)
            if r.returncode != 0 and not any(l.startswith("RESULT") for l in lines):
                print(f"FAILED {c} rc={r.returncode}: {(r.stderr or '')[-400:]}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"TIMEOUT {c}", flush=True)


if __name__ == "__main__":
    main()
