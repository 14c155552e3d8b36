"""GPU diagnostic: run the backward pass with the exact-fp32 SIMT kernels and with the tensor-core
kernels on identical inputs and print the relative difference of every intermediate buffer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unet-phasegen_b200")]

import torch  # noqa: E402

import model  # noqa: E402
from phasegen import synth  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))


def main():
    C, B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 3, int(sys.argv[2]) if len(sys.argv) > 2 else 40
    prec = sys.argv[3] if len(sys.argv) > 3 else "bf16x3"
    torch.manual_seed(11)
    net = model.UNetModel(C, 2 * C).cuda()
    synth.randomize_norm_affine(net, seed=12)
    x = torch.log1p(torch.randn(B, C, T).abs() * 2.0).cuda()
    d_out = torch.randn(B, T, 2 * C, device="cuda") * 1e-3
    exs = {}
    for p in ("fp32_simt", prec):
        ex = net.train_executor(B, T, x.device, precision=p)
        ex.load_input_cf(x)
        dn, up = net._norm_params(x.device)
        ex.run(dn, up)
        ex.backward(dn, up, d_out=d_out)
        torch.cuda.synchronize()
        exs[p] = ex
    a, b = exs["fp32_simt"], exs[prec]
    D = a.D
    print(f"C={C} T={T} B={B} precision {prec} vs fp32_simt; lengths {a.Ld}")
    print("forward out", rel(b.out, a.out))
    for i in range(D):
        print(f"up{i}:  dz {rel(b.dz_up[i].as_float(), a.dz_up[i].as_float()):.2e}  dw {rel(b.dw_up[i], a.dw_up[i]):.2e}  "
              f"din {rel(b.din_up[i], a.din_up[i]):.2e}  dgamma {rel(b.dgb_up[i][0], a.dgb_up[i][0]):.2e}")
    for i in range(D - 1, -1, -1):
        din = f"{rel(b.din_dn[i], a.din_dn[i]):.2e}" if a.din_dn[i] is not None else "-"
        dg = f"{rel(b.dgb_dn[i][0], a.dgb_dn[i][0]):.2e}" if a.dgb_dn[i] else "-"
        print(f"dn{i}:  dz {rel(b.dz_dn[i].as_float(), a.dz_dn[i].as_float()):.2e}  dw {rel(b.dw_dn[i], a.dw_dn[i]):.2e}  din {din}  dgamma {dg}")


if __name__ == "__main__":
    main()
