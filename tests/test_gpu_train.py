"""GPU parity of the training step (train.py:37-62): loss, gradients of every parameter, Adam,
against the reference-generated golden vectors and autograd through the oracle."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import unet_torch  # noqa: E402

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "unet_*.npz")))


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def _train_py_loss(pred, x, phi, C):
    """train.py:45-60 verbatim in behaviour, with torch ops on the GPU (what a user's script does)."""
    lossf = torch.nn.MSELoss()
    pp, pm = pred[:, :C], pred[:, C:]
    ang = lossf(torch.cos(pp), phi.cos()) + lossf(torch.sin(pp), phi.sin())
    return ang + lossf(pm, x) * 0.2


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_backward_matches_reference_gradients(path):
    """loss.backward() through the drop-in model (exact-fp32 kernels at these channel counts) vs
    the gradients the real reference produced (tests/golden) and the oracle's for all parameters."""
    import model
    z = np.load(path)
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    C = z["x"].shape[1]
    net = model.UNetModel(C, 2 * C).cuda()
    net.model.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in sd.items()})
    x = torch.from_numpy(z["x"]).float().cuda(); phi = torch.from_numpy(z["phi"]).float().cuda()
    pred = net.forward(x)
    loss = _train_py_loss(pred, x, phi, C)
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) < 1e-4 * abs(float(z["loss"]))
    grads = {k: p.grad for k, p in net.model.named_parameters()}
    for k in z.files:
        if k.startswith("grad::"):
            assert rel_l2(grads[k[6:]].cpu().numpy(), z[k]) < 2e-4, k
    tgt = torch.stack([torch.from_numpy(z["x"]), torch.from_numpy(z["phi"])], 1)
    _, _, _, ref = unet_torch.loss_and_grads(sd, torch.from_numpy(z["x"]), tgt)
    for k, g in ref.items():
        assert grads[k] is not None, k
        assert rel_l2(grads[k].cpu().numpy(), g.numpy()) < 2e-4, k


@pytest.mark.parametrize("prec,tol", [("bf16x3", 2e-3), ("bf16", 6e-2)])
def test_tensor_core_backward_vs_oracle(prec, tol):
    import model
    from phasegen import synth
    C, B, T = 64, 3, 40
    torch.manual_seed(11)
    net = model.UNetModel(C, 2 * C).cuda()
    synth.randomize_norm_affine(net, seed=12)
    net.train_precision = prec
    sd = {k: v.detach().cpu() for k, v in net.model.state_dict().items()}
    x = torch.log1p(torch.randn(B, C, T).abs() * 2.0)
    phi = (torch.rand(B, C, T) * 2 - 1) * np.pi
    pred = net.forward(x.cuda())
    loss = _train_py_loss(pred, x.cuda(), phi.cuda(), C)
    loss.backward()
    ref_loss, _, _, ref = unet_torch.loss_and_grads(sd, x, torch.stack([x, phi], 1))
    assert abs(loss.item() - ref_loss) < tol * abs(ref_loss)
    worst = 0.0
    for k, p in net.model.named_parameters():
        e = rel_l2(p.grad.cpu().numpy(), ref[k].numpy())
        worst = max(worst, e)
        assert e < tol, (k, e)
    print(f"{prec}: worst gradient rel-L2 {worst:.2e}")


def test_loss_kernel_and_adam_kernel():
    from phasegen import ops
    B, T, C = 3, 16, 32
    g = torch.Generator().manual_seed(5)
    out = torch.randn(B, T, 2 * C, generator=g).cuda()
    lm = torch.randn(B, T, C, generator=g).abs().cuda(); ph = ((torch.rand(B, T, C, generator=g) * 2 - 1) * 3.14159).cuda()
    d_out = torch.empty_like(out); partial = torch.empty(64, 3, device="cuda", dtype=torch.float64); loss3 = torch.zeros(4, device="cuda")
    ops.phase_loss(out, lm, ph, d_out, partial, loss3)
    o = out.double().cpu().permute(0, 2, 1).requires_grad_(True)
    tgt = torch.stack([lm.double().cpu().permute(0, 2, 1), ph.double().cpu().permute(0, 2, 1)], 1)
    loss, ang, mag = unet_torch.phase_loss(o, tgt)
    loss.backward()
    assert abs(loss3[0].item() - loss.item()) < 1e-5 * loss.item()
    assert abs((loss3[1] + loss3[2]).item() - ang.item()) < 1e-5 and abs(loss3[3].item() - mag.item()) < 1e-5
    assert rel_l2(d_out.cpu().numpy(), o.grad.permute(0, 2, 1).numpy()) < 1e-5
    # Adam vs torch.optim.Adam (train.py:26-27 defaults), three steps
    p = torch.randn(1000, generator=g).cuda(); pt = p.clone().requires_grad_(True)
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    opt = torch.optim.Adam([pt], lr=1e-3)
    for step in range(1, 4):
        gr = torch.randn(1000, generator=g).cuda()
        ops.adam_step(p, gr, m, v, 1e-3, 0.9, 0.999, 1e-8, step)
        pt.grad = gr.clone(); opt.step()
    assert float((p - pt.detach()).abs().max()) < 1e-6


def test_native_train_step_tracks_reference_loop():
    """TrainStep (all kernels native, Adam included) vs the reference loop restated on the CPU in
    float64: forward, train.py loss, autograd, torch.optim.Adam -- three steps on a fixed batch."""
    import model
    from phasegen.train import TrainStep
    C, B, T = 16, 2, 32
    torch.manual_seed(21)
    net = model.UNetModel(C, 2 * C).cuda()
    sd0 = {k: v.detach().cpu().clone() for k, v in net.model.state_dict().items()}
    x = torch.log1p(torch.randn(B, C, T).abs() * 2.0); phi = (torch.rand(B, C, T) * 2 - 1) * np.pi
    step = TrainStep(net, B, T, "cuda", precision="fp32_simt", lr=1e-3)
    lm_cl = x.permute(0, 2, 1).contiguous().cuda(); ph_cl = phi.permute(0, 2, 1).contiguous().cuda()
    losses = [float(step(lm_cl, ph_cl)[0]) for _ in range(3)]
    # reference loop
    keys = [k for k, v in sd0.items() if v.is_floating_point() and "running_" not in k]
    leaf = {k: sd0[k].double().clone().requires_grad_(True) for k in keys}
    opt = torch.optim.Adam(list(leaf.values()), lr=1e-3)
    ref_losses = []
    for _ in range(3):
        opt.zero_grad()
        full = dict(sd0); full.update(leaf)
        with torch.backends.mkldnn.flags(enabled=False):
            out = unet_torch.unet_forward(full, x, torch.float64, grad=True)
            loss, _, _ = unet_torch.phase_loss(out, torch.stack([x, phi], 1).double())
            loss.backward()
        opt.step()
        ref_losses.append(loss.item())
    assert np.allclose(losses, ref_losses, rtol=2e-4), (losses, ref_losses)
    assert ref_losses[2] < ref_losses[0]
    w = net.model.state_dict()["model.3.weight"].cpu().double()
    assert rel_l2(w.numpy(), leaf["model.3.weight"].detach().numpy()) < 1e-4
