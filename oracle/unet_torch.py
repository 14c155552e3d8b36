"""Functional torch-CPU restatement of the reference U-Net forward.  TEST INFRASTRUCTURE.

Follows /root/reference/model.py: the four nested ``UNetBlock``s built at model.py:27-34,
the layer order of model.py:85-105 and the skip concat of model.py:109-113.  It works on a
plain ``state_dict`` with the reference's key names (the inner-block dict that
``UNetModel.save`` writes, model.py:45-48), so it needs neither the reference module nor the
product module at run time.  Pinned against the real reference by ``oracle/make_golden.py``
-> ``tests/golden/unet_*.npz`` -> ``tests/test_oracle_unet.py``.

Numerics: run in float64 (default) or float32.  The float32 path must run with oneDNN off:
in this image fp32 ``conv_transpose1d`` through oneDNN is ~22 % wrong for the innermost
k=5/s=2 layer, so every call here sits inside
``torch.backends.mkldnn.flags(enabled=False)``.
"""
import torch
import torch.nn.functional as F

# (key prefix of the conv weight, key prefix of its norm or None, kind, k, stride, pad)
# in execution order; geometry from model.py:27-34 (+ the k_size+1 of model.py:94).
_P1 = "model.1.model."
_P2 = _P1 + "3.model."
_P3 = _P2 + "3.model."
LAYERS = {
    "d1": ("model.0", None, "conv", 32, 2, 16),
    "d2": (_P1 + "1", _P1 + "2", "conv", 8, 1, 2),
    "d3": (_P2 + "1", _P2 + "2", "conv", 8, 2, 1),
    "d4": (_P3 + "1", None, "conv", 4, 2, 1),
    "u4": (_P3 + "3", _P3 + "4", "convT", 5, 2, 1),
    "u3": (_P2 + "5", _P2 + "6", "convT", 8, 2, 1),
    "u2": (_P1 + "5", _P1 + "6", "convT", 8, 1, 2),
    "u1": ("model.3", "model.4", "convT", 32, 2, 16),
}
BN_EPS = 1e-5  # nn.BatchNorm default, model.py:81,83


def _bn_train(y, sd, prefix, per_clip, dt):
    """Train-mode batch norm (the reference never calls .eval()):
    biased variance over (B, L) -- or over L of each clip when ``per_clip`` (what the
    demo.py:33-42 batch-1 loop computes) -- then the affine weight/bias."""
    dims = (2,) if per_clip else (0, 2)
    mean = y.mean(dim=dims, keepdim=True)
    var = y.var(dim=dims, unbiased=False, keepdim=True)
    g = sd[prefix + ".weight"].to(dt).view(1, -1, 1)
    b = sd[prefix + ".bias"].to(dt).view(1, -1, 1)
    return (y - mean) / torch.sqrt(var + BN_EPS) * g + b


def _apply(name, x, sd, per_clip, dt):
    wkey, nkey, kind, k, s, p = LAYERS[name]
    w = sd[wkey + ".weight"].to(dt)
    y = (F.conv1d if kind == "conv" else F.conv_transpose1d)(x, w, None, s, p)
    return _bn_train(y, sd, nkey, per_clip, dt) if nkey else y


def unet_forward(sd, x, dtype=torch.float64, per_clip_bn=False, taps=None, grad=False, mkldnn=False):
    """x [B,C,T] -> [B,2C,T]; out[:, :C] is the raw phase estimate, out[:, C:] the log-mag
    estimate (train.py:45).  ``taps`` (a dict) receives the raw conv outputs and the
    normalised tensors of every layer for per-layer parity checks.  ``grad=True`` keeps the
    autograd graph (used by ``loss_and_grads``).  ``mkldnn=True`` leaves oneDNN on: only for TIMING the stock CPU path
     -- its fp32 result is numerically invalid in this image (module docstring)."""
    dt = dtype
    lrelu = lambda t: F.leaky_relu(t, 0.2)
    with torch.set_grad_enabled(grad), torch.backends.mkldnn.flags(enabled=bool(mkldnn)):
        x = x.to(dt)
        y1 = _apply("d1", x, sd, per_clip_bn, dt)                      # model.py:90, no norm
        h2 = _apply("d2", lrelu(y1), sd, per_clip_bn, dt)              # model.py:103
        h3 = _apply("d3", lrelu(h2), sd, per_clip_bn, dt)
        y4 = _apply("d4", lrelu(h3), sd, per_clip_bn, dt)              # model.py:96, no norm
        g4 = _apply("u4", F.relu(y4), sd, per_clip_bn, dt)             # model.py:97
        # in-place LeakyReLU (model.py:80) makes the skip operand LeakyReLU(h); the ReLU
        # that follows every concat (model.py:91,104) turns it into ReLU(h).
        g3 = _apply("u3", F.relu(torch.cat([lrelu(h3), g4], 1)), sd, per_clip_bn, dt)
        g2 = _apply("u2", F.relu(torch.cat([lrelu(h2), g3], 1)), sd, per_clip_bn, dt)
        out = _apply("u1", F.relu(torch.cat([lrelu(y1), g2], 1)), sd, per_clip_bn, dt)
        if taps is not None:
            taps.update(y1=y1, h2=h2, h3=h3, y4=y4, g4=g4, g3=g3, g2=g2, out=out)
    return out


def phase_loss(pred, target):
    """train.py:45-60: MSE(cos p, cos phi) + MSE(sin p, sin phi) + 0.2 * MSE(m, logmag).
    pred [B,2C,T]; target [B,2,C,T] = (logmag, phase)."""
    C = pred.shape[1] // 2
    pp, pm = pred[:, :C], pred[:, C:]
    ang = F.mse_loss(torch.cos(pp), torch.cos(target[:, 1])) + \
        F.mse_loss(torch.sin(pp), torch.sin(target[:, 1]))
    mag = F.mse_loss(pm, target[:, 0])
    return ang + 0.2 * mag, ang, mag


def loss_and_grads(sd, x, target, dtype=torch.float64):
    """train.py:42-61 on the CPU: forward (batch statistics), loss, ``loss.backward()``.
    Returns (loss, ang_loss, mag_loss, {key: grad}) for every floating-point parameter key."""
    keys = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaf = {k: sd[k].detach().to(dtype).clone().requires_grad_(True) for k in keys}
    full = dict(sd)
    full.update(leaf)
    with torch.backends.mkldnn.flags(enabled=False):
        out = unet_forward(full, x, dtype, per_clip_bn=False, grad=True)
        loss, ang, mag = phase_loss(out, target.to(dtype))
        loss.backward()
    return loss.item(), ang.item(), mag.item(), {k: v.grad.detach() for k, v in leaf.items()}


def random_state_dict(C, seed=0, dtype=torch.float32):
    """Weights with the reference's keys and PyTorch's default Conv/ConvT/BatchNorm init
    (``weights_init`` at model.py:12-20 is never called).  Built from plain nn layers so the
    oracle does not depend on the product package."""
    import torch.nn as nn
    g = torch.Generator().manual_seed(seed)
    sd = {}
    nc = {"d1": (C, 2 * C), "d2": (2 * C, 2 * C), "d3": (2 * C, 2 * C), "d4": (2 * C, 4 * C),
          "u4": (4 * C, 2 * C), "u3": (4 * C, 2 * C), "u2": (4 * C, 2 * C), "u1": (4 * C, 2 * C)}
    state = torch.random.get_rng_state()
    torch.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,), generator=g)))
    try:
        for name, (wkey, nkey, kind, k, s, p) in LAYERS.items():
            cin, cout = nc[name]
            mod = (nn.Conv1d if kind == "conv" else nn.ConvTranspose1d)(cin, cout, k, s, p, bias=False)
            sd[wkey + ".weight"] = mod.weight.detach().to(dtype).clone()
            if nkey:
                sd[nkey + ".weight"] = 1.0 + 0.1 * torch.randn(cout, dtype=dtype)
                sd[nkey + ".bias"] = 0.1 * torch.randn(cout, dtype=dtype)
                sd[nkey + ".running_mean"] = torch.zeros(cout, dtype=dtype)
                sd[nkey + ".running_var"] = torch.ones(cout, dtype=dtype)
                sd[nkey + ".num_batches_tracked"] = torch.tensor(0)
    finally:
        torch.random.set_rng_state(state)
    return sd
