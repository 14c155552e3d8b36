"""Seeded synthetic inputs of the shapes BASELINE.json names."""
import math

import torch


def frames_for(seconds, sr, hop):
    """STFT frames of a clip, rounded up to the next multiple of 8 (U-Net time-axis rule)."""
    t = 1 + int(seconds * sr) // hop
    return (t + 7) // 8 * 8


def clip_samples(seconds=4.0, sr=44100, hop=256):
    """Samples per clip such that the STFT has a multiple-of-8 frame count: (T-1)*hop."""
    return (frames_for(seconds, sr, hop) - 1) * hop


def synthetic_waves(batch, n_samples, sr=44100, seed=0, device="cpu"):
    """0.1*N(0,1) noise + 3 random sinusoids (amp U(0.05,0.3), f U(50 Hz, sr/2)), clipped to [-1,1]."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n_samples, dtype=torch.float64) / sr
    w = 0.1 * torch.randn(batch, n_samples, generator=g, dtype=torch.float64)
    amp = 0.05 + 0.25 * torch.rand(batch, 3, generator=g, dtype=torch.float64)
    f = 50.0 + (sr / 2 - 50.0) * torch.rand(batch, 3, generator=g, dtype=torch.float64)
    ph = 2 * math.pi * torch.rand(batch, 3, generator=g, dtype=torch.float64)
    for j in range(3):
        w += amp[:, j:j + 1] * torch.sin(2 * math.pi * f[:, j:j + 1] * t[None, :] + ph[:, j:j + 1])
    return w.clamp_(-1, 1).float().to(device)


def randomize_norm_affine(model, seed=0):
    """Give the norm layers non-trivial gamma/beta so the affine path is exercised."""
    import torch.nn as nn
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
