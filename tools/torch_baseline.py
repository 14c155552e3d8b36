"""GPU-LIBRARY baseline of the hot path on the same B200: what the reference's own code would run on a GPU today.

The reference U-Net is eight `nn.Conv1d` / `nn.ConvTranspose1d` calls (/root/reference/model.py:77-78,88-89,94-95,
101-102) -> cuDNN, seven `nn.BatchNorm` layers in train mode, in-place LeakyReLU/ReLU and `torch.cat` skips
(model.py:80-83,113); its STFT/ISTFT are librosa on the CPU (preproc_mdb.py:93, utils.py:40) -- the library
equivalent on a GPU is `torch.stft` / `torch.istft` (cuFFT).  This file builds exactly that from stock torch modules
and functionals: no phasegen kernel, no oracle import.  SURVEY.md section 2.2 calls it "the existing Blackwell kernel
to beat" and section 8d the secondary baseline.

Per-clip statistics: a train-mode BatchNorm on a batch-1 call (the demo.py:33-42 loop) equals an affine instance norm,
so the batched form uses `F.instance_norm(weight, bias)` -- the fastest stock way to run 256 clips with batch-1
semantics (looping over clips would be launch-bound and flatter the comparison).
"""
import math
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

# (name, kind, in-mult, out-mult, k, stride, pad, has_norm) in execution order; channel counts are multiples of C
LAYERS = (("d1", "conv", 1, 2, 32, 2, 16, False), ("d2", "conv", 2, 2, 8, 1, 2, True), ("d3", "conv", 2, 2, 8, 2, 1, True),
          ("d4", "conv", 2, 4, 4, 2, 1, False), ("u4", "convT", 4, 2, 5, 2, 1, True), ("u3", "convT", 4, 2, 8, 2, 1, True),
          ("u2", "convT", 4, 2, 8, 1, 2, True), ("u1", "convT", 4, 2, 32, 2, 16, True))


class StockUNet(nn.Module):
    """The reference architecture from stock torch layers (random init; weights are irrelevant for timing)."""

    def __init__(self, C, per_clip=True):
        super().__init__()
        self.per_clip = per_clip
        self.convs, self.norms = nn.ModuleDict(), nn.ModuleDict()
        for name, kind, ci, co, k, s, p, has_norm in LAYERS:
            mod = nn.Conv1d if kind == "conv" else nn.ConvTranspose1d
            self.convs[name] = mod(C * ci, C * co, k, s, p, bias=False)
            if has_norm:
                self.norms[name] = nn.BatchNorm1d(C * co)

    def _layer(self, name, x, times=None):
        if times is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        y = self.convs[name](x)
        if times is not None:
            e1.record()
            times.setdefault(name, []).append((e0, e1))
        if name in self.norms:
            n = self.norms[name]
            if self.per_clip:
                y = F.instance_norm(y, weight=n.weight, bias=n.bias, eps=n.eps)
            else:
                y = F.batch_norm(y, None, None, n.weight, n.bias, True, 0.0, n.eps)
        return y

    def forward(self, x, times=None):
        lrelu = lambda t: F.leaky_relu(t, 0.2)
        y1 = self._layer("d1", x, times)
        h2 = self._layer("d2", lrelu(y1), times)
        h3 = self._layer("d3", lrelu(h2), times)
        y4 = self._layer("d4", lrelu(h3), times)
        g4 = self._layer("u4", F.relu(y4), times)
        g3 = self._layer("u3", F.relu(torch.cat([h3, g4], 1)), times)
        g2 = self._layer("u2", F.relu(torch.cat([h2, g3], 1)), times)
        return self._layer("u1", F.relu(torch.cat([y1, g2], 1)), times)


def stock_pipeline(net, wave, n_fft, hop, times=None):
    """wave [B, N] -> wave [B, N]: torch.stft -> log1p|X| -> U-Net -> expm1(mag) e^{j phase} -> torch.istft -> x / max|x|."""
    C = n_fft // 2
    win = torch.hann_window(n_fft, periodic=True, device=wave.device)
    S = torch.stft(wave, n_fft, hop, window=win, center=True, pad_mode="reflect", return_complex=True)[:, 1:]   # DC row dropped
    logmag = torch.log1p(S.abs())
    out = net(logmag, times)
    phase = out[:, :C].float()
    spec = torch.polar(torch.expm1(logmag), phase)
    spec = torch.cat([torch.zeros_like(spec[:, :1]), spec], 1)
    audio = torch.istft(spec, n_fft, hop, window=win, center=True, length=wave.shape[1])
    return audio / audio.abs().amax(dim=1, keepdim=True).clamp_min(torch.finfo(torch.float32).tiny)


def _timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def inference_baseline(wave, n_fft, hop, clip_seconds, steps=3, warmup=2, sub_batch=64):
    """audio-s/s of the stock path on `wave` [B, N] (device-resident), processed in sub-batches of `sub_batch` clips
    (fp32 activations of 256 clips at once would only add allocator pressure), for fp32 (TF32 off), TF32 and bf16
    autocast.  Returns a dict of modes plus per-layer cuDNN times of the TF32 run."""
    dev = wave.device
    C = n_fft // 2
    B = wave.shape[0]
    torch.manual_seed(0)
    net = StockUNet(C, per_clip=True).to(dev)
    res = {}

    def run_all(times=None):
        with torch.no_grad():
            for i in range(0, B, sub_batch):
                stock_pipeline(net, wave[i:i + sub_batch], n_fft, hop, times)

    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for mode in ("tf32", "fp32", "bf16_autocast"):
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
            if mode == "bf16_autocast":
                def fn():
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        run_all()
            else:
                fn = run_all
            try:
                ms = _timed(fn, steps, warmup)
                res[mode] = {"value": B * clip_seconds / (ms / 1e3), "unit": "audio-s/s", "ms_per_step": ms}
            except RuntimeError as e:                      # e.g. out of memory on a shared box: report, do not die
                res[mode] = {"error": str(e).splitlines()[0][:200]}
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
        times = {}
        run_all(times)
        torch.cuda.synchronize()
        res["cudnn_tf32_conv_ms_per_step"] = {k: round(sum(a.elapsed_time(b) for a, b in v), 3) for k, v in times.items()}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    res["note"] = (f"stock torch {torch.__version__}: torch.stft/istft (cuFFT) + nn.Conv1d/ConvTranspose1d (cuDNN {torch.backends.cudnn.version()}) "
                   f"+ F.instance_norm (= train-mode BatchNorm per clip), {B} clips in sub-batches of {sub_batch}, same GPU, CUDA events; "
                   "tf32 = the reference's fp32 module with TF32 convolutions allowed (BASELINE config 2 'fp32/TF32'), "
                   "fp32 = TF32 off, bf16_autocast = outside the parity bound, listed for context")
    return res


def train_baseline(C, T, B, steps=3, warmup=2):
    """samples/s of a stock train.py step (forward, cos/sin/mag loss, backward, torch.optim.Adam) at config 3's shape."""
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(0)
    res = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for mode in ("tf32", "bf16_autocast"):
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
            try:
                net = StockUNet(C, per_clip=False).to(dev)
                opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999))
                x = torch.log1p(torch.randn(B, C, T, device=dev).abs())
                phi = (torch.rand(B, C, T, device=dev) * 2 - 1) * math.pi
                lossf = nn.MSELoss()

                def fn():
                    opt.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16_autocast"):
                        pred = net(x).float()
                    pp, pm = pred[:, :C], pred[:, C:]
                    loss = lossf(torch.cos(pp), phi.cos()) + lossf(torch.sin(pp), phi.sin()) + 0.2 * lossf(pm, x)
                    loss.backward()
                    opt.step()
                ms = _timed(fn, steps, warmup)
                res[mode] = {"value": B / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms}
                del net, opt
                torch.cuda.empty_cache()
            except RuntimeError as e:
                res[mode] = {"error": str(e).splitlines()[0][:200]}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    res["note"] = (f"stock torch train step (nn.Conv1d/ConvTranspose1d via cuDNN, nn.BatchNorm1d batch statistics, torch.optim.Adam), "
                   f"UNetModel({C},{2 * C}) shape, batch {B}, T {T}, one GPU")
    return res


if __name__ == "__main__":
    import json
    import sys
    torch.cuda.set_device(0)
    n_fft, hop, T = 1024, 256, 696
    N = (T - 1) * hop
    wave = (0.1 * torch.randn(256, N, device="cuda")).clamp_(-1, 1)
    t0 = time.time()
    out = {"inference": inference_baseline(wave, n_fft, hop, N / 44100.0), "train": train_baseline(1024, 128, 32)}
    out["wall_s"] = time.time() - t0
    json.dump(out, sys.stdout, indent=1)
    print()
