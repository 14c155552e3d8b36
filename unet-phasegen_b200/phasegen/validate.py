"""Validation metrics of train.py:69-122 on the GPU.

For every validation pair (log-magnitude, true phase) the reference rebuilds four waveforms --
ground truth, "hybrid" (true magnitude + predicted phase), "no phase" (zero phase) and Griffin-Lim
(250 iterations) -- through generate_audio (peak-normalised, utils.py:34-42) and reports the mean absolute
sample difference to the ground truth of each ("MSE", "NOPMSE", "LMSE", train.py:103-122).  Here the
whole validation batch runs at once: one U-Net forward with per-clip statistics (the reference's batch-1
loop, train.py:76), three ISTFT launches, the batched GPU Griffin-Lim; the |a - b| means are torch
reductions (plumbing, not a hot path).
"""
import torch

from . import ops
from ._lib import PG_SPEC_POLAR_LOG


def validation_report(net, val_pairs, n_fft, hop, gl_iters=250, gl_init=None, generator=None, return_audio=False):
    """``val_pairs``: float ``[V, 2, C, T]`` (log-magnitude, phase) as the reference's loader yields it
    (data.py:39-47).  Returns {"MSE", "NOPMSE", "LMSE"} (floats; LMSE is None when gl_iters == 0), plus the
    four ``[V, (T-1)*hop]`` waveform tensors under "audio" when ``return_audio``."""
    import utils as pg_utils                                   # the drop-in module (griffin_lim_batch lives there)
    if not torch.cuda.is_available():
        raise RuntimeError("validation_report needs a CUDA device: there is no CPU fallback")
    dev = next(net.parameters()).device
    vp = torch.as_tensor(val_pairs).to(dev, torch.float32)
    V, two, C, T = vp.shape
    if two != 2 or 2 * C != n_fft:
        raise RuntimeError(f"validation_report: expected [V, 2, {n_fft // 2}, T], got {tuple(vp.shape)}")
    logmag = ops.transpose(vp[:, 0].contiguous())              # [V, T, C] channels-last
    phase = ops.transpose(vp[:, 1].contiguous())
    with torch.no_grad():
        pred = net.forward_channels_last(logmag, per_clip=True, phase_only=True)
        if pred.shape[2] != C:
            pred = pred[:, :, :C].contiguous()

    def audio(ph):
        w, _ = ops.istft(logmag, ph, PG_SPEC_POLAR_LOG, n_fft, hop, normalize=True, check_finite=True)
        return w
    orig, hyb, nop = audio(phase), audio(pred), audio(None)
    out = {"MSE": float((orig - hyb).abs().mean()), "NOPMSE": float((orig - nop).abs().mean()), "LMSE": None}
    lim = None
    if gl_iters > 0:
        mag = torch.expm1(logmag)
        lim, _, _ = pg_utils.griffin_lim_batch(mag, n_fft, hop, gl_iters, init=gl_init, generator=generator)
        if not bool(torch.isfinite(lim).all()):
            raise ValueError("Audio buffer is not finite everywhere")
        peak = lim.abs().amax(dim=1, keepdim=True)
        lim = torch.where(peak >= torch.finfo(torch.float32).tiny, lim / peak.clamp_min(torch.finfo(torch.float32).tiny), lim)
        out["LMSE"] = float((orig - lim).abs().mean())
    if return_audio:
        out["audio"] = {"orig": orig, "hybrid": hyb, "no_phase": nop, "griffin_lim": lim}
    return out
