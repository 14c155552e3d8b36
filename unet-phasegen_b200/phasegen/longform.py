"""Long-form inference (BASELINE.json config 4): a long recording is cut into fixed windows of
T frames with 50 % overlap, the windows run through the magnitude -> phase -> waveform pipeline
as ONE batch (independent units: per-clip norm statistics), and the outputs are cross-faded.

The reference has no long-form path: its audio is cut offline into independent 4.064 s slices
(preproc_mdb.py:66-82,204) and never stitched.  Per-window results are exactly the pipeline's
(i.e. the reference's demo.py:33-42 chain on that window, before utils.py:42's peak
normalisation, which would make window gains inconsistent); the seam policy below is this
build's own: a periodic-Hann cross-fade, whose 50 %-overlap shifts sum to one, followed by ONE
peak normalisation of the whole recording.

Multi-GPU: windows are dealt round-robin to ranks (rank r takes windows r, r+W, ...); there is
no data-path collective, the host concatenates the returned windows.
"""
import math

import torch


def window_plan(n_samples, hop, frames=696):
    """(window length in samples, step between windows, number of windows)."""
    win = (frames - 1) * hop
    step = -(-(win // 2) // hop) * hop          # >= win/2, so at most two windows overlap anywhere
    n = 1 if n_samples <= win else 1 + math.ceil((n_samples - win) / step)
    return win, step, n


def cut_windows(wave, hop, frames=696, rank=0, world=1):
    """wave [N] (any device) -> ([n_local, win] windows of this rank, their global indices)."""
    N = wave.shape[0]
    win, step, n = window_plan(N, hop, frames)
    total = win + (n - 1) * step
    padded = torch.zeros(total, dtype=wave.dtype, device=wave.device)
    padded[:N] = wave
    idx = list(range(rank, n, world))
    if not idx:
        return padded.new_zeros(0, win), idx
    return torch.stack([padded[i * step:i * step + win] for i in idx]), idx


def stitch(windows, indices, n_windows, n_samples, hop, frames=696):
    """Cross-fade overlapping windows ([n, win] for the global `indices`) back into [n_samples]."""
    win, step, _ = window_plan(n_samples, hop, frames)
    total = win + (n_windows - 1) * step
    ov = win - step                              # samples shared by consecutive windows
    fade = torch.hann_window(2 * ov, periodic=True, dtype=torch.float64, device=windows.device)
    out = torch.zeros(total, dtype=torch.float64, device=windows.device)
    for w, i in zip(windows, indices):
        g = torch.ones(win, dtype=torch.float64, device=windows.device)
        if i > 0:
            g[:ov] = fade[:ov]                   # rising half; the falling half of window i-1 complements it
        if i < n_windows - 1:
            g[win - ov:] = fade[ov:]
        out[i * step:i * step + win] += w.double() * g
    return out[:n_samples].float()


def process_long(pipe, wave, frames=696, batch=256, peak_normalize=True):
    """wave [N] float32 on the GPU -> [N] float32, through `pipe` (a PhaseGenPipeline built with
    normalize=False) window batch by window batch on this GPU."""
    if pipe.normalize:
        raise RuntimeError("process_long needs a pipeline built with normalize=False (one global normalisation)")
    N = wave.shape[0]
    wins, idx = cut_windows(wave, pipe.hop, frames)
    outs = [pipe(wins[i:i + batch].contiguous()).clone() for i in range(0, wins.shape[0], batch)]
    y = stitch(torch.cat(outs), idx, len(idx), N, pipe.hop, frames)
    if peak_normalize:
        peak = y.abs().max()
        if float(peak) >= torch.finfo(torch.float32).tiny:
            y = y / peak
    return y
