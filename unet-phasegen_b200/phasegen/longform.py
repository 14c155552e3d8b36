"""Long-form inference (BASELINE.json config 4): a long recording is cut into fixed windows of
T frames with 50 % overlap, the windows run through the magnitude -> phase -> waveform pipeline
as ONE batch per rank (independent units: per-clip norm statistics), and the outputs are cross-faded.

The reference has no long-form path: its audio is cut offline into independent 4.064 s slices
(preproc_mdb.py:66-82,204) and never stitched.  Per-window results are exactly the pipeline's
(i.e. the reference's demo.py:33-42 chain on that window, before utils.py:42's peak
normalisation, which would make window gains inconsistent); the seam policy below is this
build's own: a periodic-Hann cross-fade, whose 50 %-overlap shifts sum to one, followed by ONE
peak normalisation of the whole recording.

Multi-GPU: windows are dealt round-robin to ranks (rank r takes windows r, r+W, ...); the data
path has no collective -- the finished windows (0.7 MB each) are all-gathered once at the end and
stitched by one kernel (pg_stitch reads the round-robin layout directly).
"""
import ctypes as C
import math

import torch

from . import _lib


def window_plan(n_samples, hop, frames=696):
    """(window length in samples, step between windows, number of windows)."""
    win = (frames - 1) * hop
    step = -(-(win // 2) // hop) * hop          # >= win/2, so at most two windows overlap anywhere
    n = 1 if n_samples <= win else 1 + math.ceil((n_samples - win) / step)
    return win, step, n


def cut_windows(wave, hop, frames=696, rank=0, world=1):
    """wave [N] (any device) -> ([n_local, win] windows of this rank, their global indices).  One strided view +
    one gather, no per-window loop."""
    N = wave.shape[0]
    win, step, n = window_plan(N, hop, frames)
    total = win + (n - 1) * step
    padded = torch.zeros(total, dtype=wave.dtype, device=wave.device)
    padded[:N] = wave
    idx = list(range(rank, n, world))
    if not idx:
        return padded.new_zeros(0, win), idx
    all_windows = padded.as_strided((n, win), (step, 1))
    return all_windows[torch.as_tensor(idx, device=wave.device)].contiguous(), idx


def stitch(windows, indices, n_windows, n_samples, hop, frames=696, world=1, per_rank=None, peak_out=None):
    """Cross-fade overlapping windows back into [n_samples].  `windows` [n, win] holds either all windows in plain
    order (world = 1; `indices` must be 0..n-1) or the all-gathered round-robin layout [world * per_rank, win].  On the
    GPU this is one pg_stitch launch; CPU tensors (host-side tests) use the equivalent vectorised torch form.
    peak_out: optional 1-element float tensor that receives max |out| (GPU path)."""
    win, step, _ = window_plan(n_samples, hop, frames)
    per_rank = (n_windows + world - 1) // world if per_rank is None else per_rank
    if world == 1 and list(indices) != list(range(n_windows)):
        order = torch.argsort(torch.as_tensor(list(indices)))
        windows = windows[order.to(windows.device)]
    if windows.is_cuda:
        windows = windows.contiguous()
        if windows.dtype != torch.float32 or windows.shape[0] < (n_windows if world == 1 else world * per_rank):
            raise RuntimeError("phasegen.longform.stitch: windows must be float32 [n_windows (or world*per_rank), win]")
        out = torch.empty(n_samples, device=windows.device, dtype=torch.float32)
        st = C.c_void_p(torch.cuda.current_stream(windows.device).cuda_stream)
        _lib.call("pg_stitch", C.c_void_p(windows.data_ptr()), n_windows, win, step, world, per_rank,
                  C.c_void_p(out.data_ptr()), n_samples, C.c_void_p(peak_out.data_ptr()) if peak_out is not None else None, st)
        return out
    # host form (float64 accumulate), windows in plain order
    if world != 1:
        slots = torch.as_tensor([(i % world) * per_rank + i // world for i in range(n_windows)])
        windows = windows[slots]
    total = win + (n_windows - 1) * step
    ov = win - step
    gain = torch.ones(n_windows, win, dtype=torch.float64)
    if ov > 0 and n_windows > 1:
        fade = torch.hann_window(2 * ov, periodic=True, dtype=torch.float64)
        gain[1:, :ov] = fade[:ov]                # rising half; the falling half of window i-1 complements it
        gain[:-1, win - ov:] = fade[ov:]
    out = torch.zeros(total, dtype=torch.float64)
    pos = (torch.arange(n_windows)[:, None] * step + torch.arange(win)[None, :]).reshape(-1)
    out.index_add_(0, pos, (windows[:n_windows].double() * gain).reshape(-1))
    return out[:n_samples].float()


def process_long(pipe, wave, frames=696, batch=256, peak_normalize=True, group=None):
    """wave [N] float32 on the GPU -> [N] float32, through `pipe` (a PhaseGenPipeline built with
    normalize=False).  With an initialised process group every rank passes the SAME wave, processes its round-robin
    share of the windows, and receives the whole stitched recording (one all-gather of the finished windows)."""
    import torch.distributed as dist
    if pipe.normalize:
        raise RuntimeError("process_long needs a pipeline built with normalize=False (one global normalisation)")
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    N = wave.shape[0]
    win, step, n = window_plan(N, pipe.hop, frames)
    wins, idx = cut_windows(wave, pipe.hop, frames, rank, world)
    per_rank = (n + world - 1) // world
    mine = torch.zeros(per_rank, win, device=wave.device, dtype=torch.float32)     # ragged last round: zero window
    for i in range(0, wins.shape[0], batch):
        pipe(wins[i:i + batch].contiguous(), wave_out=mine[i:i + min(batch, wins.shape[0] - i)])
    if world > 1:
        gathered = torch.empty(world * per_rank, win, device=wave.device, dtype=torch.float32)
        dist.all_gather_into_tensor(gathered, mine, group=group)
    else:
        gathered = mine
    peak = torch.zeros(1, device=wave.device, dtype=torch.float32)
    y = stitch(gathered, list(range(n)), n, N, pipe.hop, frames, world=world, per_rank=per_rank, peak_out=peak)
    if peak_normalize:
        st = C.c_void_p(torch.cuda.current_stream(wave.device).cuda_stream)
        _lib.call("pg_peak_normalize", C.c_void_p(y.data_ptr()), C.c_void_p(peak.data_ptr()), 1, N, st)
    return y
