"""Online training-pair producer: wave -> (log-magnitude, phase) pairs on the GPU.

Replaces the reference's offline stage for the training loop:
  chunk_audio / _chunk_and_stft      preproc_mdb.py:66-97   (regular + random chunk starts, zero pad, STFT, DC drop)
  dataset-wide standardisation       preproc_mdb.py:182     ((x - mean) / std over every stored re/im value)
  get_spec_and_angle                 data.py:39-47          (log1p|.|, angle)
The chunk gather is a torch indexing op (plumbing); the STFT, the standardisation and the magnitude/phase
conversion are one launch of the STFT kernel (``pg_stft_pairs``).  The pairs come out channels-last
``[B, T, C]`` -- the layout ``phasegen.train.TrainStep`` consumes -- so a training step needs no host
round trip and no per-step H2D copy (train.py:42,49-50,57 upload a batch every step).
"""
import numpy as np
import torch

from . import ops
from ._lib import PG_STFT_REIM


def chunk_starts(a_len, t_slice, n_random, rng=None):
    """The start offsets ``chunk_audio`` (preproc_mdb.py:66-82) visits: every ``t_slice`` samples, each
    followed by ``n_random`` uniformly random starts below ``a_len - t_slice // 1.3``."""
    rng = np.random.default_rng() if rng is None else rng
    bnd = int(a_len - t_slice // 1.3)
    out = []
    for i in range(0, a_len, t_slice):
        out.append(i)
        for _ in range(n_random):
            out.append(int(rng.integers(0, max(bnd, 1))))
    return np.asarray(out, dtype=np.int64)


class TrainingPairProducer:
    def __init__(self, audio, t_slice, n_fft, hop, device=None):
        """``audio``: mono track, float ``[a_len]`` (numpy or tensor); it is kept resident on the GPU."""
        ops.check_stft_geometry(n_fft, hop)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("TrainingPairProducer needs a CUDA device: there is no CPU fallback")
        a = torch.as_tensor(np.asarray(audio, dtype=np.float32) if not isinstance(audio, torch.Tensor) else audio)
        self.audio = a.to(dev, torch.float32).contiguous().view(-1)
        self.t_slice, self.n_fft, self.hop = int(t_slice), n_fft, hop
        self.frames = 1 + self.t_slice // hop
        self._offs = torch.arange(self.t_slice, device=dev)

    def chunks(self, starts):
        """[B, t_slice] chunks starting at ``starts`` (zero-padded past the end of the track: preproc_mdb.py:86-88)."""
        st = torch.as_tensor(np.asarray(starts, dtype=np.int64)).to(self.audio.device)
        idx = st[:, None] + self._offs[None, :]
        valid = idx < self.audio.numel()
        return torch.where(valid, self.audio[idx.clamp_max(self.audio.numel() - 1)], torch.zeros((), device=idx.device))

    def stats(self, starts, batch=64):
        """(mean, std) over every re/im value of the chunks' DC-less STFTs (preproc_mdb.py:182; population std
        like ndarray.std), accumulated in float64."""
        s = torch.zeros((), dtype=torch.float64, device=self.audio.device)
        ss = torch.zeros_like(s)
        n = 0
        starts = np.asarray(starts, dtype=np.int64)
        for i in range(0, len(starts), batch):
            re, im = ops.stft(self.chunks(starts[i:i + batch]), self.n_fft, self.hop, mode=PG_STFT_REIM)
            for p in (re, im):
                d = p.double()
                s += d.sum(); ss += (d * d).sum(); n += p.numel()
        mean = float(s / n)
        var = float(ss / n) - mean * mean
        return mean, float(np.sqrt(max(var, 0.0)))

    def pairs(self, starts, mean=0.0, std=1.0):
        """(log-magnitude, phase) ``[B, T, C]`` float32 on the GPU for the chunks at ``starts``."""
        return ops.stft_pairs(self.chunks(starts), self.n_fft, self.hop, mean, std)
