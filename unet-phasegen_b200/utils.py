"""Drop-in for the reference's ``utils.py`` signal helpers, executed on the B200 kernels.

  generate_audio(spec, sr, hop_length, is_stft=False)   utils.py:11-44  -> ISTFT kernel
  stft_nodc(wave, n_fft, hop_length)                    preproc_mdb.py:84-97 (the librosa.stft call
                                                        + DC-row drop + re/im split; new helper, the
                                                        reference calls librosa inline)
  spec_and_angle(reim)                                  data.py:39-47 on the GPU
  griffin_lim(spec, n_fft, hop_length, n_iter)          utils.py:85-134 ("next" row, see docstring)

numpy in / numpy out like the reference; the arithmetic runs in libphasegen.so on the current
CUDA device and there is no CPU fallback.  The remaining names (``View`` ... ``Pool``,
``generate_spec_img``) exist because the reference's ``model.py:7`` / ``train.py:5`` import
them; they are plain PyTorch / matplotlib helpers with the reference's behaviour.
"""
import numpy as np
import torch
import torch.nn as nn

from phasegen import ops
from phasegen._lib import PG_SPEC_CARTESIAN, PG_SPEC_POLAR_MAG, PG_STFT_LOGMAG, PG_STFT_REIM


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("phasegen utils need a CUDA device (B200): there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _frame_major(planes):
    """[2, C, T] device tensor -> two contiguous [1, T, C] planes (one transposing kernel each)."""
    return ops.transpose(planes[0:1].contiguous()), ops.transpose(planes[1:2].contiguous())


def generate_audio(spec, sr, hop_length, is_stft=False):
    """Audio from a DC-less STFT: complex ``[C, T]`` when ``is_stft`` else real ``[2, C, T]``
    (re, im).  Zero DC row, inverse STFT with n_fft = 2*C, finiteness check, peak
    normalisation; returns float32 ``[(T-1)*hop_length]``.  ``sr`` is unused (as in the reference).
    Raises ValueError on non-finite audio (librosa.util.valid_audio raises ParameterError)."""
    spec = np.asarray(spec)
    if is_stft:
        planes = np.stack([spec.real, spec.imag])
    else:
        planes = spec[:2]
    planes = torch.from_numpy(np.ascontiguousarray(planes, dtype=np.float32)).to(_device())
    n_fft = 2 * planes.shape[1]
    re, im = _frame_major(planes)
    wave, _ = ops.istft(re, im, PG_SPEC_CARTESIAN, n_fft, hop_length, normalize=True, check_finite=True)
    return wave[0].cpu().numpy()


def stft_nodc(wave, n_fft, hop_length):
    """float ``[N]`` -> float32 ``[2, n_fft/2, T]`` (re, im) with the DC row removed: what
    ``_chunk_and_stft`` (preproc_mdb.py:93-96) stores per channel."""
    w = torch.from_numpy(np.ascontiguousarray(wave, dtype=np.float32)).to(_device()).view(1, -1)
    re, im = ops.stft(w, n_fft, hop_length, mode=PG_STFT_REIM)
    return torch.cat([ops.transpose(re), ops.transpose(im)], 0).cpu().numpy()


def spec_and_angle_from_wave(wave, n_fft, hop_length):
    """wave ``[N]`` -> float32 ``[2, C, T]`` (log1p|X|, angle X): preproc_mdb.py:93 followed by
    data.py:39-47, fused in one kernel."""
    w = torch.from_numpy(np.ascontiguousarray(wave, dtype=np.float32)).to(_device()).view(1, -1)
    lm, ph = ops.stft(w, n_fft, hop_length, mode=PG_STFT_LOGMAG)
    return torch.cat([ops.transpose(lm), ops.transpose(ph)], 0).cpu().numpy()


def griffin_lim_batch(mag_fm, n_fft, hop_length, n_iter, init=None, generator=None):
    """Griffin-Lim on a batch, device tensors in and out: ``mag_fm`` float32 ``[B, T, C]`` (frame-major
    DC-less magnitudes) -> (audio ``[B, (T-1)*hop]`` un-normalised, (re, im) of the last projected
    spectrogram ``[B, T, C]``, per-clip RMSE between the last two iterates ``[B]``).  Each iteration is two
    kernels: ``pg_stft_project`` (STFT with the magnitude replaced in its epilogue, X/|X| instead of
    exp(j*angle X)) and ``pg_istft``; nothing leaves the GPU inside the loop."""
    B, T, C = mag_fm.shape
    if 2 * C != n_fft:
        raise RuntimeError(f"griffin_lim: spec has {C} rows, expected n_fft/2 = {n_fft // 2}")
    n = (T - 1) * hop_length
    if init is None:
        recon = torch.randn(B, n, device=mag_fm.device, dtype=torch.float32, generator=generator)
    else:
        recon = init.to(mag_fm.device, torch.float32).reshape(B, n).contiguous()
    bufs = (torch.empty_like(mag_fm), torch.empty_like(mag_fm))
    prev = recon
    for _ in range(int(n_iter)):
        re, im = ops.stft_project(recon, mag_fm, n_fft, hop_length, out=bufs)
        prev = recon
        recon, _ = ops.istft(re, im, PG_SPEC_CARTESIAN, n_fft, hop_length, normalize=False, check_finite=False)
    loss = torch.sqrt(torch.sum((recon - prev) ** 2, dim=1) / n) if n_iter > 0 else None
    return recon, (bufs if n_iter > 0 else None), loss


def griffin_lim(spec, n_fft, hop_length, n_iter, init=None):
    """Griffin-Lim phase retrieval on the GPU (utils.py:85-134): n_iter x (STFT -> keep phase, impose
    ``spec`` -> ISTFT), then the finiteness check and peak normalisation of utils.py:131-132.  ``spec`` is
    the DC-less magnitude ``[C, T]``; returns (audio, new_spec, loss) like the reference.  ``init``
    (optional, ``[(T-1)*hop]``) replaces the random start vector of utils.py:116.

    Deviation, stated: the reference hands the DC-less matrix straight to librosa.istft
    (utils.py:114,127), which then infers n_fft' = 2*(C-1) (2046) -- a latent bug that keeps
    its loop from converging.  This version treats ``spec`` the way
    ``generate_audio`` does (zero DC row, n_fft = 2*C), i.e. the evident intent; it is therefore
    NOT bit-comparable with the reference's Griffin-Lim output and is not on the parity path."""
    dev = _device()
    mag = torch.from_numpy(np.ascontiguousarray(np.asarray(spec), dtype=np.float32)).to(dev)
    if 2 * mag.shape[0] != n_fft:
        raise RuntimeError(f"griffin_lim: spec has {mag.shape[0]} rows, expected n_fft/2 = {n_fft // 2}")
    mag_fm = ops.transpose(mag.unsqueeze(0))                               # [1, T, C]
    if init is not None:
        init = torch.from_numpy(np.ascontiguousarray(init, dtype=np.float32)).view(1, -1)
    recon, planes, loss = griffin_lim_batch(mag_fm, n_fft, hop_length, n_iter, init=init)
    if not bool(torch.isfinite(recon).all()):
        raise ValueError("Audio buffer is not finite everywhere")
    peak = recon.abs().max()
    if float(peak) >= np.finfo(np.float32).tiny:
        recon = recon / peak
    new_spec = None
    if planes is not None:
        new_spec = (ops.transpose(planes[0])[0] + 1j * ops.transpose(planes[1])[0]).cpu().numpy()
    return recon[0].cpu().numpy(), new_spec, (float(loss[0]) if loss is not None else None)


# ------------------------------------------------------------------ plotting (utils.py:46-83,136-143)
def generate_spec_img(spec, is_stft=False, is_amp=False):
    """dB spectrogram rendered to an RGB uint8 array (needs matplotlib; visualisation only)."""
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    if is_amp:
        D = np.asarray(spec)
    else:
        z = spec if is_stft else spec[0, ...] + 1j * spec[1, ...]
        a = np.abs(z)
        D = 20.0 * np.log10(np.maximum(a, 1e-5)) - 20.0 * np.log10(max(float(a.max()), 1e-5))
        D = np.maximum(D, D.max() - 80.0)
    fig = plt.figure(figsize=(3, 2))
    plt.imshow(D, origin="lower", aspect="auto")
    plt.colorbar()
    fig.canvas.draw()
    img = np.asarray(fig.canvas.buffer_rgba())[..., :3].copy()
    plt.close()
    return img


def generate_waveplot(audio, sr):
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    fig = plt.figure(figsize=(3, 2))
    plt.plot(np.arange(len(audio)) / float(sr), audio, linewidth=0.5)
    fig.canvas.draw()
    img = np.asarray(fig.canvas.buffer_rgba())[..., :3].copy()
    plt.close()
    return img


# ------------------------------------------------------------- names model.py:7 imports (utils.py:145-262)
class View(nn.Module):
    def __init__(self, *shape):
        super().__init__()
        self.shape = shape

    def forward(self, input):
        return input.view(*self.shape)


class Flatten(nn.Module):
    def forward(self, input):
        return input.view(input.size(0), -1)


class Transpose(nn.Module):
    def __init__(self, dim0, dim1):
        super().__init__()
        self.dim0, self.dim1 = dim0, dim1

    def forward(self, input):
        return input.transpose(self.dim0, self.dim1).contiguous()


class EnergyLoss(nn.Module):
    """MSE between the amplitudes sqrt(re^2 + im^2 + 1e-10) of two [B, 2, ...] tensors."""

    def __init__(self, tensor=torch.FloatTensor):
        super().__init__()
        self.tensor = tensor
        self.loss = nn.MSELoss()

    @staticmethod
    def _calc_amp(a):
        return torch.sqrt(a[:, 0, ...] ** 2 + a[:, 1, ...] ** 2 + 1e-10)

    def __call__(self, a, b):
        return self.loss(self._calc_amp(a), self._calc_amp(b))


class GANLoss(nn.Module):
    """Least-squares GAN loss against a constant real/fake label."""

    def __init__(self, real_label=1., fake_label=0., tensor=torch.FloatTensor):
        super().__init__()
        self.tensor = tensor
        self.real_label, self.fake_label = real_label, fake_label
        self.real_var = self.fake_var = None
        self.loss = nn.MSELoss()

    def get_target(self, input, is_real):
        label = self.real_label if is_real else self.fake_label
        cached = self.real_var if is_real else self.fake_var
        if cached is None or cached.numel() != input.numel():
            cached = torch.full(input.size(), float(label), dtype=input.dtype, device=input.device)
            if is_real:
                self.real_var = cached
            else:
                self.fake_var = cached
        return cached

    def __call__(self, input, is_real):
        return self.loss(input, self.get_target(input, is_real))


class Pool(object):
    """History buffer of generated samples (image-pool trick): once full, each new sample is
    swapped with a random stored one with probability 1/2."""

    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.n = 0
        self.samples = []

    def draw(self, samples):
        if self.pool_size == 0:
            return samples
        out = []
        for s in samples:
            s = torch.unsqueeze(s, 0)
            if self.n < self.pool_size:
                self.n += 1
                self.samples.append(s)
                out.append(s)
            elif np.random.uniform() > 0.5:
                i = np.random.randint(0, self.pool_size - 1)
                out.append(self.samples[i].clone())
                self.samples[i] = s
            else:
                out.append(s)
        return torch.cat(out, 0)

    def get_samples(self, n_sample):
        if self.n < 1:
            raise RuntimeError("Empty pool!")
        if self.n == 1:
            return torch.cat([self.samples[0]], 0)
        return torch.cat([self.samples[np.random.randint(0, self.n - 1)] for _ in range(n_sample)], 0)
