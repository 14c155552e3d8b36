"""float64 numpy restatement of the reference's STFT / ISTFT arithmetic.  TEST INFRASTRUCTURE.

The reference does not own this arithmetic: it calls librosa (unpinned, ~0.5-0.6 by API
usage, absent from this image).  Each function cites the reference call site it stands in
for and restates the librosa semantics that call relies on.
"""
import numpy as np


def hann_periodic(n):
    """scipy.signal.get_window('hann', n, fftbins=True): the window librosa.stft/istft use."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def stft(y, n_fft, hop):
    """``librosa.stft(c, n_fft=n_fft, hop_length=hop)`` as called at preproc_mdb.py:93 and
    utils.py:120: center=True, reflect padding by n_fft//2, periodic Hann of length n_fft,
    frames at t*hop, FFT, keep bins 0..n_fft/2.  Returns complex128 [1+n_fft/2, T] with
    T = 1 + len(y)//hop (librosa rounds this to complex64; callers here round if needed).
    """
    y = np.asarray(y, dtype=np.float64)
    pad = n_fft // 2
    yp = np.pad(y, (pad, pad), mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    frames = yp[idx] * hann_periodic(n_fft)[:, None]
    return np.fft.rfft(frames, axis=0)


def stft_nodc_reim(y, n_fft, hop):
    """preproc_mdb.py:84-97 ``_chunk_and_stft`` for one channel: STFT, delete the DC row
    (:93), split into real/imag planes (:94-95).  Returns float32 [2, n_fft/2, T]."""
    s = stft(y, n_fft, hop).astype(np.complex64)[1:]
    return np.stack([s.real, s.imag]).astype(np.float32)


def spec_and_angle(reim):
    """data.py:39-47 ``get_spec_and_angle`` for one item: [2,C,T] (re,im) ->
    [2,C,T] (log1p|z|, angle z)."""
    z = reim[0].astype(np.float64) + 1j * reim[1].astype(np.float64)
    return np.stack([np.log1p(np.abs(z)), np.angle(z)])


def polar_to_complex(logmag, phase):
    """demo.py:39 / train.py:83: ``(exp(mag) - 1) * exp(1j * phase)``."""
    return np.expm1(np.asarray(logmag, np.float64)) * np.exp(1j * np.asarray(phase, np.float64))


def istft(spec, hop):
    """``librosa.istft(stft, hop_length=hop)`` as called at utils.py:40,114,127.
    n_fft = 2*(rows-1); per-frame inverse real FFT (imaginary parts of the DC and Nyquist
    rows do not contribute), times the periodic Hann, overlap-add into n_fft + hop*(T-1)
    samples, divide by the window sum-of-squares where it exceeds ``tiny``, trim n_fft//2
    from both ends.  Returns float64 [(T-1)*hop].
    """
    spec = np.asarray(spec, dtype=np.complex128)
    n_fft = 2 * (spec.shape[0] - 1)
    T = spec.shape[1]
    win = hann_periodic(n_fft)
    frames = np.fft.irfft(spec, n=n_fft, axis=0) * win[:, None]
    n = n_fft + hop * (T - 1)
    y = np.zeros(n)
    wss = np.zeros(n)
    w2 = win * win
    for t in range(T):
        y[t * hop:t * hop + n_fft] += frames[:, t]
        wss[t * hop:t * hop + n_fft] += w2
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2: n - n_fft // 2]


def peak_normalize(y):
    """``librosa.util.normalize(audio, norm=np.inf, axis=None)`` (utils.py:42,132):
    divide by max|y|; an all-(near-)zero signal is returned unchanged."""
    y = np.asarray(y, dtype=np.float64)
    peak = np.max(np.abs(y)) if y.size else 0.0
    if peak < np.finfo(np.float32).tiny:
        return y.copy()
    return y / peak


def generate_audio(spec, sr, hop_length, is_stft=False):
    """utils.py:11-44.  ``spec`` is complex [C,T] (is_stft) or real [2,C,T] (re,im); a zero
    DC row is prepended (:38-39), ISTFT (:40), finiteness check (:41, ValueError here where
    librosa raises ParameterError), peak normalisation (:42).  float32 out like librosa."""
    spec = np.asarray(spec)
    z = spec if is_stft else spec[0] + 1j * spec[1]
    z = np.concatenate([np.zeros((1, z.shape[1]), z.dtype), z], axis=0)
    y = istft(z, hop_length)
    if not np.isfinite(y).all():
        raise ValueError("Audio buffer is not finite everywhere")
    return peak_normalize(y).astype(np.float32)


def griffin_lim(spec, n_fft, hop_length, n_iter, rng=None):
    """utils.py:85-134.  ``spec`` is the magnitude [C,T] *without* a DC row and the reference
    hands it to istft as is (:114,127), so the inverse transform there infers
    n_fft' = 2*(C-1) while the forward STFT uses ``n_fft`` and drops its DC row (:120-121).
    That quirk is kept.  The random start vector comes from ``rng`` (np.random in the
    reference, :116)."""
    rng = np.random.default_rng(0) if rng is None else rng
    spec = np.asarray(spec, np.float64)
    audio = istft(spec, hop_length)
    recon = rng.standard_normal(audio.shape[0])
    new_spec, loss = None, None
    for _ in range(n_iter):
        rs = stft(recon, n_fft, hop_length)[1:]
        new_spec = spec * np.exp(1j * np.angle(rs))
        prev = recon
        recon = istft(new_spec, hop_length)
        m = min(len(recon), len(prev))
        loss = np.sqrt(np.sum((recon[:m] - prev[:m]) ** 2 / recon.size))
    return peak_normalize(recon).astype(np.float32), new_spec, loss
