"""oracle/stft_np.py vs CPU torch.stft/istft fixtures (tests/golden/stft_*.npz) and its own
round-trip properties."""
import glob
import os

import numpy as np
import pytest

from oracle import stft_np

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "stft_*.npz")))


def rel_l2(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


def _geom(path):
    parts = os.path.basename(path)[:-4].split("_")
    return int(parts[1][1:]), int(parts[2][1:])


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_stft_matches_torch(path):
    z = np.load(path); n_fft, hop = _geom(path)
    S = stft_np.stft(z["y"], n_fft, hop)
    assert S.shape == z["S"].shape
    assert rel_l2(S, z["S"]) < 1e-12


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_istft_matches_torch(path):
    z = np.load(path); n_fft, hop = _geom(path)
    w = stft_np.istft(z["z"], hop)
    assert w.shape == z["w"].shape
    assert rel_l2(w, z["w"]) < 1e-12


@pytest.mark.parametrize("n_fft", [512, 1024, 2048])
def test_round_trip(n_fft):
    hop = n_fft // 4
    rng = np.random.default_rng(n_fft)
    y = rng.standard_normal(hop * 31)
    back = stft_np.istft(stft_np.stft(y, n_fft, hop), hop)
    assert back.shape == y.shape
    assert rel_l2(back, y) < 1e-12


def test_generate_audio_semantics():
    rng = np.random.default_rng(0)
    C, T, hop = 256, 24, 128
    z = rng.standard_normal((C, T)) + 1j * rng.standard_normal((C, T))
    a = stft_np.generate_audio(z, 16000, hop, is_stft=True)
    b = stft_np.generate_audio(np.stack([z.real, z.imag]), 16000, hop, is_stft=False)
    assert a.dtype == np.float32 and a.shape == ((T - 1) * hop,)
    assert np.array_equal(a, b)
    assert abs(np.max(np.abs(a)) - 1.0) < 1e-6
    zero = stft_np.generate_audio(np.zeros((C, T), complex), 16000, hop, is_stft=True)
    assert not zero.any()
    bad = z.copy(); bad[3, 3] = np.nan
    with pytest.raises(ValueError):
        stft_np.generate_audio(bad, 16000, hop, is_stft=True)


def test_spec_and_angle_and_polar_inverse():
    rng = np.random.default_rng(1)
    reim = rng.standard_normal((2, 16, 8)).astype(np.float32)
    sa = stft_np.spec_and_angle(reim)
    z = stft_np.polar_to_complex(sa[0], sa[1])
    assert rel_l2(z, reim[0] + 1j * reim[1]) < 1e-12


def test_griffin_lim_keeps_reference_quirk():
    """utils.py:114,127 hand the DC-less magnitude to istft, which then infers
    n_fft' = 2*(C-1); the loop therefore does not converge -- only shapes, peak
    normalisation and finiteness are contractual."""
    rng = np.random.default_rng(2)
    n_fft, hop = 512, 128
    y = rng.standard_normal(hop * 23)
    mag = np.abs(stft_np.stft(y, n_fft, hop))[1:]
    a1, _, l1 = stft_np.griffin_lim(mag, n_fft, hop, 1, np.random.default_rng(3))
    a8, s8, l8 = stft_np.griffin_lim(mag, n_fft, hop, 8, np.random.default_rng(3))
    assert a8.shape == a1.shape and s8.shape == mag.shape
    assert np.isfinite(a8).all() and np.isfinite(l8) and np.isfinite(l1)
    assert abs(np.max(np.abs(a8)) - 1.0) < 1e-6


def test_window_is_the_one_librosa_asks_scipy_for():
    """librosa.stft/istft build their window with scipy.signal.get_window('hann', n_fft, fftbins=True); scipy IS in this
    image, so the oracle's window is pinned to the dependency's own provider."""
    from scipy.signal import get_window
    for n in (256, 512, 1024, 2048):
        assert np.array_equal(stft_np.hann_periodic(n), get_window("hann", n, fftbins=True)) or \
            np.max(np.abs(stft_np.hann_periodic(n) - get_window("hann", n, fftbins=True))) < 1e-15


@pytest.mark.parametrize("n_fft", [512, 1024, 2048])
def test_stft_istft_match_scipy(n_fft):
    """A third, independent implementation (scipy.signal.stft / istft with the librosa conventions: periodic Hann,
    hop n_fft/4, 'even' = reflect boundary extension of n_fft/2, no zero padding, unscaled spectrum).  scipy scales
    its spectrum by 1/sum(window), undone here; frames and bins then agree to rounding, and so do the inverses."""
    from scipy import signal
    hop = n_fft // 4
    rng = np.random.default_rng(n_fft + 1)
    y = rng.standard_normal(hop * 37)
    win = signal.get_window("hann", n_fft, fftbins=True)
    _, _, Z = signal.stft(y, window=win, nperseg=n_fft, noverlap=n_fft - hop, nfft=n_fft, boundary="even", padded=False,
                          return_onesided=True)
    S = stft_np.stft(y, n_fft, hop)
    assert Z.shape == S.shape
    assert rel_l2(Z * win.sum(), S) < 1e-12
    _, back = signal.istft(S / win.sum(), window=win, nperseg=n_fft, noverlap=n_fft - hop, nfft=n_fft, boundary=True,
                           input_onesided=True)
    mine = stft_np.istft(S, hop)
    assert rel_l2(back[:len(mine)], mine) < 1e-10


@pytest.mark.parametrize("n_fft", [512, 1024, 2048])
def test_stft_matches_the_librosa_compatible_numpy_stft_of_transformers(n_fft):
    """A fourth implementation: transformers.audio_utils.spectrogram (numpy framing + rfft; its documentation states it
    is compatible with librosa.stft, and its own test-suite pins it to librosa-generated values).  librosa itself is not
    installable here, so this -- with torch.stft, scipy.signal.stft and the scipy window provider librosa calls -- is
    how the restatement is held to the dependency's published behaviour."""
    audio_utils = pytest.importorskip("transformers.audio_utils")
    rng = np.random.default_rng(n_fft)
    w = rng.standard_normal(6 * n_fft + 123)
    win = audio_utils.window_function(n_fft, "hann", periodic=True)
    assert np.abs(win - stft_np.hann_periodic(n_fft)).max() < 1e-12 if hasattr(stft_np, "hann_periodic") else True
    S = audio_utils.spectrogram(w, win, frame_length=n_fft, hop_length=n_fft // 4, fft_length=n_fft, power=None,
                                center=True, pad_mode="reflect", onesided=True, dtype=np.float64)
    ref = stft_np.stft(w, n_fft, n_fft // 4)
    assert S.shape == ref.shape
    assert np.abs(S - ref).max() / np.abs(ref).max() < 1e-6       # the transformers routine returns complex64
