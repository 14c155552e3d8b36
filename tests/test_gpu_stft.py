"""GPU parity of the STFT / ISTFT kernels (through the C ABI) against oracle/stft_np.py, the
committed golden fixtures, and size-independent properties.  Tolerances are BASELINE.json's:
relative L2 <= 1e-4 on the STFT (log-)magnitude."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import stft_np  # noqa: E402

GEOMS = [(256, 64), (512, 128), (1024, 256), (2048, 512)]


def rel_l2(a, b):
    a = np.asarray(a); b = np.asarray(b)
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a = a.astype(dt); b = b.astype(dt)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _wave(n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    return (0.1 * rng.standard_normal(n) + 0.3 * np.sin(2 * np.pi * 0.013 * t) + 0.2 * np.sin(2 * np.pi * 0.21 * t + 1)).astype(np.float32)


@pytest.mark.parametrize("n_fft,hop", GEOMS)
@pytest.mark.parametrize("T", [24, 33, 8, 136])
def test_stft_matches_oracle(n_fft, hop, T):
    from phasegen import ops
    B = 3
    n = (T - 1) * hop + (hop // 3 if T == 33 else 0)           # ragged tail: extra samples that make no new frame
    w = np.stack([_wave(n, 10 * T + b) for b in range(B)])
    lm, ph = ops.stft(torch.from_numpy(w).cuda(), n_fft, hop, ops.PG_STFT_LOGMAG)
    re, im = ops.stft(torch.from_numpy(w).cuda(), n_fft, hop, ops.PG_STFT_REIM)
    assert lm.shape == (B, 1 + n // hop, n_fft // 2)
    for b in range(B):
        S = stft_np.stft(w[b], n_fft, hop)[1:]                 # DC row dropped (preproc_mdb.py:93)
        assert rel_l2(lm[b].cpu().numpy().T, np.log1p(np.abs(S))) < 1e-4
        assert rel_l2(re[b].cpu().numpy().T + 1j * im[b].cpu().numpy().T, S) < 1e-4
        z = np.exp(1j * ph[b].cpu().numpy().T.astype(np.float64)) * np.abs(S)
        assert rel_l2(z, S) < 1e-4                             # phase compared where it matters (|S|-weighted)


def test_stft_matches_golden_fixtures(golden_dir):
    from phasegen import ops
    for path in sorted(glob.glob(os.path.join(golden_dir, "stft_*.npz"))):
        z = np.load(path)
        parts = os.path.basename(path)[:-4].split("_")
        n_fft, hop = int(parts[1][1:]), int(parts[2][1:])
        re, im = ops.stft(torch.from_numpy(z["y"]).cuda()[None], n_fft, hop, ops.PG_STFT_REIM)
        got = re[0].cpu().numpy().T + 1j * im[0].cpu().numpy().T
        assert rel_l2(got, z["S"][1:]) < 1e-4, path
        # ISTFT of the fixture spectrum (DC row is zero there by construction)
        zz = z["z"][1:]
        a = torch.from_numpy(np.ascontiguousarray(zz.real.T, dtype=np.float32)).cuda()[None]
        b = torch.from_numpy(np.ascontiguousarray(zz.imag.T, dtype=np.float32)).cuda()[None]
        wv, _ = ops.istft(a, b, ops.PG_SPEC_CARTESIAN, n_fft, hop, normalize=False)
        assert rel_l2(wv[0].cpu().numpy(), z["w"]) < 1e-4, path


@pytest.mark.parametrize("n_fft,hop", GEOMS)
@pytest.mark.parametrize("T", [24, 40, 14])
def test_istft_matches_oracle_all_modes(n_fft, hop, T):
    from phasegen import ops
    C, B = n_fft // 2, 2
    rng = np.random.default_rng(n_fft + T)
    lm = np.abs(rng.standard_normal((B, T, C))).astype(np.float32)
    ph = (3.0 * rng.standard_normal((B, T, C))).astype(np.float32)    # raw, unwrapped phase like the net's
    a, b = torch.from_numpy(lm).cuda(), torch.from_numpy(ph).cuda()
    wv, peak = ops.istft(a, b, ops.PG_SPEC_POLAR_LOG, n_fft, hop, normalize=False)
    wn, _ = ops.istft(a, b, ops.PG_SPEC_POLAR_LOG, n_fft, hop, normalize=True)
    wm, _ = ops.istft(a, b, ops.PG_SPEC_POLAR_MAG, n_fft, hop, normalize=False)
    for i in range(B):
        z = stft_np.polar_to_complex(lm[i].T, ph[i].T)         # demo.py:39
        z = np.concatenate([np.zeros((1, T)), z])              # utils.py:38-39
        ref = stft_np.istft(z, hop)
        assert wv.shape[1] == ref.shape[0] == (T - 1) * hop
        assert rel_l2(wv[i].cpu().numpy(), ref) < 2e-5
        assert abs(float(peak[i]) - np.max(np.abs(ref))) < 1e-4 * np.max(np.abs(ref))
        assert rel_l2(wn[i].cpu().numpy(), stft_np.peak_normalize(ref)) < 2e-5
        zm = np.concatenate([np.zeros((1, T)), lm[i].T.astype(np.float64) * np.exp(1j * ph[i].T.astype(np.float64))])
        assert rel_l2(wm[i].cpu().numpy(), stft_np.istft(zm, hop)) < 2e-5


@pytest.mark.parametrize("n_fft,hop", GEOMS)
def test_round_trip_property(n_fft, hop):
    """STFT -> ISTFT reproduces the signal except for the DC bin this path drops; checked on a
    zero-mean-per-frame-free signal by comparing against the oracle's own DC-less round trip,
    at a size too large for the python oracle loops to be quick (full 4 s clip, batch 4)."""
    from phasegen import ops
    B, T = 4, 696
    w = torch.from_numpy(np.stack([_wave((T - 1) * hop, s) for s in range(B)])).cuda()
    re, im = ops.stft(w, n_fft, hop, ops.PG_STFT_REIM)
    back, _ = ops.istft(re, im, ops.PG_SPEC_CARTESIAN, n_fft, hop, normalize=False)
    w0 = w[0].cpu().numpy()
    S = stft_np.stft(w0, n_fft, hop); S[0] = 0
    assert rel_l2(back[0].cpu().numpy(), stft_np.istft(S, hop)) < 2e-5
    assert rel_l2(back[0].cpu().numpy(), w0) < 0.1             # DC-less, so close but not equal
    # linearity: STFT(a x + y) = a STFT(x) + STFT(y)
    re2, _ = ops.stft(2.5 * w[0:1] + w[1:2], n_fft, hop, ops.PG_STFT_REIM)
    assert rel_l2(re2[0].cpu().numpy(), 2.5 * re[0].cpu().numpy() + re[1].cpu().numpy()) < 1e-5


def test_generate_audio_semantics():
    """utils.generate_audio edge cases: (re, im) vs complex input, peak = 1, all-zero input stays
    zero, non-finite input raises (utils.py:41)."""
    import utils
    rng = np.random.default_rng(0)
    C, T, hop = 256, 24, 128
    z = (rng.standard_normal((C, T)) + 1j * rng.standard_normal((C, T))).astype(np.complex64)
    a = utils.generate_audio(z, 16000, hop, is_stft=True)
    b = utils.generate_audio(np.stack([z.real, z.imag]), 16000, hop, is_stft=False)
    ref = stft_np.generate_audio(z, 16000, hop, is_stft=True)
    assert a.dtype == np.float32 and a.shape == ((T - 1) * hop,)
    assert np.array_equal(a, b)
    assert rel_l2(a, ref) < 2e-5 and abs(np.max(np.abs(a)) - 1.0) < 1e-6
    assert not utils.generate_audio(np.zeros((C, T), np.complex64), 16000, hop, is_stft=True).any()
    bad = z.copy(); bad[3, 3] = np.nan
    with pytest.raises(ValueError):
        utils.generate_audio(bad, 16000, hop, is_stft=True)
    with pytest.raises(RuntimeError):
        utils.generate_audio(z[:100], 16000, hop, is_stft=True)   # n_fft = 200: unsupported, fails loudly


def test_stft_helpers_match_reference_formats():
    import utils
    w = _wave(127 * 512, 5)
    reim = utils.stft_nodc(w, 2048, 512)                        # what _chunk_and_stft stores
    assert reim.shape == (2, 1024, 128) and reim.dtype == np.float32
    ref = stft_np.stft_nodc_reim(w, 2048, 512)
    assert rel_l2(reim, ref) < 1e-4
    sa = utils.spec_and_angle_from_wave(w, 2048, 512)           # data.py:39-47
    assert rel_l2(sa[0], stft_np.spec_and_angle(ref)[0]) < 1e-4


def test_griffin_lim_matches_the_intended_algorithm():
    """utils.griffin_lim (fused projection kernel + ISTFT, all on the GPU) against the same loop written
    with the numpy oracle's stft/istft (zero DC row, n_fft = 2*C: the evident intent of utils.py:85-134),
    from the same start vector; batched entry point agrees with the single-clip one."""
    import utils
    n_fft, hop, T = 512, 128, 40
    rng = np.random.default_rng(11)
    w = rng.standard_normal((T - 1) * hop)
    S = stft_np.stft(w, n_fft, hop)[1:]
    mag = np.abs(S).astype(np.float32)
    mag[3, 5] = 0.0                                              # a zero-magnitude bin
    init = rng.standard_normal((T - 1) * hop).astype(np.float32)
    n_iter = 6
    recon = init.astype(np.float64)
    for _ in range(n_iter):
        rs = stft_np.stft(recon, n_fft, hop)[1:]
        new_spec = mag * np.exp(1j * np.angle(rs))
        prev = recon
        recon = stft_np.istft(np.concatenate([np.zeros((1, T)), new_spec]), hop)
    loss_ref = np.sqrt(np.sum((recon - prev) ** 2 / recon.size))
    a, spec_out, loss = utils.griffin_lim(mag, n_fft, hop, n_iter, init=init)
    assert a.dtype == np.float32 and a.shape == ((T - 1) * hop,)
    assert rel_l2(a, stft_np.peak_normalize(recon)) < 1e-4
    assert rel_l2(spec_out, new_spec) < 1e-4
    assert abs(loss - loss_ref) < 1e-4 * loss_ref
    # batch of two different clips == two single runs
    mag2 = np.abs(stft_np.stft(rng.standard_normal((T - 1) * hop), n_fft, hop)[1:]).astype(np.float32)
    init2 = rng.standard_normal((T - 1) * hop).astype(np.float32)
    mags = torch.from_numpy(np.stack([mag.T, mag2.T])).cuda().contiguous()
    rb, _, lb = utils.griffin_lim_batch(mags, n_fft, hop, n_iter, init=torch.from_numpy(np.stack([init, init2])))
    a2, _, l2 = utils.griffin_lim(mag2, n_fft, hop, n_iter, init=init2)
    rb = rb.cpu().numpy()
    assert rel_l2(rb[0] / np.abs(rb[0]).max(), a) < 1e-5 and rel_l2(rb[1] / np.abs(rb[1]).max(), a2) < 1e-5
    assert abs(float(lb[1]) - l2) < 1e-4 * l2


@pytest.mark.parametrize("n_fft,hop", [(1024, 256), (256, 64), (2048, 512)])
def test_stft_ragged_lengths_and_unaligned_clips(n_fft, hop):
    """Edge cases of the frame loader: an odd clip length (no float2 path, frame count 1 + N // hop with a
    partial last hop), a batch whose base pointer is only 4-byte aligned, the shortest legal clip (N = n_fft/2 + 1:
    every frame touches both reflections), and a frame count that is not a multiple of the frames per CTA."""
    from phasegen import ops
    rng = np.random.default_rng(n_fft)
    for N in (n_fft // 2 + 1, 5 * hop + 37, 41 * hop + 1, 33 * hop):
        w = rng.standard_normal((3, N)).astype(np.float32)
        buf = torch.zeros(3 * N + 1, device="cuda")
        buf[1:] = torch.from_numpy(w).reshape(-1).cuda()
        for wave in (torch.from_numpy(w).cuda(), buf[1:].view(3, N)):          # aligned base, then base + 4 bytes
            lm, ph = ops.stft(wave, n_fft, hop)
            assert lm.shape == (3, 1 + N // hop, n_fft // 2)
            for b in range(3):
                S = stft_np.stft(w[b], n_fft, hop)[1:]
                assert rel_l2(lm[b].cpu().numpy().T, np.log1p(np.abs(S))) < 1e-4, (N, b)
                z = np.expm1(lm[b].cpu().numpy().T.astype(np.float64)) * np.exp(1j * ph[b].cpu().numpy().T.astype(np.float64))
                assert np.linalg.norm(z - S) / np.linalg.norm(S) < 1e-4, (N, b)


@pytest.mark.parametrize("n_fft,hop", [(1024, 256), (512, 128)])
@pytest.mark.parametrize("T", [2, 3, 5, 86, 87, 171])
def test_istft_run_boundaries(n_fft, hop, T):
    """The sliding overlap-add at its seams: fewer frames than one iteration, and frame counts that end exactly on,
    one past and one short of a CTA's run of 11 * 8 - 3 = 85 hop-blocks (two CTAs per clip, halo frames on both sides)."""
    from phasegen import ops
    C = n_fft // 2
    rng = np.random.default_rng(T)
    re = rng.standard_normal((2, T, C)).astype(np.float32)
    im = rng.standard_normal((2, T, C)).astype(np.float32)
    wv, peak = ops.istft(torch.from_numpy(re).cuda(), torch.from_numpy(im).cuda(), ops.PG_SPEC_CARTESIAN, n_fft, hop, normalize=False)
    for b in range(2):
        z = np.concatenate([np.zeros((1, T)), re[b].T.astype(np.float64) + 1j * im[b].T.astype(np.float64)])
        ref = stft_np.istft(z, hop)
        assert wv.shape[1] == ref.shape[0]
        assert rel_l2(wv[b].cpu().numpy(), ref) < 2e-5
        assert abs(float(peak[b]) - np.abs(ref).max()) < 1e-4 * np.abs(ref).max()
