"""End-to-end magnitude -> phase -> waveform on the GPU vs the oracle's chain (preproc_mdb.py:93 ->
data.py:39-47 -> model.py -> demo.py:38-40), including the waveform-SNR criterion of BASELINE.json
(SNR against the clean input within 0.1 dB of the oracle's)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import stft_np, unet_torch  # noqa: E402


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def snr_db(ref, est):
    ref = np.asarray(ref, np.float64); est = np.asarray(est, np.float64)
    g = np.dot(ref, est) / max(np.dot(est, est), 1e-300)        # peak normalisation changes the gain
    return 10 * np.log10(np.sum(ref ** 2) / np.sum((ref - g * est) ** 2))


@pytest.mark.parametrize("n_fft,T,prec", [(256, 40, "bf16x3"), (256, 40, "fp32_simt"), (512, 24, "bf16x3")])
def test_pipeline_matches_oracle_chain(n_fft, T, prec):
    import model
    from phasegen import synth
    from phasegen.pipeline import PhaseGenPipeline
    hop, C, B = n_fft // 4, n_fft // 2, 2
    torch.manual_seed(1)
    net = model.UNetModel(C, 2 * C).cuda()
    synth.randomize_norm_affine(net, seed=2)
    sd = {k: v.detach().cpu() for k, v in net.model.state_dict().items()}
    wave = synth.synthetic_waves(B, (T - 1) * hop, sr=16000, seed=3, device="cuda")
    pipe = PhaseGenPipeline(net, n_fft, hop, precision=prec, per_clip=True, phase_only=True)
    audio, logmag, phase = pipe(wave, check_finite=True, return_intermediates=True)
    for b in range(B):
        w = wave[b].cpu().numpy().astype(np.float64)
        lm = np.log1p(np.abs(stft_np.stft(w, n_fft, hop)[1:]))
        out = unet_torch.unet_forward(sd, torch.from_numpy(lm)[None], torch.float64, per_clip_bn=True)[0].numpy()
        ref = stft_np.generate_audio(stft_np.polar_to_complex(lm, out[:C]), 16000, hop, is_stft=True)
        assert rel_l2(logmag[b].cpu().numpy().T, lm) < 1e-4
        assert rel_l2(phase[b].cpu().numpy().T, out[:C]) < 1e-3
        got = audio[b].cpu().numpy()
        assert rel_l2(got, ref) < 5e-3
        assert abs(snr_db(w, got) - snr_db(w, ref)) < 0.1


def test_long_form_windows_match_pipeline_and_seams_sum_to_one():
    """Config 4 (long-form): every window equals the plain pipeline on that window, the seam
    weights sum to one (a constant signal stitches back to itself), ranks partition the windows."""
    import model
    from phasegen import longform, synth
    from phasegen.pipeline import PhaseGenPipeline
    n_fft, hop, frames = 256, 64, 40
    C = n_fft // 2
    torch.manual_seed(3)
    net = model.UNetModel(C, 2 * C).cuda()
    pipe = PhaseGenPipeline(net, n_fft, hop, per_clip=True, phase_only=True, normalize=False)
    N = 5 * (frames - 1) * hop // 2 + 777                        # ragged: last window is zero-padded
    wave = synth.synthetic_waves(1, N, sr=16000, seed=4, device="cuda")[0]
    win, step, n = longform.window_plan(N, hop, frames)
    assert step % hop == 0 and 2 * step >= win and win + (n - 1) * step >= N
    wins, idx = longform.cut_windows(wave, hop, frames)
    assert wins.shape == (n, win) and idx == list(range(n))
    parts = [longform.cut_windows(wave, hop, frames, r, 3) for r in range(3)]
    assert sorted(i for _, ix in parts for i in ix) == list(range(n))
    ones = longform.stitch(torch.ones(n, win, device="cuda"), idx, n, N, hop, frames)
    assert float((ones - 1).abs().max()) < 1e-6
    y = longform.process_long(pipe, wave, frames=frames, batch=4, peak_normalize=False)
    per_window = pipe(wins).clone()
    ref = longform.stitch(per_window, idx, n, N, hop, frames)
    assert y.shape == (N,) and torch.allclose(y, ref, atol=1e-6)
    single = pipe(wins[2:3].contiguous())
    assert float((single[0] - per_window[2]).abs().max()) < 1e-5 * float(per_window[2].abs().max()) + 1e-7


@pytest.mark.parametrize("prec", ["bf16x3", "f16mix", "f16mix1"])
@pytest.mark.parametrize("n_fft,T", [(512, 1384), (1024, 696), (2048, 352)])
def test_shape_sweep_full_path_one_clip(n_fft, T, prec):
    """Config 5 shapes (n_fft 512 / hop 128 / T 1384 and n_fft 2048 / hop 512 / T 352) through the
    whole path at batch 1, against the oracle chain, in the all-three-product mode and in the bench's
    default mode (f16mix): BASELINE.json's bounds at BASELINE.json's full sizes."""
    import model
    from phasegen import synth
    from phasegen.pipeline import PhaseGenPipeline
    hop, C = n_fft // 4, n_fft // 2
    torch.manual_seed(5)
    net = model.UNetModel(C, 2 * C).cuda()
    synth.randomize_norm_affine(net, seed=6)
    sd = {k: v.detach().cpu() for k, v in net.model.state_dict().items()}
    wave = synth.synthetic_waves(1, (T - 1) * hop, sr=44100, seed=7, device="cuda")
    pipe = PhaseGenPipeline(net, n_fft, hop, precision=prec, per_clip=True, phase_only=True)
    audio, logmag, phase = pipe(wave, check_finite=True, return_intermediates=True)
    w = wave[0].cpu().numpy().astype(np.float64)
    lm = np.log1p(np.abs(stft_np.stft(w, n_fft, hop)[1:]))
    out = unet_torch.unet_forward(sd, torch.from_numpy(lm)[None], torch.float64, per_clip_bn=True)[0].numpy()
    ref = stft_np.generate_audio(stft_np.polar_to_complex(lm, out[:C]), 44100, hop, is_stft=True)
    assert rel_l2(logmag[0].cpu().numpy().T, lm) < 1e-4
    e_phase = rel_l2(phase[0].cpu().numpy().T, out[:C])
    print(f"n_fft {n_fft} T {T} {prec}: predicted-phase rel-L2 {e_phase:.2e}, wave rel-L2 {rel_l2(audio[0].cpu().numpy(), ref):.2e}")
    assert e_phase < 1e-3
    assert abs(snr_db(w, audio[0].cpu().numpy()) - snr_db(w, ref)) < 0.1


def test_full_size_batch_is_clip_independent():
    """BASELINE shape (C 512, T 696), bench precision: a clip's output does not depend on the batch it travels in
    (per-clip statistics; two-clip weight tiles, 74 CTA pairs looping over many tiles) -- 48 clips at once against
    the same clips three at a time."""
    import model
    from phasegen import synth
    from phasegen.pipeline import PhaseGenPipeline
    n_fft, hop, T, B = 1024, 256, 696, 48
    torch.manual_seed(12)
    net = model.UNetModel(n_fft // 2, n_fft).cuda()
    synth.randomize_norm_affine(net, seed=13)
    pipe = PhaseGenPipeline(net, n_fft, hop, precision="f16mix", per_clip=True, phase_only=True)
    wave = synth.synthetic_waves(B, (T - 1) * hop, sr=44100, seed=14, device="cuda")
    full = pipe(wave).clone()
    assert bool(torch.isfinite(full).all())
    for i in (0, 21, 45):
        part = pipe(wave[i:i + 3].contiguous())
        assert float((part - full[i:i + 3]).norm() / full[i:i + 3].norm()) < 1e-4, i


def test_host_buffer_call_equals_device_call():
    """PhaseGenPipeline.run_host (pinned host in/out, chunked and overlapped copies) returns exactly
    what the device-tensor call returns, ragged last chunk included."""
    import model
    from phasegen import synth
    from phasegen.pipeline import PhaseGenPipeline
    n_fft, hop, T, B = 256, 64, 24, 7
    torch.manual_seed(8)
    net = model.UNetModel(n_fft // 2, n_fft).cuda()
    pipe = PhaseGenPipeline(net, n_fft, hop)
    h_in = synth.synthetic_waves(B, (T - 1) * hop, sr=16000, seed=9).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    pipe.run_host(h_in, h_out, chunks=3)
    torch.cuda.synchronize()
    ref = torch.cat([pipe(h_in[i:i + 3].cuda()).cpu() for i in range(0, B, 3)])
    assert torch.equal(h_out, ref)
    # explicit (unequal) sub-batch sizes
    h_out2 = torch.empty_like(h_in).pin_memory()
    pipe.run_host(h_in, h_out2, chunks=[1, 4, 2])
    torch.cuda.synchronize()
    ref2 = torch.cat([pipe(h_in[a:b].cuda()).cpu() for a, b in ((0, 1), (1, 5), (5, 7))])
    assert torch.equal(h_out2, ref2)
    with pytest.raises(RuntimeError, match="add up"):
        pipe.run_host(h_in, h_out2, chunks=[3, 3])
    assert sum(pipe.suggest_chunks(B, h_in.shape[1], "cuda")) == B
    # stream of batches (pipelined=True): calls are not ordered against the current stream, each returns the event of its last
    # download; different inputs per call and alternating output buffers, six calls back to back
    ins = [synth.synthetic_waves(B, (T - 1) * hop, sr=16000, seed=20 + i).pin_memory() for i in range(3)]
    outs = [torch.empty_like(h_in).pin_memory() for _ in range(2)]
    want = [torch.cat([pipe(x[a:b].cuda()).cpu() for a, b in ((0, 3), (3, 6), (6, 7))]) for x in ins]
    torch.cuda.synchronize()
    pend, got = [None, None], []
    for i in range(6):
        k = i & 1
        if pend[k] is not None:
            pend[k][0].synchronize()
            got.append((pend[k][1], outs[k].clone()))
        pend[k] = (pipe.run_host(ins[i % 3], outs[k], chunks=[3, 3, 1], pipelined=True), i % 3)
    for k in ((6 & 1), (7 & 1)):
        pend[k][0].synchronize()
        got.append((pend[k][1], outs[k].clone()))
    assert len(got) == 6
    for j, y in got:
        assert torch.equal(y, want[j])


def test_online_training_pairs_match_the_offline_stage():
    """datapath.TrainingPairProducer vs preproc_mdb.py:66-97,182 + data.py:39-47 restated with the numpy oracle."""
    from phasegen.datapath import TrainingPairProducer, chunk_starts
    from oracle import stft_np
    n_fft, hop, t_slice = 512, 128, 127 * 128
    rng = np.random.default_rng(3)
    audio = (0.3 * rng.standard_normal(5 * t_slice + 1000)).astype(np.float32)
    starts = chunk_starts(len(audio), t_slice, 1, np.random.default_rng(4))
    assert starts[0] == 0 and len(starts) == 2 * 6 and starts.max() + t_slice > len(audio)   # the last regular chunk is padded
    prod = TrainingPairProducer(audio, t_slice, n_fft, hop)
    assert prod.frames == 128
    # oracle: chunk, zero pad, STFT, drop DC, standardise re/im over the whole set, log1p|.|, angle
    specs = []
    for s0 in starts:
        c = audio[s0:s0 + t_slice].astype(np.float64)
        c = np.pad(c, (0, t_slice - len(c)))
        specs.append(stft_np.stft(c, n_fft, hop)[1:])
    S = np.stack(specs)
    vals = np.stack([S.real, S.imag], 1)
    mean, std = vals.mean(), vals.std()
    m_gpu, s_gpu = prod.stats(starts, batch=5)
    assert abs(m_gpu - mean) < 1e-5 * std and abs(s_gpu - std) < 1e-5 * std
    Z = (S.real - mean) / std + 1j * (S.imag - mean) / std
    lm, ph = prod.pairs(starts, mean, std)
    assert lm.shape == (len(starts), 128, n_fft // 2)
    lm = lm.cpu().numpy().transpose(0, 2, 1); ph = ph.cpu().numpy().transpose(0, 2, 1)
    assert rel_l2(lm, np.log1p(np.abs(Z))) < 1e-4
    Zg = np.expm1(lm) * np.exp(1j * ph)
    assert np.linalg.norm(Zg - Z) / np.linalg.norm(Z) < 1e-4


def test_validation_report_matches_the_reference_loop():
    """phasegen.validate vs train.py:69-122 restated with the oracle (hybrid / no-phase / Griffin-Lim audio)."""
    import model
    from oracle import stft_np, unet_torch
    from phasegen import synth
    from phasegen.validate import validation_report
    C, T, V = 128, 40, 2
    n_fft, hop = 2 * C, C // 2
    torch.manual_seed(9)
    net = model.UNetModel(C, 2 * C).cuda()
    synth.randomize_norm_affine(net, seed=9)
    net.precision = "fp32_simt"
    rng = np.random.default_rng(9)
    vals = []
    for _ in range(V):
        S = stft_np.stft(rng.standard_normal((T - 1) * hop), n_fft, hop)[1:]
        vals.append(np.stack([np.log1p(np.abs(S)), np.angle(S)]))
    val = np.stack(vals).astype(np.float32)
    init = rng.standard_normal((V, (T - 1) * hop)).astype(np.float32)
    n_it = 5
    rep = validation_report(net, val, n_fft, hop, gl_iters=n_it, gl_init=torch.from_numpy(init), return_audio=True)
    sd = {k: v.detach().cpu() for k, v in net.model.state_dict().items()}
    mses, nops, lims = [], [], []
    for v in range(V):
        lm, ph = val[v, 0].astype(np.float64), val[v, 1].astype(np.float64)
        pred = unet_torch.unet_forward(sd, torch.from_numpy(lm)[None], torch.float64, per_clip_bn=True)[0, :C].numpy()
        mag = np.expm1(lm)
        orig = stft_np.generate_audio(mag * np.exp(1j * ph), 0, hop, is_stft=True)
        hyb = stft_np.generate_audio(mag * np.exp(1j * pred), 0, hop, is_stft=True)
        nop = stft_np.generate_audio(mag.astype(np.complex128), 0, hop, is_stft=True)
        recon = init[v].astype(np.float64)
        for _ in range(n_it):
            rs = stft_np.stft(recon, n_fft, hop)[1:]
            recon = stft_np.istft(np.concatenate([np.zeros((1, T)), mag * np.exp(1j * np.angle(rs))]), hop)
        lim = stft_np.peak_normalize(recon)
        assert rel_l2(rep["audio"]["hybrid"][v].cpu().numpy(), hyb) < 1e-3
        mses.extend(np.abs(orig - hyb)); nops.extend(np.abs(orig - nop)); lims.extend(np.abs(orig - lim))
    assert abs(rep["MSE"] - np.mean(mses)) < 2e-3 * np.mean(mses)
    assert abs(rep["NOPMSE"] - np.mean(nops)) < 1e-3 * np.mean(nops)
    assert abs(rep["LMSE"] - np.mean(lims)) < 1e-3 * np.mean(lims)


def test_cuda_graph_capture_replays_the_pipeline():
    """PhaseGenPipeline.capture: one cudaGraphLaunch per call, bit-identical to the eager call, for new inputs too."""
    import model
    from phasegen import _lib, synth
    from phasegen.pipeline import PhaseGenPipeline
    n_fft, hop, T, B = 1024, 256, 40, 2
    torch.manual_seed(31)
    net = model.UNetModel(n_fft // 2, n_fft).cuda()
    pipe = PhaseGenPipeline(net, n_fft, hop, precision="f16mix")
    w1 = synth.synthetic_waves(B, (T - 1) * hop, sr=44100, seed=32, device="cuda")
    w2 = synth.synthetic_waves(B, (T - 1) * hop, sr=44100, seed=33, device="cuda")
    e1, e2 = pipe(w1).clone(), pipe(w2).clone()
    g = pipe.capture(B, (T - 1) * hop)
    n0 = _lib.launches
    a1 = g(w1).clone()
    a2 = g(w2).clone()
    a1b = g(w1).clone()
    torch.cuda.synchronize()
    assert _lib.launches == n0                                  # replays launch nothing through the binding
    assert torch.equal(a1, e1) and torch.equal(a2, e2) and torch.equal(a1b, e1)
    with pytest.raises(RuntimeError, match="captured"):
        g(w1[:1])


def test_fp16_range_guard_raises_or_falls_back():
    """Activations beyond the fp16 range (first-layer weights scaled by 1e7: d1 has no norm, so its output -- an fp16
    operand of d2 in the fp16 modes -- leaves the 65504 range): a checked call raises OverflowError, or, with
    fp16_overflow="fallback", re-runs in bf16x3 (bf16 planes, fp32 range) and matches the oracle; an unchecked call leaves
    the sticky flag for range_overflow()."""
    import model
    from phasegen import synth
    from phasegen.pipeline import PhaseGenPipeline
    n_fft, hop, T, B = 256, 64, 40, 2
    C = n_fft // 2
    torch.manual_seed(51)
    net = model.UNetModel(C, 2 * C).cuda()
    synth.randomize_norm_affine(net, seed=52)
    with torch.no_grad():
        net.model._parts["down"].weight.mul_(1.0e7)
    sd = {k: v.detach().cpu() for k, v in net.model.state_dict().items()}
    wave = synth.synthetic_waves(B, (T - 1) * hop, sr=16000, seed=53, device="cuda")
    strict = PhaseGenPipeline(net, n_fft, hop, precision="f16mix", per_clip=True, phase_only=True)
    strict(wave)                                            # unchecked: runs, flag stays
    assert strict.range_overflow() and not strict.range_overflow()
    with pytest.raises(OverflowError, match="fp16 range"):
        strict(wave, check_finite=True)
    soft = PhaseGenPipeline(net, n_fft, hop, precision="f16mix", per_clip=True, phase_only=True, fp16_overflow="fallback")
    audio, logmag, phase = soft(wave, check_finite=True, return_intermediates=True)
    for b in range(B):
        w = wave[b].cpu().numpy().astype(np.float64)
        lm = np.log1p(np.abs(stft_np.stft(w, n_fft, hop)[1:]))
        out = unet_torch.unet_forward(sd, torch.from_numpy(lm)[None], torch.float64, per_clip_bn=True)[0].numpy()
        assert rel_l2(phase[b].cpu().numpy().T, out[:C]) < 1e-3
    ok = PhaseGenPipeline(model.UNetModel(C, 2 * C).cuda(), n_fft, hop, precision="f16mix", per_clip=True, phase_only=True)
    ok(wave, check_finite=True)                             # ordinary weights: no flag
    assert not ok.range_overflow()


def test_stitch_kernel_matches_host_form_and_round_robin_layout():
    """pg_stitch against the float64 host form of the same cross-fade, and its multi-rank input layout: windows dealt
    round-robin to W ranks and all-gathered (window i at slot (i mod W) * per_rank + i div W, ragged last round zero)
    stitch to exactly what the plain order gives; the peak output feeds one global normalisation."""
    from phasegen import longform
    hop, frames = 64, 40
    win, step, _ = longform.window_plan(10 ** 6, hop, frames)
    n = 11
    N = win + (n - 1) * step - 37                              # ragged tail
    g = torch.Generator().manual_seed(61)
    wins = torch.randn(n, win, generator=g)
    ref = longform.stitch(wins, list(range(n)), n, N, hop, frames)                       # host form (float64 accumulate)
    peak = torch.zeros(1, device="cuda")
    got = longform.stitch(wins.cuda(), list(range(n)), n, N, hop, frames, peak_out=peak)
    assert got.shape == (N,) and float((got.cpu() - ref).abs().max()) < 2e-6
    assert abs(float(peak) - float(ref.abs().max())) < 1e-5
    for W in (2, 3, 8):
        per_rank = -(-n // W)
        gathered = torch.zeros(W * per_rank, win)
        for i in range(n):
            gathered[(i % W) * per_rank + i // W] = wins[i]
        rr = longform.stitch(gathered.cuda(), list(range(n)), n, N, hop, frames, world=W, per_rank=per_rank)
        assert torch.equal(rr, got), W
        assert float((longform.stitch(gathered, list(range(n)), n, N, hop, frames, world=W, per_rank=per_rank) - ref).abs().max()) == 0
    with pytest.raises(RuntimeError, match="step must lie"):
        from phasegen import _lib, ops
        o = torch.zeros(100, device="cuda")
        _lib.call("pg_stitch", ops._ptr(wins.cuda()), 2, 100, 30, 1, 2, ops._ptr(o), 100, None, ops._stream())
