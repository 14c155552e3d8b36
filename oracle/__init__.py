"""CPU oracle for the magnitude -> phase -> waveform path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic the reference (LemonATsu/UNet-PhaseGen)
performs on its hot path.  It exists to *check* the CUDA product path; it is never part of
it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.

Parity status: **parity unpinned by the reference's own tests** -- the reference ships no
tests, fixtures or golden vectors.  The oracle is pinned instead
against outputs of the reference itself: ``oracle/make_golden.py`` imports the unmodified
``/root/reference/model.py`` (with stub ``matplotlib``/``librosa`` modules) in the build
container, runs it on seeded inputs and stores inputs + weights + outputs under
``tests/golden/``; ``tests/test_oracle_unet.py`` checks the restatement in
``oracle/unet_torch.py`` against those files.  The STFT/ISTFT arithmetic lives in librosa
(not vendored, not pinned, not installable here); ``oracle/stft_np.py`` restates its
published algorithm in float64 numpy and is cross-checked against ``torch.stft`` /
``torch.istft`` on the CPU.
"""
