"""wave -> STFT -> U-Net -> ISTFT -> wave with every intermediate resident in HBM.

This is the path BASELINE.json's metric is quoted on: the work of the reference's
preproc_mdb.py:93 (STFT), data.py:39-47 (log-magnitude), demo.py:33-42 (per-clip U-Net
forward, polar->complex, generate_audio) for a whole batch of clips, without the
numpy/CPU round trips the reference makes at demo.py:37-40.
"""
import torch

from . import ops
from ._lib import PG_SPEC_POLAR_LOG, PG_STFT_LOGMAG


class PhaseGenPipeline:
    def __init__(self, model, n_fft, hop, precision=None, per_clip=True, phase_only=True, normalize=True):
        ops.check_stft_geometry(n_fft, hop)
        self.model, self.n_fft, self.hop = model, n_fft, hop
        self.per_clip, self.phase_only, self.normalize = per_clip, phase_only, normalize
        self.precision = precision

    def frames(self, n_samples):
        return 1 + n_samples // self.hop

    def __call__(self, wave, check_finite=False, return_intermediates=False):
        """wave float32 [B, N] on the GPU, N = (T-1)*hop with T % 8 == 0 -> float32 [B, N]."""
        if not wave.is_cuda:
            raise RuntimeError("PhaseGenPipeline needs CUDA tensors: there is no CPU fallback")
        B, N = wave.shape
        T = self.frames(N)
        kw = {"precision": self.precision} if self.precision else {}
        ex = self.model.executor(B, T, wave.device, per_clip=self.per_clip, phase_only=self.phase_only, **kw)
        x0 = ex.x0
        if x0.lo is not None or x0.hi.dtype == torch.bfloat16:
            # the STFT kernel writes the first convolution's bf16 operand planes directly
            logmag, _ = ops.stft(wave, self.n_fft, self.hop, PG_STFT_LOGMAG, want_second=False,
                                 operand=(x0.hi, x0.lo, x0.rows * x0.ld))
        else:
            logmag, _ = ops.stft(wave, self.n_fft, self.hop, PG_STFT_LOGMAG, want_second=False)
            ex.load_input_cl(logmag)
        dn, up = self.model._norm_params(wave.device)
        out = ex.run(dn, up)                                   # [B, T, C] phase (or [B, T, 2C])
        C = self.n_fft // 2
        phase = out if out.shape[2] == C else out[:, :, :C].contiguous()
        audio, peak = ops.istft(logmag, phase, PG_SPEC_POLAR_LOG, self.n_fft, self.hop,
                                normalize=self.normalize, check_finite=check_finite)
        if return_intermediates:
            return audio, logmag, phase
        return audio
