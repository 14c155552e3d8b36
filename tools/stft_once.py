"""One launch of each n_fft-1024 inference STFT / ISTFT kernel (scalar and frame-pair forms) for an ncu capture:
ncu --set full -k regex:stft -c 4 python tools/stft_once.py [clips]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unet-phasegen_b200")]

import torch  # noqa: E402

from phasegen import ops, synth  # noqa: E402

B, n_fft, hop, T = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 1024, 256, 696
N, C = (T - 1) * hop, n_fft // 2
wave = synth.synthetic_waves(B, N, 44100, seed=3).cuda()
hi = torch.zeros(B, T, C, device="cuda", dtype=torch.float16)
lo = torch.zeros_like(hi)
phase = 3.0 * torch.randn(B, T, C, device="cuda")
ss = torch.stack([1.0 + 0.1 * torch.randn(B, C, device="cuda"), 0.2 * torch.randn(B, C, device="cuda")], dim=-1).contiguous()
out = torch.empty(B, N, device="cuda")
ops.twiddle(n_fft, wave.device)
torch.cuda.synchronize()
for flag in ("0", "1"):
    os.environ["PG_STFT_PAIR"] = flag
    lm, _ = ops.stft(wave, n_fft, hop, ops.PG_STFT_LOGMAG, want_second=False, operand=(hi, lo, T * C))
    ops.istft(lm, phase, ops.PG_SPEC_POLAR_LOG, n_fft, hop, normalize=False, check_finite=False, out=out, b_scale_shift=ss)
torch.cuda.synchronize()
print("done")
