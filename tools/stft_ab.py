"""GPU A/B: the n_fft-1024 STFT / ISTFT inference kernels, one frame per warp (PG_STFT_PAIR=0) against frame pairs on
packed fp32x2 arithmetic (default), at the BASELINE batch (256 clips x 696 frames).  CUDA-event times, inputs larger
than L2; prints ms per launch, the HBM fractions by SURVEY section 8d's strict bytes, and the difference of the outputs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unet-phasegen_b200")]

import torch  # noqa: E402

from phasegen import ops, synth  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B, n_fft, hop, T = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 1024, 256, 696
    N, C = (T - 1) * hop, n_fft // 2
    wave = synth.synthetic_waves(B, N, 44100, seed=3).cuda()
    hi = torch.zeros(B, T, C, device="cuda", dtype=torch.float16)
    lo = torch.zeros_like(hi)
    phase = 3.0 * torch.randn(B, T, C, device="cuda")
    ss = torch.stack([1.0 + 0.1 * torch.randn(B, C, device="cuda"), 0.2 * torch.randn(B, C, device="cuda")], dim=-1).contiguous()
    out = torch.empty(B, N, device="cuda")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs") or 6462.1
    res = {}
    for flag in ("0", "1"):
        os.environ["PG_STFT_PAIR"] = flag
        lm, _ = ops.stft(wave, n_fft, hop, ops.PG_STFT_LOGMAG, want_second=False, operand=(hi, lo, T * C))
        ms_s = timed(lambda: ops.stft(wave, n_fft, hop, ops.PG_STFT_LOGMAG, want_second=False, operand=(hi, lo, T * C)))
        wv, pk = ops.istft(lm, phase, ops.PG_SPEC_POLAR_LOG, n_fft, hop, normalize=False, check_finite=False, out=out, b_scale_shift=ss)
        ms_i = timed(lambda: ops.istft(lm, phase, ops.PG_SPEC_POLAR_LOG, n_fft, hop, normalize=False, check_finite=False, out=out, b_scale_shift=ss))
        res[flag] = dict(lm=lm.clone(), wv=wv.clone(), hi=hi.clone(), ms_s=ms_s, ms_i=ms_i)
        alg_s, alg_i = B * (4 * N + 4 * C * T), B * (8 * C * T + 4 * N)
        print(f"PG_STFT_PAIR={flag}: stft {ms_s:.4f} ms ({alg_s / ms_s / 1e6 / hbm:.3f} of HBM peak, strict bytes)   "
              f"istft {ms_i:.4f} ms ({alg_i / ms_i / 1e6 / hbm:.3f})", flush=True)
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    print(f"pair vs scalar: logmag rel {rel(res['1']['lm'], res['0']['lm']):.2e}  wave rel {rel(res['1']['wv'], res['0']['wv']):.2e}  "
          f"hi plane max diff {float((res['1']['hi'].float() - res['0']['hi'].float()).abs().max()):.2e}")
    print(f"speed-up: stft {res['0']['ms_s'] / res['1']['ms_s']:.2f}x  istft {res['0']['ms_i'] / res['1']['ms_i']:.2f}x")


if __name__ == "__main__":
    main()
