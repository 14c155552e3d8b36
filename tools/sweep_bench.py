"""Throughput of the full path at the STFT shape sweep of BASELINE.json config 5 (n_fft 512/1024/2048, hop n_fft/4,
4 s clips @ 44.1 kHz), resident inputs, CUDA events.   python tools/sweep_bench.py [--clips 256] [--precision f16mix]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unet-phasegen_b200")]
import torch
import model as pg_model
from phasegen import synth
from phasegen.pipeline import PhaseGenPipeline

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=256)
ap.add_argument("--precision", default="f16mix")
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
for n_fft in (512, 1024, 2048):
    hop, C = n_fft // 4, n_fft // 2
    T = synth.frames_for(4.0, 44100, hop); N = (T - 1) * hop
    B = a.clips if n_fft < 2048 else a.clips // 2               # C = 1024 activations: halve the batch to bound memory
    torch.manual_seed(0)
    net = pg_model.UNetModel(C, 2 * C).cuda()
    pipe = PhaseGenPipeline(net, n_fft, hop, precision=a.precision, per_clip=True, phase_only=True)
    wave = synth.synthetic_waves(B, N, 44100, seed=1).cuda()
    for _ in range(3):
        pipe(wave)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        pipe(wave)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"n_fft {n_fft} hop {hop} T {T} C {C} clips {B} {a.precision}: {ms:8.2f} ms/step  {B * N / 44100 / (ms / 1e3):10.0f} audio-s/s", flush=True)
    del net, pipe, wave
    torch.cuda.empty_cache()
