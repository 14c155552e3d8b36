"""A/B of the inference pipeline at the bench shape in ONE process (same GPU, same thermal state): fused epilogues vs the
two-pass norm form, interleaved, with CUDA-event times per kernel family.

    python tools/pipe_ab.py [--B 256] [--reps 3] [--steps 5] [--prec f16mix]
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unet-phasegen_b200")]

import torch  # noqa: E402

import model as pg_model  # noqa: E402
from phasegen import _lib, synth  # noqa: E402
from phasegen.pipeline import PhaseGenPipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--prec", default="f16mix")
    a = ap.parse_args()
    n_fft, hop, T = 1024, 256, 696
    N = (T - 1) * hop
    torch.manual_seed(1234)
    net = pg_model.UNetModel(n_fft // 2, n_fft).cuda()
    synth.randomize_norm_affine(net, seed=7)
    wave = synth.synthetic_waves(a.B, N, 44100, seed=100).cuda()
    pipes = {"fused": PhaseGenPipeline(net, n_fft, hop, precision=a.prec), "two-pass": PhaseGenPipeline(net, n_fft, hop, precision=a.prec, executor_kw={"fuse": False})}
    for p in pipes.values():
        for _ in range(3):
            p(wave)
    torch.cuda.synchronize()
    for rep in range(a.reps):
        for name, p in pipes.items():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                p(wave)
            e1.record(); torch.cuda.synchronize()
            print(f"rep {rep} {name:9s} {e0.elapsed_time(e1) / a.steps:8.3f} ms/step", flush=True)
    # per-entry-point times (events around every C-ABI call of one step)
    orig = _lib.call
    for name, p in pipes.items():
        evs = []

        def timed(fn, *args, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = orig(fn, *args, **kw); e1.record()
            evs.append((fn, e0, e1))
            return r
        _lib.call = timed
        p(wave); torch.cuda.synchronize(); evs.clear()
        p(wave); torch.cuda.synchronize()
        _lib.call = orig
        agg = collections.OrderedDict()
        seq = []
        for fn, e0, e1 in evs:
            ms = e0.elapsed_time(e1)
            agg[fn] = agg.get(fn, 0.0) + ms
            if fn == "pg_conv_tc":
                seq.append(round(ms, 3))
        print(name, {k: round(v, 3) for k, v in agg.items()}, "convs:", seq, "sum", round(sum(agg.values()), 3), flush=True)


if __name__ == "__main__":
    main()
