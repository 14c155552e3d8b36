"""GPU diagnostic for the end-to-end path: where do the milliseconds between the resident step and the host-buffer step go?
(1) the resident 256-clip step; (2) the same clips as the ramped sub-batches run_host uses, still device-resident (sub-batch
tiling loss, no copies); (3) the resident step while a side stream keeps both PCIe directions busy (DMA interference);
(4) run_host one batch at a time; (5) run_host as a stream of batches, 1 sub-batch and ramped, over 5 and 40 steps (the
pipeline fill/drain is inside the timed region and amortises with the step count)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unet-phasegen_b200")]

import torch  # noqa: E402

import bench  # noqa: E402
import model as pg_model  # noqa: E402
from phasegen import synth  # noqa: E402
from phasegen.pipeline import PhaseGenPipeline  # noqa: E402


def timed(fn, n, finish=None):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    if finish:
        finish()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    T, N, clip_s = bench.workload_geometry()
    B, C = 256, bench.N_FFT // 2
    torch.manual_seed(1234)
    net = pg_model.UNetModel(C, 2 * C).cuda()
    synth.randomize_norm_affine(net, seed=7)
    pipe = PhaseGenPipeline(net, bench.N_FFT, bench.HOP, precision=bench.DEFAULT_PRECISION, per_clip=True, phase_only=True, normalize=True)
    host_in = synth.synthetic_waves(B, N, bench.SR, seed=100).pin_memory()
    host_outs = [torch.empty(B, N).pin_memory() for _ in range(2)]
    wave = host_in.cuda()
    sizes = pipe.suggest_chunks(B, N, wave.device)
    for _ in range(3):
        pipe(wave)
    print(f"(1) resident, one batch of {B}:            {timed(lambda: pipe(wave), 10):7.3f} ms")
    parts = list(torch.split(wave, sizes))
    outs = [torch.empty_like(p) for p in parts]

    def chunked():
        for p, o in zip(parts, outs):
            pipe(p, wave_out=o)
    for _ in range(2):
        chunked()
    print(f"(2) resident, sub-batches {sizes}: {timed(chunked, 10):7.3f} ms")
    side_a, side_b = torch.cuda.Stream(), torch.cuda.Stream()
    d_a, d_b = torch.empty_like(wave), torch.empty_like(wave)
    stop = [False]

    def with_copies():
        with torch.cuda.stream(side_a):
            d_a.copy_(host_in, non_blocking=True)
        with torch.cuda.stream(side_b):
            host_outs[1].copy_(d_b, non_blocking=True)
        pipe(wave)
    for _ in range(2):
        with_copies()
    print(f"(3) resident + 182 MB H2D and D2H in flight: {timed(with_copies, 10):7.3f} ms")
    torch.cuda.synchronize()
    for _ in range(2):
        pipe.run_host(host_in, host_outs[0], chunks=sizes)
    print(f"(4) run_host, one batch at a time, ramped:  {timed(lambda: pipe.run_host(host_in, host_outs[0], chunks=sizes), 10):7.3f} ms")
    for chunks, name in ((1, "1 sub-batch"), (sizes, "ramped")):
        for steps in (5, 40):
            pend, i = [None, None], [0]

            def step():
                k = i[0] & 1
                if pend[k] is not None:
                    pend[k].synchronize()
                pend[k] = pipe.run_host(host_in, host_outs[k], chunks=chunks, pipelined=True)
                i[0] += 1

            def finish():
                for ev in pend:
                    if ev is not None:
                        torch.cuda.current_stream().wait_event(ev)
            for _ in range(2):
                step()
            finish()
            print(f"(5) run_host stream of batches, {name:11s}, {steps:2d} steps: {timed(step, steps, finish):7.3f} ms")


if __name__ == "__main__":
    main()
