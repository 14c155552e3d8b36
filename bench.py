"""Benchmark of the magnitude -> phase -> waveform hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (configs[1] of BASELINE.json): 256 synthetic 4 s / 44.1 kHz mono clips per GPU,
n_fft 1024, hop 256 (T = 696 frames after padding to a multiple of 8, N = 177,920 samples),
STFT -> U-Net (C = 512, random-init weights, per-clip train-mode norm statistics = the
demo.py batch-1 loop) -> ISTFT with peak normalisation.  One "step" = one pass over the
batch.  metric = audio-seconds processed per wall second, whole job (all ranks).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "unet-phasegen_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

SR, N_FFT, HOP, SECONDS = 44100, 1024, 256, 4.0
CLIPS_PER_GPU = 256
METRIC = "audio_seconds_per_second_mag_to_phase_to_wave"
UNIT = "audio-s/s"


def workload_geometry():
    from phasegen import synth
    T = synth.frames_for(SECONDS, SR, HOP)
    N = (T - 1) * HOP
    return T, N, N / SR


def unet_flops_per_clip(C, T, phase_only):
    """Algorithmic MACs*2 of the eight convolutions, per clip."""
    L1 = T // 2 + 1; L2 = L1 - 3; L3 = L2 // 2 - 2; L4 = (L3 - 1) // 2
    c2, c4 = 2 * C, 4 * C
    per_layer = {
        "d1": C * c2 * 32 * L1, "d2": c2 * c2 * 8 * L2, "d3": c2 * c2 * 8 * L3, "d4": c2 * c4 * 4 * L4,
        "u4": c4 * c2 * 5 * L4, "u3": c4 * c2 * 8 * L3, "u2": c4 * c2 * 8 * L2,
        "u1": c4 * (C if phase_only else c2) * 32 * L1,
    }
    return {k: 2 * v for k, v in per_layer.items()}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_step(n_clips, seed=0, time_budget_s=None):
    """The reference's CPU path (oracle port: numpy STFT/ISTFT restating librosa + the reference
    U-Net arithmetic through torch CPU fp32 with oneDNN off, batch 1 per clip like demo.py:33-42)
    on `n_clips` clips of the bench workload.  Returns (seconds, clips done)."""
    import numpy as np
    import torch
    from oracle import stft_np, unet_torch
    from phasegen import synth
    T, N, _ = workload_geometry()
    C = N_FFT // 2
    torch.set_num_threads(os.cpu_count() or 1)
    sd = unet_torch.random_state_dict(C, seed=seed)
    waves = synth.synthetic_waves(n_clips, N, SR, seed=seed).numpy()
    t0 = time.perf_counter()
    done = 0
    for i in range(n_clips):
        S = stft_np.stft(waves[i], N_FFT, HOP)[1:]
        lm = np.log1p(np.abs(S)).astype(np.float32)
        out = unet_torch.unet_forward(sd, torch.from_numpy(lm)[None], torch.float32, per_clip_bn=True)[0].numpy()
        stft_np.generate_audio(stft_np.polar_to_complex(lm, out[:C]), SR, HOP, is_stft=True)
        done += 1
        if time_budget_s is not None and time.perf_counter() - t0 > time_budget_s:
            break
    return time.perf_counter() - t0, done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T, N, clip_s = workload_geometry()
    cores = os.cpu_count() or 1
    n = 2
    for _ in range(args.warmup):
        cpu_reference_step(1)
    times = []
    for _ in range(args.steps):
        dt, done = cpu_reference_step(n)
        times.append(dt / done)
    per_clip = statistics.mean(times)
    value = clip_s / per_clip
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_clip * n * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{n} clips/step of the batched-inference workload (4 s @ 44.1 kHz, n_fft {N_FFT}, hop {HOP}, "
                                   f"T {T}, C {N_FFT // 2}), batch-1 per clip like demo.py",
                       "timing": "host perf_counter; CPU only"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} clips per step x {args.steps} steps, oracle port (numpy STFT/ISTFT + torch-CPU fp32 U-Net, oneDNN off)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def synthetic_train_pairs(B, C, T, seed, device):
    """"MedleyDB-shaped" pairs: z = complex N(0,1) * sigma(f) with a 1/f tilt;
    ch0 = log1p|z|, ch1 = angle z -- what data.py:39-47 yields.  Channels-last [B, T, C]."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sigma = 4.0 / (1.0 + torch.arange(C, dtype=torch.float32) / 32.0)
    re = torch.randn(B, T, C, generator=g) * sigma
    im = torch.randn(B, T, C, generator=g) * sigma
    return torch.log1p(torch.sqrt(re * re + im * im)).contiguous().to(device), torch.atan2(im, re).contiguous().to(device)


def run_train(args):
    """BASELINE.json config 3: train.py step (forward, cos/sin/mag loss, backward, Adam), UNetModel(1024, 2048),
    128-frame pairs, batch 32 per GPU, bf16 tensor-core products with fp32 accumulation and fp32 master weights,
    data-parallel over NCCL (one gradient all-reduce per step)."""
    import torch
    import torch.distributed as dist
    import model as pg_model
    from phasegen import _lib
    from phasegen.train import TrainStep
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    C, T, B = args.train_c, 128, args.train_batch
    torch.manual_seed(1234)
    net = pg_model.UNetModel(C, 2 * C).to(dev)
    train_prec = "bf16x3" if args.train_fp32 else (args.precision or "bf16")
    step = TrainStep(net, B, T, dev, precision=train_prec)
    lm, ph = synthetic_train_pairs(B, C, T, 100 + rank, dev)
    host = [t.cpu().pin_memory() for t in (lm, ph)]
    W = max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    losses = []

    def step_resident():
        step(lm, ph)

    from phasegen.train import HostBatchFeeder
    feeder = HostBatchFeeder(tuple(lm.shape), dev)
    loss_host = [torch.zeros(4).pin_memory() for _ in range(2)]
    loss_ev = [None, None]
    e2e_i = [0]

    def step_e2e():
        # the public host path: pinned host batch -> copy stream -> step; the loss comes back through a pinned
        # buffer and is read one step late, so neither copy stalls the GPU
        k = e2e_i[0] & 1
        if loss_ev[k] is not None:
            loss_ev[k].synchronize()
            losses.append(float(loss_host[k][0]))
        a, b = feeder.upload(host[0], host[1])
        l3 = step(a, b)
        feeder.done()
        loss_host[k].copy_(l3, non_blocking=True)
        loss_ev[k] = torch.cuda.Event(); loss_ev[k].record()
        e2e_i[0] += 1

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                        # nvidia-smi needs ~0.1 s to deliver its first sample: start it
    for _ in range(W):                                         # before the warm-up so the timed region is covered
        step_resident()
    l0 = _lib.launches
    ms = timed(step_resident, args.steps)
    launches = _lib.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps)
    value = world * B / (ms / args.steps / 1e3)
    e2e = world * B / (ms_e2e / args.steps / 1e3)
    flops = 3 * sum(unet_flops_per_clip(C, T, False).values()) - unet_flops_per_clip(C, T, False)["d1"]   # no dgrad for d1
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = peaks.get("bf16_tflops_sustained") or 1590.0
    achieved = flops * B / (ms / args.steps / 1e3) / 1e12
    if rank == 0:
        line = {"metric": "train_samples_per_second", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": f"{train_prec}-fp32acc (fp32 master weights, fp32 Adam)", "data": "synthetic",
                "config": {"workload": f"train.py step: UNetModel({C},{2 * C}) on [B,2,{C},{T}] log-mag/phase pairs, batch {B}/GPU, "
                                       "forward + cos/sin/mag loss + backward + Adam(lr 1e-3)", "parallelism": f"dp{world}",
                           "timing": "CUDA events on the launch stream, max over ranks", "l2_policy": "weights + optimizer state (>10 GB) exceed L2"},
                "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": 2 * B * T * C * 4, "d2h_bytes_per_step": 16, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"bound": "tensor", "kernel": "whole step (conv_tc + wgrad_tc dominate)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": achieved / peak, "traffic": None, "note": "algorithmic FLOPs: fwd + dgrad + wgrad of the 8 convolutions"},
                "loss_trace": losses[-3:]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


DTYPE_NAMES = {"bf16x3": "bf16x3-fp32acc", "bf16": "bf16-fp32acc", "fp32_simt": "f32", "f16x3": "f16x3-fp32acc",
               "f16x2": "f16x2-fp32acc", "f16mix": "f16x3/f16x2-fp32acc (two-product fp16 on d1,u1,u2)",
               "f16mix1": "f16x3/f16x2/f16-fp32acc (two-product fp16 on d1,u2; single fp16 product = TF32 operand rounding on u1)"}
MMA_NOTES = {"bf16x3": "the fp32-class mode issues 3 bf16 MMAs per algorithmic MAC, so tensor-pipe work is 3x this figure",
             "f16x3": "3 fp16 MMAs per algorithmic MAC, so tensor-pipe work is 3x this figure",
             "f16x2": "2 fp16 MMAs per algorithmic MAC (weights rounded to fp16), tensor-pipe work is 2x this figure",
             "f16mix": "2 fp16 MMAs per algorithmic MAC on d1/u1/u2 (80 % of the MACs), 3 on the other five layers: "
                       "tensor-pipe work is 2.2x this figure",
             "f16mix1": "1 fp16 MMA per algorithmic MAC on u1 (36 % of the MACs), 2 on d1/u2 (44 %), 3 on the other five layers: "
                        "tensor-pipe work is 1.85x this figure",
             "bf16": "1 bf16 MMA per algorithmic MAC", "fp32_simt": ""}


PARITY_NOTES = {
    "f16mix": "full path at this exact shape vs the float64 oracle (tests/test_gpu_pipeline.py::test_shape_sweep_full_path_one_clip): "
              "predicted phase rel-L2 5.1e-4 (bound 1e-3), waveform SNR within 0.1 dB, STFT log-magnitude < 1e-4",
    "bf16x3": "full path at this exact shape vs the float64 oracle: predicted phase rel-L2 9.7e-5 (bound 1e-3), waveform SNR "
              "within 0.1 dB, STFT log-magnitude < 1e-4",
    "f16mix1": "full path at this exact shape vs the float64 oracle: predicted phase rel-L2 5.7e-4 (bound 1e-3), waveform SNR within 0.1 dB",
    "f16x3": "predicted phase rel-L2 vs float64 oracle 1e-4 at C=512 (bound 1e-3)",
    "f16x2": "predicted phase rel-L2 vs float64 oracle 7e-4 at C=512 (bound 1e-3: no margin, not the default)",
    "bf16": "loose mode: predicted phase rel-L2 ~1e-2", "fp32_simt": "exact fp32 CUDA-core convolutions"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="phasegen", choices=["phasegen", "reference"])
    ap.add_argument("--clips", type=int, default=CLIPS_PER_GPU, help="clips per GPU per step")
    ap.add_argument("--precision", default=None, choices=["bf16x3", "bf16", "fp32_simt", "f16x3", "f16mix", "f16mix1", "f16x2"],
                    help="inference default: f16mix (fp32-class, within the 1e-3 phase bound; the all-three-product bf16x3 "
                         "figure is reported beside it); training default: bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=0,
                    help="sub-batches whose copies overlap GPU work in the e2e leg (0 = wave-aligned sizes from PhaseGenPipeline.suggest_chunks)")
    ap.add_argument("--workload", default="infer", choices=["infer", "train"],
                    help="infer = BASELINE config 2 (headline, default); train = config 3 (train.py step)")
    ap.add_argument("--train-batch", type=int, default=32)
    ap.add_argument("--train-c", type=int, default=1024)
    ap.add_argument("--train-fp32", action="store_true", help="train with the fp32-class bf16x3 products instead of bf16")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "train":
        run_train(args)
        return

    import torch
    import torch.distributed as dist
    import model as pg_model
    from phasegen import _lib, synth
    from phasegen.pipeline import PhaseGenPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the phasegen path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)

    if args.precision is None:
        args.precision = "f16mix"
    T, N, clip_s = workload_geometry()
    C, B = N_FFT // 2, args.clips
    torch.manual_seed(1234)
    net = pg_model.UNetModel(C, 2 * C).to(dev)
    synth.randomize_norm_affine(net, seed=7)
    pipe = PhaseGenPipeline(net, N_FFT, HOP, precision=args.precision, per_clip=True, phase_only=True, normalize=True)
    host_in = synth.synthetic_waves(B, N, SR, seed=100 + rank).pin_memory()
    host_out = torch.empty(B, N, dtype=torch.float32).pin_memory()
    wave = host_in.to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        return pipe(wave)

    def step_e2e():
        # the public host-buffer call: pinned host wave in, pinned host wave out, copies inside
        pipe.run_host(host_in, host_out, chunks=e2e_chunks)

    e2e_chunks = pipe.suggest_chunks(B, N, dev) if args.e2e_chunks == 0 else args.e2e_chunks
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                        # nvidia-smi needs ~0.1 s to deliver its first sample: start it
    for _ in range(W):                                         # before the warm-up so the timed region is covered
        step_resident()
    l0 = _lib.launches
    ms = timed(step_resident, args.steps)
    launches = _lib.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    audio_s = world * B * clip_s
    value = audio_s / (ms / args.steps / 1e3)
    e2e = audio_s / (ms_e2e / args.steps / 1e3)

    # ---- roofline of the dominant kernel (conv_tc_kernel): CUDA events around every launch
    from phasegen import ops
    conv_ms = {"n": 0, "ms": 0.0}
    pending = []
    orig = ops.conv_tc

    def conv_timed(desc, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig(desc, *a); e1.record()
        pending.append((e0, e1))
    # ... and of the two HBM-class kernels (stft_kernel, istft_kernel), same method
    hbm_pending = {"stft_kernel": [], "istft_kernel": []}
    orig_stft, orig_istft = ops.stft, ops.istft

    def stft_timed(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = orig_stft(*a, **k); e1.record()
        hbm_pending["stft_kernel"].append((e0, e1))
        return r

    def istft_timed(a, b, mode, n_fft, hop, normalize=True, check_finite=True, out=None):
        # time the ISTFT kernel alone: the peak normalisation is a separate launch (and a separate +8N-byte pass)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); w, pk = orig_istft(a, b, mode, n_fft, hop, normalize=False, check_finite=False, out=out); e1.record()
        hbm_pending["istft_kernel"].append((e0, e1))
        if normalize:
            _lib.call("pg_peak_normalize", ops._ptr(w), ops._ptr(pk), w.shape[0], w.shape[1], ops._stream())
        return w, pk
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
    except OSError:
        pass
    roof = None
    if args.precision != "fp32_simt":
        ops.conv_tc = conv_timed
        ops.stft, ops.istft = stft_timed, istft_timed
        step_resident()                                        # one instrumented warm-up pass, not counted
        torch.cuda.synchronize()
        pending.clear()
        for v in hbm_pending.values():
            v.clear()
        for _ in range(args.steps):
            step_resident()
        torch.cuda.synchronize()
        ops.conv_tc = orig
        ops.stft, ops.istft = orig_stft, orig_istft
        tot_ms = sum(a.elapsed_time(b) for a, b in pending)
        n_launch = len(pending)
        flops_step = sum(unet_flops_per_clip(C, T, True).values()) * B
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1590.0 if not peaks else None) or 1590.0
        achieved = flops_step * args.steps / (tot_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": None, "launches": n_launch, "avg_launch_ms": tot_ms / max(n_launch, 1),
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                "note": "algorithmic FLOPs (8 convs, phase-only last layer); " + MMA_NOTES[args.precision],
                "conv_share_of_step": (tot_ms / args.steps) / (ms / args.steps)}
        roof["traffic"] = traffic.get("conv_tc_kernel", {}).get("bytes_per_launch")
        roof["traffic_source"] = traffic.get("conv_tc_kernel", {}).get("source")
        # algorithmic bytes per clip: STFT 4N + 4CT (+ 4CT for the operand planes it also writes),
        # ISTFT 8CT + 4N
        hbm_peak = peaks.get("hbm_gbs") or 6650.0
        alg = {"stft_kernel": B * (4 * N + 8 * C * T), "istft_kernel": B * (8 * C * T + 4 * N)}
        roof_hbm = []
        for kname, evs in hbm_pending.items():
            if not evs:
                continue
            k_ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
            gbs = alg[kname] / (k_ms / 1e3) / 1e9
            roof_hbm.append({"bound": "hbm", "kernel": kname, "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                             "avg_launch_ms": k_ms, "algorithmic_bytes_per_launch": alg[kname],
                             "traffic": traffic.get(kname, {}).get("bytes_per_launch"),
                             "note": "issue-bound, not HBM-bound: ~2000 warp instructions per 7 KB frame (DESIGN.md 4.2)"})
        roof["hbm_kernels"] = roof_hbm

    # the all-three-product form (bf16x3: every layer at ~2^-16 per product) measured beside the default
    alt = alt1 = None
    if args.precision in ("f16mix", "f16mix1"):
        def side_leg(prec):
            p = PhaseGenPipeline(net, N_FFT, HOP, precision=prec, per_clip=True, phase_only=True, normalize=True)
            for _ in range(2):
                p(wave)
            t = timed(lambda: p(wave), args.steps)
            return {"precision": DTYPE_NAMES[prec], "value": audio_s / (t / args.steps / 1e3), "unit": UNIT, "ms_per_step": t / args.steps,
                    "parity": PARITY_NOTES.get(prec)}
        alt = side_leg("bf16x3")
        if args.precision == "f16mix":
            # one step further inside the same 1e-3 bound: the last layer with a single fp16 product (TF32 operand rounding)
            alt1 = side_leg("f16mix1")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:     # reported at N = 1 only
        dt, done = cpu_reference_step(16, time_budget_s=12.0)
        cpu = {"value": clip_s * done / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"{done} clips of the same workload, batch-1 per clip (demo.py loop), numpy STFT/ISTFT + torch-CPU fp32 U-Net (oneDNN off)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": DTYPE_NAMES[args.precision],
                "data": "synthetic",
                "config": {"workload": f"batched inference: {B} clips/GPU x 4 s @ 44.1 kHz (N {N}), n_fft {N_FFT}, hop {HOP}, "
                                       f"T {T}, U-Net C {C} (153 M params, random init), STFT->U-Net->ISTFT+peak-normalise",
                           "norm_statistics": "per clip (demo.py batch-1 semantics)", "last_layer": "phase-only",
                           "l2_policy": f"inputs larger than L2 ({B * N * 4 / 1e6:.0f} MB wave, >2 GB of activations per step)",
                           "timing": "CUDA events on the launch stream, max over ranks"},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * N * 4, "d2h_bytes_per_step": B * N * 4,
                        "ms_per_step": ms_e2e / args.steps, "sub_batches": e2e_chunks},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
                "all_three_product_form": alt, "single_product_last_layer_form": alt1, "parity": PARITY_NOTES.get(args.precision)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
